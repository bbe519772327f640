// 3x3 convolution on token-major (NHWC, bf16) data as an implicit GEMM on tcgen05, without im2col:
// a tile of TH x TW output pixels plus its 1-pixel halo is staged ONCE in shared memory as a K-major UMMA
// operand image whose "rows" are halo positions in raster order (pitch LW = TW+2).  Because consecutive rows of a
// SWIZZLE_NONE image are 16 bytes apart, the A operand of filter tap (dy,dx) is the same image with the descriptor
// start address advanced by (dy*LW+dx)*16 bytes -- nine taps x Cin/16 k-steps accumulate into one TMEM tile.
// Rows that fall on halo columns are computed and discarded ((TH-1)*LW+TW <= 128).
// The weights of one N-slice (all 9 taps) stay resident in shared memory for the whole persistent CTA.
// Epilogue: +bias, *scale, +residual, and for the up-sampler nn.PixelShuffle(2) folded into the store address.
// Replaces RDSTB.conv (LFF, reference rdst_variations.py:444-445), conv_after_body (:1349-1350) and the UpSampler
// convs + PixelShuffle (networks/common.py:129-132).
#include "common.cuh"
#include "umma.cuh"
#include "tma.cuh"

namespace rdst {
using namespace umma;

__device__ __forceinline__ uint32_t cpk2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 32 contiguous bytes (16 bf16 channels) in one 256-bit store: a full sector per lane
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                              uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
__device__ __forceinline__ float2 cup2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

constexpr int CONV_NP_MAX = 200;      // staged halo positions (2*LW+2+128 for TW<=32)

struct ConvGeom {
  int B, H, W, TW, TH, LW, NP, ntx, nty;
  int shuffle;
  float out_scale;
  float out_bias;      // NT == 16 ("last conv" mode): img = (acc + bias0) * out_scale + out_bias, fp32 NCHW, 1 channel
};

// K-major SWIZZLE_128B operand (rows of 128 bytes = 64 bf16, 8-row groups of 1024 bytes): what a TMA box with a
// 64-element inner extent and CU_TENSOR_MAP_SWIZZLE_128B leaves in shared memory.  A k-step of 16 advances the start by 32 B.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// same for 64-byte rows (32 bf16): SWIZZLE_64B, 8-row groups of 512 bytes
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}

template <int CIN, int NT>
struct ConvCfg {
  static constexpr int NCH = CIN / 8;
  static constexpr int W_BYTES = 9 * CIN * NT * 2;
  static constexpr int A_BYTES = NCH * CONV_NP_MAX * 16;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_A = W_BYTES;                  // staging buffers: tile n+1 is staged while tile n's MMAs run
  static constexpr int NSTG = CIN == 64 ? 3 : 2;         // Cin = 64 (TMA staging): a ring of three, loads run two tiles ahead
  static constexpr int OFF_BIAS = OFF_A + NSTG * A_BYTES;
  static constexpr int SMEM = OFF_BIAS + NT * 4;
  static constexpr int TMEM_COLS = NT <= 16 ? 32 : 2 * NT;   // two accumulators: tile n+1 runs under tile n's epilogue
  static_assert(NCH % 4 == 0 && (NT == 16 || NT == 32 || NT == 64 || NT == 128), "shape");
};

constexpr int CONV_THREADS = 288;     // warps 0..7 stage and run the epilogue, warp 8 only issues MMAs

template <int CIN, int NT, bool DBG>       // DBG: clock64() phase stamps (rdst_debug_conv_timing)
__global__ void __launch_bounds__(CONV_THREADS)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __nv_bfloat16* __restrict__ X, int64_t ldx,
                  const uint8_t* __restrict__ wimg,
                  const float* __restrict__ bias, const __nv_bfloat16* __restrict__ R, int64_t ldr,
                  __nv_bfloat16* __restrict__ Y, int64_t ldy, ConvGeom g, unsigned long long* __restrict__ dbg) {
  using K = ConvCfg<CIN, NT>;
  // Cin = 64 (round 2): the halo tile is ONE TMA box, SWIZZLE_128B = the K-major A operand (row = halo position, tap (dy,dx)
  // = descriptor start advanced by (dy*LW+dx) rows; the hardware derives the swizzle phase from the address, so unaligned
  // row offsets need no base offset -- verified with the reconstruction conv).  Probe builds without the loads: the 16-byte
  // cp.async staging cost 152 of 369 us in the 80x64 up-conv, 35 of 96 us and 16 of 64 us in the 40x32 convs.
  constexpr bool TMA_A = CIN == 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];           // accumulator b complete
  __shared__ uint64_t full[3];          // TMA_A: halo tile of staging buffer b has landed
  __shared__ uint64_t wbar;             // weights landed (bulk async copies)
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.y;
  uint8_t* sW = smem + K::OFF_W;
  uint8_t* sA = smem + K::OFF_A;
  float* sBias = reinterpret_cast<float*>(smem + K::OFF_BIAS);

  if (warp == 0) tmem_alloc<K::TMEM_COLS>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&full[2], 1);
    mbar_init(&wbar, 1);
    fence_mbar_init();
    // the filter slice arrives by bulk async copies that overlap the staging of the first halo tile
    const uint8_t* src = wimg + (size_t)slice * K::W_BYTES;
    mbar_arrive_expect_tx(&wbar, K::W_BYTES);
    for (int off = 0; off < K::W_BYTES; off += 32768) bulk_g2s(sW + off, src + off, min(32768, K::W_BYTES - off), &wbar);
  }
  {
    if (!TMA_A)         // (TMA_A: stale rows beyond the staged positions only feed accumulator rows that are never stored)
      for (int i = tid; i < 2 * K::A_BYTES / 16; i += CONV_THREADS) *reinterpret_cast<uint4*>(sA + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < NT; i += CONV_THREADS) sBias[i] = (NT == 16) ? 0.f : bias[slice * NT + i];
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t aA = smem_u32(sA), aW = smem_u32(sW);
  const uint32_t lboA = (uint32_t)g.NP * 16;
  const int nps = (g.TH + 2) * g.LW;                       // halo positions actually staged
  const int row = tid & 127, part = tid >> 7;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int64_t ntiles = (int64_t)g.B * g.nty * g.ntx;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);

  // stage the halo tile of `tile` into staging buffer `buf` as a K-major image.  cp.async (LDGSTS) with zero fill for
  // the padding ring: nothing waits on the loads until the tile's MMAs are about to be issued (a register-staged copy
  // serialised three L2 round trips per warp, and the warp that also issued MMAs started its share last)
  auto stage = [&](int64_t tile, int buf) {
    if (TMA_A) {
      if (warp == 0) {
        if (elect_one()) {
          const int b = (int)(tile / (g.nty * g.ntx));
          const int tr = (int)(tile - (int64_t)b * g.nty * g.ntx);
          const int y0 = (tr / g.ntx) * g.TH, x0 = (tr % g.ntx) * g.TW;
          mbar_arrive_expect_tx(&full[buf], (uint32_t)nps * 128u);
          tma::load_4d(sA + (size_t)buf * K::A_BYTES, &mapX, 0, x0 - 1, y0 - 1, b, &full[buf]);
        }
        __syncwarp();
      }
      return;
    }
    if (warp >= 8) return;
    const int b = (int)(tile / (g.nty * g.ntx));
    const int tr = (int)(tile - (int64_t)b * g.nty * g.ntx);
    const int y0 = (tr / g.ntx) * g.TH, x0 = (tr % g.ntx) * g.TW;
    uint8_t* dst = sA + (size_t)buf * K::A_BYTES;
#pragma unroll 1
    for (int pg = warp; pg * 8 < nps; pg += 8) {
      const int pos = pg * 8 + (lane & 7);
      if (pos >= nps) continue;
      const int hy = pos / g.LW, hx = pos - hy * g.LW;
      const int y = y0 - 1 + hy, x = x0 - 1 + hx;
      const bool ok = y >= 0 && y < g.H && x >= 0 && x < g.W;
      const __nv_bfloat16* src = X + (ok ? (((int64_t)b * g.H + y) * g.W + x) * ldx : 0);
#pragma unroll
      for (int j = 0; j < K::NCH / 4; ++j)
        cp_async16(dst + (size_t)((lane >> 3) + 4 * j) * lboA + pos * 16, reinterpret_cast<const uint4*>(src) + (lane >> 3) + 4 * j,
                   ok ? 16u : 0u);
    }
    cp_async_commit();
  };
  auto issue = [&](int sb, int buf) {        // warp 8, one elected lane: staging buffer `sb` -> accumulator `buf`
    constexpr uint32_t idesc = make_idesc_bf16(128, NT, false, false);
    const uint32_t ab = aA + sb * K::A_BYTES;
    const uint32_t acc = tmem_u + buf * NT;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const uint32_t r0 = (uint32_t)((tap / 3) * g.LW + (tap % 3));
      const uint32_t a0 = ab + r0 * 16;
      const uint32_t w0 = aW + tap * (CIN * NT * 2);
#pragma unroll
      for (int ks = 0; ks < CIN / 16; ++ks)
        mma_bf16_ss(acc, TMA_A ? make_smem_desc_sw128(ab + r0 * 128 + ks * 32) : make_smem_desc(a0 + ks * 2 * lboA, lboA, 128),
                    make_smem_desc(w0 + ks * 2 * (NT * 16), NT * 16, 128), idesc, (tap | ks) > 0);
    }
    commit(&bar[buf]);
  };

  int dbg_n = 0;
  const bool dbg_on = DBG && dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 32;
#define RDST_TSTAMP()                                                         \
  do {                                                                        \
    if (DBG && dbg_on && dbg_n < 128) dbg[dbg_n++] = clock64();               \
  } while (0)
  pdl_launch_dependents();
  pdl_wait();                    // only the filter slice was touched so far; activations come from the previous kernel
  int buf = 0, sb = 0;                   // accumulator / staging buffer of the current tile (sb == buf unless TMA_A)
  uint32_t par0 = 0, par1 = 0, fpar = 0; // fpar: bit s = phase parity of full[s]
  if ((int64_t)blockIdx.x < ntiles) {
    stage(blockIdx.x, 0);
    if (TMA_A) {
      if ((int64_t)blockIdx.x + gridDim.x < ntiles) stage((int64_t)blockIdx.x + gridDim.x, 1);
      mbar_wait(&full[0], 0);
      fpar ^= 1u;
    }
    cp_async_wait_all();
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (warp_u == 8) {
      mbar_wait(&wbar, 0);             // filter slice has landed
      fence_after_sync();
      if (elect_one()) issue(0, 0);
      __syncwarp();
    }
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1, sb = TMA_A ? (sb + 1) % 3 : sb ^ 1) {
    const int b = (int)(tile / (g.nty * g.ntx));
    const int tr = (int)(tile - (int64_t)b * g.nty * g.ntx);
    const int y0 = (tr / g.ntx) * g.TH, x0 = (tr % g.ntx) * g.TW;
    const int64_t next = tile + gridDim.x;
    RDST_TSTAMP();   // tile start
    const int sn = TMA_A ? (sb + 1) % 3 : sb ^ 1;  // staging buffer of the next tile
    if (TMA_A) {                                   // two tiles ahead, into the buffer the previous tile's (complete) MMAs read
      if (next + gridDim.x < ntiles) stage(next + gridDim.x, (sb + 2) % 3);
    } else if (next < ntiles) {
      stage(next, sn);                             // in flight under the MMAs of the current tile
    }
    RDST_TSTAMP();   // next staged
    if (buf == 0) { mbar_wait(&bar[0], par0); par0 ^= 1; } else { mbar_wait(&bar[1], par1); par1 ^= 1; }
    RDST_TSTAMP();   // MMAs done
    if (TMA_A && next < ntiles) {
      mbar_wait(&full[sn], (fpar >> sn) & 1u);
      fpar ^= 1u << sn;
    }
    cp_async_wait_all();
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();          // next halo tile staged; every thread has drained the other accumulator (previous epilogue)
    fence_after_sync();
    if (next < ntiles && warp_u == 8) {            // the next tile's MMAs run under this tile's epilogue
      if (elect_one()) issue(sn, buf ^ 1);
      __syncwarp();
    }
    if (warp >= 8) continue;
    const uint32_t acc_addr = lane_addr + buf * NT;
    // ---------------- epilogue ----------------
    if constexpr (NT == 16) {
      // last conv: one real output channel, fp32 image
      const int oy = row / g.LW, ox = row - oy * g.LW;
      const int y = y0 + oy, x = x0 + ox;
      const bool ok = part == 0 && ox < g.TW && oy < g.TH && y < g.H && x < g.W;
      uint32_t v[4];
      tmem_ld_x4(acc_addr, v);
      wait_ld();
      if (ok)
        reinterpret_cast<float*>(Y)[((int64_t)b * g.H + y) * g.W + x] = (__uint_as_float(v[0]) + sBias[0]) * g.out_scale + g.out_bias;
    } else {
      const int oy = row / g.LW, ox = row - oy * g.LW;
      const int y = y0 + oy, x = x0 + ox;
      const bool ok = ox < g.TW && oy < g.TH && y < g.H && x < g.W;
      constexpr int NC = NT / 2;                            // columns per thread
      const int cb = part * NC;
      // shuffle: accumulator columns are (sub-pixel s, channel) with 64 channels per s
      const int sp = (slice * NT + cb) >> 6;                // sub-pixel index 2*dy+dx of this thread's columns
      int64_t tout;
      if (g.shuffle) tout = ((int64_t)(b * 2 * g.H + 2 * y + (sp >> 1))) * (2 * g.W) + 2 * x + (sp & 1);
      else tout = ((int64_t)b * g.H + y) * g.W + x;
      const int ncol0 = g.shuffle ? ((slice * NT + cb) & 63) : slice * NT + cb;   // first output channel of this thread
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(acc_addr + cb + c0, v);
        wait_ld();
        if (ok) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = (__uint_as_float(v[j]) + sBias[cb + c0 + j]) * g.out_scale;
          if (R != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(R + tout * ldr + ncol0 + c0);
            const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float2 t = cup2(rw[j]); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
          }
          st_global_256(Y + tout * ldy + ncol0 + c0, cpk2(f[0], f[1]), cpk2(f[2], f[3]), cpk2(f[4], f[5]), cpk2(f[6], f[7]),
                        cpk2(f[8], f[9]), cpk2(f[10], f[11]), cpk2(f[12], f[13]), cpk2(f[14], f[15]));
        }
      }
    }
    RDST_TSTAMP();   // epilogue done
  }
#undef RDST_TSTAMP
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<K::TMEM_COLS>(tmem);
}


// ------------------------------------------------------------------------------------------------------------------
// Final 64 -> 1 convolution as a "tap GEMM": P[pos][tap] = sum_c x[pos][c] * w[tap][c] for every staged halo position
// (one K = 64 GEMM with the 9 taps as N, padded to 16: 4 MMAs per 128 positions instead of the 36 of the implicit-GEMM
// form, which spent the tensor pipe on 15 zero output columns and was bound by its A reads), then the 9-point sum
// out[y][x] = sum_tap P[(y+dy, x+dx)][tap] on the CUDA cores from shared memory.  HBM-bound: 128 B read per pixel.
// Round 2: (a) the halo tile ((TH+2) x (TW+2) positions x 64 channels, up to 512 positions = 64 KB) arrives as ONE TMA
// box, SWIZZLE_128B, i.e. directly as a K-major A operand (row = position); out-of-range coordinates zero-fill = the
// convolution's padding.  The round-1 form staged it with 16-byte cp.async per thread and got 1.8 TB/s: with the loads
// removed the same kernel ran in 80 us, with the epilogue removed in 259 of 262 us (probe builds) -- per-thread async
// copies top out at ~10 B/clk/SM.  (b) a tile is up to four 128-position MMA blocks instead of two (halo over-read
// 1.23x instead of 1.77x, the fixed per-tile chain is paid per ~416 pixels).  (c) warp 8 is producer + MMA issuer on
// mbarriers (3-stage ring, two accumulator sets), warps 0-7 the epilogue: no CTA-wide barrier in the loop.
// ------------------------------------------------------------------------------------------------------------------
struct LastCfg {
  static constexpr int NBLK_MAX = 4;                                 // blocks of 128 halo positions per tile
  static constexpr int NP_MAX = 128 * NBLK_MAX;
  static constexpr int W_BYTES = 64 * 16 * 2;                        // [8][16 taps][8] bf16 (SWIZZLE_NONE B operand)
  static constexpr int NSTAGE = 3;                                   // halo tiles in flight
  // per launch: nblk position blocks -> stage = nblk * 16 KB (1024-byte aligned), [9][nblk*128] fp32 tap products
  __host__ __device__ static constexpr int a_bytes(int nblk) { return nblk * 128 * 128; }
  __host__ __device__ static constexpr int off_w(int nblk) { return NSTAGE * a_bytes(nblk); }
  __host__ __device__ static constexpr int off_p(int nblk) { return off_w(nblk) + W_BYTES; }
  __host__ __device__ static constexpr int smem(int nblk) { return off_p(nblk) + 9 * nblk * 128 * 4; }
  static constexpr int TMEM_COLS = 128;                              // 2 accumulator sets x 4 position blocks x 16 taps
};

__global__ void __launch_bounds__(CONV_THREADS)
last_conv_tap_kernel(const __grid_constant__ CUtensorMap mapX, const uint8_t* __restrict__ wimg, float* __restrict__ img, ConvGeom g) {
  using K = LastCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[K::NSTAGE], freeb[K::NSTAGE], tfull[2], tempty[2];
  __shared__ uint64_t wbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nps = (g.TH + 2) * g.LW;                       // staged halo positions
  const int nblk = (nps + 127) >> 7;                       // MMA blocks of 128 positions (the last one reads past nps: rows are
  const int NPp = nblk * 128;                              //   independent, those products are never summed)
  const int a_bytes = K::a_bytes(nblk);
  uint8_t* sA = smem;
  uint8_t* sW = smem + K::off_w(nblk);
  float* sP = reinterpret_cast<float*>(smem + K::off_p(nblk));

  if (warp == 0) tmem_alloc<K::TMEM_COLS>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < K::NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&freeb[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    mbar_init(&wbar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&wbar, K::W_BYTES);
    bulk_g2s(sW, wimg, K::W_BYTES, &wbar);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const int64_t ntiles = (int64_t)g.B * g.nty * g.ntx;
  const int ntl = (int)((ntiles - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);   // tiles of this CTA
  const int tpi = g.nty * g.ntx;

  pdl_launch_dependents();
  pdl_wait();
  if (warp == 8) {
    // ------------------------------- producer + MMA issuer -------------------------------
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t aA = smem_u32(sA), aW = smem_u32(sW);
    auto tma_tile = [&](int k) {                             // elected lane: halo box of this CTA's k-th tile -> stage k % NSTAGE
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      const int b = (int)(tile / tpi);
      const int tr = (int)(tile - (int64_t)b * tpi);
      const int y0 = (tr / g.ntx) * g.TH, x0 = (tr % g.ntx) * g.TW;
      const int s = k % K::NSTAGE;
      mbar_arrive_expect_tx(&full[s], (uint32_t)nps * 128u);
      tma::load_4d(sA + (size_t)s * a_bytes, &mapX, 0, x0 - 1, y0 - 1, b, &full[s]);
    };
    if (elect_one())
      for (int k = 0; k < ntl && k < K::NSTAGE; ++k) tma_tile(k);
    __syncwarp();
    mbar_wait(&wbar, 0);
    constexpr uint32_t idesc = make_idesc_bf16(128, 16, false, false);
    for (int i = 0; i < ntl; ++i) {
      const int s = i % K::NSTAGE, set = i & 1;
      mbar_wait(&full[s], (uint32_t)(i / K::NSTAGE) & 1u);
      if (i >= 2) mbar_wait(&tempty[set], (uint32_t)((i >> 1) - 1) & 1u);     // epilogue of tile i-2 has read this accumulator set
      fence_after_sync();
      if (elect_one()) {
        const uint32_t ab = aA + (uint32_t)s * (uint32_t)a_bytes;
#pragma unroll 1
        for (int blk = 0; blk < nblk; ++blk)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ss(tmem_u + set * 64 + blk * 16, make_smem_desc_sw128(ab + blk * (128 * 128) + ks * 32),
                        make_smem_desc(aW + ks * 2 * (16 * 16), 16 * 16, 128), idesc, ks > 0);
        commit(&tfull[set]);
        commit(&freeb[s]);
      }
      __syncwarp();
      if (i >= 1 && i + 2 < ntl) {                           // refill the stage of tile i-1 (its MMAs were issued an iteration ago)
        mbar_wait(&freeb[(i - 1) % K::NSTAGE], (uint32_t)((i - 1) / K::NSTAGE) & 1u);
        if (elect_one()) tma_tile(i + 2);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- epilogue: tap products -> shared memory -> 9-point sums -------------------------------
    const int row = tid & 127, part = tid >> 7;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int nout = g.TH * g.TW;
    const int twsh = 31 - __clz(g.TW);                       // TW is a power of two
    for (int i = 0; i < ntl; ++i) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
      const int b = (int)(tile / tpi);
      const int tr = (int)(tile - (int64_t)b * tpi);
      const int y0 = (tr / g.ntx) * g.TH, x0 = (tr % g.ntx) * g.TW;
      const int set = i & 1;
      mbar_wait(&tfull[set], (uint32_t)(i >> 1) & 1u);
      fence_after_sync();
      for (int blk = part; blk < nblk; blk += 2) {
        uint32_t v[16];
        tmem_ld_x16(lane_addr + set * 64 + blk * 16, v);
        wait_ld();
        const int pos = blk * 128 + row;
#pragma unroll
        for (int t = 0; t < 9; ++t) sP[t * NPp + pos] = __uint_as_float(v[t]);
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[set]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int o = tid; o < nout; o += 256) {
        const int oy = o >> twsh, ox = o & (g.TW - 1);
        const int y = y0 + oy, x = x0 + ox;
        if (y < g.H && x < g.W) {
          const float* pp = sP + oy * g.LW + ox;
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) acc += pp[t * NPp + (t / 3) * g.LW + (t % 3)];
          img[((int64_t)b * g.H + y) * g.W + x] = acc * g.out_scale + g.out_bias;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");        // sP may be overwritten by the next tile
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<K::TMEM_COLS>(tmem);
}


// ------------------------------------------------------------------------------------------------------------------
// LFF (Cin = 160 -> 64) on CTA pairs.  The 9 x 160 x 64 filter (184 KB) does not fit one SM next to the halo tiles; the
// single-CTA kernel therefore splits N over two CTAs and every CTA re-reads the whole A image for 32 output columns
// (the MMAs are bound by the 4 KB/k-step A read).  Here two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 per
// k-step: M = 256 = the two CTAs' own 128-position tiles, N = 64 with each CTA holding 32 columns of B -- every A
// byte is read once for all 64 output columns.  Leader (cluster rank 0) issues, completion is multicast to both.
// ------------------------------------------------------------------------------------------------------------------
struct PairCfg {
  static constexpr int CIN = 160, NT = 32, NOUT = 64;
  static constexpr int NCH = CIN / 8;
  static constexpr int W_BYTES = 9 * CIN * NT * 2;               // this CTA's half of the filter
  // halo tile = three TMA boxes: channels [0,64) and [64,128) as SWIZZLE_128B panels (128-byte rows = positions),
  // channels [128,160) as a SWIZZLE_64B panel (64-byte rows).  Each IS a K-major A operand; tap (dy,dx) = start + rows.
  static constexpr int PANEL = CONV_NP_MAX * 128;                // 25 600 B (a multiple of 1024)
  static constexpr int A_BYTES = (2 * PANEL + CONV_NP_MAX * 64 + 1023) / 1024 * 1024;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_A = W_BYTES;
  static constexpr int OFF_BIAS = OFF_A + 2 * A_BYTES;
  static_assert(OFF_A % 1024 == 0 && PANEL % 1024 == 0, "swizzled panels need 1024-byte alignment");
  static constexpr int SMEM = OFF_BIAS + NOUT * 4;
  static constexpr int TMEM_COLS = 128;                           // two accumulators of 64 columns
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const uint8_t* __restrict__ wimg,
                    const float* __restrict__ bias, const __nv_bfloat16* __restrict__ R, int64_t ldr,
                    __nv_bfloat16* __restrict__ Y, int64_t ldy, ConvGeom g) {
  using K = PairCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint64_t full[2];           // halo tile of staging buffer b has landed (TMA complete_tx)
  __shared__ uint64_t wbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  uint8_t* sW = smem + K::OFF_W;
  uint8_t* sA = smem + K::OFF_A;
  float* sBias = reinterpret_cast<float*>(smem + K::OFF_BIAS);

  if (warp == 0) tmem_alloc_pair<K::TMEM_COLS>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&wbar, 1);
    fence_mbar_init();
    const uint8_t* src = wimg + (size_t)rank * K::W_BYTES;
    mbar_arrive_expect_tx(&wbar, K::W_BYTES);
    for (int off = 0; off < K::W_BYTES; off += 32768) bulk_g2s(sW + off, src + off, min(32768, K::W_BYTES - off), &wbar);
  }
  // (no zero fill of the staging buffers: MMA rows beyond the staged positions only feed accumulator rows that are never
  // stored -- rows are independent -- and the padding ring is the TMA's out-of-range zero fill)
  for (int i = tid; i < K::NOUT; i += CONV_THREADS) sBias[i] = bias[i];
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t aA = smem_u32(sA), aW = smem_u32(sW);
  const int nps = (g.TH + 2) * g.LW;
  const int row = tid & 127, part = tid >> 7;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int64_t ntiles = (int64_t)g.B * g.nty * g.ntx;
  const int64_t npairs = (ntiles + 1) / 2;
  const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const bool leader = rank == 0;

  // round 2: the halo tile arrives by TMA (round 1 staged it with 16-byte cp.async per thread, ~10 B/clk/SM: the kernel ran
  // in 58 us with the loads removed and 81 us with them -- probe builds)
  auto stage = [&](int64_t tile, int buf) {      // warp 0, one elected lane; returns whether a tile was requested
    if (tile >= ntiles) return;
    const int b = (int)(tile / (g.nty * g.ntx));
    const int tr = (int)(tile - (int64_t)b * g.nty * g.ntx);
    const int y0 = (tr / g.ntx) * g.TH, x0 = (tr % g.ntx) * g.TW;
    uint8_t* dst = sA + (size_t)buf * K::A_BYTES;
    mbar_arrive_expect_tx(&full[buf], (uint32_t)nps * 320u);
    tma::load_4d(dst, &mapA, 0, x0 - 1, y0 - 1, b, &full[buf]);
    tma::load_4d(dst + K::PANEL, &mapA, 64, x0 - 1, y0 - 1, b, &full[buf]);
    tma::load_4d(dst + 2 * K::PANEL, &mapB, 128, x0 - 1, y0 - 1, b, &full[buf]);
  };
  auto issue = [&](int buf) {        // leader CTA, warp 8, one elected lane
    constexpr uint32_t idesc = make_idesc_bf16(256, K::NOUT, false, false);
    const uint32_t ab = aA + buf * K::A_BYTES;
    const uint32_t acc = tmem_u + buf * K::NOUT;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const uint32_t r0 = (uint32_t)((tap / 3) * g.LW + (tap % 3));        // first halo row of this tap
      const uint32_t w0 = aW + tap * (K::CIN * K::NT * 2);
#pragma unroll
      for (int ks = 0; ks < K::CIN / 16; ++ks) {
        const uint64_t da = ks < 8 ? make_smem_desc_sw128(ab + (ks >> 2) * K::PANEL + r0 * 128 + (ks & 3) * 32)
                                   : make_smem_desc_sw64(ab + 2 * K::PANEL + r0 * 64 + (ks & 1) * 32);
        mma_bf16_ss_pair(acc, da, make_smem_desc(w0 + ks * 2 * (K::NT * 16), K::NT * 16, 128), idesc, (tap | ks) > 0);
      }
    }
    commit_pair(&bar[buf]);
  };

  pdl_launch_dependents();
  pdl_wait();
  int buf = 0;
  uint32_t par0 = 0, par1 = 0;
  uint32_t fpar0 = 0, fpar1 = 0;
  if (warp == 0) {                               // both staging buffers are requested up front; afterwards a buffer is refilled
    if (elect_one()) {                           // (pair k+2) the moment the MMAs that read it (pair k) have completed
      stage(2 * cid + rank, 0);
      stage(2 * (cid + ncl) + rank, 1);
    }
    __syncwarp();
  }
  if (2 * cid + rank < ntiles) { mbar_wait(&full[0], fpar0); fpar0 ^= 1; }
  if (warp_u == 8) mbar_wait(&wbar, 0);          // this CTA's half of the filter has landed
  fence_proxy_async();
  fence_before_sync();
  cluster_sync();                                // both halo tiles staged, both filter halves resident
  fence_after_sync();
  if (cid < npairs && leader && warp_u == 8) {
    if (elect_one()) issue(0);
    __syncwarp();
  }
  for (int64_t pr = cid; pr < npairs; pr += ncl, buf ^= 1) {
    const int64_t tile = 2 * pr + rank;
    const bool have = tile < ntiles;
    const int b = (int)(tile / (g.nty * g.ntx));
    const int tr = (int)(tile - (int64_t)b * g.nty * g.ntx);
    const int y0 = (tr / g.ntx) * g.TH, x0 = (tr % g.ntx) * g.TW;
    const int64_t npr = pr + ncl;
    if (buf == 0) { mbar_wait(&bar[0], par0); par0 ^= 1; } else { mbar_wait(&bar[1], par1); par1 ^= 1; }
    if (warp == 0) {                             // this pair's MMAs are done with buffer `buf`: request the pair after next
      if (elect_one()) stage(2 * (npr + ncl) + rank, buf);
      __syncwarp();
    }
    if (2 * npr + rank < ntiles) {               // the next halo tile has landed
      if (buf == 0) { mbar_wait(&full[1], fpar1); fpar1 ^= 1; } else { mbar_wait(&full[0], fpar0); fpar0 ^= 1; }
    }
    fence_proxy_async();
    fence_before_sync();
    cluster_sync();           // both CTAs: next halo tiles staged, other accumulator drained
    fence_after_sync();
    if (npr < npairs && leader && warp_u == 8) {
      if (elect_one()) issue(buf ^ 1);
      __syncwarp();
    }
    if (warp >= 8 || !have) continue;
    const uint32_t acc_addr = lane_addr + buf * K::NOUT;
    const int oy = row / g.LW, ox = row - oy * g.LW;
    const int y = y0 + oy, x = x0 + ox;
    const bool ok = ox < g.TW && oy < g.TH && y < g.H && x < g.W;
    constexpr int NC = K::NOUT / 2;                        // columns per thread
    const int cb = part * NC;
    const int64_t tout = ((int64_t)b * g.H + y) * g.W + x;
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(acc_addr + cb + c0, v);
      wait_ld();
      if (ok) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = (__uint_as_float(v[j]) + sBias[cb + c0 + j]) * g.out_scale;
        if (R != nullptr) {
          const uint4* rp = reinterpret_cast<const uint4*>(R + tout * ldr + cb + c0);
          const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
          const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float2 t = cup2(rw[j]); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
        }
        st_global_256(Y + tout * ldy + cb + c0, cpk2(f[0], f[1]), cpk2(f[2], f[3]), cpk2(f[4], f[5]), cpk2(f[6], f[7]),
                      cpk2(f[8], f[9]), cpk2(f[10], f[11]), cpk2(f[12], f[13]), cpk2(f[14], f[15]));
      }
    }
  }
  fence_before_sync();
  cluster_sync();             // the peer's tensor core may still be reading this CTA's shared memory / TMEM
  if (warp == 0) tmem_dealloc_pair<K::TMEM_COLS>(tmem);
}

static int launch_conv_pair(const void* x, int64_t ldx, const void* wimg, const float* bias, const void* r, int64_t ldr,
                            void* y, int64_t ldy, ConvGeom g, int sms, cudaStream_t st) {
  using K = PairCfg;
  cudaError_t e = cudaFuncSetAttribute(conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
  if (e != cudaSuccess) { set_error("rdst_conv3x3_fwd_bf16_tc: smem attr (%d B): %s", K::SMEM, cudaGetErrorString(e)); return RDST_E_CUDA; }
  const int64_t ntiles = (int64_t)g.B * g.nty * g.ntx;
  const int64_t npairs = (ntiles + 1) / 2;
  int64_t ncl = sms / 2;
  if (ncl > npairs) ncl = npairs;
  if (ncl < 1) ncl = 1;
  const CUtensorMap* ma = get_act_tmap(x, ldx, g.B, g.H, g.W, K::CIN, g.LW, g.TH + 2, 64);
  const CUtensorMap* mb = get_act_tmap(x, ldx, g.B, g.H, g.W, K::CIN, g.LW, g.TH + 2, 32);
  if (!ma || !mb) return RDST_E_CUDA;
  e = launch_pdl(conv3x3_pair_kernel, dim3((unsigned)(2 * ncl)), dim3(CONV_THREADS), (size_t)K::SMEM, st,
                 *ma, *mb, (const uint8_t*)wimg, bias, (const __nv_bfloat16*)r, ldr, (__nv_bfloat16*)y, ldy, g);
  if (e != cudaSuccess) { set_error("rdst_conv3x3_fwd_bf16_tc (pair): launch: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  return RDST_OK;
}

static unsigned long long* g_conv_dbg = nullptr;

template <int CIN, int NT>
static int launch_conv(const void* x, int64_t ldx, const void* wimg, const float* bias, const void* r, int64_t ldr,
                       void* y, int64_t ldy, ConvGeom g, int nslices, int sms, cudaStream_t st) {
  using K = ConvCfg<CIN, NT>;
  auto k = g_conv_dbg ? conv3x3_tc_kernel<CIN, NT, true> : conv3x3_tc_kernel<CIN, NT, false>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
  if (e != cudaSuccess) { set_error("rdst_conv3x3_fwd_bf16_tc: smem attr (%d B): %s", K::SMEM, cudaGetErrorString(e)); return RDST_E_CUDA; }
  int occ = (int)(232448 / (K::SMEM + 2048));          // shared memory, TMEM columns and a cap of 4 CTAs per SM
  if (occ > 512 / K::TMEM_COLS) occ = 512 / K::TMEM_COLS;
  if (occ < 1) occ = 1;
  if (occ > 4) occ = 4;
  const int64_t ntiles = (int64_t)g.B * g.nty * g.ntx;
  int64_t gx = (int64_t)occ * sms / nslices;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  dim3 grid((unsigned)gx, (unsigned)nslices);
  const CUtensorMap* mx = get_act_tmap(x, ldx, g.B, g.H, g.W, CIN == 64 ? 64 : CIN, g.LW, g.TH + 2, 64);   // (used when CIN == 64)
  if (!mx) return RDST_E_CUDA;
  e = launch_pdl(k, grid, dim3(CONV_THREADS), (size_t)K::SMEM, st, *mx, (const __nv_bfloat16*)x, ldx, (const uint8_t*)wimg, bias,
                 (const __nv_bfloat16*)r, ldr, (__nv_bfloat16*)y, ldy, g, g_conv_dbg);
  if (e != cudaSuccess) { set_error("rdst_conv3x3_fwd_bf16_tc: launch: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  return RDST_OK;
}

}  // namespace rdst

extern "C" int rdst_debug_conv_timing(void* device_buffer_128_u64) {
  rdst::g_conv_dbg = (unsigned long long*)device_buffer_128_u64;
  return RDST_OK;
}

extern "C" int rdst_last_conv_fwd_bf16_tc(const void* x, int64_t ldx, const void* wimg, float bias, float out_scale,
                                          float out_bias, float* img, int B, int H, int W, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && wimg && img, "rdst_last_conv_fwd_bf16_tc: null pointer");
  RDST_REQUIRE(B >= 0 && H > 0 && W > 0 && ldx >= 64 && ldx % 8 == 0 && ((uintptr_t)x % 16 == 0),
               "rdst_last_conv_fwd_bf16_tc: bad shape / alignment");
  if (B == 0) return RDST_OK;
  ConvGeom g{};
  g.B = B; g.H = H; g.W = W; g.shuffle = 0; g.out_scale = out_scale; g.out_bias = out_bias;
  // tile: TW in {8,16,32,64} output columns, TH rows with (TH+2)*(TW+2) <= 512 staged halo positions (and >= 128: one full
  // MMA block); pick the shape that stages the fewest positions for the whole image (halo over-read + ragged edges)
  int64_t best = -1;
  for (int tw = 8; tw <= 64; tw *= 2) {
    const int lw = tw + 2;
    int th = LastCfg::NP_MAX / lw - 2;
    if (th > H) th = H;
    while ((th + 2) * lw < 128) ++th;                   // tiny images: pad the tile downwards (rows beyond H are masked)
    if (th < 1 || (th + 2) * lw > LastCfg::NP_MAX) continue;
    const int64_t cost = (int64_t)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((th + 2) * lw + 64);
    if (best < 0 || cost <= best) { best = cost; g.TW = tw; g.TH = th; }
  }
  RDST_REQUIRE(best >= 0, "rdst_last_conv_fwd_bf16_tc: no tile shape for H=%d W=%d", H, W);
  g.LW = g.TW + 2;
  g.NP = ((g.TH + 2) * g.LW + 127) / 128 * 128;
  g.ntx = (W + g.TW - 1) / g.TW;
  g.nty = (H + g.TH - 1) / g.TH;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g.out_bias = out_bias + bias * out_scale;      // scalar conv bias folded into the output affine (no device read)
  {
    using K = LastCfg;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = ((g.TH + 2) * g.LW + 127) / 128;
    const int smem_bytes = K::smem(nblk) + 1024;          // + slack: the dynamic segment is aligned to 1024 B by the kernel's declaration
    const CUtensorMap* mx = get_act_tmap(x, ldx, B, H, W, 64, g.LW, g.TH + 2);     // halo box: 64 channels x LW x (TH+2)
    if (!mx) return RDST_E_CUDA;
    cudaError_t e = cudaFuncSetAttribute(last_conv_tap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::smem(K::NBLK_MAX) + 1024);
    if (e != cudaSuccess) { set_error("rdst_last_conv_fwd_bf16_tc: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
    int occ = (int)(232448 / (smem_bytes + 2048));
    if (occ > 512 / K::TMEM_COLS) occ = 512 / K::TMEM_COLS;
    if (occ < 1) occ = 1;
    const int64_t ntiles = (int64_t)g.B * g.nty * g.ntx;
    int64_t gx = (int64_t)occ * sms;
    if (gx > ntiles) gx = ntiles;
    e = launch_pdl(last_conv_tap_kernel, dim3((unsigned)gx), dim3(CONV_THREADS), (size_t)smem_bytes, st,
                   *mx, (const uint8_t*)wimg, img, g);
    if (e != cudaSuccess) { set_error("rdst_last_conv_fwd_bf16_tc: launch: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  }
  RDST_CHECK_LAUNCH("rdst_last_conv_fwd_bf16_tc");
  return RDST_OK;
}

extern "C" int rdst_conv3x3_fwd_bf16_tc(const void* x, int64_t ldx, const void* wimg, const float* bias,
                                        const void* resid, int64_t ldr, void* y, int64_t ldy, int B, int H, int W,
                                        int Cin, int N, float out_scale, int shuffle, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && wimg && bias && y, "rdst_conv3x3_fwd_bf16_tc: null pointer");
  RDST_REQUIRE(B >= 0 && H > 0 && W > 0, "rdst_conv3x3_fwd_bf16_tc: bad shape");
  RDST_REQUIRE((Cin == 160 && N == 64) || (Cin == 64 && (N == 64 || N == 256)),
               "rdst_conv3x3_fwd_bf16_tc: supported (Cin,N) = (160,64), (64,64), (64,256); got (%d,%d)", Cin, N);
  RDST_REQUIRE(shuffle == 0 || (shuffle == 2 && N == 256 && resid == nullptr),
               "rdst_conv3x3_fwd_bf16_tc: shuffle=2 needs N=256 and no residual");
  RDST_REQUIRE(!(N == 256 && shuffle == 0), "rdst_conv3x3_fwd_bf16_tc: N=256 is only supported with shuffle=2");
  RDST_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ldx % 8 == 0 && ldy % 8 == 0 &&
                   (resid == nullptr || (((uintptr_t)resid % 16 == 0) && ldr % 8 == 0)),
               "rdst_conv3x3_fwd_bf16_tc: pointers must be 16-byte aligned, strides multiples of 8 elements");
  RDST_REQUIRE(ldx >= Cin, "rdst_conv3x3_fwd_bf16_tc: ldx < Cin");
  RDST_REQUIRE(((uintptr_t)y % 32 == 0) && ldy % 16 == 0,
               "rdst_conv3x3_fwd_bf16_tc: y must be 32-byte aligned with a row stride multiple of 16 elements (256-bit stores)");
  if (B == 0) return RDST_OK;
  ConvGeom g{};
  g.B = B; g.H = H; g.W = W; g.shuffle = shuffle; g.out_scale = out_scale;
  // tile shape: TW in {8,16,24,32}, TH = largest with (TH-1)*(TW+2)+TW <= 128; pick the fewest tiles
  int64_t best = -1;
  for (int tw = 8; tw <= 32; tw += 8) {
    const int th = (128 - tw) / (tw + 2) + 1;
    const int64_t nt = (int64_t)((W + tw - 1) / tw) * ((H + th - 1) / th);
    if (best < 0 || nt < best) { best = nt; g.TW = tw; g.TH = th; }
  }
  g.LW = g.TW + 2;
  g.NP = (2 * g.LW + 2 + 128 + 7) / 8 * 8;
  g.ntx = (W + g.TW - 1) / g.TW;
  g.nty = (H + g.TH - 1) / g.TH;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (Cin == 160) rc = g_conv_dbg ? launch_conv<160, 32>(x, ldx, wimg, bias, resid, ldr, y, ldy, g, 2, sms, st)
                                  : launch_conv_pair(x, ldx, wimg, bias, resid, ldr, y, ldy, g, sms, st);
  else if (N == 256) rc = launch_conv<64, 128>(x, ldx, wimg, bias, resid, ldr, y, ldy, g, 2, sms, st);
  else rc = launch_conv<64, 64>(x, ldx, wimg, bias, resid, ldr, y, ldy, g, 1, sms, st);
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_conv3x3_fwd_bf16_tc");
  return RDST_OK;
}
