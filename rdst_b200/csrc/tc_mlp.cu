// Fused Swin MLP on tcgen05:   Y = X + fc2( GELU( fc1( LNhat(X) ) ) )          (bf16 operands, fp32 accumulate)
//
// Replaces norm2 + Mlp + residual of SwinTransformerBlock.forward (reference swin_transformer_sr.py:272, :23-29).
// One persistent CTA per SM keeps both weight matrices resident in shared memory as ready-made UMMA operand
// images and walks over 128-token tiles:
//   P0  the 128 token rows of a tile arrive by TMA (one box per 64-channel panel, SWIZZLE_128B) into one of two raw
//       tiles, a whole tile ahead; finished tiles leave the same way (TMA store), so no thread ever waits on HBM
//   P1  thread = (token row, channel quarter): one pass over the landed rows -> partial LayerNorm sums -> exchange ->
//       normalise from registers -> packed bf16 pairs -> tcgen05.st: the A operand of fc1 lives in TMEM, so
//       the MMAs read only the weights from shared memory (SS-mode A reads were the bottleneck: 4 KB per k-step)
//   P2  fc1 as two N-halves of tcgen05.mma (A from TMEM), each committed to its own mbarrier; issue is warp-uniform
//       (elect.sync) so descriptors stay in uniform registers
//   P3  per half: tcgen05.ld -> +bias -> GELU -> packed bf16 -> tcgen05.st (hidden stays in TMEM as the A operand of
//       fc2); fc2's K-half is issued as soon as its hidden half is staged
//   P5  tcgen05.ld of the fc2 accumulator -> +bias +residual (raw tile) in the row mapping -> in-place bf16 ->
//       coalesced copy out; TAIL variant: LayerNorm(y) -> TMEM -> growth projection MMA -> dense slice
// Hidden activations never leave the SM.  LayerNorm gamma/beta are folded into fc1 by the host.
#include "common.cuh"
#include "umma.cuh"
#include "tma.cuh"

namespace rdst {
using namespace umma;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// GELU(x) = x*Phi(x) with Phi(x) ~ 0.5*(1+tanh(x*(a+b x^2+c x^4))), a/b/c fitted to the exact erf form
// (max abs deviation 2.6e-5 on |x|<=8); EXACT evaluates erff instead.
// Two GELUs per instruction on packed fp16 (two-term tanh form emitting 2*GELU, the 1/2 lives in the fc2 image; see
// tc_mlp2.cu): a + bias pairs in, packed fp16 pair out.  The
// fp16 chain is more accurate than rounding the exact value to bf16 (max 3.9e-3 vs 3.1e-2 on |x|<=10), and the hidden
// activations stay fp16 (fc2 runs with fp16 A and fp16 weights).
template <bool EXACT>
__device__ __forceinline__ uint32_t gelu_pair(float a, float b) {
  if (EXACT) {
    __half2 r = __floats2half2_rn(2.0f * gelu_erf(a), 2.0f * gelu_erf(b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  const __half2 x = __floats2half2_rn(a, b);
  const __half2 u = __hmul2(x, x);
  const __half2 p = __hfma2(u, __float2half2_rn(3.470089e-02f), __float2half2_rn(8.0015708e-01f));
  const __half2 inner = __hmul2(x, p);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&inner)));
  const __half2 r = __hfma2(x, *reinterpret_cast<const __half2*>(&t), x);
  return *reinterpret_cast<const uint32_t*>(&r);
}

template <int CP, int HP>
struct MlpCfg {
  static constexpr int NCH = CP / 8;                      // 16-byte chunks per activation row
  static constexpr int NCHK = (HP + 63) / 64;             // fc1 N-chunks (== fc2 K-chunks) of 64 hidden units
  static constexpr int LASTW = HP - 64 * (NCHK - 1);      // width of the last chunk (64 or 48)
  static constexpr int W1_BYTES = HP * CP * 2;
  static constexpr int W2_BYTES = CP * HP * 2;
  static constexpr int NP = CP > 64 ? 2 : 1;              // raw tile: 64-channel panels [128 rows][128 B], SWIZZLE_128B
  static constexpr int PANEL = 128 * 128;
  static constexpr int XT_BYTES = NP * PANEL;
  static constexpr int WT_BYTES = 32 * CP * 2;
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
  static constexpr int OFF_XT = OFF_W2 + W2_BYTES;         // raw bf16 tile (residual source, later the output staging)
  static constexpr int OFF_B1 = OFF_XT + 2 * XT_BYTES;     // two raw tiles: tile n+1 lands (TMA) while tile n computes
  static constexpr int OFF_B2 = OFF_B1 + HP * 4;
  static constexpr int OFF_STAT = OFF_B2 + CP * 4;         // [128] (mean, rstd)
  static constexpr int OFF_XCH = OFF_STAT + 128 * 8;       // [4][128] (sum, sumsq) exchange for the tail LayerNorm
  static constexpr int OFF_WT = OFF_XCH + 4 * 128 * 8;     // fused DenseSTLayer tail: Linear(C -> 30, padded 32) image
  static constexpr int OFF_BT = OFF_WT + WT_BYTES;
  static constexpr int SMEM_PLAIN = OFF_WT;
  static constexpr int SMEM_TAIL = OFF_BT + 32 * 4;
  // TMEM columns.  A operands live in TMEM (packed bf16 pairs, lane = token row): only weights are read from smem.
  static constexpr int TM_FC1 = 0;                         // fc1 accumulator [0,HP); later the tail accumulator [0,32)
  static constexpr int TM_FC2 = 256;                       // fc2 accumulator [256,256+CP)
  static constexpr int TM_XH = 256;                        // normalised input [256,256+CP/2): dead before fc2 starts (in-order MMAs)
  static constexpr int TM_HID = 384;                       // GELU(hidden) packed [384,384+HP/2)
  static_assert(HP % 16 == 0 && CP % 32 == 0 && LASTW % 16 == 0 && LASTW > 0 && NCHK <= 4, "tile shape");
  static_assert(HP <= 240 && CP <= 128 && TM_HID + HP / 2 <= 512, "TMEM budget");
  static_assert(OFF_XT % 1024 == 0, "raw tiles must be 1024-byte aligned (SWIZZLE_128B)");
};

// byte offset of 16-byte chunk c (8 channels) of token row `row` inside a raw tile (TMA SWIZZLE_128B panels)
__device__ __forceinline__ uint32_t mlp_xt_off(int row, int c) {
  return (uint32_t)((c >> 3) * (128 * 128) + row * 128 + (((c & 7) ^ (row & 7)) << 4));
}

// b1 rides inside fc1: x-hat carries ones at pad channels 60/61 and rows 60/61 of the resident W1 image hold b1 as a
// fp16 hi/lo pair (error <= 2^-22 |b1|), patched once per CTA.  (Round 2: fc1 runs on fp16 operands in both kernels.)
__device__ __forceinline__ uint32_t mlp_bias_hi_lo(float b) {
  const __half hi = __float2half_rn(b);
  const __half lo = __float2half_rn(b - __half2float(hi));
  return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
}
__device__ __forceinline__ uint32_t mlp_pack_f16x2(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// store 8 packed columns held in a larger register array
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = v[i];
  tmem_st_x8(taddr, a);
}

struct TailArgs {
  const uint8_t* wtimg;      // [CP/8][32][8] bf16, LN gamma folded
  const float* bt;           // [32]
  __nv_bfloat16* dense;      // &D[0][64 + 32 j]
  int64_t ldd;
  float scale;               // dense_scale
};

constexpr int MLP_THREADS = 512;      // 16 warps: four threads (one per warpgroup) share a token row / TMEM lane

template <int CP, int HP, bool EXACT, bool TAIL, bool DBG>      // DBG: clock64() phase stamps (rdst_debug_mlp_timing)
__global__ void __launch_bounds__(MLP_THREADS, 1)
stl_mlp_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
               const uint8_t* __restrict__ w1img, const uint8_t* __restrict__ w2img,
               const float* __restrict__ b1, const float* __restrict__ b2, int64_t T, int creal, TailArgs ta,
               unsigned long long* __restrict__ dbg) {
  using C = MlpCfg<CP, HP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];          // [0..3] fc1 chunks, [4] fc2, [5] tail, [6] weights landed
  __shared__ uint64_t xbar[2];          // raw tile landed, one per buffer
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sB2 = reinterpret_cast<float*>(smem + C::OFF_B2);
  float2* sXch = reinterpret_cast<float2*>(smem + C::OFF_XCH);

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    fence_mbar_init();
    // resident weights (ready-made operand images) arrive by bulk async copies that overlap the prologue and the
    // first tile's load + LayerNorm; the issuer waits on bars[6] once before its first tcgen05.mma
    mbar_arrive_expect_tx(&bars[6], C::W1_BYTES + C::W2_BYTES + (TAIL ? C::WT_BYTES : 0));
    for (int off = 0; off < C::W1_BYTES; off += 32768)
      bulk_g2s(smem + C::OFF_W1 + off, w1img + off, min(32768, C::W1_BYTES - off), &bars[6]);
    for (int off = 0; off < C::W2_BYTES; off += 32768)
      bulk_g2s(smem + C::OFF_W2 + off, w2img + off, min(32768, C::W2_BYTES - off), &bars[6]);
    if (TAIL) bulk_g2s(smem + C::OFF_WT, ta.wtimg, C::WT_BYTES, &bars[6]);
  }
  for (int i = tid; i < CP; i += MLP_THREADS) sB2[i] = b2[i];
  if (TAIL) {
    for (int i = tid; i < 32; i += MLP_THREADS) reinterpret_cast<float*>(smem + C::OFF_BT)[i] = ta.bt[i];
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  const int row = tid & 127, qtr = tid >> 7;             // token row (= TMEM lane) and column quarter of this thread
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  // warp-uniform values for the MMA issuer (warp 0): descriptors stay in uniform registers
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t aW1 = smem_u32(smem + C::OFF_W1), aW2 = smem_u32(smem + C::OFF_W2), aWT = smem_u32(smem + C::OFF_WT);
  const float inv_c = 1.0f / (float)creal;
  const int64_t ntiles = (T + 127) / 128;
  uint32_t parity = 0;
  int dbg_n = 0;
  const bool dbg_on = DBG && dbg != nullptr && blockIdx.x == 0 && (tid & 127) == 0 && qtr < 2;
#define RDST_TSTAMP()                                                         \
  do {                                                                        \
    if (DBG && dbg_on && dbg_n < 64) dbg[qtr * 64 + dbg_n++] = clock64();     \
  } while (0)
  // All TMA traffic is issued by one elected lane of warp 4 under a warp-uniform branch (coordinates and descriptors
  // stay in uniform registers; elect.sync picks the same lane every time, which the bulk-group waits rely on).
  const bool tma_warp = warp_u == 4;
  auto load_tile = [&](int64_t tile, int b) {           // rows beyond T are out of range -> zero fill
    if (tile >= ntiles) return;
    mbar_arrive_expect_tx(&xbar[b], C::XT_BYTES);
#pragma unroll
    for (int pnl = 0; pnl < C::NP; ++pnl)
      tma::load_4d(smem + C::OFF_XT + b * C::XT_BYTES + pnl * C::PANEL, &mapX, pnl * 64, (int)(tile * 128), 0, 0, &xbar[b]);
  };
  // fc1 runs in 64-column chunks (the last one may be 48 wide), each committed to its own mbarrier, so the GELU of
  // chunk c starts while the tensor pipe works on chunk c+1; fc2 accumulates over the same chunks as K-slices.
  // tcgen05.mma issue blocks while the tensor-pipe queue is full, so at most two chunks are queued ahead and the
  // issuing warp rotates (an issuer that also owns token rows would otherwise be late for every barrier).
  // All fc1 chunks are issued before the first fc2 chunk: the fc2 accumulator aliases the normalised input.
  auto issue_fc1 = [&](int c) {
    constexpr uint32_t idw = make_idesc_f16(128, 64, false, false);
    constexpr uint32_t idl = make_idesc_f16(128, C::LASTW, false, false);
#pragma unroll
    for (int ks = 0; ks < CP / 16; ++ks)
      mma_ts(tmem_u + C::TM_FC1 + 64 * c, tmem_u + C::TM_XH + ks * 8,
             make_smem_desc(aW1 + 64 * c * 16 + ks * 2 * (HP * 16), HP * 16, 128), c == C::NCHK - 1 ? idl : idw, ks > 0);
    commit(&bars[c]);
  };
  auto issue_fc2 = [&](int c) {
    constexpr uint32_t id2 = make_idesc_f16(128, CP, false, false);       // hidden and W2 are fp16
    const int ks0 = 4 * c, ks1 = c == C::NCHK - 1 ? HP / 16 : 4 * c + 4;
    for (int ks = ks0; ks < ks1; ++ks)
      mma_ts(tmem_u + C::TM_FC2, tmem_u + C::TM_HID + ks * 8, make_smem_desc(aW2 + ks * 2 * (CP * 16), CP * 16, 128), id2, ks > 0);
    if (c == C::NCHK - 1) commit(&bars[4]);
  };
  pdl_launch_dependents();
  pdl_wait();                    // prologue above touched only weights; the rows below come from the previous kernel
  if (tma_warp) {
    if (elect_one()) load_tile(blockIdx.x, 0);
    __syncwarp();
  }
  int buf = 0;
  uint32_t ph_x0 = 0, ph_x1 = 0;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, parity ^= 1, buf ^= 1) {
    const int64_t t0 = tile * 128;
    uint8_t* sXT = smem + C::OFF_XT + buf * C::XT_BYTES;
    RDST_TSTAMP();   // tile start
    if (buf == 0) { mbar_wait(&xbar[0], ph_x0 & 1); ph_x0++; } else { mbar_wait(&xbar[1], ph_x1 & 1); ph_x1++; }
    // ---------------- P1: one pass over the landed rows: partial LayerNorm sums -> exchange -> A operand in TMEM ------
    constexpr int NCQ = C::NCH / 4;                           // 16-byte chunks per thread (2, 3 or 4)
    float fv[NCQ * 8];
    {
      uint4 rv[NCQ];
#pragma unroll
      for (int cc = 0; cc < NCQ; ++cc) rv[cc] = *reinterpret_cast<const uint4*>(sXT + mlp_xt_off(row, qtr * NCQ + cc));
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int cc = 0; cc < NCQ; ++cc) {
        const uint32_t w4[4] = {rv[cc].x, rv[cc].y, rv[cc].z, rv[cc].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w4[q]);
          fv[cc * 8 + 2 * q] = f.x; fv[cc * 8 + 2 * q + 1] = f.y;     // unpacked once, normalised from registers
          s0 += f.x; s1 += f.y;
          q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
        }
      }
      sXch[qtr * 128 + row] = make_float2(s0 + s1, q0 + q1);   // zero pads add nothing to either sum
    }
    __syncthreads();
    RDST_TSTAMP();   // P1a done
    {
      const float2 p0 = sXch[row], p1 = sXch[128 + row], p2 = sXch[256 + row], p3 = sXch[384 + row];
      const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) * inv_c;
      const float var = ((p0.y + p1.y) + (p2.y + p3.y)) * inv_c - mean * mean;
      const float rstd = rsqrtf(fmaxf(var, 0.f) + 1e-5f);
      const float nb = -mean * rstd;
      uint32_t o[NCQ * 4];
#pragma unroll
      for (int e = 0; e < NCQ * 4; ++e) o[e] = mlp_pack_f16x2(fmaf(fv[2 * e], rstd, nb), fmaf(fv[2 * e + 1], rstd, nb));
      if (qtr == 7 / NCQ) o[(7 % NCQ) * 4 + 2] = 0x3C003C00u;      // ones at pad channels 60, 61 (folded fc1 bias)
      const uint32_t dst = lane_addr + C::TM_XH + qtr * NCQ * 4;
#pragma unroll
      for (int c0 = 0; c0 + 8 <= NCQ * 4; c0 += 8) tmem_st8(dst + c0, o + c0);
      if ((NCQ * 4) % 8 != 0) {
        uint32_t a4[4] = {o[NCQ * 4 - 4], o[NCQ * 4 - 3], o[NCQ * 4 - 2], o[NCQ * 4 - 1]};
        tmem_st_x4(dst + NCQ * 4 - 4, a4);
      }
      wait_st();
    }
    fence_before_sync();
    __syncthreads();
    RDST_TSTAMP();   // P1b done
    // ---------------- P2: fc1 (two N-halves), A from TMEM ----------------
    if (warp_u == 0) {
      if (tile == (int64_t)blockIdx.x) {
        mbar_wait(&bars[6], 0);  // weights have landed (first tile only)
        for (int n = lane; n < HP; n += 32)
          *reinterpret_cast<uint32_t*>(smem + C::OFF_W1 + (7 * HP + n) * 16 + 8) = mlp_bias_hi_lo(b1[n]);
        fence_proxy_async();
        __syncwarp();
      }
      fence_after_sync();
      if (elect_one()) {
        issue_fc1(0);
        if (C::NCHK > 1) issue_fc1(1);
      }
      __syncwarp();
    }
    if (tma_warp) {                // next tile -> other buffer, under fc1 (its last store has long been read out)
      if (elect_one()) {
        bulk_wait_read();
        load_tile(tile + gridDim.x, buf ^ 1);
      }
      __syncwarp();
    }
    // ---------------- P3: GELU epilogue per chunk -> packed fp16 hidden in TMEM; fc2 K-chunk issued behind it ----------------
#pragma unroll
    for (int h = 0; h < C::NCHK; ++h) {
      const int hbase = 64 * h;
      const int hw = h == C::NCHK - 1 ? C::LASTW : 64;
      const int qw = (hw / 4 + 7) / 8 * 8;                     // columns per quarter (multiple of 8), last one may be short
      const int cbeg = hbase + min(qtr * qw, hw);
      const int cend = hbase + min((qtr + 1) * qw, hw);
      RDST_TSTAMP();   // before fc1 half wait
      mbar_wait(&bars[h], parity);
      fence_after_sync();
      RDST_TSTAMP();   // fc1 half ready
      {
        // all accumulator columns of this thread are requested up front (one wait), then GELU -> packed hidden
        constexpr int MAXC = 16;
        uint32_t v[MAXC];
#pragma unroll
        for (int q = 0; q < MAXC / 8; ++q)
          if (cbeg + q * 8 < cend) {
            uint32_t t8[8];
            tmem_ld_x8(lane_addr + C::TM_FC1 + cbeg + q * 8, t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q * 8 + e] = t8[e];
          }
        wait_ld();
#pragma unroll
        for (int q = 0; q < MAXC / 8; ++q)
          if (cbeg + q * 8 < cend) {
            const int c0 = cbeg + q * 8;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = gelu_pair<EXACT>(__uint_as_float(v[q * 8 + 2 * j]), __uint_as_float(v[q * 8 + 2 * j + 1]));
            tmem_st_x4(lane_addr + C::TM_HID + c0 / 2, o);
          }
      }
      wait_st();
      RDST_TSTAMP();   // GELU half done
      fence_before_sync();
      __syncthreads();
      if (warp_u == 1 + 4 * h) {            // rotating issuer: warps 1, 5, 9, 13
        fence_after_sync();
        if (elect_one()) {
          if (h + 2 < C::NCHK) issue_fc1(h + 2);
          if (h >= C::NCHK - 3) {             // every fc1 chunk has been issued: fc2 K-slices of the finished chunks
            if (h == C::NCHK - 3 || (C::NCHK < 3 && h == 0)) {
#pragma unroll
              for (int c2 = 0; c2 <= h; ++c2) issue_fc2(c2);
            } else {
              issue_fc2(h);
            }
          }
        }
        __syncwarp();
      }
    }
    // ---------------- P5: fc2 epilogue in the row mapping: y = acc + b2 + x (raw tile) ----------------
    RDST_TSTAMP();   // before fc2 wait
    mbar_wait(&bars[4], parity);
    fence_after_sync();
    RDST_TSTAMP();   // fc2 ready
    {
      constexpr int NC = CP / 4;                               // columns per thread (16, 24 or 32)
      const int cb = qtr * NC;
      float y[NC];
      float s1 = 0.f, s2 = 0.f;
      uint32_t acc[NC];
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 8) {
        uint32_t t8[8];
        tmem_ld_x8(lane_addr + C::TM_FC2 + cb + c0, t8);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[c0 + e] = t8[e];
      }
      wait_ld();
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 8) {
        const int ch = (cb + c0) / 8;
        uint8_t* xp = sXT + mlp_xt_off(row, ch);
        const uint4 xv = *reinterpret_cast<const uint4*>(xp);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 xf = unpack_bf16x2(xw[e]);
          const int k = c0 + 2 * e;
          y[k] = __uint_as_float(acc[k]) + sB2[cb + k] + xf.x;
          y[k + 1] = __uint_as_float(acc[k + 1]) + sB2[cb + k + 1] + xf.y;
          s1 += y[k] + y[k + 1];
          s2 += y[k] * y[k] + y[k + 1] * y[k + 1];
          o[e] = pack_bf16x2(y[k], y[k + 1]);
        }
        if (!TAIL) *reinterpret_cast<uint4*>(xp) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (TAIL) {
        // LayerNorm of the block output over the full row: its four column quarters live in four warpgroups
        sXch[qtr * 128 + row] = make_float2(s1, s2);
        __syncthreads();
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 p = sXch[q * 128 + row]; t1 += p.x; t2 += p.y; }
        const float mean = t1 * inv_c;
        const float var = fmaxf(t2 * inv_c - mean * mean, 0.f);   // pads are exact zeros: they add nothing
        const float rstd = rsqrtf(var + 1e-5f);
        uint32_t o[NC / 2];
#pragma unroll
        for (int k = 0; k < NC; k += 2) o[k / 2] = pack_bf16x2((y[k] - mean) * rstd, (y[k + 1] - mean) * rstd);
        const uint32_t dst = lane_addr + C::TM_XH + cb / 2;
#pragma unroll
        for (int c0 = 0; c0 + 8 <= NC / 2; c0 += 8) tmem_st8(dst + c0, o + c0);
        if ((NC / 2) % 8 != 0) {
          uint32_t a4[4] = {o[NC / 2 - 4], o[NC / 2 - 3], o[NC / 2 - 2], o[NC / 2 - 1]};
          tmem_st_x4(dst + NC / 2 - 4, a4);
        }
        wait_st();
      }
    }
    if (!TAIL) fence_proxy_async();      // the finished rows are read by the TMA store (async proxy)
    fence_before_sync();
    __syncthreads();
    RDST_TSTAMP();   // P5 done
    if (TAIL) {
      if (warp_u == 0) {
        fence_after_sync();
        if (elect_one()) {
          constexpr uint32_t idt = make_idesc_bf16(128, 32, false, false);
#pragma unroll
          for (int ks = 0; ks < CP / 16; ++ks)
            mma_ts(tmem_u + C::TM_FC1, tmem_u + C::TM_XH + ks * 8, make_smem_desc(aWT + ks * 2 * (32 * 16), 32 * 16, 128), idt, ks > 0);
          commit(&bars[5]);
        }
        __syncwarp();
      }
      mbar_wait(&bars[5], parity);
      fence_after_sync();
      {
        // each quarter writes 8 of the 32 growth columns of its row (16 bytes)
        const float* sBT = reinterpret_cast<const float*>(smem + C::OFF_BT);
        uint32_t v[8];
        tmem_ld_x8(lane_addr + C::TM_FC1 + 8 * qtr, v);
        wait_ld();
        const int64_t t = t0 + row;
        if (t < T) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = pack_bf16x2((__uint_as_float(v[2 * j]) + sBT[8 * qtr + 2 * j]) * ta.scale,
                               (__uint_as_float(v[2 * j + 1]) + sBT[8 * qtr + 2 * j + 1]) * ta.scale);
          reinterpret_cast<uint4*>(ta.dense + t * ta.ldd)[qtr] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      fence_before_sync();
    } else {
      if (tma_warp) {              // rows beyond T are out of range and dropped
        if (elect_one()) {
#pragma unroll
          for (int pnl = 0; pnl < C::NP; ++pnl) tma::store_4d(&mapY, pnl * 64, (int)t0, 0, 0, sXT + pnl * C::PANEL);
          bulk_commit();
        }
        __syncwarp();
      }
    }
    RDST_TSTAMP();   // tile done
  }
#undef RDST_TSTAMP
  if (tma_warp) {
    if (elect_one()) bulk_wait_read();
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static unsigned long long* g_mlp_dbg = nullptr;

template <int CP, int HP>
static int launch_mlp(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w1, const void* w2, const float* b1,
                      const float* b2, int64_t T, int creal, int exact_gelu, const TailArgs* tail, int sms, cudaStream_t st) {
  using C = MlpCfg<CP, HP>;
  const int64_t ntiles = (T + 127) / 128;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  TailArgs ta{};
  int smem = C::SMEM_PLAIN;
  // activations as [T][CP] "images" of one row: box = 128 tokens x 64 channels
  const CUtensorMap* mx = get_act_tmap(x, ldx, 1, 1, (int)T, CP, 128, 1);
  const CUtensorMap* my = tail ? mx : get_act_tmap(y, ldy, 1, 1, (int)T, CP, 128, 1);
  if (!mx || !my) return RDST_E_CUDA;
  void (*k)(const CUtensorMap, const CUtensorMap, const uint8_t*, const uint8_t*, const float*,
            const float*, int64_t, int, TailArgs, unsigned long long*);
  if (tail) {
    ta = *tail;
    smem = C::SMEM_TAIL;
    k = exact_gelu ? stl_mlp_kernel<CP, HP, true, true, false> : stl_mlp_kernel<CP, HP, false, true, false>;
    if (g_mlp_dbg && !exact_gelu) k = stl_mlp_kernel<CP, HP, false, true, true>;
  } else {
    k = exact_gelu ? stl_mlp_kernel<CP, HP, true, false, false> : stl_mlp_kernel<CP, HP, false, false, false>;
    if (g_mlp_dbg && !exact_gelu) k = stl_mlp_kernel<CP, HP, false, false, true>;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("rdst_stl_mlp_fwd_bf16: smem attr (%d B): %s", smem, cudaGetErrorString(e)); return RDST_E_CUDA; }
  e = launch_pdl(k, dim3(grid), dim3(MLP_THREADS), (size_t)smem, st, *mx, *my,
                 (const uint8_t*)w1, (const uint8_t*)w2, b1, b2, T, creal, ta, g_mlp_dbg);
  if (e != cudaSuccess) { set_error("rdst_stl_mlp_fwd_bf16: launch: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  return RDST_OK;
}

// warp-specialised kernel (tc_mlp2.cu), the default
int mlp_v2_dispatch(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w1img, const void* w2img, const float* b1,
                    const float* b2, int64_t T, int C, int exact_gelu, const void* wtimg, const float* bt, void* dense, int64_t ldd,
                    float scale, int has_tail, int sms, cudaStream_t st);
static int g_mlp_variant = 2;

static int mlp_dispatch(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w1img, const void* w2img,
                        const float* b1, const float* b2, int64_t T, int C, int exact_gelu, const TailArgs* tail,
                        void* stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (g_mlp_variant == 2)
    return mlp_v2_dispatch(x, ldx, y, ldy, w1img, w2img, b1, b2, T, C, exact_gelu, tail ? tail->wtimg : nullptr,
                           tail ? tail->bt : nullptr, tail ? (void*)tail->dense : nullptr, tail ? tail->ldd : 0,
                           tail ? tail->scale : 0.f, tail != nullptr, sms, st);
  switch (C) {
    case 60:  return launch_mlp<64, 128>(x, ldx, y, ldy, w1img, w2img, b1, b2, T, 60, exact_gelu, tail, sms, st);
    case 90:  return launch_mlp<96, 192>(x, ldx, y, ldy, w1img, w2img, b1, b2, T, 90, exact_gelu, tail, sms, st);
    case 120: return launch_mlp<128, 240>(x, ldx, y, ldy, w1img, w2img, b1, b2, T, 120, exact_gelu, tail, sms, st);
    default: set_error("rdst_stl_mlp_fwd_bf16: C=%d unsupported (60, 90, 120 with mlp_ratio 2)", C); return RDST_E_UNSUPPORTED;
  }
}

}  // namespace rdst

extern "C" int rdst_debug_mlp_variant(int variant) {
  if (variant != 1 && variant != 2) { rdst::set_error("rdst_debug_mlp_variant: 1 (lock-step kernel) or 2 (warp-specialised)"); return RDST_E_INVALID; }
  rdst::g_mlp_variant = variant;
  return RDST_OK;
}

extern "C" int rdst_debug_mlp_timing(void* device_buffer_128_u64) {
  rdst::g_mlp_dbg = (unsigned long long*)device_buffer_128_u64;
  return RDST_OK;
}

extern "C" int rdst_stl_mlp_fwd_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w1img,
                                     const void* w2img, const float* b1, const float* b2, int64_t T, int C,
                                     int exact_gelu, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && y && w1img && w2img && b1 && b2, "rdst_stl_mlp_fwd_bf16: null pointer");
  RDST_REQUIRE(T >= 0, "rdst_stl_mlp_fwd_bf16: negative T");
  RDST_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ldx % 8 == 0 && ldy % 8 == 0,
               "rdst_stl_mlp_fwd_bf16: x/y must be 16-byte aligned with row strides multiple of 8 elements");
  RDST_REQUIRE(C == 60 || C == 90 || C == 120, "rdst_stl_mlp_fwd_bf16: C=%d unsupported (60, 90, 120)", C);
  RDST_REQUIRE(ldx >= 64 + 32 * ((C - 60) / 30) && ldy >= 64 + 32 * ((C - 60) / 30), "rdst_stl_mlp_fwd_bf16: ld too small");
  if (T == 0) return RDST_OK;
  int rc = mlp_dispatch(x, ldx, y, ldy, w1img, w2img, b1, b2, T, C, exact_gelu, nullptr, stream);
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_stl_mlp_fwd_bf16");
  return RDST_OK;
}

extern "C" int rdst_stl_mlp_tail_fwd_bf16(const void* x, int64_t ldx, const void* w1img, const void* w2img,
                                          const float* b1, const float* b2, const void* wtimg, const float* bt,
                                          void* dense, int64_t ldd, float dense_scale, int64_t T, int C,
                                          int exact_gelu, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && w1img && w2img && b1 && b2 && wtimg && bt && dense, "rdst_stl_mlp_tail_fwd_bf16: null pointer");
  RDST_REQUIRE(T >= 0, "rdst_stl_mlp_tail_fwd_bf16: negative T");
  RDST_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)dense % 16 == 0) && ldx % 8 == 0 && ldd % 8 == 0,
               "rdst_stl_mlp_tail_fwd_bf16: x/dense must be 16-byte aligned with row strides multiple of 8 elements");
  RDST_REQUIRE(C == 60 || C == 90 || C == 120, "rdst_stl_mlp_tail_fwd_bf16: C=%d unsupported (60, 90, 120)", C);
  RDST_REQUIRE(ldx >= 64 + 32 * ((C - 60) / 30) && ldd >= 32, "rdst_stl_mlp_tail_fwd_bf16: ld too small");
  if (T == 0) return RDST_OK;
  TailArgs ta{(const uint8_t*)wtimg, bt, (__nv_bfloat16*)dense, ldd, dense_scale};
  int rc = mlp_dispatch(x, ldx, nullptr, 0, w1img, w2img, b1, b2, T, C, exact_gelu, &ta, stream);
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_stl_mlp_tail_fwd_bf16");
  return RDST_OK;
}
