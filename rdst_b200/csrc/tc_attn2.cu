// Fused shifted-window attention on tcgen05, warp-specialised pipeline (round 2):
//     Y = X + proj( softmax( q k^T + rel_pos_bias + shift_mask ) v ),   [q|k|v] = qkv( LNhat(X) )
// Replaces norm1 + roll + window_partition + WindowAttention + window_reverse + roll + residual of
// SwinTransformerBlock.forward (reference swin_transformer_sr.py:239-271, :110-141, :211-232).
//
// Same arithmetic, operand images and TMEM/TMA plumbing as the first kernel (tc_attn.cu); what changes is WHO does
// what WHEN.  The first kernel ran every phase on all 512 threads in lock step (LN -> 3 head iterations -> proj ->
// epilogue), so the tensor pipe idled while the CUDA cores worked and vice versa (ncu: tensor pipe 24 %, issue slots
// 36 %).  Here one persistent CTA per SM is five roles that only meet at mbarriers, each working on a different
// (tile, head) at any moment:
//
//   warpgroup A  (warps 0-3,  thread = token row)  LayerNorm of tile n+1 -> x^ (TMEM);  proj epilogue of tile n-1
//                                                   (+residual, in place in the staging tile);  all TMA traffic
//   warpgroup B  (warps 4-7,  thread = token row)  drain of the qkv accumulator of head g: Q -> TMEM (packed bf16),
//                                                   K / V -> shared-memory operand images;  O / rowsum of head g-2
//   warpgroups C, D (warps 8-15, thread = token row, all 64 keys of the row)  softmax of even / odd heads
//   warp 16      one elected lane issues EVERY tcgen05.mma in one static order, so all ordering constraints between
//                MMAs (P is overwritten by the next S, Q by the next qkv, ...) hold by program order:
//                    iteration g (global head index 6 n + h):  S(g) ; PV(g-1) ; qkv(g+2) ; [proj halves of tile n-1]
//                qkv of the first two heads of tile n+1 are issued during heads 4, 5 of tile n: tiles overlap.
//
// A head needs no barrier wider than four warps: softmax rows are thread-private (no max exchange), hand-offs are
// mbarrier arrivals (one per warp) on the consumer side and tcgen05.commit on the MMA side.
// TMEM (512 columns; C = 120):  x^ [0,64) | qkv acc / packed Q, 2 slots [64,192) | S / P, 2 slots [192,320) |
// O, 2 slots [320,384) | normalised O = A of proj [384,448) | proj accumulator (one N-half at a time) [448,512).
// C = 60 / 90 are narrower and run THREE head slots (Cfg::NS) with a separate output staging tile.
// Shared memory: resident weight images | landing tile L (LayerNorm source of the NEXT tile) | staging tile E (the
// residual rows of the CURRENT tile, fetched a second time from L2 right before its epilogue, then output staging) |
// K / V images (2 slots) | bias table.  C = 120 has no room for a third tile buffer: the second fetch (an L2 hit, the
// rows were read a few microseconds earlier) is what lets LayerNorm run a whole tile ahead.
#include "common.cuh"
#include "umma.cuh"
#include "tma.cuh"

namespace rdst {
using namespace umma;

namespace a2 {

__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 up2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ uint32_t pk2h(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t bias_hi_lo(float b) {          // see tc_attn.cu: bias as a bf16 hi/lo pair inside the MMA
  const __nv_bfloat16 hi = __float2bfloat16_rn(b);
  const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
  return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}
__device__ __forceinline__ void tmem_st_x64(uint32_t taddr, const uint32_t* v) {
  uint32_t a[32], b[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = v[i]; b[i] = v[32 + i]; }
  tmem_st_x32(taddr, a);
  tmem_st_x32(taddr + 32, b);
}

template <int C_>
struct Cfg {
  static constexpr int C = C_;
  static constexpr int CP = C == 60 ? 64 : (C == 90 ? 96 : 128);
  static constexpr int HD = C / 6;
  static constexpr int NH = (3 * HD + 15) / 16 * 16;        // qkv columns of one head (padded): 32 / 48 / 64
  static constexpr int HDP = HD <= 16 ? 16 : 32;            // K of S = Q K^T
  static constexpr int HDV = HD <= 16 ? 16 : 32;            // N of O = P V
  static constexpr int HDO = HD == 20 ? 20 : 16;            // stride of one head inside the proj A operand
  static constexpr int KPROJ = (6 * HDO + 15) / 16 * 16;    // 96 / 96 / 128
  static constexpr int NCH = CP / 8;                        // 16-byte chunks per token row
  static constexpr int TBL = 15 * 24;                       // bias table per head (packed fp16 pairs), row pitch 24
  static constexpr int WQKV_BYTES = 6 * NH * CP * 2;
  static constexpr int WPROJ_BYTES = CP * KPROJ * 2;
  static constexpr int NP = CP > 64 ? 2 : 1;
  static constexpr int PANEL = 128 * 128;
  static constexpr int XT_BYTES = NP * PANEL;
  // K / V images: only the chunks that hold data are allocated.  For head_dim 20 the MMAs read a fourth 8-wide chunk
  // (K = 32 / N = 32): it aliases whatever follows (another image or the bias table: finite bits); the matching Q pads
  // are exact zeros and the matching O columns are never read.
  static constexpr int KCH = (HD + 7) / 8;                  // 2 / 2 / 3
  static constexpr int VCH = (HD + 8) / 8;                  // 2 / 2 / 3   (head_dim values + the ones column)
  static constexpr int BK_BYTES = KCH * 2048;
  static constexpr int BV_BYTES = VCH * 2048;
  static constexpr int KV_BYTES = BK_BYTES + BV_BYTES;      // per slot
  // Head slots in flight (qkv accumulator + S/P columns + O columns + K/V images each).  P is written in place over S, so
  // a slot's next S can only follow its PV: with two slots a softmax warpgroup waits one MMA round trip (~1100 cycles)
  // per head for its next logits.  C = 60 / 90 have the TMEM and shared memory for a third slot, which hides that round
  // trip (the next S of a warpgroup then sits on a slot whose PV was issued half a softmax earlier).
  // Measured (176 x 40 x 32, us per launch, 2 vs 3 slots): C = 60: 61.4 vs 59.7; C = 90: 67.0 vs 79.5 (its proj then needs
  // three 32-column blocks to fit TMEM, and the role-B chain drain(g+1) <- PV(g-NS) gets longer) -> three slots only at C = 60.
  static constexpr int NS = C == 60 ? 3 : 2;
  static constexpr int NBLK = CP == 64 ? 1 : (CP == 96 && NS == 3 ? 3 : 2);   // proj runs as NBLK column blocks through one accumulator
  static constexpr int NPC = CP / NBLK;                     // 64 / 48 (32 with three slots) / 64
  static constexpr bool YBUF = NS == 3;                     // room for a separate output staging tile (else E is reused)
  static constexpr int OFF_WQKV = 0;
  static constexpr int OFF_WPROJ = OFF_WQKV + WQKV_BYTES;
  static constexpr int OFF_L = OFF_WPROJ + WPROJ_BYTES;     // landing tile (LayerNorm source)
  static constexpr int OFF_E = OFF_L + XT_BYTES;            // residual rows (and output staging when there is no Y)
  static constexpr int OFF_Y = OFF_E + XT_BYTES;            // output staging
  static constexpr int OFF_KV = OFF_Y + (YBUF ? XT_BYTES : 0);     // [NS slots][K | V]
  static constexpr int OFF_TAB = OFF_KV + NS * KV_BYTES;
  static constexpr int OFF_BARS = OFF_TAB + 6 * TBL * 4;    // mbarriers + TMEM base live in dynamic shared memory too: a static
  static constexpr int SMEM = OFF_BARS + 40 * 8;            // __shared__ would cost a whole 1024-byte alignment unit
  // TMEM columns: x^ | qkv accumulators / packed Q | S / P | O | normalised O (A of proj) | proj accumulator (one block)
  static constexpr int TM_XH = 0, TM_QKV = CP / 2, TM_S = TM_QKV + NS * NH, TM_O = TM_S + NS * 64, TM_AP = TM_O + NS * HDV,
                       TM_PROJ = TM_AP + KPROJ / 2;
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(OFF_L % 1024 == 0 && OFF_E % 1024 == 0 && OFF_Y % 1024 == 0 && OFF_KV % 1024 == 0, "TMA tiles need 1024-byte alignment");
  static_assert(TM_PROJ + NPC <= 512 && NPC % 16 == 0, "TMEM map");
};

__device__ __forceinline__ uint32_t xt_off(int row, int c) {          // 16-byte chunk c of token row `row` (SWIZZLE_128B panels)
  return (uint32_t)((c >> 3) * (128 * 128) + row * 128 + (((c & 7) ^ (row & 7)) << 4));
}

struct Geom {
  int H, W, shift, nwx, nw_img;
  int nwt;
};
__device__ __forceinline__ int row_iy(int irow) { return ((irow >> 5) & 1) * 4 + ((irow >> 2) & 3); }
__device__ __forceinline__ int row_ix(int irow) { return ((irow >> 4) & 1) * 4 + (irow & 3); }
__device__ __forceinline__ int win_region(const Geom& g, int wy, int wx, int iy, int ix) {
  const int hs = wy * 8 + iy, ws = wx * 8 + ix;             // coordinates on the shifted frame (calculate_mask, :321-341)
  const int rh = hs < g.H - 8 ? 0 : (hs < g.H - g.shift ? 1 : 2);
  const int rw = ws < g.W - 8 ? 0 : (ws < g.W - g.shift ? 1 : 2);
  return rh * 3 + rw;
}

// 4 role warpgroups + one warpgroup whose first warp issues the MMAs (registers are allocated per warpgroup, so a 17th
// warp costs as much as four).  640 threads = 96 registers each; every role is written to fit (no setmaxnreg: ptxas
// refuses to spill at all once it is used, and the MMA role does not fit the 32 registers a useful hand-over needs).
constexpr int THREADS = 640;
enum Bar {
  B_W = 0, B_LFULL, B_EFULL, B_XH_READY, B_XH_FREE, B_AP_READY, B_PROJ_FULL, B_PROJ_DRAINED,
  B_QKV_FULL = 8, B_QK_DRAINED = 11, B_V_DRAINED = 14, B_S_FULL = 17, B_P_READY = 20, B_O_FULL = 23,      // one per slot (<= 3)
  B_AP_FREE = 26, B_WPROJ = 27, B_STAGED = 28, B_WHEAD = 29 /* 6: qkv weights of one head each */, NBARS = 35
};

__device__ __forceinline__ void wgA_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// one arrival per warp: every lane's prior tcgen05.ld/st has completed (wait::ld / wait::st by the caller)
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

template <int C_, bool DBG>
__global__ void __launch_bounds__(THREADS, 1)
stl_attn2_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                 const uint8_t* __restrict__ wqkv_img, const uint8_t* __restrict__ wproj_img,
                 const float* __restrict__ bqkv, const float* __restrict__ bproj, const float* __restrict__ table,
                 Geom geo, float mask_val, unsigned long long* __restrict__ dbg) {
  using K = Cfg<C_>;
  constexpr int CP = K::CP, HD = K::HD, NH = K::NH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BARS);
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smem + K::OFF_BARS + 38 * 8);
  static_assert(NBARS <= 38, "barrier block");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2;                         // 0 = A, 1 = B, 2 = C, 3 = D, 4 = MMA warp
  const int row = tid & 127;
  uint8_t* const sL = smem + K::OFF_L;
  uint8_t* const sE = smem + K::OFF_E;
  uint8_t* const sY = smem + (K::YBUF ? K::OFF_Y : K::OFF_E);      // output staging (= E when there is no room for both)

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < NBARS; ++i) {
      const bool w4 = i == B_XH_READY || i == B_AP_READY || i == B_PROJ_DRAINED || (i >= B_QK_DRAINED && i < B_S_FULL) ||
                      (i >= B_P_READY && i < B_O_FULL);
      mbar_init(&bars[i], w4 ? 4 : (i == B_XH_FREE ? K::NS : 1));  // x^ is free when ALL slot issuers are past their last qkv
    }
    fence_mbar_init();
    // resident weights: one bulk copy and one barrier PER HEAD, so that the first tile starts as soon as head 0's 16 KB are
    // there instead of after all 130 KB (every CTA pulls the same image out of L2 at launch: ~8000 cycles at C = 120)
    constexpr int WH = NH * CP * 2;
    static_assert(WH <= 32768 && K::WPROJ_BYTES <= 32768, "one bulk copy each");
    for (int h = 0; h < 6; ++h) {
      mbar_arrive_expect_tx(&bars[B_WHEAD + h], WH);
      bulk_g2s(smem + K::OFF_WQKV + h * WH, wqkv_img + h * WH, WH, &bars[B_WHEAD + h]);
    }
    mbar_arrive_expect_tx(&bars[B_WPROJ], K::WPROJ_BYTES);
    bulk_g2s(smem + K::OFF_WPROJ, wproj_img, K::WPROJ_BYTES, &bars[B_WPROJ]);
  }
  // K / V images start as zeros (pads must be, and stay, zero / finite); bias table -> shared memory
  for (int i = tid; i < K::NS * K::KV_BYTES / 16; i += THREADS)
    *reinterpret_cast<uint4*>(smem + K::OFF_KV + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 6 * K::TBL; i += THREADS)
    reinterpret_cast<uint32_t*>(smem + K::OFF_TAB)[i] = reinterpret_cast<const uint32_t*>(table)[i];
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int ntiles = (geo.nwt + 1) / 2;
  const int NT = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA (>= 1)
  const int G = 6 * NT;                                                                // global head count
  int dbg_n = 0;
  const bool dbg_on = DBG && dbg != nullptr && blockIdx.x == 0 && (tid & 127) == 0;
#define A2_STAMP()                                                                            \
  do {                                                                                        \
    if (DBG && dbg_on && dbg_n < 255) dbg[wg * 256 + 1 + dbg_n++] = clock64();                \
  } while (0)

  auto load_tile = [&](int tile, uint8_t* dst, uint64_t* bar) {      // one lane: both windows of `tile`, 4 boxes x NP panels each
    mbar_arrive_expect_tx(bar, K::XT_BYTES);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int win = tile * 2 + w;
      const int b = win < geo.nwt ? win / geo.nw_img : geo.nwt / geo.nw_img;     // beyond the end: out of range -> zero fill
      const int wl = win < geo.nwt ? win - b * geo.nw_img : 0;
      const int wy = wl / geo.nwx, wx = wl - wy * geo.nwx;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int hh = wy * 8 + 4 * (q >> 1) + geo.shift; if (hh >= geo.H) hh -= geo.H;   // shifted[h'] = x[(h'+s) mod H]  (:245)
        int ww = wx * 8 + 4 * (q & 1) + geo.shift; if (ww >= geo.W) ww -= geo.W;
#pragma unroll
        for (int pnl = 0; pnl < K::NP; ++pnl)
          tma::load_4d(dst + pnl * K::PANEL + (w * 64 + q * 16) * 128, &mapX, pnl * 64, ww, hh, b, bar);
      }
    }
  };
  auto store_tile = [&](int tile, const uint8_t* src) {
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int win = tile * 2 + w;
      if (win >= geo.nwt) break;
      const int b = win / geo.nw_img;
      const int wl = win - b * geo.nw_img;
      const int wy = wl / geo.nwx, wx = wl - wy * geo.nwx;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int hh = wy * 8 + 4 * (q >> 1) + geo.shift; if (hh >= geo.H) hh -= geo.H;
        int ww = wx * 8 + 4 * (q & 1) + geo.shift; if (ww >= geo.W) ww -= geo.W;
#pragma unroll
        for (int pnl = 0; pnl < K::NP; ++pnl)
          tma::store_4d(&mapY, pnl * 64, ww, hh, b, src + pnl * K::PANEL + (w * 64 + q * 16) * 128);
      }
    }
    bulk_commit();
  };

  pdl_launch_dependents();
  pdl_wait();                    // everything above touched only weights; from here on we read the producer's output

  if (wg == 0) {
    // =============================== role A: LayerNorm (one tile ahead), proj epilogue, TMA ===============================
    // Per period t the role touches two tiles: LayerNorm of tile t+1 (its x^ must be in TMEM before the issuers reach
    // heads 0/1 of that tile, i.e. about two thirds into tile t) and the epilogue of tile t-1 (its proj blocks arrive
    // early in tile t).  The two are interleaved in the order their inputs become available:
    //     statistics(t+1)  ->  epilogue block 0 (t-1)  ->  normalise(t+1) -> x^, next landing  ->  epilogue block 1 (t-1), store
    // (doing the whole epilogue first left x^ ~4000 cycles late at every tile boundary, measured.)
    const float inv_c = 1.0f / (float)C_;
    float ln_rstd = 0.f, ln_nb = 0.f;
    auto ln_stats = [&](int n) {                           // pass 1 over the landed rows of tile n
      mbar_wait(&bars[B_LFULL], n & 1);
      A2_STAMP();   // A: tile landed
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int c = 0; c < K::NCH; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(sL + xt_off(row, c));
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = up2(w4[q]);
          s0 += f.x; s1 += f.y;
          q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
        }
      }
      const float mean = (s0 + s1) * inv_c;
      const float var = (q0 + q1) * inv_c - mean * mean;
      ln_rstd = rsqrtf(fmaxf(var, 0.f) + 1e-5f);
      ln_nb = -mean * ln_rstd;
    };
    auto ln_write = [&](int n) {                           // pass 2: normalised bf16 rows of tile n -> TMEM, then L is free
      if (n >= 1) { mbar_wait(&bars[B_XH_FREE], (n - 1) & 1); fence_after_sync(); }   // both issuers are past the previous tile's last qkv
      A2_STAMP();   // A: XH free
#pragma unroll
      for (int c0 = 0; c0 < K::NCH; c0 += 4) {             // 4 chunks = 32 channels = 16 packed columns per store
        uint32_t a[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(sL + xt_off(row, c0 + c));
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = up2(w4[q]);
            a[4 * c + q] = pk2(fmaf(f.x, ln_rstd, ln_nb), fmaf(f.y, ln_rstd, ln_nb));
          }
        }
        if (c0 == 4) a[14] = 0x3F803F80u;                  // pad channels 60, 61 carry the ones of the folded qkv bias
        tmem_st_x16(lane_addr + K::TM_XH + 4 * c0, a);
      }
      wgA_sync();                                          // every row of L has been read twice: the next tile may land
      if (warp == 0 && n + 1 < NT) {
        if (elect_one()) load_tile(blockIdx.x + (n + 1) * gridDim.x, sL, &bars[B_LFULL]);
        __syncwarp();
      }
      wait_st();
      warp_arrive(&bars[B_XH_READY], lane);
    };
    auto epi_block = [&](int pn, int hf) {                 // y = proj + bias + x for one column block -> staging tile
      mbar_wait(&bars[B_PROJ_FULL], (pn * K::NBLK + hf) & 1);
      fence_after_sync();
      A2_STAMP();   // A: proj block ready
      uint32_t acc[K::NPC];
#pragma unroll
      for (int c0 = 0; c0 < K::NPC; c0 += 16) {
        uint32_t t[16];
        tmem_ld_x16(lane_addr + K::TM_PROJ + c0, t);
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[c0 + e] = t[e];
      }
      wait_ld();
      warp_arrive(&bars[B_PROJ_DRAINED], lane);            // the accumulator may be overwritten by the next block
      if (hf == 0) {
        mbar_wait(&bars[B_EFULL], pn & 1);
        if (K::YBUF && pn >= 1) {                          // the store of the previous tile (issued a whole tile ago) has left Y
          if (warp == 0) {
            if (elect_one()) bulk_wait_read();
            __syncwarp();
          }
          wgA_sync();
        }
      }
#pragma unroll
      for (int c0 = 0; c0 < K::NPC; c0 += 8) {
        const uint32_t off = xt_off(row, (hf * K::NPC + c0) >> 3);
        const uint4 xv = *reinterpret_cast<const uint4*>(sE + off);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 xf = up2(xw[e]);
          y[e] = pk2(__uint_as_float(acc[c0 + 2 * e]) + xf.x, __uint_as_float(acc[c0 + 2 * e + 1]) + xf.y);
        }
        *reinterpret_cast<uint4*>(sY + off) = make_uint4(y[0], y[1], y[2], y[3]);
      }
    };
    if (warp == 0) {
      if (elect_one()) load_tile(blockIdx.x, sL, &bars[B_LFULL]);
      __syncwarp();
    }
#pragma unroll 1
    for (int t = -1; t <= NT; ++t) {                       // t = -1: LayerNorm of the first tile only
      if (t + 1 < NT) ln_stats(t + 1);
      // two slots: x^ of tile t+1 can only be written once head 5 of tile t has its qkv (about a third into tile t), the
      // first proj block of tile t-1 arrives earlier.  Three slots: all qkv of tile t are issued during tile t-1, x^ first.
      if (K::NS == 2 && t >= 1) epi_block(t - 1, 0);
      if (t + 1 < NT) ln_write(t + 1);
      if (t >= 1) {
#pragma unroll
        for (int hf = (K::NS == 2 ? 1 : 0); hf < K::NBLK; ++hf) epi_block(t - 1, hf);
        fence_proxy_async();                               // the finished rows are read by the TMA store (async proxy)
        wgA_sync();
        A2_STAMP();   // A: y staged
      }
      if (K::YBUF) {
        // E and Y are separate: the store needs no wait here (Y is written again a whole tile later), and the residual rows
        // of tile t (second fetch, an L2 hit) can land in E right away
        if (warp == 0) {
          if (elect_one()) {
            if (t >= 1) store_tile(blockIdx.x + (t - 1) * gridDim.x, sY);
            if (t >= 0 && t < NT) load_tile(blockIdx.x + t * gridDim.x, sE, &bars[B_EFULL]);
          }
          __syncwarp();
        }
      } else {
        if (t >= 1 && tid == 0) mbar_arrive(&bars[B_STAGED]);   // warp 19 stores the tile and refills E (below)
      }
    }
    if (K::YBUF && warp == 0) {
      if (elect_one()) bulk_wait_read();
      __syncwarp();
    }
  } else if (wg == 1) {
    // =============================== role B: qkv drain (head g), O / rowsum (head g-2) ===============================
    // Iteration g:  drain(g)  ->  [PV(g-2) complete]  load O(g-2)  ->  write V(g), release it  ->  normalise O(g-2) -> AP.
    // V is released before the normalised O is stored because that store may have to wait for the proj of the previous
    // tile (AP is single-buffered) and PV(g) must not wait with it.  When O(g-2) is already there at the top of the
    // iteration (qkv(g) late, e.g. at a tile boundary) it is handled first, so that the last head of a tile -- which
    // triggers the proj chain everybody downstream waits for -- never queues behind a late qkv.
    uint4 vimg[K::VCH];
    constexpr int NCO = (HD + 8) / 8 * 8;                      // head_dim values + the row-sum column: 16 / 16 / 24
    uint32_t fo[NCO];
    auto o_load = [&](int gp) {
      const int s = gp % K::NS;
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 8) {
        uint32_t t[8];
        tmem_ld_x8(lane_addr + K::TM_O + K::HDV * s + c0, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) fo[c0 + e] = t[e];
      }
      wait_ld();                                               // (the V_DRAINED arrival of head gp+2 also tells the issuer that O is free)
      fence_before_sync();
    };
    auto o_store = [&](int gp) {
      const int np = gp / 6, hp = gp - 6 * np;
      if (hp == 0 && np >= 1) {                                 // proj of the previous tile has read the normalised O
        mbar_wait(&bars[B_AP_FREE], (np - 1) & 1);              // (one completion per tile: a waiter never lags two phases)
        fence_after_sync();
      }
      const float inv = 1.0f / __uint_as_float(fo[HD]);
      auto ov = [&](int d) { return __uint_as_float(fo[d]) * inv; };
      if (K::HDO == 16) {
        uint32_t a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int d0 = 2 * e, d1 = d0 + 1;
          a[e] = pk2(d0 < HD ? ov(d0) : 0.f, d1 < HD ? ov(d1) : 0.f);
        }
        // ones of the folded proj bias: C=60 k = 10, 11 (head 0); C=90 k = 15 and 31 (heads 0, 1)
        if (C_ == 60 && hp == 0) a[5] = 0x3F803F80u;
        if (C_ == 90 && hp < 2) a[7] = (a[7] & 0xFFFFu) | 0x3F800000u;
        tmem_st_x8(lane_addr + K::TM_AP + 8 * hp, a);
      } else {                                                   // head_dim 20: 10 packed columns per head
        uint32_t a[8], b[2];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = pk2(ov(2 * e), ov(2 * e + 1));
        b[0] = pk2(ov(16), ov(17));
        b[1] = pk2(ov(18), ov(19));
        tmem_st_x8(lane_addr + K::TM_AP + 10 * hp, a);
        tmem_st_x2(lane_addr + K::TM_AP + 10 * hp + 8, b);
        if (hp == 5) {                                           // k = 120, 121: ones of the folded proj bias; rest of the pad zero
          uint32_t zz[4] = {0x3F803F80u, 0, 0, 0};
          tmem_st_x4(lane_addr + K::TM_AP + 60, zz);
        }
      }
      if (hp == 5) {
        wait_st();
        warp_arrive(&bars[B_AP_READY], lane);
        A2_STAMP();   // B: normalised O complete
      }
    };
#pragma unroll 1
    for (int g = 0; g < G + K::NS; ++g) {
      const int s = g % K::NS;                                   // slot of head g (and of head g - NS, whose O is handled here)
      const uint32_t ph = (g / K::NS) & 1;
      uint8_t* sBk = smem + K::OFF_KV + s * K::KV_BYTES;
      uint8_t* sBv = sBk + K::BK_BYTES;
      bool o_early = false;
      if (g >= K::NS && g < G && __all_sync(0xffffffffu, mbar_test(&bars[B_O_FULL + s], ph ^ 1))) {
        fence_after_sync();
        o_load(g - K::NS);
        o_store(g - K::NS);
        o_early = true;
      }
      if (g < G) {
        mbar_wait(&bars[B_QKV_FULL + s], ph);
        fence_after_sync();
        A2_STAMP();   // B: qkv ready
        constexpr int NC = (3 * HD + 7) / 8 * 8;                // 32 / 48 / 64 accumulator columns: q | k | v
        uint32_t f[NC];
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 16) {
          uint32_t t[16];
          tmem_ld_x16(lane_addr + K::TM_QKV + NH * s + c0, t);
#pragma unroll
          for (int e = 0; e < 16; ++e) f[c0 + e] = t[e];
        }
        wait_ld();
        auto val = [&](int i) { return __uint_as_float(f[i]); };
        // q -> packed bf16 back into the consumed accumulator columns (A operand of S = Q K^T), pads exact zeros
        uint32_t qp[K::HDP / 2];
#pragma unroll
        for (int e = 0; e < K::HDP / 2; ++e) {
          const int d0 = 2 * e, d1 = d0 + 1;
          qp[e] = pk2h(d0 < HD ? val(d0) : 0.f, d1 < HD ? val(d1) : 0.f);      // fp16: S accumulates in fp16 (below)
        }
        if constexpr (K::HDP / 2 == 8) {
          tmem_st_x8(lane_addr + K::TM_QKV + NH * s, qp);
        } else {
          tmem_st_x16(lane_addr + K::TM_QKV + NH * s, qp);
        }
        // k -> K-major bf16 image (the previous S of this slot completed before this head's qkv did)
#pragma unroll
        for (int c8 = 0; c8 < K::KCH; ++c8) {
          uint32_t w4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int d0 = c8 * 8 + 2 * q, d1 = d0 + 1;
            w4[q] = pk2h(d0 < HD ? val(HD + d0) : 0.f, d1 < HD ? val(HD + d1) : 0.f);
          }
          *reinterpret_cast<uint4*>(sBk + c8 * 2048 + row * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
        // v (+ a column of ones: the PV MMA then accumulates the softmax row sum) -> fp16, written below
#pragma unroll
        for (int c8 = 0; c8 < K::VCH; ++c8) {
          uint32_t w4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int d0 = c8 * 8 + 2 * q, d1 = d0 + 1;
            const float a = d0 < HD ? val(2 * HD + d0) : (d0 == HD ? 1.f : 0.f);
            const float b = d1 < HD ? val(2 * HD + d1) : (d1 == HD ? 1.f : 0.f);
            w4[q] = pk2h(a, b);
          }
          vimg[c8] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
        wait_st();
        fence_proxy_async();
        warp_arrive(&bars[B_QK_DRAINED + s], lane);
        A2_STAMP();   // B: q, k drained
      }
      if (g >= K::NS && !o_early) {
        mbar_wait(&bars[B_O_FULL + s], ph ^ 1);                  // PV(g-NS) complete: O ready, the V image is free
        fence_after_sync();
        o_load(g - K::NS);
      }
      if (g < G) {
#pragma unroll
        for (int c8 = 0; c8 < K::VCH; ++c8) *reinterpret_cast<uint4*>(sBv + c8 * 2048 + row * 16) = vimg[c8];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_V_DRAINED + s]);
      }
      if (g >= K::NS && !o_early) o_store(g - K::NS);
    }
  } else if (wg < 4) {
    // =============================== roles C / D: softmax of even / odd heads, thread = one row x 64 keys ===============================
    const int wsm = wg - 2;                                      // this warpgroup takes heads g = wsm (mod 2)
    const uint32_t* sTab2 = reinterpret_cast<const uint32_t*>(smem + K::OFF_TAB);
    const int wsel = row >> 6, irow = row & 63, iy = row_iy(irow), ix = row_ix(irow);
    __half2 mk[4];
    bool masked = false;
    for (int g = wsm; g < G; g += 2) {
      const int n = g / 6, h = g - 6 * n;
      const int s = g % K::NS;                                   // head slot
      const uint32_t tS = lane_addr + K::TM_S + 64 * s;
      if (h == wsm) {
        // shift mask of this tile (edge windows only): the region borders of calculate_mask (:321-341) cut a window
        // exactly between the 4x4 boxes that define the row order, so the mask of 64 keys is four per-box constants
        masked = false;
        const int win = (blockIdx.x + n * gridDim.x) * 2 + wsel;
        if (geo.shift > 0 && win < geo.nwt) {
          const int wl = win % geo.nw_img;
          const int wy = wl / geo.nwx, wx = wl - wy * geo.nwx;
          if (wy == geo.H / 8 - 1 || wx == geo.nwx - 1) {
            masked = true;
            const int mine = win_region(geo, wy, wx, iy, ix);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              mk[q] = __float2half2_rn(win_region(geo, wy, wx, 4 * (q >> 1), 4 * (q & 1)) != mine ? mask_val : 0.f);
          }
        }
      }
      mbar_wait(&bars[B_S_FULL + s], (g / K::NS) & 1);
      fence_after_sync();
      A2_STAMP();   // C/D: S ready
      uint32_t p[32];                                            // half2 logits, then probabilities, in place
      tmem_ld_x32_pack16(tS, p);                                 // 64 fp16 logits, keys 2j / 2j+1 in register j
      wait_ld();
      // logits -> half2, + relative-position bias pair (one 32-bit table read per two keys), + mask
      const uint32_t* tb = sTab2 + h * K::TBL + (iy + 7) * 24 + ix + 7;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int kk = 2 * j, q = kk >> 4;                       // keys 2j, 2j+1 in box order
        const int jy = (q >> 1) * 4 + ((kk >> 2) & 3), jx = (q & 1) * 4 + (kk & 3);
        const uint32_t bp = tb[-(jy * 24 + jx)];
        __half2 t = __hadd2(*reinterpret_cast<const __half2*>(&p[j]), *reinterpret_cast<const __half2*>(&bp));
        if (masked) t = __hadd2(t, mk[j >> 3]);
        p[j] = *reinterpret_cast<const uint32_t*>(&t);
      }
      auto H2 = [&](int j) { return *reinterpret_cast<const __half2*>(&p[j]); };
      __half2 m4[4] = {H2(0), H2(1), H2(2), H2(3)};
#pragma unroll
      for (int j = 4; j < 32; ++j) m4[j & 3] = __hmax2(m4[j & 3], H2(j));
      const __half2 mm = __hmax2(__hmax2(m4[0], m4[1]), __hmax2(m4[2], m4[3]));
      const __half mx = __hmax(__low2half(mm), __high2half(mm));
      const __half2 mx2 = __half2half2(mx);
      // P stays fp16 (V is fp16 as well); the row sum comes out of the PV MMA (ones column of V)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const __half2 d = __hsub2(H2(j), mx2);
        p[j] = ex2_h2(*reinterpret_cast<const uint32_t*>(&d));
      }
      tmem_st_x32(tS, p);                                        // P overwrites the first 32 columns of S
      wait_st();
      warp_arrive(&bars[B_P_READY + s], lane);
      A2_STAMP();   // C/D: P written
    }
  }
  if (!K::YBUF && warp == 19) {
    // =============================== E-side TMA: store of finished tiles, second fetch of the residual rows ===============================
    // A TMA store needs ~3000 cycles before it has read its 32 KB out of shared memory (16 small boxes; measured), and E
    // cannot take the next tile's residual rows earlier.  On the LayerNorm/epilogue role that wait delayed every proj
    // drain behind it; here it costs nothing.
    for (int t = 0; t <= NT; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      if (t >= 1) mbar_wait(&bars[B_STAGED], (t - 1) & 1);   // y of tile t-1 is complete in E (writers fenced the async proxy)
      if (elect_one()) {
        if (t >= 1) { store_tile(tile - gridDim.x, sE); bulk_wait_read(); }
        if (t < NT) load_tile(tile, sE, &bars[B_EFULL]);     // residual rows of tile t (an L2 hit) for its epilogue
      }
      __syncwarp();
    }
  }
  if (wg == 4 && warp <= 16 + K::NS) {
    // =============================== MMA issue: three warps, each with its own static program ===============================
    // The tensor pipe executes MMAs in the order they are issued, whoever issues them.  All MMAs that touch the resources
    // of head slot s (its qkv accumulator / packed Q, its S / P columns, its O columns, its K / V images) are issued by
    // ONE warp (16 + s), so every "X before Y" on a slot holds by that warp's program order:
    //     iteration j (heads of this slot, NS apart):  PV(j) ;  S(j+NS) ;  qkv(j+2 NS)
    // PV(j) waits for the softmax of head j; S(j+NS) right behind it reuses the S/P columns PV(j) has just read; qkv(j+2 NS)
    // reuses the accumulator whose Q operand S(j+NS) has just read (role B copied the rest into registers before it
    // released q/k).  Warp 16 + NS issues the proj blocks.  Splitting the issue role matters because the control path of an issuer (barrier
    // probes at ~100-150 cycles each, elect, commit) costs ~2000 cycles per head when one warp does everything (measured),
    // i.e. it, not the tensor pipe (~650 cycles of MMAs per head), paced the first version of this kernel.
    // The whole warp runs the loop with warp-uniform values and only the tcgen05 instructions sit under elect.sync:
    // descriptors then live in uniform registers (one lane of a divergent branch costs a ~60-cycle waterfall per MMA).
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t aWqkv = smem_u32(smem + K::OFF_WQKV), aWproj = smem_u32(smem + K::OFF_WPROJ);
    const uint32_t aKV = smem_u32(smem + K::OFF_KV);
    constexpr uint32_t idq = make_idesc_bf16(128, NH, false, false);
    // S = Q K^T on fp16 operands with an fp16 accumulator: the softmax needs its logits as fp16 pairs anyway, and an fp16
    // accumulator comes back from TMEM two values per register (tcgen05.ld .pack::16b) -- no 32 F2FP per row and head, half
    // the registers.  K = 32 is two k-steps, i.e. two fp16 roundings instead of one.
    constexpr uint32_t ids = make_idesc_f16_acc16(128, 64);
    constexpr uint32_t idv = make_idesc_f16(128, K::HDV, false, true);
    constexpr uint32_t idp = make_idesc_bf16(128, K::NPC, false, false);
    if (warp == 16 + K::NS) {
      // proj weights landed: fold the bias into the spare K rows of the resident image (bf16 hi/lo pair, see tc_attn.cu)
      mbar_wait(&bars[B_WPROJ], 0);
      for (int nn = lane; nn < CP; nn += 32) {
        const uint32_t hl = bias_hi_lo(bproj[nn]);
        uint8_t* wp = smem + K::OFF_WPROJ;
        if (C_ == 60) *reinterpret_cast<uint32_t*>(wp + (1 * CP + nn) * 16 + 4) = hl;            // k = 10, 11
        else if (C_ == 120) *reinterpret_cast<uint32_t*>(wp + (15 * CP + nn) * 16) = hl;         // k = 120, 121
        else {                                                                                   // k = 15, 31
          *reinterpret_cast<uint16_t*>(wp + (1 * CP + nn) * 16 + 14) = (uint16_t)(hl & 0xFFFFu);
          *reinterpret_cast<uint16_t*>(wp + (3 * CP + nn) * 16 + 14) = (uint16_t)(hl >> 16);
        }
      }
      fence_proxy_async();
      __syncwarp();
      for (int k = 0; k < NT * K::NBLK; ++k) {
        const int pn = k / K::NBLK, hf = k - pn * K::NBLK;
        if (hf == 0) mbar_wait(&bars[B_AP_READY], pn & 1);
        if (k >= 1) mbar_wait(&bars[B_PROJ_DRAINED], (k - 1) & 1);               // previous block is in registers
        fence_after_sync();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < K::KPROJ / 16; ++ks)
            mma_ts(tm + K::TM_PROJ, tm + K::TM_AP + ks * 8,
                   make_smem_desc(aWproj + hf * K::NPC * 16 + ks * 2 * (CP * 16), CP * 16, 128), idp, ks > 0);
          commit(&bars[B_PROJ_FULL]);
          if (hf == K::NBLK - 1) commit(&bars[B_AP_FREE]);       // the normalised O of this tile has been read
        }
        __syncwarp();
      }
    } else {
      const int s = warp - 16;                                   // head slot of this issuer
      constexpr int NS = K::NS;
      const uint32_t tQ = tm + K::TM_QKV + NH * s, tS = tm + K::TM_S + 64 * s, tO = tm + K::TM_O + K::HDV * s;
      const uint32_t aBk = aKV + s * K::KV_BYTES, aBv = aBk + K::BK_BYTES;
      uint64_t* const bQKV = &bars[B_QKV_FULL + s];
      uint64_t* const bQK = &bars[B_QK_DRAINED + s];
      uint64_t* const bV = &bars[B_V_DRAINED + s];
      uint64_t* const bS = &bars[B_S_FULL + s];
      uint64_t* const bP = &bars[B_P_READY + s];
      uint64_t* const bO = &bars[B_O_FULL + s];
      auto qkv_mmas = [&](int h) {                               // elected lane only
        const uint32_t wb = aWqkv + h * (NH * CP * 2);
#pragma unroll
        for (int ks = 0; ks < CP / 16; ++ks)
          mma_ts(tQ, tm + K::TM_XH + ks * 8, make_smem_desc(wb + ks * 2 * (NH * 16), NH * 16, 128), idq, ks > 0);
        commit(bQKV);
        if (h >= 6 - NS) commit(&bars[B_XH_FREE]);               // last qkv of this slot in the tile (NS arrivals free x^)
      };
      auto s_mmas = [&]() {
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
          for (int ks = 0; ks < K::HDP / 16; ++ks)
            mma_bf16_ts_masked(tS, tQ + ks * 8, make_smem_desc(aBk + w * 1024 + ks * 4096, 2048, 128), ids, ks > 0,
                               w ? 0xFFFFFFFFu : 0u, w ? 0xFFFFFFFFu : 0u, w ? 0u : 0xFFFFFFFFu, w ? 0u : 0xFFFFFFFFu);
        commit(bS);
      };
      auto pv_mmas = [&]() {
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ts_masked(tO, tS + ks * 8, make_smem_desc(aBv + w * 1024 + ks * 256, 128, 2048), idv, ks > 0,
                               w ? 0xFFFFFFFFu : 0u, w ? 0xFFFFFFFFu : 0u, w ? 0u : 0xFFFFFFFFu, w ? 0u : 0xFFFFFFFFu);
        commit(bO);
      };
      auto prep_head = [&](int h) {                              // whole warp, first tile only: head h's weights landed ->
        mbar_wait(&bars[B_WHEAD + h], 0);                        // fold its qkv bias into K rows 60 / 61 of the resident image
        for (int nn = lane; nn < NH; nn += 32)
          *reinterpret_cast<uint32_t*>(smem + K::OFF_WQKV + h * (NH * CP * 2) + (7 * NH + nn) * 16 + 8) = bias_hi_lo(bqkv[h * NH + nn]);
        fence_proxy_async();
        __syncwarp();
      };
      // head g of this slot is its (g / NS)-th: that is the phase index of all per-slot barriers.
      //     iteration j:  PV(j) ;  S(j+NS) ;  qkv(j+2 NS)
      // prologue: qkv(s) ; S(s) ; qkv(s+NS)
      prep_head(s);
      mbar_wait(&bars[B_XH_READY], 0);
      fence_after_sync();
      if (elect_one()) qkv_mmas(s);
      __syncwarp();
      prep_head(s + NS);
      mbar_wait(bQK, 0);
      fence_after_sync();
      if (elect_one()) { s_mmas(); qkv_mmas(s + NS); }
      __syncwarp();
      for (int j = s; j < G; j += NS) {
        const uint32_t ph = (j / NS) & 1;
        const bool has_s = j + NS < G;
        const int g4 = j + 2 * NS, h4 = g4 % 6;
        const bool q_fused = g4 < G && h4 >= NS;                 // qkv of the same tile as S(j+NS): goes out in the same breath
        if (g4 < 6) prep_head(g4);                               // (first tile: that head's weights and bias)
        // Everything that is normally long complete is waited for FIRST (V of head j -- which also means O of head j-NS is
        // in role B's registers -- and q/k of head j+NS), so that the only wait left on the softmax -> PV -> S chain is P.
        mbar_wait(bV, ph);
        const bool qk_early = has_s && __all_sync(0xffffffffu, mbar_test(bQK, ph ^ 1)) != 0;
        mbar_wait(bP, ph);                                       // softmax of head j delivered P
        fence_after_sync();
        A2_STAMP();   // MMA: P, V ready
        if (elect_one()) {
          pv_mmas();
          if (qk_early) {                                        // PV(j), S(j+NS) (+ qkv) in one go
            s_mmas();
            if (q_fused) qkv_mmas(h4);
          }
        }
        __syncwarp();
        A2_STAMP();   // MMA: PV (+ S) issued
        if (has_s) {
          if (!qk_early) {                                       // role B was late with q/k of head j+NS
            mbar_wait(bQK, ph ^ 1);
            fence_after_sync();
            if (elect_one()) {
              s_mmas();
              if (q_fused) qkv_mmas(h4);
            }
            __syncwarp();
          }
          A2_STAMP();   // MMA: S (+ qkv) issued
          if (g4 < G && h4 < NS) {
            // first head of this slot in the next tile: x^ of that tile has to be in TMEM.  This warp has nothing else
            // to do until the softmax of head j+NS delivers (> 1000 cycles), so it simply waits here; everything the
            // LayerNorm role needs on the way (all issuers past their last qkv of the tile, the proj blocks of the tile
            // before) is issued by other warps or earlier in this program.
            mbar_wait(&bars[B_XH_READY], (g4 / 6) & 1);
            fence_after_sync();
            if (elect_one()) qkv_mmas(h4);
            __syncwarp();
          }
        }
      }
    }
  }
#undef A2_STAMP
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static unsigned long long* g_dbg = nullptr;

template <int C_>
static int launch(const void* x, int64_t ldx, void* y, int64_t ldy, const void* wqkv, const void* wproj,
                  const float* bqkv, const float* bproj, const float* table, int B, int H, int W, int shift,
                  int sms, cudaStream_t st) {
  using K = Cfg<C_>;
  Geom g;
  g.H = H; g.W = W; g.shift = shift; g.nwx = W / 8; g.nw_img = (H / 8) * (W / 8);
  g.nwt = B * g.nw_img;
  const int64_t ntiles = ((int64_t)g.nwt + 1) / 2;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  const CUtensorMap* mx = get_act_tmap(x, ldx, B, H, W, K::CP, 4, 4);
  const CUtensorMap* my = get_act_tmap(y, ldy, B, H, W, K::CP, 4, 4);
  if (!mx || !my) return RDST_E_CUDA;
  auto k = g_dbg ? stl_attn2_kernel<C_, true> : stl_attn2_kernel<C_, false>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
  if (e != cudaSuccess) { set_error("rdst_stl_attn_fwd_bf16: smem attr (%d B): %s", K::SMEM, cudaGetErrorString(e)); return RDST_E_CUDA; }
  e = launch_pdl(k, dim3(grid), dim3(THREADS), (size_t)K::SMEM, st, *mx, *my,
                 (const uint8_t*)wqkv, (const uint8_t*)wproj, bqkv, bproj, table, g, -100.0f * 1.4426950408889634f, g_dbg);
  if (e != cudaSuccess) { set_error("rdst_stl_attn_fwd_bf16: launch: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  return RDST_OK;
}

}  // namespace a2

// round-1 kernel (tc_attn.cu), kept selectable for A/B timing
int attn_v1_dispatch(const void* x, int64_t ldx, void* y, int64_t ldy, const void* wqkv_img, const void* wproj_img,
                     const float* bqkv, const float* bproj, const float* table, int B, int H, int W, int C, int shift,
                     int sms, cudaStream_t st);
static int g_attn_variant = 2;

}  // namespace rdst

extern "C" int rdst_debug_attn_variant(int variant) {
  if (variant != 1 && variant != 2) { rdst::set_error("rdst_debug_attn_variant: 1 (lock-step kernel) or 2 (warp-specialised)"); return RDST_E_INVALID; }
  rdst::g_attn_variant = variant;
  return RDST_OK;
}

extern "C" int rdst_debug_attn2_timing(void* device_buffer_1280_u64) {
  rdst::a2::g_dbg = (unsigned long long*)device_buffer_1280_u64;
  return RDST_OK;
}

extern "C" int rdst_stl_attn_fwd_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, const void* wqkv_img,
                                      const void* wproj_img, const float* bqkv, const float* bproj, const float* table,
                                      int B, int H, int W, int C, int shift, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && y && wqkv_img && wproj_img && bqkv && bproj && table, "rdst_stl_attn_fwd_bf16: null pointer");
  RDST_REQUIRE(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0,
               "rdst_stl_attn_fwd_bf16: H=%d W=%d must be positive multiples of the window size 8", H, W);
  RDST_REQUIRE(shift == 0 || shift == 4, "rdst_stl_attn_fwd_bf16: shift must be 0 or 4");
  RDST_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ldx % 8 == 0 && ldy % 8 == 0,
               "rdst_stl_attn_fwd_bf16: x/y must be 16-byte aligned with row strides multiple of 8 elements");
  RDST_REQUIRE(x != y, "rdst_stl_attn_fwd_bf16: in-place operation is not supported");
  RDST_REQUIRE(C == 60 || C == 90 || C == 120, "rdst_stl_attn_fwd_bf16: C=%d unsupported (60, 90, 120 with 6 heads)", C);
  const int cp = C == 60 ? 64 : (C == 90 ? 96 : 128);
  RDST_REQUIRE(ldx >= cp && ldy >= cp, "rdst_stl_attn_fwd_bf16: row stride smaller than the padded width %d", cp);
  if (B <= 0) return RDST_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (g_attn_variant == 1) {
    rc = attn_v1_dispatch(x, ldx, y, ldy, wqkv_img, wproj_img, bqkv, bproj, table, B, H, W, C, shift, sms, st);
  } else {
    switch (C) {
      case 60:  rc = a2::launch<60>(x, ldx, y, ldy, wqkv_img, wproj_img, bqkv, bproj, table, B, H, W, shift, sms, st); break;
      case 90:  rc = a2::launch<90>(x, ldx, y, ldy, wqkv_img, wproj_img, bqkv, bproj, table, B, H, W, shift, sms, st); break;
      default:  rc = a2::launch<120>(x, ldx, y, ldy, wqkv_img, wproj_img, bqkv, bproj, table, B, H, W, shift, sms, st); break;
    }
  }
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_stl_attn_fwd_bf16");
  return RDST_OK;
}
