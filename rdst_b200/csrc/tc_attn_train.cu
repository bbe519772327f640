// Shifted-window attention of the TRAINING path on tcgen05 (precision 'bf16'): forward and backward of
//   O = softmax(q k^T + B_rel + mask) v     per 8x8 window and head,   q | k | v rows in an fp32 [T][ldq] buffer
// (reference WindowAttention.forward, swin_transformer_sr.py:110-141, with roll / window_partition / window_reverse,
// :32-59, :239-271, and the shifted-window mask, :211-232, evaluated by index arithmetic while the operands are staged).
// Same mixed-precision rule as csrc/tc_train.cu: fp32 in HBM, operands rounded to bf16 on their way into shared memory,
// fp32 accumulation in TMEM, softmax / its backward in fp32 registers.
//
// Tile = two windows = 128 tokens = the 128 TMEM lanes; two heads are in flight (threads 0-127 / 128-255, thread = token
// row).  The per-window block structure comes from the MMA's disable-output-lane mask: every product is issued twice,
// once per window, with the other window's 64 lanes masked, so S, dP are [128][64] (columns = keys of the row's own
// window) and dK / dV rows only see queries of their own window.  One bf16 image [d/8][token][8] per operand serves both as
// a K-major operand (contraction over head_dim: S = q k^T, dP = dO v^T) and as an MN-major operand (contraction over
// tokens: O = P v, dQ = dS k, dK = dS^T q, dV = P^T dO); likewise the P / dS images [key/8][token][8] are read K-major
// for O / dQ and MN-major (transposed) for dV / dK.
// Forward stores the row log-sum-exp (lse) so that backward recomputes P = exp(S - lse) without a second reduction.
#include "common.cuh"
#include "umma.cuh"

namespace rdst {
using namespace umma;

namespace {

constexpr int NT = 256;
constexpr int HEADS = 6;

struct AttnArgs {
  const float* qkv; int64_t ldq;
  const float* table;                 // [225][6]
  float* out; int64_t ldo;            // forward output [T][ldo]
  float* lse;                         // [T][6]
  const float* dout; int64_t ldd;     // backward input [T][ldd]
  float* dqkv; int64_t ldg;           // backward output, same layout as qkv
  float* dtable;                      // [225][6], accumulated
  int B, H, W, C, shift, nwin;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// token index and mask region of tile row r (window = 2*tile + r/64); -1 if the window does not exist
__device__ __forceinline__ int64_t row_token(const AttnArgs& a, int tile, int r, int& iy, int& ix, int& reg) {
  const int wi = tile * 2 + (r >> 6), i = r & 63;
  iy = i >> 3;
  ix = i & 7;
  reg = 0;
  if (wi >= a.nwin) return -1;
  const int nwx = a.W >> 3, nw_img = (a.H >> 3) * nwx;
  const int b = wi / nw_img, wl = wi - b * nw_img;
  const int wy = wl / nwx, wx = wl - wy * nwx;
  const int hs = wy * 8 + iy, ws = wx * 8 + ix;             // coordinates on the shifted frame
  int hh = hs + a.shift; if (hh >= a.H) hh -= a.H;          // shifted[h'] = x[(h'+s) mod H]
  int ww = ws + a.shift; if (ww >= a.W) ww -= a.W;
  if (a.shift > 0) {
    const int rh = hs < a.H - 8 ? 0 : (hs < a.H - a.shift ? 1 : 2);
    const int rw = ws < a.W - 8 ? 0 : (ws < a.W - a.shift ? 1 : 2);
    reg = rh * 3 + rw;
  }
  return ((int64_t)b * a.H + hh) * a.W + ww;
}

// HD fp32 values at p (or zeros) -> bf16 image units [d/8][row][8] (pads zero)
template <int HD, int HDP>
__device__ __forceinline__ void stage_row(uint8_t* img, int row, const float* p) {
  float v[HDP];
#pragma unroll
  for (int d = 0; d < HDP; ++d) v[d] = (p != nullptr && d < HD) ? __ldg(p + d) : 0.f;
#pragma unroll
  for (int d8 = 0; d8 < HDP / 8; ++d8) {
    uint4 u;
    u.x = pack2(v[d8 * 8 + 0], v[d8 * 8 + 1]);
    u.y = pack2(v[d8 * 8 + 2], v[d8 * 8 + 3]);
    u.z = pack2(v[d8 * 8 + 4], v[d8 * 8 + 5]);
    u.w = pack2(v[d8 * 8 + 6], v[d8 * 8 + 7]);
    *reinterpret_cast<uint4*>(img + d8 * 2048 + row * 16) = u;
  }
}

#define WMASK(w) (w) ? 0xFFFFFFFFu : 0u, (w) ? 0xFFFFFFFFu : 0u, (w) ? 0u : 0xFFFFFFFFu, (w) ? 0u : 0xFFFFFFFFu

// scores of the row's own window: 64 fp32 values from TMEM columns [col, col+64), + relative-position bias + shift mask
__device__ __forceinline__ void load_scores(uint32_t taddr, float (&s)[64], const float* stab, const int* sreg_w, int iy, int ix,
                                            int reg, int shift) {
  uint32_t v[32];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    tmem_ld_x32(taddr + half * 32, v);
    wait_ld();
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const int j = half * 32 + jj;
      float x = __uint_as_float(v[jj]) + stab[(iy - (j >> 3) + 7) * 15 + (ix - (j & 7) + 7)];
      if (shift > 0 && sreg_w[j] != reg) x += -100.0f;
      s[j] = x;
    }
  }
}

// 64 fp32 row values -> bf16 units of a [key/8][token][8] image
__device__ __forceinline__ void store_row64(uint8_t* img, int row, const float (&p)[64]) {
#pragma unroll
  for (int j8 = 0; j8 < 8; ++j8) {
    uint4 u;
    u.x = pack2(p[j8 * 8 + 0], p[j8 * 8 + 1]);
    u.y = pack2(p[j8 * 8 + 2], p[j8 * 8 + 3]);
    u.z = pack2(p[j8 * 8 + 4], p[j8 * 8 + 5]);
    u.w = pack2(p[j8 * 8 + 6], p[j8 * 8 + 7]);
    *reinterpret_cast<uint4*>(img + j8 * 2048 + row * 16) = u;
  }
}

template <int HDP>
__device__ __forceinline__ void load_acc(uint32_t taddr, float (&o)[HDP]) {
  if constexpr (HDP == 16) {
    uint32_t v[16];
    tmem_ld_x16(taddr, v);
    wait_ld();
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = __uint_as_float(v[d]);
  } else {
    uint32_t v[32];
    tmem_ld_x32(taddr, v);
    wait_ld();
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = __uint_as_float(v[d]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(NT, 1) attn_train_fwd_kernel(const AttnArgs a) {
  constexpr int HDP = HD <= 16 ? 16 : 32;
  constexpr int IMG = 128 * HDP * 2;                 // one operand image
  constexpr int SLOT = 3 * IMG + 16384;              // Q, K, V, P of one in-flight head
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float stab[HEADS][225];
  __shared__ int sreg[128];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, g = tid >> 7;
  uint8_t* sQ = smem + g * SLOT;
  uint8_t* sK = sQ + IMG;
  uint8_t* sV = sK + IMG;
  uint8_t* sP = sV + IMG;

  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int e = tid; e < HEADS * 225; e += NT) stab[e % HEADS][e / HEADS] = __ldg(a.table + e);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t phase = 0;
  const int ntiles = (a.nwin + 1) >> 1;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int iy, ix, reg;
    const int64_t t = row_token(a, tile, row, iy, ix, reg);
    if (g == 0) sreg[row] = reg;
    const float* qrow = t >= 0 ? a.qkv + t * a.ldq : nullptr;
    for (int rnd = 0; rnd < HEADS / 2; ++rnd) {
      const int h = rnd * 2 + g;
      stage_row<HD, HDP>(sQ, row, qrow ? qrow + h * HD : nullptr);
      stage_row<HD, HDP>(sK, row, qrow ? qrow + a.C + h * HD : nullptr);
      stage_row<HD, HDP>(sV, row, qrow ? qrow + 2 * a.C + h * HD : nullptr);
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (warp == 0) {
        if (elect_one()) {
          constexpr uint32_t ids = make_idesc_bf16(128, 64, false, false);
          for (int gg = 0; gg < 2; ++gg) {
            const uint32_t q = smem_u32(smem + gg * SLOT), k = q + IMG;
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
              for (int ks = 0; ks < HDP / 16; ++ks)
                mma_bf16_ss_masked(tmem + gg * 64, make_smem_desc(q + ks * 4096, 2048, 128),
                                   make_smem_desc(k + w * 1024 + ks * 4096, 2048, 128), ids, ks > 0, WMASK(w));
          }
          commit(&bar);
        }
        __syncwarp();
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      // ---- softmax of the row (thread = token row of head h) ----
      {
        float s[64];
        load_scores(tmem + lane_base + g * 64, s, stab[h], sreg + (row & 64), iy, ix, reg, a.shift);
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) mx = fmaxf(mx, s[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) { s[j] = __expf(s[j] - mx); sum += s[j]; }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int j = 0; j < 64; ++j) s[j] *= inv;
        store_row64(sP, row, s);
        if (t >= 0 && a.lse) a.lse[t * HEADS + h] = mx + __logf(sum);
      }
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (warp == 0) {
        if (elect_one()) {
          constexpr uint32_t ido = make_idesc_bf16(128, HDP, false, true);
          for (int gg = 0; gg < 2; ++gg) {
            const uint32_t v = smem_u32(smem + gg * SLOT) + 2 * IMG, p = v + IMG;
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)          // 64 keys of the window, 16 per step
                mma_bf16_ss_masked(tmem + 128 + gg * 32, make_smem_desc(p + ks * 4096, 2048, 128),
                                   make_smem_desc(v + w * 1024 + ks * 256, 128, 2048), ido, ks > 0, WMASK(w));
          }
          commit(&bar);
        }
        __syncwarp();
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      {
        float o[HDP];
        load_acc<HDP>(tmem + lane_base + 128 + g * 32, o);
        if (t >= 0) {
          float* orow = a.out + t * a.ldo + h * HD;
#pragma unroll
          for (int d = 0; d < HD; ++d) orow[d] = o[d];
        }
      }
      fence_before_sync();
      __syncthreads();          // images and accumulators are reused by the next pair of heads
      fence_after_sync();
    }
  }
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(NT, 1) attn_train_bwd_kernel(const AttnArgs a) {
  constexpr int HDP = HD <= 16 ? 16 : 32;
  constexpr int IMG = 128 * HDP * 2;
  constexpr int PRE = 16384;                         // readable bytes in front of / behind the transposed P, dS operands
  constexpr int SLOT = 4 * IMG + 2 * 16384;          // Q, K, V, dO, P, dS of one in-flight head
  constexpr int TS = 128 + 3 * HDP;                  // TMEM columns per head: S | dP | dQ | dK | dV
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float stab[HEADS][225];
  __shared__ float sdtab[HEADS][225];
  __shared__ float swtab[NT / 32][232];              // per-warp private copy of the table gradient of the head in flight
  __shared__ int sreg[128];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, g = tid >> 7;
  uint8_t* base = smem + PRE;                        // (the masked-off half of a transposed operand reads +-16 KB around it)
  uint8_t* sQ = base + g * SLOT;
  uint8_t* sK = sQ + IMG;
  uint8_t* sV = sK + IMG;
  uint8_t* sG = sV + IMG;                            // dO
  uint8_t* sP = sG + IMG;
  uint8_t* sD = sP + 16384;                          // dS

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int e = tid; e < HEADS * 225; e += NT) {
    stab[e % HEADS][e / HEADS] = __ldg(a.table + e);
    sdtab[e % HEADS][e / HEADS] = 0.f;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t phase = 0;
  const int ntiles = (a.nwin + 1) >> 1;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int iy, ix, reg;
    const int64_t t = row_token(a, tile, row, iy, ix, reg);
    if (g == 0) sreg[row] = reg;
    const float* qrow = t >= 0 ? a.qkv + t * a.ldq : nullptr;
    const float* grow = t >= 0 ? a.dout + t * a.ldd : nullptr;
    for (int rnd = 0; rnd < HEADS / 2; ++rnd) {
      const int h = rnd * 2 + g;
      stage_row<HD, HDP>(sQ, row, qrow ? qrow + h * HD : nullptr);
      stage_row<HD, HDP>(sK, row, qrow ? qrow + a.C + h * HD : nullptr);
      stage_row<HD, HDP>(sV, row, qrow ? qrow + 2 * a.C + h * HD : nullptr);
      stage_row<HD, HDP>(sG, row, grow ? grow + h * HD : nullptr);
      const float lse = t >= 0 ? __ldg(a.lse + t * HEADS + h) : 0.f;
      for (int e = tid & 31; e < 225; e += 32) swtab[warp][e] = 0.f;
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      // ---- S = q k^T and dP = dO v^T (contraction over head_dim) ----
      if (warp == 0) {
        if (elect_one()) {
          constexpr uint32_t ids = make_idesc_bf16(128, 64, false, false);
          for (int gg = 0; gg < 2; ++gg) {
            const uint32_t q = smem_u32(base + gg * SLOT), k = q + IMG, v = k + IMG, go = v + IMG;
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
              for (int ks = 0; ks < HDP / 16; ++ks) {
                mma_bf16_ss_masked(tmem + gg * TS, make_smem_desc(q + ks * 4096, 2048, 128),
                                   make_smem_desc(k + w * 1024 + ks * 4096, 2048, 128), ids, ks > 0, WMASK(w));
                mma_bf16_ss_masked(tmem + gg * TS + 64, make_smem_desc(go + ks * 4096, 2048, 128),
                                   make_smem_desc(v + w * 1024 + ks * 4096, 2048, 128), ids, ks > 0, WMASK(w));
              }
          }
          commit(&bar);
        }
        __syncwarp();
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      // ---- P = exp(S - lse), dS = P * (dP - sum_j P dP) (thread = token row of head h) ----
      {
        float p[64];
        load_scores(tmem + lane_base + g * TS, p, stab[h], sreg + (row & 64), iy, ix, reg, a.shift);
#pragma unroll
        for (int j = 0; j < 64; ++j) p[j] = __expf(p[j] - lse);
        store_row64(sP, row, p);
        float delta = 0.f;
        uint32_t v[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld_x32(tmem + lane_base + g * TS + 64 + half * 32, v);
          wait_ld();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) delta = fmaf(p[half * 32 + jj], __uint_as_float(v[jj]), delta);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld_x32(tmem + lane_base + g * TS + 64 + half * 32, v);
          wait_ld();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int j = half * 32 + jj;
            const float ds = p[j] * (__uint_as_float(v[jj]) - delta);
            p[j] = ds;
            // relative-position-table gradient: for one key j the 32 rows of a warp hit 32 different bins, so a plain
            // read-modify-write on the warp's private table needs no atomics (shared fp32 atomics are CAS loops)
            float* bin = &swtab[warp][(iy - (j >> 3) + 7) * 15 + (ix - (j & 7) + 7)];
            if (t >= 0) *bin += ds;
            __syncwarp();
          }
        }
        store_row64(sD, row, p);
      }
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      for (int e = row; e < 225; e += 128)
        sdtab[h][e] += (swtab[4 * g][e] + swtab[4 * g + 1][e]) + (swtab[4 * g + 2][e] + swtab[4 * g + 3][e]);
      // ---- dQ = dS k, dK = dS^T q, dV = P^T dO (contraction over the 64 tokens of the window) ----
      if (warp == 0) {
        if (elect_one()) {
          constexpr uint32_t idn = make_idesc_bf16(128, HDP, false, true);     // A K-major (rows = queries)
          constexpr uint32_t idt = make_idesc_bf16(128, HDP, true, true);      // A MN-major (rows = keys: transposed image)
          for (int gg = 0; gg < 2; ++gg) {
            const uint32_t q = smem_u32(base + gg * SLOT), k = q + IMG, go = k + 2 * IMG, pp = go + IMG, dd = pp + 16384;
            const uint32_t acc = tmem + gg * TS + 128;
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                // dQ rows = queries, K = keys of window w
                mma_bf16_ss_masked(acc, make_smem_desc(dd + ks * 4096, 2048, 128),
                                   make_smem_desc(k + w * 1024 + ks * 256, 128, 2048), idn, ks > 0, WMASK(w));
                // transposed images: element (k = query of window w, m = key row) at (k/8)*128 + (m/8)*2048 + ..., start
                // moved so that key row 64w + j reads image column j of query row 64w + k
                const uint32_t toff = (uint32_t)(w * 1024 + ks * 256) - (uint32_t)(w * 16384);
                mma_bf16_ss_masked(acc + HDP, make_smem_desc(dd + toff, 128, 2048),
                                   make_smem_desc(q + w * 1024 + ks * 256, 128, 2048), idt, ks > 0, WMASK(w));
                mma_bf16_ss_masked(acc + 2 * HDP, make_smem_desc(pp + toff, 128, 2048),
                                   make_smem_desc(go + w * 1024 + ks * 256, 128, 2048), idt, ks > 0, WMASK(w));
              }
          }
          commit(&bar);
        }
        __syncwarp();
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
      {
        float o[HDP];
        float* drow = t >= 0 ? a.dqkv + t * a.ldg + h * HD : nullptr;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          load_acc<HDP>(tmem + lane_base + g * TS + 128 + part * HDP, o);
          if (drow) {
#pragma unroll
            for (int d = 0; d < HD; ++d) drow[part * a.C + d] = o[d];
          }
        }
      }
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
    }
  }
  for (int e = tid; e < HEADS * 225; e += NT) {
    const float v = sdtab[e % HEADS][e / HEADS];
    if (v != 0.f) atomicAdd(a.dtable + e, v);
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int sm_count_attn() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int HD>
int launch_fwd(const AttnArgs& a, cudaStream_t st) {
  constexpr int HDP = HD <= 16 ? 16 : 32;
  const size_t smem = 2 * (3 * 128 * HDP * 2 + 16384);
  cudaError_t e = cudaFuncSetAttribute(attn_train_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("rdst_window_attention_tc_fwd: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  const int tiles = (a.nwin + 1) / 2;
  attn_train_fwd_kernel<HD><<<tiles < sm_count_attn() ? tiles : sm_count_attn(), NT, smem, st>>>(a);
  return RDST_OK;
}

template <int HD>
int launch_bwd(const AttnArgs& a, cudaStream_t st) {
  constexpr int HDP = HD <= 16 ? 16 : 32;
  const size_t smem = 16384 + 2 * (4 * 128 * HDP * 2 + 2 * 16384) + 16384;
  cudaError_t e = cudaFuncSetAttribute(attn_train_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("rdst_window_attention_tc_bwd: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  const int tiles = (a.nwin + 1) / 2;
  attn_train_bwd_kernel<HD><<<tiles < sm_count_attn() ? tiles : sm_count_attn(), NT, smem, st>>>(a);
  return RDST_OK;
}

int check_geom(const char* who, int B, int H, int W, int C, int shift) {
  if (!(B >= 0 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0)) { set_error("%s: H=%d W=%d must be positive multiples of 8", who, H, W); return RDST_E_INVALID; }
  if (!(C == 60 || C == 90 || C == 120)) { set_error("%s: C=%d unsupported (60, 90, 120 with 6 heads)", who, C); return RDST_E_UNSUPPORTED; }
  if (!(shift == 0 || shift == 4)) { set_error("%s: shift must be 0 or 4", who); return RDST_E_INVALID; }
  return RDST_OK;
}

}  // namespace
}  // namespace rdst

extern "C" int rdst_window_attention_tc_fwd(const float* qkv, int64_t ldq, const float* table, float* out, int64_t ldo,
                                            float* lse, int B, int H, int W, int C, int shift, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(qkv && table && out, "rdst_window_attention_tc_fwd: null pointer");
  if (int rc = check_geom("rdst_window_attention_tc_fwd", B, H, W, C, shift)) return rc;
  RDST_REQUIRE(ldq >= 3 * C && ldo >= C, "rdst_window_attention_tc_fwd: bad leading dimension");
  if (B == 0) return RDST_OK;
  AttnArgs a{};
  a.qkv = qkv; a.ldq = ldq; a.table = table; a.out = out; a.ldo = ldo; a.lse = lse;
  a.B = B; a.H = H; a.W = W; a.C = C; a.shift = shift; a.nwin = B * (H / 8) * (W / 8);
  int rc = C == 60 ? launch_fwd<10>(a, (cudaStream_t)stream) : C == 90 ? launch_fwd<15>(a, (cudaStream_t)stream)
                                                                       : launch_fwd<20>(a, (cudaStream_t)stream);
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_window_attention_tc_fwd");
  return RDST_OK;
}

extern "C" int rdst_window_attention_tc_bwd(const float* qkv, int64_t ldq, const float* table, const float* lse,
                                            const float* dout, int64_t ldo, float* dqkv, int64_t ldg, float* dtable, int B,
                                            int H, int W, int C, int shift, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(qkv && table && lse && dout && dqkv && dtable, "rdst_window_attention_tc_bwd: null pointer");
  if (int rc = check_geom("rdst_window_attention_tc_bwd", B, H, W, C, shift)) return rc;
  RDST_REQUIRE(ldq >= 3 * C && ldo >= C && ldg >= 3 * C, "rdst_window_attention_tc_bwd: bad leading dimension");
  if (B == 0) return RDST_OK;
  AttnArgs a{};
  a.qkv = qkv; a.ldq = ldq; a.table = table; a.lse = const_cast<float*>(lse); a.dout = dout; a.ldd = ldo;
  a.dqkv = dqkv; a.ldg = ldg; a.dtable = dtable;
  a.B = B; a.H = H; a.W = W; a.C = C; a.shift = shift; a.nwin = B * (H / 8) * (W / 8);
  int rc = C == 60 ? launch_bwd<10>(a, (cudaStream_t)stream) : C == 90 ? launch_bwd<15>(a, (cudaStream_t)stream)
                                                                       : launch_bwd<20>(a, (cudaStream_t)stream);
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_window_attention_tc_bwd");
  return RDST_OK;
}
