// Tensor-core GEMMs of the TRAINING path (precision 'bf16'): fp32 activations / weights / gradients in HBM, operands
// rounded to bf16 while they are staged into shared memory as UMMA operand images, fp32 accumulation in TMEM
// (tcgen05.mma kind::f16), fp32 epilogue.  This is the mixed-precision rule of bf16 autocast training: every GEMM
// runs on the tensor cores in bf16, everything that is stored or reduced (saved activations, LayerNorm statistics,
// weight gradients, Adam state) stays fp32.
//
//   rdst_gemm_tc     Y[T][N] = scale * (LNhat(X)[T][K] . Wop + bias) + R      -- forward Linear, Linear data gradient
//                    (weight given transposed: MN-major B operand) and 3x3 convolution forward / data gradient
//                    (implicit GEMM: the A image of a K-chunk is gathered from the 3x3 neighbourhood while staging;
//                    PixelShuffle(2) folded into the store).
//   rdst_gemm_tn_tc  dW[N][K] += dY^T . X, db[n] += sum_t dY[t][n]             -- weight / bias gradients.  Tokens are the
//                    reduction dimension of the MMA: both operands are MN-major images ([token][channel] rows as they lie
//                    in HBM), the bias gradient rides along as a column of ones appended to X, token ranges are split
//                    over CTAs and combined with vectorised fp32 reductions (red.global.add.v4.f32).
//
// Operand images are SWIZZLE_NONE core-matrix layouts (umma.cuh); each was validated in isolation by
// rdst_umma_selftest and tests/test_gpu_train_tc.py compares both entry points with torch matmuls on bf16-rounded
// operands.  Replaces, for the training config, the nn.Linear / nn.Conv2d forward+backward GEMMs of
// swin_transformer_sr.py:117,139,24-27 and rdst_variations.py:339,444, common.py:6-9,129-132.
#include "common.cuh"
#include "umma.cuh"

namespace rdst {
using namespace umma;

namespace {

constexpr int TM = 128;          // token rows per tile (= TMEM lanes)
constexpr int NTHREADS = 256;

struct GemmArgs {
  const float* x; int64_t ldx;
  const float* w; int64_t ldw; int w_mn;
  const float* bias; const float* resid; int64_t ldr;
  float* y; int64_t ldy;
  int64_t T; int K, N, Np;       // Np = N rounded up to 16
  int KC, nkc;                   // K-chunk (multiple of 16, <= 256 or the whole padded K) and number of chunks
  int ln_creal; float out_scale;
  int conv, B, H, W, Cin, shuffle;
};

__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
  r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
  return r;
}

// 8 consecutive floats starting at p, of which the first `valid` (0..8) are real; the rest read as 0.  Activation rows are
// always 16-byte aligned (checked on the host); weight rows of odd width (C = 90) take the scalar path.
__device__ __forceinline__ void load8(const float* p, int valid, float4& a, float4& b) {
  a = make_float4(0.f, 0.f, 0.f, 0.f);
  b = a;
  if (valid >= 8 && ((uintptr_t)p & 15) == 0) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  } else if (valid > 0) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = j < valid ? __ldg(p + j) : 0.f;
    a = make_float4(t[0], t[1], t[2], t[3]);
    b = make_float4(t[4], t[5], t[6], t[7]);
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// Y = scale * (LNhat(X) . Wop + bias) + R
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_mean[TM], s_rstd[TM];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = a.KC, Np = a.Np;
  uint8_t* sA = smem;                               // K-major [KC/8][128][8] bf16
  uint8_t* sB = smem + (size_t)TM * KC * 2;         // K-major [KC/8][Np][8]  or  MN-major [Np/8][KC/8][8(k)][8(n)]

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t phase = 0;

  const int r8 = lane & 7, c4 = lane >> 3;          // staging block: 8 rows x 4 chunks of 8 columns per warp pass
  const int nch = KC >> 3;                          // 8-column chunks per K-chunk
  const int chg = (nch + 3) >> 2;                   // chunk groups of 4
  const int64_t ntiles = (a.T + TM - 1) / TM;
  const int HW = a.H * a.W;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t0 = tile * TM;
    // ---- LayerNorm statistics of the tile's rows (one warp per row, coalesced) ----
    if (a.ln_creal > 0) {
      for (int r = warp; r < TM; r += NTHREADS / 32) {
        const int64_t t = t0 + r;
        float s = 0.f, ss = 0.f;
        if (t < a.T) {
          const float* row = a.x + t * a.ldx;
          for (int k = lane * 4; k < a.K; k += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + k));
            s += v.x + v.y + v.z + v.w;
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) {
          const float inv = 1.f / (float)a.ln_creal;
          const float mean = s * inv;
          const float var = fmaxf(ss * inv - mean * mean, 0.f);
          s_mean[r] = mean;
          s_rstd[r] = rsqrtf(var + 1e-5f);
        }
      }
      __syncthreads();
    }
    for (int kc = 0; kc < a.nkc; ++kc) {
      const int k0 = kc * KC;
      if (kc > 0) {                                  // the previous chunk's MMAs still read sA / sB
        mbar_wait(&bar, phase);
        phase ^= 1;
        fence_after_sync();
      }
      // ---- stage A: rows = tokens, K-major image ----
      for (int blk = warp; blk < 16 * chg; blk += NTHREADS / 32) {
        const int r = (blk % 16) * 8 + r8;
        const int ch = (blk / 16) * 4 + c4;
        if (ch >= nch) continue;
        const int k = k0 + ch * 8;                   // first of 8 columns
        const int64_t t = t0 + r;
        float4 v0, v1;
        if (t >= a.T || k >= a.K) {
          v0 = make_float4(0.f, 0.f, 0.f, 0.f);
          v1 = v0;
        } else if (!a.conv) {
          load8(a.x + t * a.ldx + k, a.K - k, v0, v1);
          if (a.ln_creal > 0) {
            const float m = s_mean[r], rs = s_rstd[r];
            v0 = make_float4((v0.x - m) * rs, (v0.y - m) * rs, (v0.z - m) * rs, (v0.w - m) * rs);
            v1 = make_float4((v1.x - m) * rs, (v1.y - m) * rs, (v1.z - m) * rs, (v1.w - m) * rs);
          }
        } else {
          const int tap = k / a.Cin, ci = k - tap * a.Cin;
          const int b = (int)(t / HW), rem = (int)(t - (int64_t)b * HW);
          const int py = rem / a.W + tap / 3 - 1, px = rem % a.W + tap % 3 - 1;
          if (py < 0 || py >= a.H || px < 0 || px >= a.W) {
            v0 = make_float4(0.f, 0.f, 0.f, 0.f);
            v1 = v0;
          } else {
            load8(a.x + ((int64_t)b * HW + (int64_t)py * a.W + px) * a.ldx + ci, 8, v0, v1);
          }
        }
        *reinterpret_cast<uint4*>(sA + (size_t)ch * (TM * 16) + r * 16) = pack8(v0, v1);
      }
      // ---- stage B (weights): once per CTA when there is a single K-chunk ----
      if (a.nkc > 1 || tile == blockIdx.x) {
        if (!a.w_mn) {
          // W [N][ldw] (K contiguous) -> K-major image: (k/8)*(Np*16) + n*16
          const int ng = Np >> 3;
          for (int blk = warp; blk < ng * chg; blk += NTHREADS / 32) {
            const int n = (blk % ng) * 8 + r8;
            const int ch = (blk / ng) * 4 + c4;
            if (ch >= nch) continue;
            const int k = k0 + ch * 8;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (n < a.N && k < a.K) load8(a.w + (int64_t)n * a.ldw + k, a.K - k, v0, v1);
            *reinterpret_cast<uint4*>(sB + (size_t)ch * (Np * 16) + n * 16) = pack8(v0, v1);
          }
        } else {
          // W [K][ldw] (N contiguous) -> MN-major image: (k/8)*128 + (n/8)*(nch*128) + (k%8)*16
          const int ncg = ((Np >> 3) + 3) >> 2;
          for (int blk = warp; blk < nch * ncg; blk += NTHREADS / 32) {
            const int kk = (blk % nch) * 8 + r8;         // k within the chunk
            const int n8 = (blk / nch) * 4 + c4;
            if (n8 >= (Np >> 3)) continue;
            const int k = k0 + kk, n = n8 * 8;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (k < a.K && n < a.N) load8(a.w + (int64_t)k * a.ldw + n, a.N - n, v0, v1);
            *reinterpret_cast<uint4*>(sB + (size_t)(kk >> 3) * 128 + (size_t)n8 * (nch * 128) + (kk & 7) * 16) = pack8(v0, v1);
          }
        }
      }
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      // ---- MMAs: N split into pieces of <= 256 columns ----
      if (warp == 0) {
        if (elect_one()) {
          for (int n0 = 0; n0 < Np; n0 += 256) {
            const int ncols = min(256, Np - n0);
            const uint32_t idesc = make_idesc_bf16(TM, ncols, false, a.w_mn != 0);
            for (int ks = 0; ks < KC / 16; ++ks) {
              const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2 * (TM * 16), TM * 16, 128);
              uint64_t db;
              if (!a.w_mn) db = make_smem_desc(smem_u32(sB) + n0 * 16 + ks * 2 * (Np * 16), Np * 16, 128);
              else db = make_smem_desc(smem_u32(sB) + (n0 >> 3) * (nch * 128) + ks * 2 * 128, 128, nch * 128);
              mma_bf16_ss(tmem + n0, da, db, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
            }
          }
          commit(&bar);
        }
        __syncwarp();
      }
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---- epilogue: thread = token row (TMEM lane), the two thread halves alternate over 16-column groups ----
    {
      const int r = tid & 127, half = tid >> 7;
      const int64_t t = t0 + r;
      const bool live = t < a.T;
      int64_t yrow = t;
      int py = 0, px = 0, bimg = 0;
      if (a.shuffle && live) {
        bimg = (int)(t / HW);
        const int rem = (int)(t - (int64_t)bimg * HW);
        py = rem / a.W;
        px = rem % a.W;
      }
      for (int g = half; g * 16 < a.N; g += 2) {
        const int n0 = g * 16;
        uint32_t v[16];
        __syncwarp();
        tmem_ld_x16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + n0, v);
        wait_ld();
        if (!live) continue;
        int ncol = n0;
        if (a.shuffle) {                                   // PixelShuffle(2): out channel n = s*64 + c, s = 2*dy+dx
          const int s = n0 >> 6;
          yrow = ((int64_t)bimg * (2 * a.H) + (2 * py + (s >> 1))) * (2 * a.W) + (2 * px + (s & 1));
          ncol = n0 & 63;
        }
        float* yp = a.y + yrow * a.ldy + ncol;
        const float* rp = a.resid ? a.resid + t * a.ldr + n0 : nullptr;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float acc = __uint_as_float(v[j]);
          if (a.bias && n0 + j < a.N) acc += __ldg(a.bias + n0 + j);
          o[j] = acc * a.out_scale;
        }
        if (n0 + 16 <= a.N) {
          if (rp) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 rv = *reinterpret_cast<const float4*>(rp + j);
              o[j] += rv.x; o[j + 1] += rv.y; o[j + 2] += rv.z; o[j + 3] += rv.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + j < a.N) yp[j] = o[j] + (rp ? rp[j] : 0.f);
        }
      }
    }
    fence_before_sync();
    __syncthreads();          // all TMEM reads of this tile done before the next tile's MMAs overwrite the accumulator
    fence_after_sync();
  }
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// dW[N][K] += dY^T X   (+ db)
// ------------------------------------------------------------------------------------------------------------------
struct TnArgs {
  const float* dy; int64_t ldy; const float* x; int64_t ldx; float* dw; float* db;
  int64_t T; int N, K;
  int KCW;                      // columns of X (k) per CTA, multiple of 16, <= 240
  int64_t t_per_split;          // multiple of 128
  int conv, B, H, W, Cin;
};

__global__ void __launch_bounds__(NTHREADS, 1) gemm_tn_tc_kernel(const TnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TM;                   // first output row (n) of this CTA
  const int k0 = blockIdx.y * a.KCW;                // first output column (k)
  const bool with_bias = a.db != nullptr && blockIdx.y == 0;
  const int NB = a.KCW + (with_bias ? 16 : 0);      // MMA N: k columns (+ the ones column group)
  // MN-major images, token = MMA K dimension: element (t, c) at (t/8)*128 + (c/8)*2048 + (t%8)*16 + (c%8)*2
  uint8_t* sA = smem;                               // dY: 128 tokens x 128 n
  uint8_t* sB = smem + (size_t)TM * 128 * 2;        // X : 128 tokens x NB k

  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  uint32_t phase = 0;

  const int r8 = lane & 7, c4 = lane >> 3;
  const int64_t ts = (int64_t)blockIdx.z * a.t_per_split;
  const int64_t te = min(a.T, ts + a.t_per_split);
  const int HW = a.H * a.W;
  const int nbc = NB >> 3;                          // 8-column chunks of the B image
  bool first = true;
  for (int64_t t0 = ts; t0 < te; t0 += TM) {
    if (!first) {
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
    }
    // ---- A image: dY[t][m0 .. m0+128) ----
    for (int blk = warp; blk < 16 * 4; blk += NTHREADS / 32) {
      const int tt = (blk % 16) * 8 + r8;
      const int ch = (blk / 16) * 4 + c4;           // 16 chunks of 8 n
      const int64_t t = t0 + tt;
      const int n = m0 + ch * 8;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (t < te && n < a.N) load8(a.dy + t * a.ldy + n, a.N - n, v0, v1);
      *reinterpret_cast<uint4*>(sA + (size_t)(tt >> 3) * 128 + (size_t)ch * 2048 + (tt & 7) * 16) = pack8(v0, v1);
    }
    // ---- B image: X[t][k0 .. k0+KCW) (3x3 gather in conv mode) + ones column ----
    for (int blk = warp; blk < 16 * ((nbc + 3) >> 2); blk += NTHREADS / 32) {
      const int tt = (blk % 16) * 8 + r8;
      const int ch = (blk / 16) * 4 + c4;
      if (ch >= nbc) continue;
      const int64_t t = t0 + tt;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (ch * 8 >= a.KCW) {
        if (ch * 8 == a.KCW && t < te) v0.x = 1.f;  // the ones column: accumulates db[n] = sum_t dY[t][n]
      } else {
        const int k = k0 + ch * 8;
        if (t < te && k < a.K) {
          if (!a.conv) {
            load8(a.x + t * a.ldx + k, a.K - k, v0, v1);
          } else {
            const int tap = k / a.Cin, ci = k - tap * a.Cin;
            const int b = (int)(t / HW), rem = (int)(t - (int64_t)b * HW);
            const int py = rem / a.W + tap / 3 - 1, px = rem % a.W + tap % 3 - 1;
            if (py >= 0 && py < a.H && px >= 0 && px < a.W)
              load8(a.x + ((int64_t)b * HW + (int64_t)py * a.W + px) * a.ldx + ci, 8, v0, v1);
          }
        }
      }
      *reinterpret_cast<uint4*>(sB + (size_t)(tt >> 3) * 128 + (size_t)ch * 2048 + (tt & 7) * 16) = pack8(v0, v1);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(TM, NB, true, true);
        for (int ks = 0; ks < TM / 16; ++ks) {
          const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2 * 128, 128, 2048);
          const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2 * 128, 128, 2048);
          mma_bf16_ss(tmem, da, db, idesc, (!first || ks > 0) ? 1u : 0u);
        }
        commit(&bar);
      }
      __syncwarp();
    }
    first = false;
  }
  if (first) {                                       // empty token range: nothing to add
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
    return;
  }
  mbar_wait(&bar, phase);
  fence_after_sync();
  // ---- epilogue: thread = output row n, fp32 reductions into dW / db ----
  {
    const int r = tid & 127, half = tid >> 7;
    const int n = m0 + r;
    const bool vec = (a.K & 3) == 0;
    for (int g = half; g * 16 < NB; g += 2) {
      const int c0 = g * 16;
      uint32_t v[16];
      __syncwarp();
      tmem_ld_x16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
      wait_ld();
      if (n >= a.N) continue;
      if (c0 >= a.KCW) {                             // ones column group
        atomicAdd(a.db + n, __uint_as_float(v[0]));
        continue;
      }
      float* p = a.dw + (int64_t)n * a.K + k0 + c0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int k = k0 + c0 + j;
        if (vec && k + 4 <= a.K) {
          red_add_v4(p + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                     __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k + q < a.K) atomicAdd(p + j + q, __uint_as_float(v[j + q]));
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace
}  // namespace rdst

extern "C" int rdst_gemm_tc(const float* x, int64_t ldx, const float* w, int64_t ldw, int w_mn_major, const float* bias,
                            const float* resid, int64_t ldr, float* y, int64_t ldy, int64_t T, int K, int N, int ln_creal,
                            float out_scale, int conv, int B, int H, int W, int Cin, int shuffle, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && w && y, "rdst_gemm_tc: null pointer");
  RDST_REQUIRE(T >= 0 && K > 0 && N > 0 && N <= 512, "rdst_gemm_tc: bad shape (T=%lld K=%d N=%d; N <= 512)", (long long)T, K, N);
  RDST_REQUIRE((ldx & 3) == 0 && (ldy & 3) == 0 && (!resid || (ldr & 3) == 0) &&
                   ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                   (!resid || ((uintptr_t)resid & 15) == 0),
               "rdst_gemm_tc: activation rows must be 16-byte aligned (pointers and leading dimensions multiples of 4 floats)");
  RDST_REQUIRE(!conv || (Cin > 0 && Cin % 16 == 0 && K == 9 * Cin && T == (int64_t)B * H * W && !w_mn_major && !ln_creal),
               "rdst_gemm_tc: conv mode needs Cin %% 16 == 0, K == 9*Cin, T == B*H*W, K-major weights, no LayerNorm");
  RDST_REQUIRE(!shuffle || (conv && shuffle == 2 && N == 256 && !resid), "rdst_gemm_tc: shuffle=2 needs conv mode, N == 256, no residual");
  RDST_REQUIRE(!ln_creal || (ln_creal <= K && K <= 512 && !conv), "rdst_gemm_tc: bad LayerNorm width");
  if (T == 0) return RDST_OK;
  GemmArgs a{};
  a.x = x; a.ldx = ldx; a.w = w; a.ldw = ldw; a.w_mn = w_mn_major; a.bias = bias; a.resid = resid; a.ldr = ldr;
  a.y = y; a.ldy = ldy; a.T = T; a.K = K; a.N = N; a.Np = (N + 15) / 16 * 16;
  a.ln_creal = ln_creal; a.out_scale = out_scale;
  a.conv = conv; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.shuffle = shuffle;
  const int Kp = (K + 15) / 16 * 16;
  if (conv) {
    a.KC = 9 * Cin <= 256 ? 9 * Cin : (3 * Cin <= 256 ? 3 * Cin : Cin);   // whole taps per chunk: 9, 3 (a filter row) or 1
    RDST_REQUIRE(a.KC <= 256, "rdst_gemm_tc: conv mode supports Cin <= 256");
  } else {
    a.KC = Kp <= 384 ? Kp : 256;
  }
  a.nkc = (Kp + a.KC - 1) / a.KC;
  const size_t smem = (size_t)TM * a.KC * 2 + (size_t)a.Np * a.KC * 2;
  RDST_REQUIRE(smem <= 220 * 1024, "rdst_gemm_tc: operand images need %zu bytes of shared memory (K=%d N=%d)", smem, K, N);
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) { set_error("rdst_gemm_tc: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  const int64_t tiles = (T + TM - 1) / TM;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  gemm_tc_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_gemm_tc");
  return RDST_OK;
}

extern "C" int rdst_gemm_tn_tc(const float* dy, int64_t ldy, const float* x, int64_t ldx, float* dw, float* db, int64_t T,
                               int N, int K, int conv, int B, int H, int W, int Cin, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(dy && x && dw, "rdst_gemm_tn_tc: null pointer");
  RDST_REQUIRE(T >= 0 && N > 0 && K > 0, "rdst_gemm_tn_tc: bad shape");
  RDST_REQUIRE((ldx & 3) == 0 && (ldy & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 &&
                   ((uintptr_t)dw & 15) == 0,
               "rdst_gemm_tn_tc: rows must be 16-byte aligned (pointers and leading dimensions multiples of 4 floats)");
  RDST_REQUIRE(!conv || (Cin > 0 && Cin % 16 == 0 && K == 9 * Cin && T == (int64_t)B * H * W),
               "rdst_gemm_tn_tc: conv mode needs Cin %% 16 == 0, K == 9*Cin and T == B*H*W");
  if (T == 0) return RDST_OK;
  TnArgs a{};
  a.dy = dy; a.ldy = ldy; a.x = x; a.ldx = ldx; a.dw = dw; a.db = db; a.T = T; a.N = N; a.K = K;
  a.conv = conv; a.B = B; a.H = H; a.W = W; a.Cin = Cin;
  const int Kp = (K + 15) / 16 * 16;
  const int kchunks = (Kp + 239) / 240;
  a.KCW = ((Kp + kchunks - 1) / kchunks + 15) / 16 * 16;
  const int kc = (Kp + a.KCW - 1) / a.KCW;
  const int mt = (N + TM - 1) / TM;
  int64_t splits = (2 * sm_count() + mt * kc - 1) / (mt * kc);          // about two CTAs' worth of work per SM in flight
  const int64_t max_splits = (T + TM - 1) / TM;                         // at least one 128-token stage per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.t_per_split = ((T + splits - 1) / splits + TM - 1) / TM * TM;
  splits = (T + a.t_per_split - 1) / a.t_per_split;
  const size_t smem = (size_t)TM * 128 * 2 + (size_t)TM * (a.KCW + 16) * 2;
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  if (e != cudaSuccess) { set_error("rdst_gemm_tn_tc: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  dim3 grid((unsigned)mt, (unsigned)kc, (unsigned)splits);
  gemm_tn_tc_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_gemm_tn_tc");
  return RDST_OK;
}
