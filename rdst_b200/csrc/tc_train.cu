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
constexpr int NT_GEMM = 512;         // one tile per SM at the training batch: threads buy memory-level parallelism
constexpr int NT_TN = 512;

struct GemmArgs {
  const float* x; int64_t ldx;
  const float* w; int64_t ldw; int w_mn, w_vec;
  const float* bias; const float* resid; int64_t ldr;
  const float* aux; int64_t lda;   // epilogue: acc *= gelu'(aux[t][n])  (GELU backward fused into the data-gradient GEMM)
  const float* lnx; int64_t ldlx; int lnb_creal;   // epilogue: LayerNorm-hat backward of the row w.r.t. lnx (N <= 128)
  const float* resid2; int64_t ldr2;
  float* y; int64_t ldy;
  int64_t T; int K, N, Np;       // Np = N rounded up to 16
  int KC, nkc;                   // K-chunk (multiple of 16, <= 256 or the whole padded K) and number of chunks
  int a_op, ln_creal;            // A prologue: 0 none, 1 LayerNorm-hat over ln_creal real channels, 2 exact-erf GELU
  float out_scale;
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

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// Activation rows are 16-byte aligned and readable up to the next multiple of 8 columns (checked on the host), so a chunk
// is always two vector loads; columns >= valid are zeroed afterwards (pads may hold anything, 0 * NaN must not reach the MMA)
__device__ __forceinline__ void mask8(int valid, float4& a, float4& b) {
  if (valid < 8) {
    if (valid < 1) a.x = 0.f;
    if (valid < 2) a.y = 0.f;
    if (valid < 3) a.z = 0.f;
    if (valid < 4) a.w = 0.f;
    if (valid < 5) b.x = 0.f;
    if (valid < 6) b.y = 0.f;
    if (valid < 7) b.z = 0.f;
    b.w = 0.f;
  }
}

// weight rows: vector loads when every row is 16-byte aligned (vec), else scalar (rows of 90 floats, C = 90)
__device__ __forceinline__ void load_w8(const float* p, int valid, bool vec, float4& a, float4& b) {
  if (vec && valid >= 8) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  } else {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = j < valid ? __ldg(p + j) : 0.f;
    a = make_float4(t[0], t[1], t[2], t[3]);
    b = make_float4(t[4], t[5], t[6], t[7]);
  }
}

// address of the 8-column chunk starting at column k of token t's GEMM row (3x3 gather in conv mode); nullptr = zeros
__device__ __forceinline__ const float* act_chunk(const float* x, int64_t ldx, int64_t t, int64_t T, int k, int K, int conv,
                                                  int Cin, int H, int W, int HW) {
  if (t >= T || k >= K) return nullptr;
  if (!conv) return x + t * ldx + k;
  const int tap = k / Cin, ci = k - tap * Cin;
  const int b = (int)(t / HW), rem = (int)(t - (int64_t)b * HW);
  const int yy = rem / W;
  const int py = yy + tap / 3 - 1, px = rem - yy * W + tap % 3 - 1;
  if (py < 0 || py >= H || px < 0 || px >= W) return nullptr;
  return x + ((int64_t)b * HW + (int64_t)py * W + px) * ldx + ci;
}

__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * __expf(-0.5f * x * x);
}

__device__ __forceinline__ void apply_op(int op, float m, float rs, float4& a, float4& b) {
  if (op == 1) {
    a = make_float4((a.x - m) * rs, (a.y - m) * rs, (a.z - m) * rs, (a.w - m) * rs);
    b = make_float4((b.x - m) * rs, (b.y - m) * rs, (b.z - m) * rs, (b.w - m) * rs);
  } else if (op == 2) {
    a = make_float4(gelu_erf(a.x), gelu_erf(a.y), gelu_erf(a.z), gelu_erf(a.w));
    b = make_float4(gelu_erf(b.x), gelu_erf(b.y), gelu_erf(b.z), gelu_erf(b.w));
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// LayerNorm statistics of 128 consecutive rows (one warp per row, four rows in flight): mean / rstd over the K stored
// values divided by creal real channels (pads are zero), as rdst_linear_fwd.
template <int NTHREADS>
__device__ __forceinline__ void row_stats(const float* x, int64_t ldx, int64_t t0, int64_t T, int K, int creal, float* s_mean,
                                          float* s_rstd, int warp, int lane) {
  const float inv = 1.f / (float)creal;
  for (int r0 = warp * 4; r0 < TM; r0 += (NTHREADS / 32) * 4) {
    float s[4], ss[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = ss[u] = 0.f;
      const int64_t t = t0 + r0 + u;
      if (t < T) {
        const float* row = x + t * ldx;
        for (int k = lane * 4; k < K; k += 128) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + k));
          s[u] += v.x + v.y + v.z + v.w;
          ss[u] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
        ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], o);
      }
    }
    if (lane < 4) {                            // lane u publishes row r0+u (all lanes hold all four sums)
      const float sv = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      const float qv = lane == 0 ? ss[0] : lane == 1 ? ss[1] : lane == 2 ? ss[2] : ss[3];
      const float mean = sv * inv;
      const float var = fmaxf(qv * inv - mean * mean, 0.f);
      s_mean[r0 + lane] = mean;
      s_rstd[r0 + lane] = rsqrtf(var + 1e-5f);
    }
  }
}

constexpr int U = 4;                      // independent chunk loads in flight per thread while staging

// ------------------------------------------------------------------------------------------------------------------
// Y = scale * (op(X) . Wop + bias) [* gelu'(aux)] + R
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT_GEMM, 1) gemm_tc_kernel(const GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_mean[TM], s_rstd[TM];
  __shared__ float s_bias[512];
  __shared__ float2 s_red[2][NT_GEMM / 128][TM];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = a.KC, Np = a.Np;
  uint8_t* sA = smem;                               // K-major [KC/8][128][8] bf16
  uint8_t* sB = smem + (size_t)TM * KC * 2;         // K-major [KC/8][Np][8]  or  MN-major [Np/8][KC/8][8(k)][8(n)]

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int n = tid; n < Np; n += NT_GEMM) s_bias[n] = (a.bias && n < a.N) ? __ldg(a.bias + n) : 0.f;
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
  constexpr int NW = NT_GEMM / 32;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t0 = tile * TM;
    if (a.a_op == 1) {
      row_stats<NT_GEMM>(a.x, a.ldx, t0, a.T, a.K, a.ln_creal, s_mean, s_rstd, warp, lane);
      __syncthreads();
    }
    for (int kc = 0; kc < a.nkc; ++kc) {
      const int k0 = kc * KC;
      if (kc > 0) {                                  // the previous chunk's MMAs still read sA / sB
        mbar_wait(&bar, phase);
        phase ^= 1;
        fence_after_sync();
      }
      // ---- stage A: rows = tokens, K-major image; U chunk loads in flight per thread ----
      for (int b0 = warp; b0 < 16 * chg; b0 += NW * U) {
        float4 v0[U], v1[U];
        int valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int blk = b0 + u * NW;
          const int r = (blk & 15) * 8 + r8, ch = (blk >> 4) * 4 + c4;
          const int k = k0 + ch * 8;
          const float* p = (blk < 16 * chg && ch < nch)
                               ? act_chunk(a.x, a.ldx, t0 + r, a.T, k, a.K, a.conv, a.Cin, a.H, a.W, HW) : nullptr;
          valid[u] = p ? min(8, a.K - k) : 0;
          if (p) {
            v0[u] = __ldg(reinterpret_cast<const float4*>(p));
            v1[u] = __ldg(reinterpret_cast<const float4*>(p) + 1);
          } else {
            v0[u] = zero4();
            v1[u] = zero4();
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int blk = b0 + u * NW;
          const int r = (blk & 15) * 8 + r8, ch = (blk >> 4) * 4 + c4;
          if (blk < 16 * chg && ch < nch) {
            if (valid[u] > 0) {
              mask8(valid[u], v0[u], v1[u]);
              apply_op(a.a_op, s_mean[r], s_rstd[r], v0[u], v1[u]);
            }
            *reinterpret_cast<uint4*>(sA + (size_t)ch * (TM * 16) + r * 16) = pack8(v0[u], v1[u]);
          }
        }
      }
      // ---- stage B (weights): once per CTA when there is a single K-chunk ----
      if (a.nkc > 1 || tile == blockIdx.x) {
        const bool vec = a.w_vec != 0;
        if (!a.w_mn) {
          // W [N][ldw] (K contiguous) -> K-major image: (k/8)*(Np*16) + n*16
          const int ng = Np >> 3;
          for (int b0 = warp; b0 < ng * chg; b0 += NW * U) {
            float4 v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int blk = b0 + u * NW;
              const int n = (blk % ng) * 8 + r8, ch = (blk / ng) * 4 + c4;
              const int k = k0 + ch * 8;
              v0[u] = zero4();
              v1[u] = zero4();
              if (blk < ng * chg && ch < nch && n < a.N && k < a.K) load_w8(a.w + (int64_t)n * a.ldw + k, a.K - k, vec, v0[u], v1[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int blk = b0 + u * NW;
              const int n = (blk % ng) * 8 + r8, ch = (blk / ng) * 4 + c4;
              if (blk < ng * chg && ch < nch) *reinterpret_cast<uint4*>(sB + (size_t)ch * (Np * 16) + n * 16) = pack8(v0[u], v1[u]);
            }
          }
        } else {
          // W [K][ldw] (N contiguous) -> MN-major image: (k/8)*128 + (n/8)*(nch*128) + (k%8)*16
          const int ncg = ((Np >> 3) + 3) >> 2;
          for (int b0 = warp; b0 < nch * ncg; b0 += NW * U) {
            float4 v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int blk = b0 + u * NW;
              const int kk = (blk % nch) * 8 + r8, n8 = (blk / nch) * 4 + c4;
              const int k = k0 + kk, n = n8 * 8;
              v0[u] = zero4();
              v1[u] = zero4();
              if (blk < nch * ncg && n8 < (Np >> 3) && k < a.K && n < a.N)
                load_w8(a.w + (int64_t)k * a.ldw + n, a.N - n, vec, v0[u], v1[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int blk = b0 + u * NW;
              const int kk = (blk % nch) * 8 + r8, n8 = (blk / nch) * 4 + c4;
              if (blk < nch * ncg && n8 < (Np >> 3))
                *reinterpret_cast<uint4*>(sB + (size_t)(kk >> 3) * 128 + (size_t)n8 * (nch * 128) + (kk & 7) * 16) = pack8(v0[u], v1[u]);
            }
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
    // ---- epilogue A: the accumulator row is d(xhat); LayerNorm-hat backward w.r.t. the row of lnx, + residual(s):
    //      dx = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)) over the real channels, pads 0   (rdst_lnhat_bwd) ----
    if (a.lnx) {
      const int r = tid & 127, quarter = tid >> 7;
      const int64_t t = t0 + r;
      const bool live = t < a.T;
      float av[2][16], xv[2][16];
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n0 = (quarter + 4 * i) * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) { av[i][j] = 0.f; xv[i][j] = 0.f; }
        if (n0 < a.N) {                                   // warp-uniform
          uint32_t v[16];
          __syncwarp();
          tmem_ld_x16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + n0, v);
          wait_ld();
          if (live) {
            const float4* xp = reinterpret_cast<const float4*>(a.lnx + t * a.ldlx + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 x4 = __ldg(xp + j);
              xv[i][4 * j] = x4.x; xv[i][4 * j + 1] = x4.y; xv[i][4 * j + 2] = x4.z; xv[i][4 * j + 3] = x4.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              av[i][j] = __uint_as_float(v[j]) * a.out_scale;
              s += xv[i][j];
              ss += xv[i][j] * xv[i][j];
            }
          }
        }
      }
      s_red[0][quarter][r] = make_float2(s, ss);
      __syncthreads();
      const float inv = 1.f / (float)a.lnb_creal;
      float S = 0.f, SS = 0.f;
#pragma unroll
      for (int q = 0; q < NT_GEMM / 128; ++q) { const float2 p = s_red[0][q][r]; S += p.x; SS += p.y; }
      const float mean = S * inv;
      const float rstd = rsqrtf(fmaxf(SS * inv - mean * mean, 0.f) + 1e-5f);
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n0 = (quarter + 4 * i) * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + j;
          const bool real = n < 60 || (n >= 64 && ((n - 64) & 31) < 30);
          xv[i][j] = (xv[i][j] - mean) * rstd;            // xhat
          if (real && n < a.N) { m1 += av[i][j]; m2 = fmaf(av[i][j], xv[i][j], m2); }
        }
      }
      s_red[1][quarter][r] = make_float2(m1, m2);
      __syncthreads();
      float M1 = 0.f, M2 = 0.f;
#pragma unroll
      for (int q = 0; q < NT_GEMM / 128; ++q) { const float2 p = s_red[1][q][r]; M1 += p.x; M2 += p.y; }
      M1 *= inv;
      M2 *= inv;
      if (live) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int n0 = (quarter + 4 * i) * 16;
          if (n0 < a.N) {
            float* yp = a.y + t * a.ldy + n0;
            const float4* rp = a.resid ? reinterpret_cast<const float4*>(a.resid + t * a.ldr + n0) : nullptr;
            const float4* rp2 = a.resid2 ? reinterpret_cast<const float4*>(a.resid2 + t * a.ldr2 + n0) : nullptr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int n = n0 + 4 * j + e;
                const bool real = n < 60 || (n >= 64 && ((n - 64) & 31) < 30);
                o[e] = real ? rstd * (av[i][4 * j + e] - M1 - xv[i][4 * j + e] * M2) : 0.f;
              }
              if (rp) { const float4 q4 = rp[j]; o[0] += q4.x; o[1] += q4.y; o[2] += q4.z; o[3] += q4.w; }
              if (rp2) { const float4 q4 = rp2[j]; o[0] += q4.x; o[1] += q4.y; o[2] += q4.z; o[3] += q4.w; }
              *reinterpret_cast<float4*>(yp + 4 * j) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
    } else
    // ---- epilogue B: thread = token row (TMEM lane), the four thread quarters alternate over 16-column groups ----
    {
      const int r = tid & 127, quarter = tid >> 7;
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
      for (int g = quarter; g * 16 < a.N; g += NT_GEMM / 128) {
        const int n0 = g * 16;
        const bool full = n0 + 16 <= a.N;
        const float* rp = (a.resid && live) ? a.resid + t * a.ldr + n0 : nullptr;
        const float* ap = (a.aux && live) ? a.aux + t * a.lda + n0 : nullptr;
        float4 rv[4], hv[4];                              // residual / GELU operand: in flight while TMEM is read
        if (full) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rv[j] = rp ? __ldg(reinterpret_cast<const float4*>(rp) + j) : zero4();
            if (ap) hv[j] = __ldg(reinterpret_cast<const float4*>(ap) + j);
          }
        }
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
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = (__uint_as_float(v[j]) + s_bias[n0 + j]) * a.out_scale;
        if (full) {
          if (ap) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o[4 * j] *= gelu_grad(hv[j].x); o[4 * j + 1] *= gelu_grad(hv[j].y);
              o[4 * j + 2] *= gelu_grad(hv[j].z); o[4 * j + 3] *= gelu_grad(hv[j].w);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(yp + 4 * j) =
                make_float4(o[4 * j] + rv[j].x, o[4 * j + 1] + rv[j].y, o[4 * j + 2] + rv[j].z, o[4 * j + 3] + rv[j].w);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + j < a.N) yp[j] = o[j] * (ap ? gelu_grad(ap[j]) : 1.f) + (rp ? rp[j] : 0.f);
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
// dW[N][K] += dY^T op(X)   (+ db)
// ------------------------------------------------------------------------------------------------------------------
struct TnArgs {
  const float* dy; int64_t ldy; const float* x; int64_t ldx; float* dw; float* db;
  int64_t T; int N, K;
  int KCW;                      // columns of X (k) per CTA, multiple of 16, <= 240
  int MT;                       // 128-row tiles of dW (n) per CTA: X is staged once for all of them (MT * (KCW+16) <= 512)
  int64_t t_per_split;          // multiple of 128
  int conv, B, H, W, Cin;
  int x_op, x_creal;            // prologue on X: 0 none, 1 LayerNorm-hat (x_creal real channels), 2 exact-erf GELU
};

__global__ void __launch_bounds__(NT_TN, 1) gemm_tn_tc_kernel(const TnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_mean[TM], s_rstd[TM];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MT = a.MT;
  const int m0 = blockIdx.x * MT * TM;              // first output row (n) of this CTA
  const int k0 = blockIdx.y * a.KCW;                // first output column (k)
  const bool with_bias = a.db != nullptr && blockIdx.y == 0;
  const int NB = a.KCW + (with_bias ? 16 : 0);      // MMA N: k columns (+ the ones column group)
  // MN-major images, token = MMA K dimension: element (t, c) at (t/8)*128 + (c/8)*2048 + (t%8)*16 + (c%8)*2
  uint8_t* sA = smem;                               // dY: MT images of 128 tokens x 128 n
  uint8_t* sB = smem + (size_t)MT * TM * 128 * 2;   // X : 128 tokens x NB k

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
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
  const int nbg = (nbc + 3) >> 2;
  constexpr int NW = NT_TN / 32;
  bool first = true;
  for (int64_t t0 = ts; t0 < te; t0 += TM) {
    if (!first) {
      mbar_wait(&bar, phase);
      phase ^= 1;
      fence_after_sync();
    }
    // ---- A images: dY[t][m0 .. m0 + MT*128) ----
    for (int b0 = warp; b0 < MT * 64; b0 += NW * U) {
      float4 v0[U], v1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int blk = b0 + u * NW;
        const int tt = (blk & 15) * 8 + r8, ch = (blk >> 4) * 4 + c4;       // MT*16 chunks of 8 n
        const int64_t t = t0 + tt;
        const int n = m0 + ch * 8;
        v0[u] = zero4();
        v1[u] = zero4();
        if (blk < MT * 64 && t < te && n < a.N) {
          const float* p = a.dy + t * a.ldy + n;
          v0[u] = __ldg(reinterpret_cast<const float4*>(p));
          v1[u] = __ldg(reinterpret_cast<const float4*>(p) + 1);
          mask8(a.N - n, v0[u], v1[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int blk = b0 + u * NW;
        const int tt = (blk & 15) * 8 + r8, ch = (blk >> 4) * 4 + c4;
        if (blk < MT * 64)
          *reinterpret_cast<uint4*>(sA + (size_t)(tt >> 3) * 128 + (size_t)ch * 2048 + (tt & 7) * 16) = pack8(v0[u], v1[u]);
      }
    }
    if (a.x_op == 1) {
      row_stats<NT_TN>(a.x, a.ldx, t0, te, a.K, a.x_creal, s_mean, s_rstd, warp, lane);
      __syncthreads();
    }
    // ---- B image: op(X)[t][k0 .. k0+KCW) (3x3 gather in conv mode) + ones column ----
    for (int b0 = warp; b0 < 16 * nbg; b0 += NW * U) {
      float4 v0[U], v1[U];
      int valid[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int blk = b0 + u * NW;
        const int tt = (blk & 15) * 8 + r8, ch = (blk >> 4) * 4 + c4;
        const int k = k0 + ch * 8;
        const float* p = (blk < 16 * nbg && ch * 8 < a.KCW)
                             ? act_chunk(a.x, a.ldx, t0 + tt, te, k, a.K, a.conv, a.Cin, a.H, a.W, HW) : nullptr;
        valid[u] = p ? min(8, a.K - k) : 0;
        if (p) {
          v0[u] = __ldg(reinterpret_cast<const float4*>(p));
          v1[u] = __ldg(reinterpret_cast<const float4*>(p) + 1);
        } else {
          v0[u] = zero4();
          v1[u] = zero4();
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int blk = b0 + u * NW;
        const int tt = (blk & 15) * 8 + r8, ch = (blk >> 4) * 4 + c4;
        if (blk < 16 * nbg && ch < nbc) {
          if (valid[u] > 0) {
            mask8(valid[u], v0[u], v1[u]);
            apply_op(a.x_op, s_mean[tt], s_rstd[tt], v0[u], v1[u]);
            if (a.x_op == 1) mask8(valid[u], v0[u], v1[u]);
          }
          if (ch * 8 == a.KCW && t0 + tt < te) v0[u].x = 1.f;   // the ones column: accumulates db[n] = sum_t dY[t][n]
          *reinterpret_cast<uint4*>(sB + (size_t)(tt >> 3) * 128 + (size_t)ch * 2048 + (tt & 7) * 16) = pack8(v0[u], v1[u]);
        }
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(TM, NB, true, true);
        for (int m = 0; m < MT; ++m)
          for (int ks = 0; ks < TM / 16; ++ks) {
            const uint64_t da = make_smem_desc(smem_u32(sA) + m * (TM * 128 * 2) + ks * 2 * 128, 128, 2048);
            const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2 * 128, 128, 2048);
            mma_bf16_ss(tmem + m * NB, da, db, idesc, (!first || ks > 0) ? 1u : 0u);
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
    if (warp == 0) tmem_dealloc<512>(tmem);
    return;
  }
  mbar_wait(&bar, phase);
  fence_after_sync();
  // ---- epilogue: thread = output row n of tile m, fp32 reductions into dW / db ----
  {
    const int r = tid & 127, quarter = tid >> 7;
    const bool vec = (a.K & 3) == 0;
    const int ng = NB >> 4;
    for (int gi = quarter; gi < MT * ng; gi += NT_TN / 128) {
      const int m = gi / ng, c0 = (gi - m * ng) * 16;
      const int n = m0 + m * TM + r;
      uint32_t v[16];
      __syncwarp();
      tmem_ld_x16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + m * NB + c0, v);
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
  if (warp == 0) tmem_dealloc<512>(tmem);
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
                            const float* resid, int64_t ldr, const float* gelu_aux, int64_t lda, float* y, int64_t ldy,
                            int64_t T, int K, int N, int a_op, int ln_creal, float out_scale, int conv, int B, int H, int W,
                            int Cin, int shuffle, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && w && y, "rdst_gemm_tc: null pointer");
  RDST_REQUIRE(T >= 0 && K > 0 && N > 0 && N <= 512, "rdst_gemm_tc: bad shape (T=%lld K=%d N=%d; N <= 512)", (long long)T, K, N);
  RDST_REQUIRE((ldx & 3) == 0 && (ldy & 3) == 0 && (!resid || (ldr & 3) == 0) &&
                   ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                   (!resid || ((uintptr_t)resid & 15) == 0),
               "rdst_gemm_tc: activation rows must be 16-byte aligned (pointers and leading dimensions multiples of 4 floats)");
  RDST_REQUIRE(!conv || (Cin > 0 && Cin % 16 == 0 && K == 9 * Cin && T == (int64_t)B * H * W && !w_mn_major && !a_op && ldx >= Cin),
               "rdst_gemm_tc: conv mode needs Cin %% 16 == 0, K == 9*Cin, T == B*H*W, K-major weights, no prologue");
  RDST_REQUIRE(conv || ldx >= (K + 7) / 8 * 8, "rdst_gemm_tc: rows of X must be readable up to the next multiple of 8 columns");
  RDST_REQUIRE(a_op >= 0 && a_op <= 2 && (!gelu_aux || ((lda & 3) == 0 && ((uintptr_t)gelu_aux & 15) == 0)), "rdst_gemm_tc: bad a_op / aux");
  RDST_REQUIRE(!shuffle || (conv && shuffle == 2 && N == 256 && !resid), "rdst_gemm_tc: shuffle=2 needs conv mode, N == 256, no residual");
  RDST_REQUIRE(a_op != 1 || (ln_creal > 0 && ln_creal <= K && (K & 3) == 0), "rdst_gemm_tc: bad LayerNorm width");
  if (T == 0) return RDST_OK;
  GemmArgs a{};
  a.x = x; a.ldx = ldx; a.w = w; a.ldw = ldw; a.w_mn = w_mn_major; a.bias = bias; a.resid = resid; a.ldr = ldr;
  a.w_vec = ((ldw & 3) == 0 && ((uintptr_t)w & 15) == 0) ? 1 : 0;
  a.aux = gelu_aux; a.lda = lda; a.a_op = a_op;
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
  RDST_REQUIRE(smem <= 212 * 1024, "rdst_gemm_tc: operand images need %zu bytes of shared memory (K=%d N=%d)", smem, K, N);
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  if (e != cudaSuccess) { set_error("rdst_gemm_tc: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  const int64_t tiles = (T + TM - 1) / TM;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  gemm_tc_kernel<<<grid, NT_GEMM, smem, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_gemm_tc");
  return RDST_OK;
}

extern "C" int rdst_gemm_tc_lnbwd(const float* dy, int64_t ldy, const float* w, int64_t ldw, const float* x, int64_t ldx,
                                  const float* resid, int64_t ldr, const float* resid2, int64_t ldr2, float* dx, int64_t ldo,
                                  int64_t T, int K, int N, int creal, float scale, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(dy && w && x && dx, "rdst_gemm_tc_lnbwd: null pointer");
  RDST_REQUIRE(T >= 0 && K > 0 && K <= 384 && N > 0 && N <= 128 && (N & 15) == 0 && creal > 0 && creal <= N,
               "rdst_gemm_tc_lnbwd: bad shape (T=%lld K=%d N=%d creal=%d; K <= 384, N <= 128, N %% 16 == 0)", (long long)T, K, N, creal);
  RDST_REQUIRE((ldy & 3) == 0 && (ldx & 3) == 0 && (ldo & 3) == 0 && (!resid || (ldr & 3) == 0) && (!resid2 || (ldr2 & 3) == 0) &&
                   ((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dx & 15) == 0 &&
                   (!resid || ((uintptr_t)resid & 15) == 0) && (!resid2 || ((uintptr_t)resid2 & 15) == 0) &&
                   ldy >= (K + 7) / 8 * 8 && ldx >= N && ldo >= N,
               "rdst_gemm_tc_lnbwd: rows must be 16-byte aligned and wide enough");
  if (T == 0) return RDST_OK;
  GemmArgs a{};
  a.x = dy; a.ldx = ldy; a.w = w; a.ldw = ldw; a.w_mn = 1; a.w_vec = ((ldw & 3) == 0 && ((uintptr_t)w & 15) == 0) ? 1 : 0;
  a.resid = resid; a.ldr = ldr; a.resid2 = resid2; a.ldr2 = ldr2; a.lnx = x; a.ldlx = ldx; a.lnb_creal = creal;
  a.y = dx; a.ldy = ldo; a.T = T; a.K = K; a.N = N; a.Np = N; a.out_scale = scale;
  a.KC = (K + 15) / 16 * 16;
  a.nkc = 1;
  const size_t smem = (size_t)TM * a.KC * 2 + (size_t)a.Np * a.KC * 2;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  if (e != cudaSuccess) { set_error("rdst_gemm_tc_lnbwd: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  const int64_t tiles = (T + TM - 1) / TM;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  gemm_tc_kernel<<<grid, NT_GEMM, smem, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_gemm_tc_lnbwd");
  return RDST_OK;
}

extern "C" int rdst_gemm_tn_tc(const float* dy, int64_t ldy, const float* x, int64_t ldx, float* dw, float* db, int64_t T,
                               int N, int K, int conv, int B, int H, int W, int Cin, int x_op, int x_creal, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(dy && x && dw, "rdst_gemm_tn_tc: null pointer");
  RDST_REQUIRE(T >= 0 && N > 0 && K > 0, "rdst_gemm_tn_tc: bad shape");
  RDST_REQUIRE((ldx & 3) == 0 && (ldy & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 &&
                   ((uintptr_t)dw & 15) == 0,
               "rdst_gemm_tn_tc: rows must be 16-byte aligned (pointers and leading dimensions multiples of 4 floats)");
  RDST_REQUIRE(!conv || (Cin > 0 && Cin % 16 == 0 && K == 9 * Cin && T == (int64_t)B * H * W && ldx >= Cin && !x_op),
               "rdst_gemm_tn_tc: conv mode needs Cin %% 16 == 0, K == 9*Cin, T == B*H*W and no prologue");
  RDST_REQUIRE(ldy >= (N + 7) / 8 * 8 && (conv || ldx >= (K + 7) / 8 * 8),
               "rdst_gemm_tn_tc: rows of dY / X must be readable up to the next multiple of 8 columns");
  RDST_REQUIRE(x_op >= 0 && x_op <= 2 && (x_op != 1 || (x_creal > 0 && x_creal <= K && K <= 240 && (K & 3) == 0)),
               "rdst_gemm_tn_tc: bad prologue (LayerNorm needs the whole row in one column chunk: K <= 240)");
  if (T == 0) return RDST_OK;
  TnArgs a{};
  a.dy = dy; a.ldy = ldy; a.x = x; a.ldx = ldx; a.dw = dw; a.db = db; a.T = T; a.N = N; a.K = K;
  a.conv = conv; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.x_op = x_op; a.x_creal = x_creal;
  const int Kp = (K + 15) / 16 * 16;
  const int kchunks = (Kp + 239) / 240;
  a.KCW = ((Kp + kchunks - 1) / kchunks + 15) / 16 * 16;
  const int kc = (Kp + a.KCW - 1) / a.KCW;
  const int mt_all = (N + TM - 1) / TM;
  a.MT = mt_all;                                                        // all row tiles of dW in one CTA when TMEM holds them
  while (a.MT > 1 && (a.MT * (a.KCW + 16) > 512 || a.MT > 3)) --a.MT;
  const int mt = (mt_all + a.MT - 1) / a.MT;
  int64_t splits = (sm_count() + mt * kc - 1) / (mt * kc);              // one CTA per SM
  const int64_t max_splits = (T + TM - 1) / TM;                         // at least one 128-token stage per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.t_per_split = ((T + splits - 1) / splits + TM - 1) / TM * TM;
  splits = (T + a.t_per_split - 1) / a.t_per_split;
  const size_t smem = (size_t)a.MT * TM * 128 * 2 + (size_t)TM * (a.KCW + 16) * 2;
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  if (e != cudaSuccess) { set_error("rdst_gemm_tn_tc: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  dim3 grid((unsigned)mt, (unsigned)kc, (unsigned)splits);
  gemm_tn_tc_kernel<<<grid, NT_TN, smem, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_gemm_tn_tc");
  return RDST_OK;
}
