// Backward kernels (fp32, CUDA cores) for the training path of the drop-in module.
// Together with the forward kernels of simt_kernels.cu (which also serve as the data-gradient kernels:
// dX = dY . W is rdst_linear_fwd with the transposed weight, conv dgrad is rdst_conv3x3_fwd with the flipped,
// transposed filter) they implement the gradient of every layer on the RDST path:
//   rdst_gemm_tn_acc        dW[n][k] += sum_t dY[t][n] * X[t][k]   (+ column sums for the bias), optional 3x3 gather
//   rdst_lnhat_fwd / _bwd   affine-free LayerNorm (affine is folded into the next Linear by the host)
//   rdst_layernorm_bwd      LayerNorm with affine (patch_embed.norm, final norm)
//   rdst_gelu_fwd / _bwd    exact erf GELU
//   rdst_window_attention_bwd   gradient of the windowed attention core incl. the relative-position table
// Reference semantics: autograd of the modules cited in include/rdst_b200.h.
#include "common.cuh"

namespace rdst {

// real (non-pad) channel test for the padded dense layout: trunk [0,60), growth blocks [64+32g, +30)
__device__ __forceinline__ bool dense_real(int p) { return p < 60 || (p >= 64 && ((p - 64) & 31) < 30); }

// ------------------------------------------------------------------------------------------------
// dW += dY^T X   (reduction over tokens, split over blockIdx.z, fp32 atomics)
// ------------------------------------------------------------------------------------------------
struct TnArgs {
  const float* dy; int64_t ldy;
  const float* x; int64_t ldx;
  float* dw; float* db;
  int64_t T; int N; int K;
  int B, H, W, Cin, conv;     // conv: X rows are gathered 3x3 neighbourhoods, K = 9*Cin
  int64_t t_per_split;
};

__global__ void __launch_bounds__(256) gemm_tn_kernel(TnArgs a) {
  __shared__ float As[16][64 + 4];     // dY tile  [t][n]
  __shared__ float Bs[16][64 + 4];     // X tile   [t][k]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int64_t tbeg = (int64_t)blockIdx.z * a.t_per_split;
  const int64_t tend = tbeg + a.t_per_split < a.T ? tbeg + a.t_per_split : a.T;
  const int tx = tid & 15, ty = tid >> 4;      // thread computes n = n0 + ty*4.., k = k0 + tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;                            // column sum for n = n0 + (tid & 63) (threads with tid < 64, k-tile 0)
  const int hw = a.H * a.W;

  for (int64_t t0 = tbeg; t0 < tend; t0 += 16) {
    // stage 16 tokens x 64 columns of dY and of X
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;           // 0..1023
      const int tt = idx >> 6, c = idx & 63;
      const int64_t t = t0 + tt;
      float va = 0.f, vb = 0.f;
      if (t < tend) {
        if (n0 + c < a.N) va = __ldg(a.dy + t * a.ldy + n0 + c);
        if (k0 + c < a.K) {
          if (a.conv) {
            const int gk = k0 + c, tap = gk / a.Cin, ch = gk - tap * a.Cin;
            const int b = (int)(t / hw), rem = (int)(t % hw), y = rem / a.W + tap / 3 - 1, x = rem % a.W + tap % 3 - 1;
            if (y >= 0 && y < a.H && x >= 0 && x < a.W)
              vb = __ldg(a.x + (((int64_t)b * a.H + y) * a.W + x) * a.ldx + ch);
          } else {
            vb = __ldg(a.x + t * a.ldx + k0 + c);
          }
        }
      }
      As[tt][c] = va;
      Bs[tt][c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < 16; ++tt) {
      const float4 av = *reinterpret_cast<const float4*>(&As[tt][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[tt][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (a.db != nullptr && blockIdx.y == 0 && tid < 64) {
#pragma unroll
      for (int tt = 0; tt < 16; ++tt) bsum += As[tt][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= a.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < a.K) atomicAdd(a.dw + (int64_t)n * a.K + k, acc[i][j]);
    }
  }
  if (a.db != nullptr && blockIdx.y == 0 && tid < 64 && n0 + tid < a.N) atomicAdd(a.db + n0 + tid, bsum);
}

// ------------------------------------------------------------------------------------------------
// affine-free LayerNorm forward / backward (warp per token)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lnhat_fwd_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
                                                        int64_t ldy, int64_t T, int K, int creal, int dense_layout) {
  const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const float* row = x + t * ldx;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += row[k];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)creal;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) { const float d = row[k] - mean; ss += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  ss -= (float)(K - creal) * mean * mean;
  const float rstd = rsqrtf(fmaxf(ss, 0.f) / (float)creal + 1e-5f);
  float* yr = y + t * ldy;
  for (int k = lane; k < K; k += 32) {
    const bool real = dense_layout ? dense_real(k) : (k < creal);
    yr[k] = real ? (row[k] - mean) * rstd : 0.f;
  }
}

// dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)) (+ resid), means over the real channels
__global__ void __launch_bounds__(256) lnhat_bwd_kernel(const float* __restrict__ dxh, int64_t ldd, const float* __restrict__ x,
                                                        int64_t ldx, const float* __restrict__ resid, int64_t ldr,
                                                        const float* resid2, int64_t ldr2,
                                                        float* dx, int64_t ldo, int64_t T, int K, int creal,
                                                        int dense_layout) {
  const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const float* row = x + t * ldx;
  const float* drow = dxh + t * ldd;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += row[k];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)creal;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) { const float d = row[k] - mean; ss += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  ss -= (float)(K - creal) * mean * mean;
  const float rstd = rsqrtf(fmaxf(ss, 0.f) / (float)creal + 1e-5f);
  float m1 = 0.f, m2 = 0.f;
  for (int k = lane; k < K; k += 32) {
    const bool real = dense_layout ? dense_real(k) : (k < creal);
    if (real) { const float g = drow[k]; m1 += g; m2 += g * (row[k] - mean) * rstd; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
  m1 /= (float)creal; m2 /= (float)creal;
  float* orow = dx + t * ldo;
  for (int k = lane; k < K; k += 32) {
    const bool real = dense_layout ? dense_real(k) : (k < creal);
    float v = real ? rstd * (drow[k] - m1 - (row[k] - mean) * rstd * m2) : 0.f;
    if (resid) v += resid[t * ldr + k];
    if (resid2) v += resid2[t * ldr2 + k];
    orow[k] = v;
  }
}

// LayerNorm with affine: y = (xhat*gamma + beta)*scale.  dx, dgamma += , dbeta += (atomics per block)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, int64_t ldd, const float* __restrict__ x,
                                                            int64_t ldx, const float* __restrict__ gamma, float* __restrict__ dx,
                                                            int64_t ldo, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int64_t T, int creal, float scale) {
  __shared__ float sg[128], sb[128];
  for (int i = threadIdx.x; i < 128; i += 256) { sg[i] = 0.f; sb[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < T; t += (int64_t)gridDim.x * 8) {
    const float* row = x + t * ldx;
    const float* drow = dy + t * ldd;
    float s = 0.f;
    for (int k = lane; k < creal; k += 32) s += row[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)creal;
    float ss = 0.f;
    for (int k = lane; k < creal; k += 32) { const float d = row[k] - mean; ss += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)creal + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
    for (int k = lane; k < creal; k += 32) {
      const float xh = (row[k] - mean) * rstd, g = drow[k] * scale;
      atomicAdd(&sg[k], g * xh);
      atomicAdd(&sb[k], g);
      const float gh = g * gamma[k];
      m1 += gh; m2 += gh * xh;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
    m1 /= (float)creal; m2 /= (float)creal;
    if (dx) {
      float* orow = dx + t * ldo;
      for (int k = lane; k < creal; k += 32) {
        const float xh = (row[k] - mean) * rstd;
        orow[k] = rstd * (drow[k] * scale * gamma[k] - m1 - xh * m2);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < creal; i += 256) { atomicAdd(dgamma + i, sg[i]); atomicAdd(dbeta + i, sb[i]); }
}

// ------------------------------------------------------------------------------------------------
// GELU (exact erf) forward / backward, elementwise over a [T][N] matrix with row strides
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gelu_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t ldd,
                                                   float* __restrict__ out, int64_t ldo, int64_t T, int N) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= T * N) return;
  const int64_t t = i / N; const int n = (int)(i % N);
  const float v = x[t * ldx + n];
  if (dy == nullptr) {
    out[t * ldo + n] = gelu_erf(v);
  } else {
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
    out[t * ldo + n] = dy[t * ldd + n] * (cdf + v * pdf);
  }
}

// ------------------------------------------------------------------------------------------------
// window attention backward: one block per (window, head), one thread per query token
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(64) window_attention_bwd_kernel(const float* __restrict__ qkv, int64_t ldq,
                                                                  const float* __restrict__ table, const float* __restrict__ dout,
                                                                  int64_t ldo, float* __restrict__ dqkv, int64_t ldg,
                                                                  float* __restrict__ dtable, int H, int W, int C, int heads,
                                                                  int shift) {
  // q / k / v / dO rows padded to a multiple of 4 floats (pads zero): the row a loop iteration needs is the same for all
  // threads (broadcast), so one 16-byte shared-memory load feeds four FMAs
  constexpr int HP = (HD + 3) / 4 * 4;
  extern __shared__ __align__(16) float smf[];
  float (*sq)[HP] = reinterpret_cast<float (*)[HP]>(smf);
  float (*sk)[HP] = sq + 64;
  float (*sv)[HP] = sk + 64;
  float (*sgo)[HP] = sv + 64;
  float (*sP)[65] = reinterpret_cast<float (*)[65]>(smf + 4 * 64 * HP);
  float (*sdS)[65] = sP + 64;
  float* stab = smf + 4 * 64 * HP + 2 * 64 * 65;
  float* sdtab = stab + 225;                        // two private copies (one per warp): plain read-modify-write, no atomics
  int* sreg = reinterpret_cast<int*>(sdtab + 2 * 225);
  const int wi = blockIdx.x, h = blockIdx.y, i = threadIdx.x;
  const int nwx = W >> 3, nw_img = (H >> 3) * nwx;
  const int b = wi / nw_img, wl = wi - b * nw_img;
  const int wy = wl / nwx, wx = wl - wy * nwx;
  const int iy = i >> 3, ix = i & 7;
  const int hs = wy * 8 + iy, ws = wx * 8 + ix;
  int hh = hs + shift; if (hh >= H) hh -= H;
  int ww = ws + shift; if (ww >= W) ww -= W;
  const int64_t t = ((int64_t)b * H + hh) * W + ww;
  for (int e = i; e < 225; e += 64) { stab[e] = table[e * heads + h]; sdtab[e] = 0.f; sdtab[225 + e] = 0.f; }
  int reg = 0;
  if (shift > 0) {
    const int rh = hs < H - 8 ? 0 : (hs < H - shift ? 1 : 2);
    const int rw = ws < W - 8 ? 0 : (ws < W - shift ? 1 : 2);
    reg = rh * 3 + rw;
  }
  sreg[i] = reg;
  const float* row = qkv + t * ldq + h * HD;
#pragma unroll
  for (int d = 0; d < HP; ++d) {
    sq[i][d] = d < HD ? row[d] : 0.f;
    sk[i][d] = d < HD ? row[C + d] : 0.f;
    sv[i][d] = d < HD ? row[2 * C + d] : 0.f;
    sgo[i][d] = d < HD ? dout[t * ldo + h * HD + d] : 0.f;
  }
  __syncthreads();
  // phase 1 (thread = query i): probabilities P_i., dP_i. = dO_i . v_j, dS_i. = P (dP - sum_k P_k dP_k), dq_i
  {
    float q[HP], go[HP];
#pragma unroll
    for (int d = 0; d < HP; ++d) { q[d] = sq[i][d]; go[d] = sgo[i][d]; }
    float mx = -INFINITY;
    for (int j = 0; j < 64; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < HP; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&sk[j][d]);
        acc = fmaf(q[d], kv.x, acc); acc = fmaf(q[d + 1], kv.y, acc); acc = fmaf(q[d + 2], kv.z, acc); acc = fmaf(q[d + 3], kv.w, acc);
      }
      acc += stab[(iy - (j >> 3) + 7) * 15 + (ix - (j & 7) + 7)];
      if (shift > 0 && sreg[j] != reg) acc += -100.0f;
      sP[i][j] = acc;
      mx = fmaxf(mx, acc);
    }
    float sum = 0.f;
    for (int j = 0; j < 64; ++j) { const float e = expf(sP[i][j] - mx); sP[i][j] = e; sum += e; }
    const float inv = 1.0f / sum;
    float dotsum = 0.f;
    for (int j = 0; j < 64; ++j) {
      const float pj = sP[i][j] * inv;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < HP; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&sv[j][d]);
        acc = fmaf(go[d], vv.x, acc); acc = fmaf(go[d + 1], vv.y, acc); acc = fmaf(go[d + 2], vv.z, acc); acc = fmaf(go[d + 3], vv.w, acc);
      }
      sP[i][j] = pj;
      sdS[i][j] = acc;
      dotsum = fmaf(pj, acc, dotsum);
    }
    float dq[HP];
#pragma unroll
    for (int d = 0; d < HP; ++d) dq[d] = 0.f;
    for (int j = 0; j < 64; ++j) {
      const float ds = sP[i][j] * (sdS[i][j] - dotsum);
      sdS[i][j] = ds;
#pragma unroll
      for (int d = 0; d < HP; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&sk[j][d]);
        dq[d] = fmaf(ds, kv.x, dq[d]); dq[d + 1] = fmaf(ds, kv.y, dq[d + 1]); dq[d + 2] = fmaf(ds, kv.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kv.w, dq[d + 3]);
      }
      // for one key j the 32 query rows of a warp hit 32 different table entries
      sdtab[(i >> 5) * 225 + (iy - (j >> 3) + 7) * 15 + (ix - (j & 7) + 7)] += ds;
      __syncwarp();
    }
    float* grow = dqkv + t * ldg + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) grow[d] = dq[d];
  }
  __syncthreads();
  // phase 2 (thread = key j = i): dk_j = sum_i dS_ij q_i ; dv_j = sum_i P_ij dO_i
  {
    float dk[HP], dv[HP];
#pragma unroll
    for (int d = 0; d < HP; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int r = 0; r < 64; ++r) {
      const float ds = sdS[r][i], pr = sP[r][i];
#pragma unroll
      for (int d = 0; d < HP; d += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(&sq[r][d]);
        const float4 gv = *reinterpret_cast<const float4*>(&sgo[r][d]);
        dk[d] = fmaf(ds, qv.x, dk[d]); dk[d + 1] = fmaf(ds, qv.y, dk[d + 1]); dk[d + 2] = fmaf(ds, qv.z, dk[d + 2]); dk[d + 3] = fmaf(ds, qv.w, dk[d + 3]);
        dv[d] = fmaf(pr, gv.x, dv[d]); dv[d + 1] = fmaf(pr, gv.y, dv[d + 1]); dv[d + 2] = fmaf(pr, gv.z, dv[d + 2]); dv[d + 3] = fmaf(pr, gv.w, dv[d + 3]);
      }
    }
    float* grow = dqkv + t * ldg + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) { grow[C + d] = dk[d]; grow[2 * C + d] = dv[d]; }
  }
  for (int e = i; e < 225; e += 64)
    if (sdtab[e] + sdtab[225 + e] != 0.f) atomicAdd(dtable + e * heads + h, sdtab[e] + sdtab[225 + e]);
}

template <int HD>
static void launch_wattn_bwd(dim3 grid, cudaStream_t st, const float* qkv, int64_t ldq, const float* table, const float* dout,
                             int64_t ldo, float* dqkv, int64_t ldg, float* dtable, int H, int W, int C, int heads, int shift) {
  const size_t smem = (size_t)(4 * 64 * ((HD + 3) / 4 * 4) + 2 * 64 * 65 + 3 * 225 + 64) * sizeof(float);
  cudaFuncSetAttribute(window_attention_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  window_attention_bwd_kernel<HD><<<grid, 64, smem, st>>>(qkv, ldq, table, dout, ldo, dqkv, ldg, dtable, H, W, C, heads, shift);
}

// y[t][n] += alpha * x[t][n]
__global__ void __launch_bounds__(256) axpy_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy,
                                                   int64_t T, int N, float alpha) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= T * N) return;
  const int64_t t = i / N; const int n = (int)(i % N);
  y[t * ldy + n] += alpha * x[t * ldx + n];
}

// inverse of the pixel-shuffle store of rdst_conv3x3_fwd(shuffle=2): z[t][s*G + c] = u[(b, 2y+dy, 2x+dx)][c], s = 2dy+dx
__global__ void __launch_bounds__(256) unshuffle_kernel(const float* __restrict__ u, int64_t ldu, float* __restrict__ z, int64_t ldz,
                                                        int B, int H, int W, int G) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t total = (int64_t)B * H * W * 4 * G;
  if (i >= total) return;
  const int c = (int)(i % G);
  const int s = (int)((i / G) % 4);
  const int64_t t = i / (4 * G);
  const int hw = H * W;
  const int b = (int)(t / hw), rem = (int)(t % hw), y = rem / W, x = rem % W;
  const int64_t tu = ((int64_t)(b * 2 * H + 2 * y + (s >> 1))) * (2 * W) + 2 * x + (s & 1);
  z[t * ldz + s * G + c] = u[tu * ldu + c];
}

}  // namespace rdst

using namespace rdst;

extern "C" int rdst_axpy(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t T, int N, float alpha, void* stream) {
  RDST_REQUIRE(x && y && T >= 0 && N > 0, "rdst_axpy: bad argument");
  if (T == 0) return RDST_OK;
  axpy_kernel<<<(unsigned)((T * N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, T, N, alpha);
  RDST_CHECK_LAUNCH("rdst_axpy");
  return RDST_OK;
}

extern "C" int rdst_pixel_unshuffle2(const float* u, int64_t ldu, float* z, int64_t ldz, int B, int H, int W, int G,
                                     void* stream) {
  RDST_REQUIRE(u && z && B >= 0 && H > 0 && W > 0 && G > 0, "rdst_pixel_unshuffle2: bad argument");
  if (B == 0) return RDST_OK;
  const int64_t total = (int64_t)B * H * W * 4 * G;
  unshuffle_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, ldu, z, ldz, B, H, W, G);
  RDST_CHECK_LAUNCH("rdst_pixel_unshuffle2");
  return RDST_OK;
}

extern "C" int rdst_gemm_tn_acc(const float* dy, int64_t ldy, const float* x, int64_t ldx, float* dw, float* db, int64_t T,
                                int N, int K, int conv, int B, int H, int W, int Cin, void* stream) {
  RDST_REQUIRE(dy && x && dw, "rdst_gemm_tn_acc: null pointer");
  RDST_REQUIRE(T >= 0 && N > 0 && K > 0, "rdst_gemm_tn_acc: bad shape");
  RDST_REQUIRE(!conv || (Cin > 0 && K == 9 * Cin && T == (int64_t)B * H * W),
               "rdst_gemm_tn_acc: conv mode needs K == 9*Cin and T == B*H*W");
  if (T == 0) return RDST_OK;
  TnArgs a{};
  a.dy = dy; a.ldy = ldy; a.x = x; a.ldx = ldx; a.dw = dw; a.db = db; a.T = T; a.N = N; a.K = K;
  a.conv = conv; a.B = B; a.H = H; a.W = W; a.Cin = Cin;
  const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
  int splits = (int)((T + 255) / 256);
  const int max_splits = (148 * 8 + tiles - 1) / tiles;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.t_per_split = ((T + splits - 1) / splits + 15) / 16 * 16;
  splits = (int)((T + a.t_per_split - 1) / a.t_per_split);
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((K + 63) / 64), (unsigned)splits);
  gemm_tn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_gemm_tn_acc");
  return RDST_OK;
}

extern "C" int rdst_lnhat_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t T, int K, int creal,
                              int dense_layout, void* stream) {
  RDST_REQUIRE(x && y, "rdst_lnhat_fwd: null pointer");
  RDST_REQUIRE(T >= 0 && K > 0 && creal > 0 && creal <= K && ldx >= K && ldy >= K, "rdst_lnhat_fwd: bad shape");
  if (T == 0) return RDST_OK;
  lnhat_fwd_kernel<<<(unsigned)((T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, T, K, creal, dense_layout);
  RDST_CHECK_LAUNCH("rdst_lnhat_fwd");
  return RDST_OK;
}

extern "C" int rdst_lnhat_bwd(const float* dxhat, int64_t ldd, const float* x, int64_t ldx, const float* resid, int64_t ldr,
                              const float* resid2, int64_t ldr2, float* dx, int64_t ldo, int64_t T, int K, int creal,
                              int dense_layout, void* stream) {
  RDST_REQUIRE(dxhat && x && dx, "rdst_lnhat_bwd: null pointer");
  RDST_REQUIRE(T >= 0 && K > 0 && creal > 0 && creal <= K, "rdst_lnhat_bwd: bad shape");
  if (T == 0) return RDST_OK;
  lnhat_bwd_kernel<<<(unsigned)((T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dxhat, ldd, x, ldx, resid, ldr, resid2, ldr2, dx, ldo,
                                                                             T, K, creal, dense_layout);
  RDST_CHECK_LAUNCH("rdst_lnhat_bwd");
  return RDST_OK;
}

extern "C" int rdst_layernorm_bwd(const float* dy, int64_t ldd, const float* x, int64_t ldx, const float* gamma, float* dx,
                                  int64_t ldo, float* dgamma, float* dbeta, int64_t T, int creal, float scale, void* stream) {
  RDST_REQUIRE(dy && x && gamma && dgamma && dbeta, "rdst_layernorm_bwd: null pointer");
  RDST_REQUIRE(T >= 0 && creal > 0 && creal <= 128, "rdst_layernorm_bwd: bad shape (creal <= 128)");
  if (T == 0) return RDST_OK;
  unsigned grid = (unsigned)((T + 7) / 8);
  if (grid > 1184) grid = 1184;
  layernorm_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, ldd, x, ldx, gamma, dx, ldo, dgamma, dbeta, T, creal, scale);
  RDST_CHECK_LAUNCH("rdst_layernorm_bwd");
  return RDST_OK;
}

extern "C" int rdst_gelu_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t T, int N, void* stream) {
  RDST_REQUIRE(x && y && T >= 0 && N > 0, "rdst_gelu_fwd: bad argument");
  if (T == 0) return RDST_OK;
  gelu_kernel<<<(unsigned)((T * N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, nullptr, 0, y, ldy, T, N);
  RDST_CHECK_LAUNCH("rdst_gelu_fwd");
  return RDST_OK;
}

extern "C" int rdst_gelu_bwd(const float* x, int64_t ldx, const float* dy, int64_t ldd, float* dx, int64_t ldo, int64_t T,
                             int N, void* stream) {
  RDST_REQUIRE(x && dy && dx && T >= 0 && N > 0, "rdst_gelu_bwd: bad argument");
  if (T == 0) return RDST_OK;
  gelu_kernel<<<(unsigned)((T * N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, dy, ldd, dx, ldo, T, N);
  RDST_CHECK_LAUNCH("rdst_gelu_bwd");
  return RDST_OK;
}

extern "C" int rdst_window_attention_bwd(const float* qkv, int64_t ldq, const float* table, const float* dout, int64_t ldo,
                                         float* dqkv, int64_t ldg, float* dtable, int B, int H, int W, int C, int heads,
                                         int shift, void* stream) {
  RDST_REQUIRE(qkv && table && dout && dqkv && dtable, "rdst_window_attention_bwd: null pointer");
  RDST_REQUIRE(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0,
               "rdst_window_attention_bwd: H=%d W=%d must be positive multiples of the window size 8", H, W);
  RDST_REQUIRE(heads > 0 && C % heads == 0 && (shift == 0 || shift == 4), "rdst_window_attention_bwd: bad C/heads/shift");
  if (B <= 0) return RDST_OK;
  dim3 grid((unsigned)(B * (H / 8) * (W / 8)), (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C / heads) {
    case 10: launch_wattn_bwd<10>(grid, st, qkv, ldq, table, dout, ldo, dqkv, ldg, dtable, H, W, C, heads, shift); break;
    case 15: launch_wattn_bwd<15>(grid, st, qkv, ldq, table, dout, ldo, dqkv, ldg, dtable, H, W, C, heads, shift); break;
    case 20: launch_wattn_bwd<20>(grid, st, qkv, ldq, table, dout, ldo, dqkv, ldg, dtable, H, W, C, heads, shift); break;
    default: set_error("rdst_window_attention_bwd: head_dim %d unsupported (10,15,20)", C / heads); return RDST_E_UNSUPPORTED;
  }
  RDST_CHECK_LAUNCH("rdst_window_attention_bwd");
  return RDST_OK;
}
