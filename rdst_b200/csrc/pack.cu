// Weight packing for the training path, forward and backward, as two small kernels per Linear instead of ~40 torch ops:
//   Wp[pn(n)][pk(k)] = rs(n) * W[n][k] * gamma[k]            bp[pn(n)] = rs(n) * (b[n] + sum_k W[n][k] * beta[k])
// i.e. the LayerNorm affine in front of the Linear folded into it (the GEMM kernels normalise without affine), the
// attention scale head_dim^-0.5 folded into the q rows (rs(n) = q_scale for n < q_rows, else 1; reference
// swin_transformer_sr.py:120), and rows / columns moved to their stored channel positions in the padded dense-block
// layout (include/rdst_b200.h: 60 trunk channels at [0,60), growth group g at [64+32g, +30)).  The backward kernel turns the
// gradients of the packed tensors into the gradients of the reference-named parameters:
//   dW = rs * (gamma * dWp + dbp (x) beta),  db = rs * dbp,  dgamma[k] += sum_n rs W dWp,  dbeta[k] += sum_n rs W dbp.
// Same arithmetic as rdst_b200/packing.py (pack_stl / pack_dstl_tail), which stays the inference-time packer.
#include "common.cuh"

namespace rdst {
namespace {

__device__ __forceinline__ int chan_pos(int c, int scatter) {
  if (!scatter || c < 60) return c;
  const int g = (c - 60) / 30;
  return 64 + 32 * g + (c - 60) - 30 * g;
}

__global__ void __launch_bounds__(128) pack_linear_fwd_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float* __restrict__ Wp, float* __restrict__ bp, int N, int K,
                                                              int ldp, int srows, int scols, int q_rows, float q_scale) {
  const int n = blockIdx.x;
  const float rs = n < q_rows ? q_scale : 1.f;
  const int pn = chan_pos(n, srows);
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) {
    const float w = W[(size_t)n * K + k];
    Wp[(size_t)pn * ldp + chan_pos(k, scols)] = rs * w * (gamma ? gamma[k] : 1.f);
    if (beta) acc += w * beta[k];
  }
  __shared__ float red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) bp[pn] = rs * ((b ? b[n] : 0.f) + red[0] + red[1] + red[2] + red[3]);
}

// block = 32 columns (k) x 8 row groups; dW written coalesced along k, dgamma / dbeta reduced over n
__global__ void __launch_bounds__(256) pack_linear_bwd_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ dWp,
                                                              const float* __restrict__ dbp, float* __restrict__ dW,
                                                              float* __restrict__ db, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int N, int K, int ldp, int srows,
                                                              int scols, int q_rows, float q_scale) {
  const int kx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;
  float ag = 0.f, ab = 0.f;
  if (k < K) {
    const int pk = chan_pos(k, scols);
    const float g = gamma ? gamma[k] : 1.f, be = beta ? beta[k] : 0.f;
    for (int n = blockIdx.y * 8 + ry; n < N; n += 8 * gridDim.y) {
      const float rs = n < q_rows ? q_scale : 1.f;
      const int pn = chan_pos(n, srows);
      const float gw = dWp[(size_t)pn * ldp + pk], gb = dbp[pn];
      const float w = W[(size_t)n * K + k];
      dW[(size_t)n * K + k] = rs * (g * gw + gb * be);
      ag += rs * w * gw;
      ab += rs * w * gb;
      if (blockIdx.x == 0 && kx == 0 && db) db[n] = rs * gb;
    }
  }
  __shared__ float sg[8][33], sb[8][33];
  sg[ry][kx] = ag;
  sb[ry][kx] = ab;
  __syncthreads();
  if (ry == 0 && k < K && dgamma) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { tg += sg[j][kx]; tb += sb[j][kx]; }
    atomicAdd(dgamma + k, tg);          // row groups are split over blockIdx.y; dgamma / dbeta arrive zeroed
    atomicAdd(dbeta + k, tb);
  }
}

// ---- batched form: every Linear of one RDSTB in a single launch (descriptors travel in the kernel parameter space) ----
struct PackBatch {
  RdstPackDesc d[RDST_PACK_MAX];
};

__global__ void __launch_bounds__(128) pack_batch_fwd_kernel(const __grid_constant__ PackBatch pb) {
  const RdstPackDesc& p = pb.d[blockIdx.y];
  const int n = blockIdx.x;
  if (n >= p.N) return;
  const float rs = n < p.q_rows ? p.q_scale : 1.f;
  const int pn = chan_pos(n, p.scatter_rows);
  float acc = 0.f;
  for (int k = threadIdx.x; k < p.K; k += 128) {
    const float w = p.W[(size_t)n * p.K + k];
    p.Wp[(size_t)pn * p.ldp + chan_pos(k, p.scatter_cols)] = rs * w * (p.gamma ? p.gamma[k] : 1.f);
    if (p.beta) acc += w * p.beta[k];
  }
  __shared__ float red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) p.bp[pn] = rs * ((p.b ? p.b[n] : 0.f) + red[0] + red[1] + red[2] + red[3]);
}

__global__ void __launch_bounds__(256) pack_batch_bwd_kernel(const __grid_constant__ PackBatch pb) {
  const RdstPackDesc& p = pb.d[blockIdx.z];
  const int kx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;
  if (blockIdx.x * 32 >= p.K || blockIdx.y * 64 >= p.N) return;           // whole block outside this Linear
  float ag = 0.f, ab = 0.f;
  if (k < p.K) {
    const int pk = chan_pos(k, p.scatter_cols);
    const float g = p.gamma ? p.gamma[k] : 1.f, be = p.beta ? p.beta[k] : 0.f;
    for (int n = blockIdx.y * 64 + ry; n < min(p.N, blockIdx.y * 64 + 64); n += 8) {
      const float rs = n < p.q_rows ? p.q_scale : 1.f;
      const int pn = chan_pos(n, p.scatter_rows);
      const float gw = p.dWp[(size_t)pn * p.ldp + pk], gb = p.dbp[pn];
      const float w = p.W[(size_t)n * p.K + k];
      p.dW[(size_t)n * p.K + k] = rs * (g * gw + gb * be);
      ag += rs * w * gw;
      ab += rs * w * gb;
      if (blockIdx.x == 0 && kx == 0 && p.db) p.db[n] = rs * gb;
    }
  }
  __shared__ float sg[8][33], sb[8][33];
  sg[ry][kx] = ag;
  sb[ry][kx] = ab;
  __syncthreads();
  if (ry == 0 && k < p.K && p.dgamma) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { tg += sg[j][kx]; tb += sb[j][kx]; }
    atomicAdd(p.dgamma + k, tg);
    atomicAdd(p.dbeta + k, tb);
  }
}

}  // namespace
}  // namespace rdst

extern "C" int rdst_pack_linear_batch(const RdstPackDesc* descs, int n, int backward, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(descs && n > 0 && n <= RDST_PACK_MAX, "rdst_pack_linear_batch: need 1..%d descriptors", RDST_PACK_MAX);
  PackBatch pb;
  int maxN = 0, maxK = 0;
  for (int i = 0; i < n; ++i) {
    const RdstPackDesc& p = descs[i];
    RDST_REQUIRE(p.W && p.N > 0 && p.K > 0 && p.ldp >= p.K && (p.gamma != nullptr) == (p.beta != nullptr),
                 "rdst_pack_linear_batch: bad descriptor %d", i);
    if (!backward) RDST_REQUIRE(p.Wp && p.bp, "rdst_pack_linear_batch: descriptor %d: null packed pointers", i);
    else RDST_REQUIRE(p.dWp && p.dbp && p.dW && (p.gamma != nullptr) == (p.dgamma != nullptr) && (p.dgamma != nullptr) == (p.dbeta != nullptr),
                      "rdst_pack_linear_batch: descriptor %d: bad gradient pointers", i);
    pb.d[i] = p;
    maxN = p.N > maxN ? p.N : maxN;
    maxK = p.K > maxK ? p.K : maxK;
  }
  if (!backward) {
    pack_batch_fwd_kernel<<<dim3((unsigned)maxN, (unsigned)n), 128, 0, (cudaStream_t)stream>>>(pb);
  } else {
    pack_batch_bwd_kernel<<<dim3((unsigned)((maxK + 31) / 32), (unsigned)((maxN + 63) / 64), (unsigned)n), 256, 0,
                            (cudaStream_t)stream>>>(pb);
  }
  RDST_CHECK_LAUNCH("rdst_pack_linear_batch");
  return RDST_OK;
}

extern "C" int rdst_pack_linear_fwd(const float* W, const float* b, const float* gamma, const float* beta, float* Wp, float* bp,
                                    int N, int K, int ldp, int scatter_rows, int scatter_cols, int q_rows, float q_scale,
                                    void* stream) {
  using namespace rdst;
  RDST_REQUIRE(W && Wp && bp, "rdst_pack_linear_fwd: null pointer");
  RDST_REQUIRE(N > 0 && K > 0 && ldp >= K && (gamma != nullptr) == (beta != nullptr), "rdst_pack_linear_fwd: bad argument");
  pack_linear_fwd_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(W, b, gamma, beta, Wp, bp, N, K, ldp, scatter_rows, scatter_cols,
                                                              q_rows, q_scale);
  RDST_CHECK_LAUNCH("rdst_pack_linear_fwd");
  return RDST_OK;
}

extern "C" int rdst_pack_linear_bwd(const float* W, const float* gamma, const float* beta, const float* dWp, const float* dbp,
                                    float* dW, float* db, float* dgamma, float* dbeta, int N, int K, int ldp, int scatter_rows,
                                    int scatter_cols, int q_rows, float q_scale, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(W && dWp && dbp && dW, "rdst_pack_linear_bwd: null pointer");
  RDST_REQUIRE(N > 0 && K > 0 && ldp >= K && (gamma != nullptr) == (beta != nullptr) && (gamma != nullptr) == (dgamma != nullptr) &&
                   (dgamma != nullptr) == (dbeta != nullptr),
               "rdst_pack_linear_bwd: bad argument");
  const dim3 grid((unsigned)((K + 31) / 32), (unsigned)((N + 63) / 64));
  pack_linear_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, gamma, beta, dWp, dbp, dW, db, dgamma, dbeta, N, K,
                                                                         ldp, scatter_rows, scatter_cols, q_rows, q_scale);
  RDST_CHECK_LAUNCH("rdst_pack_linear_bwd");
  return RDST_OK;
}
