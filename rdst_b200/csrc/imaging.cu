// LR synthesis and image metrics on the device: the callers either side of the network.
//   rdst_bicubic_resize_f32  cv2.resize(..., interpolation=cv2.INTER_CUBIC) of single-channel fp32 images, as the datasets
//                            call it for LR synthesis and for the bicubic "res" image (datasets/basic_dataset.py:65-123,
//                            :258-301): separable 4-tap cubic convolution, A = -0.75, replicated borders, no antialiasing
//   rdst_sqdiff_sum_f64      per-image sum of squared differences in fp64 -> PSNR (metrics/sr_metrics.py:8-9)
//   rdst_ssim_sum_f64        per-image sum of the SSIM map over the valid region (7x7 uniform window, sample covariance,
//                            K1 = 0.01, K2 = 0.03, data_range 1: skimage structural_similarity as :12-13 calls it)
// All three are HBM-bound: every source pixel is read once from DRAM (neighbouring threads share taps through L1/L2).
#include "common.cuh"

namespace rdst {

// out[b][y][x] = sum_r cy[y][r] * ( sum_k cx[x][k] * src[b][iy[y][r]][ix[x][k]] ), plain fp32 multiplies and adds in this order
// (no fused multiply-add): the same arithmetic as oracle/imaging_oracle.py, which is pinned to cv2 within a few ulp.
__global__ void __launch_bounds__(256) bicubic_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                      const int* __restrict__ ix, const float* __restrict__ cx,
                                                      const int* __restrict__ iy, const float* __restrict__ cy,
                                                      int B, int Hs, int Ws, int Hd, int Wd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= Wd) return;
  const int4 jx = *reinterpret_cast<const int4*>(ix + 4 * x);
  const float4 ax = *reinterpret_cast<const float4*>(cx + 4 * x);
  const int4 jy = *reinterpret_cast<const int4*>(iy + 4 * y);
  const float4 ay = *reinterpret_cast<const float4*>(cy + 4 * y);
  const float* img = src + (size_t)b * Hs * Ws;
  const int rows[4] = {jy.x, jy.y, jy.z, jy.w};
  const float wy[4] = {ay.x, ay.y, ay.z, ay.w};
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* p = img + (size_t)rows[r] * Ws;
    float h = __fmul_rn(__ldg(p + jx.x), ax.x);
    h = __fadd_rn(h, __fmul_rn(__ldg(p + jx.y), ax.y));
    h = __fadd_rn(h, __fmul_rn(__ldg(p + jx.z), ax.z));
    h = __fadd_rn(h, __fmul_rn(__ldg(p + jx.w), ax.w));
    const float t = __fmul_rn(h, wy[r]);
    acc = r == 0 ? t : __fadd_rn(acc, t);
  }
  dst[((size_t)b * Hd + y) * Wd + x] = acc;
}

__global__ void __launch_bounds__(256) sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     double* __restrict__ out, int64_t n) {
  const int img = blockIdx.y;
  const float* pa = a + (size_t)img * n;
  const float* pb = b + (size_t)img * n;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = (double)pa[i] - (double)pb[i];
    s += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out + img, t);
  }
}

// one thread per pixel of the valid region [3, H-3) x [3, W-3); window sums in fp64
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   double* __restrict__ out, int H, int W) {
  const int img = blockIdx.z;
  const int x = 3 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = 3 + blockIdx.y * 8 + (threadIdx.x >> 5);
  const float* pa = a + (size_t)img * H * W;
  const float* pb = b + (size_t)img * H * W;
  double s = 0.0;
  if (x < W - 3 && y < H - 3) {
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    for (int dy = -3; dy <= 3; ++dy)
#pragma unroll
      for (int dx = -3; dx <= 3; ++dx) {
        const double u = pa[(size_t)(y + dy) * W + x + dx], v = pb[(size_t)(y + dy) * W + x + dx];
        sx += u; sy += v; sxx += u * u; syy += v * v; sxy += u * v;
      }
    const double NP = 49.0, cov_norm = NP / (NP - 1.0);
    const double ux = sx / NP, uy = sy / NP;
    const double vx = cov_norm * (sxx / NP - ux * ux), vy = cov_norm * (syy / NP - uy * uy), vxy = cov_norm * (sxy / NP - ux * uy);
    const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;              // (K1 R)^2, (K2 R)^2 with data_range R = 1
    s = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out + img, t);
  }
}

}  // namespace rdst

extern "C" int rdst_bicubic_resize_f32(const float* src, float* dst, const int* ix, const float* cx, const int* iy,
                                       const float* cy, int B, int Hs, int Ws, int Hd, int Wd, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(src && dst && ix && cx && iy && cy, "rdst_bicubic_resize_f32: null pointer");
  RDST_REQUIRE(B >= 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && Hd <= 65535 && B <= 65535, "rdst_bicubic_resize_f32: bad shape");
  RDST_REQUIRE(((uintptr_t)ix % 16 == 0) && ((uintptr_t)cx % 16 == 0) && ((uintptr_t)iy % 16 == 0) && ((uintptr_t)cy % 16 == 0),
               "rdst_bicubic_resize_f32: tap tables must be 16-byte aligned");
  if (B == 0) return RDST_OK;
  dim3 grid((Wd + 255) / 256, Hd, B);
  bicubic_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, ix, cx, iy, cy, B, Hs, Ws, Hd, Wd);
  RDST_CHECK_LAUNCH("rdst_bicubic_resize_f32");
  return RDST_OK;
}

extern "C" int rdst_sqdiff_sum_f64(const float* a, const float* b, double* out_zeroed, int B, int64_t n_per_image, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(a && b && out_zeroed && B >= 0 && B <= 65535 && n_per_image > 0, "rdst_sqdiff_sum_f64: bad arguments");
  if (B == 0) return RDST_OK;
  const int64_t want = (n_per_image + 256 * 8 - 1) / (256 * 8);
  dim3 grid((unsigned)(want < 1 ? 1 : (want > 592 ? 592 : want)), B);
  sqdiff_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, out_zeroed, n_per_image);
  RDST_CHECK_LAUNCH("rdst_sqdiff_sum_f64");
  return RDST_OK;
}

extern "C" int rdst_ssim_sum_f64(const float* a, const float* b, double* out_zeroed, int B, int H, int W, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(a && b && out_zeroed && B >= 0 && B <= 65535, "rdst_ssim_sum_f64: bad arguments");
  RDST_REQUIRE(H >= 7 && W >= 7, "rdst_ssim_sum_f64: images must be at least 7x7 (window size), got %dx%d", H, W);
  if (B == 0) return RDST_OK;
  dim3 grid((W - 6 + 31) / 32, (H - 6 + 7) / 8, B);
  ssim_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, out_zeroed, H, W);
  RDST_CHECK_LAUNCH("rdst_ssim_sum_f64");
  return RDST_OK;
}
