// CUDA-core (fp32 FFMA) kernels of librdst_b200: the fp32 mode of the drop-in module and the
// bring-up / validation path of the bf16 mode.  One generic tiled GEMM serves nn.Linear (with fused
// LayerNorm prologue, GELU, residual epilogue) and the 3x3 convolutions (implicit GEMM over 9 taps,
// optional pixel-shuffle store).  See include/rdst_b200.h for the contract of each entry point.
#include "common.cuh"

namespace rdst {

// ------------------------------------------------------------------------------------------------
// generic GEMM:  Y = epi( A . W^T ),  A rows are tokens (linear) or gathered 3x3 neighbourhoods (conv)
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
  const void* x; int64_t ldx;
  const float* w; const float* bias;
  const void* r; int64_t ldr;
  void* y; int64_t ldy;
  int64_t T; int K; int N;
  int ln_creal; int act; float out_scale;
  int B, H, W, Cin, shuffle;
};

constexpr int BM = 64, BN = 64, BK = 16;

template <typename TA, bool CONV>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float s_mean[BM], s_rstd[BM];

  const TA* __restrict__ X = reinterpret_cast<const TA*>(a.x);
  const TA* __restrict__ R = reinterpret_cast<const TA*>(a.r);
  TA* __restrict__ Y = reinterpret_cast<TA*>(a.y);
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ktot = CONV ? 9 * a.Cin : a.K;

  // ---- LayerNorm statistics of this block's rows (two-pass, pads inside [0,K) are zero) ----
  if (!CONV && a.ln_creal > 0) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < BM; r += 8) {
      const int64_t t = m0 + r;
      float mean = 0.f, rstd = 0.f;
      if (t < a.T) {
        const TA* row = X + t * a.ldx;
        float s = 0.f;
        for (int k = lane; k < a.K; k += 32) s += ld_act(row + k);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        mean = s / (float)a.ln_creal;
        float ss = 0.f;
        for (int k = lane; k < a.K; k += 32) { float d = ld_act(row + k) - mean; ss += d * d; }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        ss -= (float)(a.K - a.ln_creal) * mean * mean;      // remove the zero pads' contribution
        rstd = rsqrtf(fmaxf(ss, 0.f) / (float)a.ln_creal + 1e-5f);
      }
      if (lane == 0) { s_mean[r] = mean; s_rstd[r] = rstd; }
    }
    __syncthreads();
  }

  // rows this thread stages into As: r = (tid>>4) + 16*i
  int rb[4], ry[4], rx[4];
  if (CONV) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t t = m0 + (tid >> 4) + 16 * i;
      if (t < a.T) {
        int hw = a.H * a.W;
        rb[i] = (int)(t / hw);
        int rem = (int)(t % hw);
        ry[i] = rem / a.W; rx[i] = rem % a.W;
      } else { rb[i] = -1; ry[i] = 0; rx[i] = 0; }
    }
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tx = tid & 15, ty = tid >> 4;
  const int kk = tid & 15;

  for (int kt = 0; kt < Ktot; kt += BK) {
    const int gk = kt + kk;
    // ---- A tile ----
    if (CONV) {
      const int tap = kt / a.Cin;              // Cin % 16 == 0 => a K-chunk never straddles taps
      const int c = gk - tap * a.Cin;
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        int yy = ry[i] + dy, xx = rx[i] + dx;
        if (rb[i] >= 0 && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W)
          v = ld_act(X + ((int64_t)(rb[i] * a.H + yy) * a.W + xx) * a.ldx + c);
        As[kk][(tid >> 4) + 16 * i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (tid >> 4) + 16 * i;
        const int64_t t = m0 + r;
        float v = 0.f;
        if (t < a.T && gk < a.K) {
          v = ld_act(X + t * a.ldx + gk);
          if (a.ln_creal > 0) v = (v - s_mean[r]) * s_rstd[r];
        }
        As[kk][r] = v;
      }
    }
    // ---- W tile ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nl = (tid >> 4) + 16 * i;
      const int n = n0 + nl;
      Bs[kk][nl] = (n < a.N && gk < Ktot) ? __ldg(a.w + (int64_t)n * Ktot + gk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t t = m0 + ty * 4 + i;
    if (t >= a.T) continue;
    int64_t obase = t;
    int ob = 0, oy = 0, ox = 0;
    if (CONV && a.shuffle) {
      int hw = a.H * a.W;
      ob = (int)(t / hw);
      int rem = (int)(t % hw);
      oy = rem / a.W; ox = rem % a.W;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j] + (a.bias ? __ldg(a.bias + n) : 0.f);
      if (a.act == 1) v = gelu_erf(v);
      else if (a.act == 2) v = v > 0.f ? v : 0.2f * v;      // nn.LeakyReLU(0.2) of the '3conv' fusion (rdst_variations.py:424-426)
      v *= a.out_scale;
      if (CONV && a.shuffle) {
        const int G = a.N >> 2;
        const int s = n / G, c = n - s * G;
        const int64_t ot = ((int64_t)(ob * 2 * a.H + 2 * oy + (s >> 1))) * (2 * a.W) + 2 * ox + (s & 1);
        st_act(Y + ot * a.ldy + c, v);
      } else {
        if (R) v += ld_act(R + obase * a.ldr + n);
        st_act(Y + obase * a.ldy + n, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// window attention core: one block per (window, head), one thread per query token
// ------------------------------------------------------------------------------------------------
template <typename TA, int HD>
__global__ void __launch_bounds__(64) window_attention_simt_kernel(
    const TA* __restrict__ qkv, int64_t ldq, const float* __restrict__ table, TA* __restrict__ out, int64_t ldo,
    int H, int W, int C, int heads, int shift) {
  // K / V rows padded to a multiple of 4 floats (pads zero): every thread reads the same row (broadcast), so one
  // 16-byte shared-memory load feeds four FMAs instead of one
  constexpr int HP = (HD + 3) / 4 * 4;
  __shared__ __align__(16) float sk[64][HP];
  __shared__ __align__(16) float sv[64][HP];
  __shared__ float stab[225];
  __shared__ int sreg[64];

  const int wi = blockIdx.x, h = blockIdx.y, i = threadIdx.x;
  const int nwx = W >> 3, nw_img = (H >> 3) * nwx;
  const int b = wi / nw_img, wl = wi - b * nw_img;
  const int wy = wl / nwx, wx = wl - wy * nwx;
  const int iy = i >> 3, ix = i & 7;
  const int hs = wy * 8 + iy, ws = wx * 8 + ix;                 // coordinates on the shifted frame
  int hh = hs + shift; if (hh >= H) hh -= H;                    // shifted[h'] = x[(h'+s) mod H]
  int ww = ws + shift; if (ww >= W) ww -= W;
  const int64_t t = ((int64_t)b * H + hh) * W + ww;

  for (int e = i; e < 225; e += 64) stab[e] = table[e * heads + h];
  int reg = 0;
  if (shift > 0) {
    const int rh = hs < H - 8 ? 0 : (hs < H - shift ? 1 : 2);
    const int rw = ws < W - 8 ? 0 : (ws < W - shift ? 1 : 2);
    reg = rh * 3 + rw;
  }
  sreg[i] = reg;

  const TA* row = qkv + t * ldq + h * HD;
  float q[HP];
#pragma unroll
  for (int d = 0; d < HP; ++d) {
    q[d] = d < HD ? ld_act(row + d) : 0.f;
    sk[i][d] = d < HD ? ld_act(row + C + d) : 0.f;
    sv[i][d] = d < HD ? ld_act(row + 2 * C + d) : 0.f;
  }
  __syncthreads();

  float s[64];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < HP; d += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(&sk[j][d]);
      acc = fmaf(q[d], kv.x, acc); acc = fmaf(q[d + 1], kv.y, acc); acc = fmaf(q[d + 2], kv.z, acc); acc = fmaf(q[d + 3], kv.w, acc);
    }
    const int jy = j >> 3, jx = j & 7;
    acc += stab[(iy - jy + 7) * 15 + (ix - jx + 7)];
    if (shift > 0 && sreg[j] != reg) acc += -100.0f;
    s[j] = acc;
    mx = fmaxf(mx, acc);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.0f / sum;
  float o[HP];
#pragma unroll
  for (int d = 0; d < HP; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const float p = s[j] * inv;
#pragma unroll
    for (int d = 0; d < HP; d += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(&sv[j][d]);
      o[d] = fmaf(p, vv.x, o[d]); o[d + 1] = fmaf(p, vv.y, o[d + 1]); o[d + 2] = fmaf(p, vv.z, o[d + 2]); o[d + 3] = fmaf(p, vv.w, o[d + 3]);
    }
  }
  TA* orow = out + t * ldo + h * HD;
#pragma unroll
  for (int d = 0; d < HD; ++d) st_act(orow + d, o[d]);
}

// ------------------------------------------------------------------------------------------------
// head: conv3x3 (1 -> 60) + LayerNorm(60)
// ------------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(128) head_kernel(const float* __restrict__ img, float in_scale, float in_bias,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   TA* __restrict__ feat0, int64_t ldf, TA* __restrict__ dense,
                                                   int64_t ldd, int B, int H, int W) {
  __shared__ float sw[60 * 9], sb[60], sg[60], sbe[60];
  for (int e = threadIdx.x; e < 540; e += blockDim.x) sw[e] = w[e];
  for (int e = threadIdx.x; e < 60; e += blockDim.x) { sb[e] = bias[e]; sg[e] = gamma[e]; sbe[e] = beta[e]; }
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t T = (int64_t)B * H * W;
  if (t >= T) return;
  const int hw = H * W;
  const int b = (int)(t / hw), rem = (int)(t % hw), y = rem / W, x = rem % W;
  float px[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    // zero padding applies to the sub_mean output, i.e. pad value is 0 (not in_bias)
    px[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[((int64_t)b * H + yy) * W + xx] * in_scale + in_bias : 0.f;
  }
  float f[60];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 60; ++c) {
    float acc = sb[c];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc = fmaf(px[k], sw[c * 9 + k], acc);
    f[c] = acc; s += acc;
  }
  const float mean = s * (1.0f / 60.0f);
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 60; ++c) { const float d = f[c] - mean; ss += d * d; }
  const float rstd = rsqrtf(ss * (1.0f / 60.0f) + 1e-5f);
  TA* fr = feat0 + t * ldf;
  TA* dr = dense + t * ldd;
  if constexpr (sizeof(TA) == 2) {
    // bf16 storage: 16-byte vector stores (8 channels each) instead of 128 scalar stores per token
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      uint32_t pf[4], pd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c8 * 8 + 2 * q;
        const float f0 = c < 60 ? f[c] : 0.f, f1 = c + 1 < 60 ? f[c + 1] : 0.f;
        const float d0 = c < 60 ? (f[c] - mean) * rstd * sg[c] + sbe[c] : 0.f;
        const float d1 = c + 1 < 60 ? (f[c + 1] - mean) * rstd * sg[c + 1] + sbe[c + 1] : 0.f;
        __nv_bfloat162 a = __floats2bfloat162_rn(f0, f1), b = __floats2bfloat162_rn(d0, d1);
        pf[q] = *reinterpret_cast<uint32_t*>(&a);
        pd[q] = *reinterpret_cast<uint32_t*>(&b);
      }
      reinterpret_cast<uint4*>(fr)[c8] = make_uint4(pf[0], pf[1], pf[2], pf[3]);
      reinterpret_cast<uint4*>(dr)[c8] = make_uint4(pd[0], pd[1], pd[2], pd[3]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 60; ++c) {
      st_act(fr + c, f[c]);
      st_act(dr + c, (f[c] - mean) * rstd * sg[c] + sbe[c]);
    }
#pragma unroll
    for (int c = 60; c < 64; ++c) { st_act(fr + c, 0.f); st_act(dr + c, 0.f); }
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone LayerNorm (warp per token)
// ------------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256) layernorm_kernel(const TA* __restrict__ x, int64_t ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        TA* __restrict__ y, int64_t ldy, int64_t T, int creal,
                                                        float out_scale) {
  const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const TA* row = x + t * ldx;
  float s = 0.f;
  for (int k = lane; k < creal; k += 32) s += ld_act(row + k);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)creal;
  float ss = 0.f;
  for (int k = lane; k < creal; k += 32) { const float d = ld_act(row + k) - mean; ss += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)creal + 1e-5f);
  TA* yr = y + t * ldy;
  for (int k = lane; k < creal; k += 32)
    st_act(yr + k, ((ld_act(row + k) - mean) * rstd * gamma[k] + beta[k]) * out_scale);
}

// bf16, <= 64 channels, 16-byte aligned rows: eight lanes per token, one 16-byte load and store per lane (the
// generic kernel's scalar 2-byte accesses ran at a tenth of the HBM rate)
__global__ void __launch_bounds__(256) layernorm_bf16_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 __nv_bfloat16* __restrict__ y, int64_t ldy, int64_t T,
                                                                 int creal, float out_scale) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const int64_t t = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
  const bool ok = t < T;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (ok) v = __ldg(reinterpret_cast<const uint4*>(x + t * ldx) + sub);
  const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
  float f[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w4[q]);
    const float2 p = __bfloat1622float2(h);
    f[2 * q] = p.x; f[2 * q + 1] = p.y;
  }
  const int k0 = sub * 8;
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += (k0 + e < creal) ? f[e] : 0.f;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float mean = s / (float)creal;
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { const float d = (k0 + e < creal) ? f[e] - mean : 0.f; ss += d * d; }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  const float rstd = rsqrtf(ss / (float)creal + 1e-5f);
  if (!ok || k0 >= creal) return;
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float r[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k0 + 2 * q + h;
      r[h] = k < creal ? ((f[2 * q + h] - mean) * rstd * gamma[k] + beta[k]) * out_scale : 0.f;
    }
    __nv_bfloat162 pk = __floats2bfloat162_rn(r[0], r[1]);
    o[q] = *reinterpret_cast<uint32_t*>(&pk);
  }
  // channels beyond creal inside the last chunk keep their previous value in the generic kernel; here they are pads
  // of a padded [T][64] buffer and are written as zero only when the chunk is partially real
  *(reinterpret_cast<uint4*>(y + t * ldy) + sub) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------
// last conv: 3x3, Cin -> 1, fp32 NCHW image out (warp per output pixel)
// ------------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256) last_conv_kernel(const TA* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                        float bias, float out_scale, float out_bias,
                                                        float* __restrict__ img, int B, int H, int W, int Cin) {
  extern __shared__ float sw[];   // [9][Cin]
  for (int e = threadIdx.x; e < 9 * Cin; e += blockDim.x) sw[e] = w[e];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t T = (int64_t)B * H * W;
  if (p >= T) return;
  const int hw = H * W;
  const int b = (int)(p / hw), rem = (int)(p % hw), y = rem / W, xx0 = rem % W;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = xx0 + k % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const TA* row = x + (((int64_t)b * H + yy) * W + xx) * ldx;
    for (int c = lane; c < Cin; c += 32) acc = fmaf(ld_act(row + c), sw[k * Cin + c], acc);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) img[p] = (acc + bias) * out_scale + out_bias;
}

}  // namespace rdst

// =================================================================================================
// C ABI
// =================================================================================================
using namespace rdst;

extern "C" int rdst_linear_fwd(const void* x, int64_t ldx, const float* w, const float* bias, const void* resid,
                               int64_t ldr, void* y, int64_t ldy, int64_t T, int K, int N, int ln_creal, int act,
                               float out_scale, int dtype, void* stream) {
  RDST_REQUIRE(x && w && y, "rdst_linear_fwd: null pointer");
  RDST_REQUIRE(T >= 0 && K > 0 && N > 0 && ldx >= K && ldy >= N, "rdst_linear_fwd: bad shape T=%lld K=%d N=%d ldx=%lld ldy=%lld",
               (long long)T, K, N, (long long)ldx, (long long)ldy);
  RDST_REQUIRE(ln_creal >= 0 && ln_creal <= K, "rdst_linear_fwd: ln_creal=%d out of range for K=%d", ln_creal, K);
  RDST_REQUIRE(act >= 0 && act <= 2, "rdst_linear_fwd: act must be 0 (none), 1 (GELU) or 2 (LeakyReLU 0.2)");
  RDST_REQUIRE(dtype == RDST_F32 || dtype == RDST_BF16, "rdst_linear_fwd: bad dtype %d", dtype);
  if (T == 0) return RDST_OK;
  GemmArgs a{};
  a.x = x; a.ldx = ldx; a.w = w; a.bias = bias; a.r = resid; a.ldr = ldr; a.y = y; a.ldy = ldy;
  a.T = T; a.K = K; a.N = N; a.ln_creal = ln_creal; a.act = act; a.out_scale = out_scale;
  dim3 grid((unsigned)((T + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  if (dtype == RDST_F32) gemm_simt_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else gemm_simt_kernel<__nv_bfloat16, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_linear_fwd");
  return RDST_OK;
}

extern "C" int rdst_conv3x3_fwd(const void* x, int64_t ldx, const float* w, const float* bias, const void* resid,
                                int64_t ldr, void* y, int64_t ldy, int B, int H, int W, int Cin, int N,
                                float out_scale, int shuffle, int dtype, void* stream) {
  RDST_REQUIRE(x && w && y, "rdst_conv3x3_fwd: null pointer");
  RDST_REQUIRE(B >= 0 && H > 0 && W > 0 && Cin > 0 && N > 0, "rdst_conv3x3_fwd: bad shape");
  RDST_REQUIRE(Cin % 16 == 0 && ldx >= Cin, "rdst_conv3x3_fwd: Cin (%d) must be a multiple of 16 and <= ldx", Cin);
  RDST_REQUIRE(shuffle == 0 || (shuffle == 2 && N % 4 == 0 && resid == nullptr),
               "rdst_conv3x3_fwd: shuffle must be 0 or 2 (N %% 4 == 0, no residual)");
  RDST_REQUIRE(dtype == RDST_F32 || dtype == RDST_BF16, "rdst_conv3x3_fwd: bad dtype %d", dtype);
  if (B == 0) return RDST_OK;
  GemmArgs a{};
  a.x = x; a.ldx = ldx; a.w = w; a.bias = bias; a.r = resid; a.ldr = ldr; a.y = y; a.ldy = ldy;
  a.T = (int64_t)B * H * W; a.K = 9 * Cin; a.N = N; a.out_scale = out_scale;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.shuffle = shuffle;
  dim3 grid((unsigned)((a.T + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  if (dtype == RDST_F32) gemm_simt_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else gemm_simt_kernel<__nv_bfloat16, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_conv3x3_fwd");
  return RDST_OK;
}

extern "C" int rdst_conv3x3_act_fwd(const void* x, int64_t ldx, const float* w, const float* bias, void* y, int64_t ldy,
                                    int B, int H, int W, int Cin, int N, int act, int dtype, void* stream) {
  RDST_REQUIRE(x && w && y, "rdst_conv3x3_act_fwd: null pointer");
  RDST_REQUIRE(B >= 0 && H > 0 && W > 0 && Cin > 0 && N > 0, "rdst_conv3x3_act_fwd: bad shape");
  RDST_REQUIRE(Cin % 16 == 0 && ldx >= Cin && ldy >= N, "rdst_conv3x3_act_fwd: Cin (%d) must be a multiple of 16 and <= ldx", Cin);
  RDST_REQUIRE(act >= 0 && act <= 2, "rdst_conv3x3_act_fwd: act must be 0 (none), 1 (GELU) or 2 (LeakyReLU 0.2)");
  RDST_REQUIRE(dtype == RDST_F32 || dtype == RDST_BF16, "rdst_conv3x3_act_fwd: bad dtype %d", dtype);
  if (B == 0) return RDST_OK;
  GemmArgs a{};
  a.x = x; a.ldx = ldx; a.w = w; a.bias = bias; a.r = nullptr; a.ldr = 0; a.y = y; a.ldy = ldy;
  a.T = (int64_t)B * H * W; a.K = 9 * Cin; a.N = N; a.out_scale = 1.0f; a.act = act;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.shuffle = 0;
  dim3 grid((unsigned)((a.T + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  if (dtype == RDST_F32) gemm_simt_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else gemm_simt_kernel<__nv_bfloat16, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  RDST_CHECK_LAUNCH("rdst_conv3x3_act_fwd");
  return RDST_OK;
}

template <typename TA>
static int launch_wattn(const void* qkv, int64_t ldq, const float* table, void* out, int64_t ldo, int B, int H, int W,
                        int C, int heads, int shift, cudaStream_t st) {
  dim3 grid((unsigned)(B * (H / 8) * (W / 8)), (unsigned)heads);
  const TA* q = reinterpret_cast<const TA*>(qkv);
  TA* o = reinterpret_cast<TA*>(out);
  switch (C / heads) {
    case 5:  window_attention_simt_kernel<TA, 5><<<grid, 64, 0, st>>>(q, ldq, table, o, ldo, H, W, C, heads, shift); break;
    case 10: window_attention_simt_kernel<TA, 10><<<grid, 64, 0, st>>>(q, ldq, table, o, ldo, H, W, C, heads, shift); break;
    case 15: window_attention_simt_kernel<TA, 15><<<grid, 64, 0, st>>>(q, ldq, table, o, ldo, H, W, C, heads, shift); break;
    case 20: window_attention_simt_kernel<TA, 20><<<grid, 64, 0, st>>>(q, ldq, table, o, ldo, H, W, C, heads, shift); break;
    case 30: window_attention_simt_kernel<TA, 30><<<grid, 64, 0, st>>>(q, ldq, table, o, ldo, H, W, C, heads, shift); break;
    default: set_error("rdst_window_attention_fwd: head_dim %d unsupported (5,10,15,20,30)", C / heads); return RDST_E_UNSUPPORTED;
  }
  return RDST_OK;
}

extern "C" int rdst_window_attention_fwd(const void* qkv, int64_t ldq, const float* table, void* out, int64_t ldo,
                                         int B, int H, int W, int C, int heads, int shift, int dtype, void* stream) {
  RDST_REQUIRE(qkv && table && out, "rdst_window_attention_fwd: null pointer");
  RDST_REQUIRE(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0,
               "rdst_window_attention_fwd: H=%d W=%d must be positive multiples of the window size 8", H, W);
  RDST_REQUIRE(heads > 0 && C % heads == 0 && ldq >= 3 * C && ldo >= C, "rdst_window_attention_fwd: bad C/heads/ld");
  RDST_REQUIRE(shift == 0 || shift == 4, "rdst_window_attention_fwd: shift must be 0 or 4");
  RDST_REQUIRE(dtype == RDST_F32 || dtype == RDST_BF16, "rdst_window_attention_fwd: bad dtype %d", dtype);
  if (B == 0) return RDST_OK;
  // reference: when min(H,W) <= window the block runs unshifted (swin_transformer_sr.py:188-191) -- the caller decides.
  int rc = dtype == RDST_F32 ? launch_wattn<float>(qkv, ldq, table, out, ldo, B, H, W, C, heads, shift, (cudaStream_t)stream)
                             : launch_wattn<__nv_bfloat16>(qkv, ldq, table, out, ldo, B, H, W, C, heads, shift, (cudaStream_t)stream);
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_window_attention_fwd");
  return RDST_OK;
}

extern "C" int rdst_head_fwd(const float* img, float in_scale, float in_bias, const float* w, const float* bias,
                             const float* gamma, const float* beta, void* feat0, int64_t ldf, void* dense, int64_t ldd,
                             int B, int H, int W, int dtype, void* stream) {
  RDST_REQUIRE(img && w && bias && gamma && beta && feat0 && dense, "rdst_head_fwd: null pointer");
  RDST_REQUIRE(B >= 0 && H > 0 && W > 0 && ldf >= 64 && ldd >= 64, "rdst_head_fwd: bad shape");
  RDST_REQUIRE(dtype == RDST_F32 || dtype == RDST_BF16, "rdst_head_fwd: bad dtype %d", dtype);
  if (B == 0) return RDST_OK;
  const int64_t T = (int64_t)B * H * W;
  const unsigned grid = (unsigned)((T + 127) / 128);
  if (dtype == RDST_F32)
    head_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(img, in_scale, in_bias, w, bias, gamma, beta,
                                                               (float*)feat0, ldf, (float*)dense, ldd, B, H, W);
  else
    head_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(img, in_scale, in_bias, w, bias, gamma, beta,
                                                                       (__nv_bfloat16*)feat0, ldf, (__nv_bfloat16*)dense, ldd, B, H, W);
  RDST_CHECK_LAUNCH("rdst_head_fwd");
  return RDST_OK;
}

extern "C" int rdst_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y,
                                  int64_t ldy, int64_t T, int creal, float out_scale, int dtype, void* stream) {
  RDST_REQUIRE(x && gamma && beta && y, "rdst_layernorm_fwd: null pointer");
  RDST_REQUIRE(T >= 0 && creal > 0 && ldx >= creal && ldy >= creal, "rdst_layernorm_fwd: bad shape");
  RDST_REQUIRE(dtype == RDST_F32 || dtype == RDST_BF16, "rdst_layernorm_fwd: bad dtype %d", dtype);
  if (T == 0) return RDST_OK;
  const unsigned grid = (unsigned)((T + 7) / 8);
  if (dtype == RDST_F32)
    layernorm_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ldx, gamma, beta, (float*)y, ldy, T, creal, out_scale);
  else if (creal <= 64 && ldx % 8 == 0 && ldy % 8 == 0 && ldy >= (creal + 7) / 8 * 8 && ((uintptr_t)x % 16 == 0) &&
           ((uintptr_t)y % 16 == 0))
    layernorm_bf16_vec_kernel<<<(unsigned)((T + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, ldx, gamma, beta, (__nv_bfloat16*)y, ldy, T, creal, out_scale);
  else
    layernorm_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, gamma, beta,
                                                                            (__nv_bfloat16*)y, ldy, T, creal, out_scale);
  RDST_CHECK_LAUNCH("rdst_layernorm_fwd");
  return RDST_OK;
}

extern "C" int rdst_last_conv_fwd(const void* x, int64_t ldx, const float* w, float bias, float out_scale,
                                  float out_bias, float* img, int B, int H, int W, int Cin, int dtype, void* stream) {
  RDST_REQUIRE(x && w && img, "rdst_last_conv_fwd: null pointer");
  RDST_REQUIRE(B >= 0 && H > 0 && W > 0 && Cin > 0 && Cin <= 1024 && ldx >= Cin, "rdst_last_conv_fwd: bad shape");
  RDST_REQUIRE(dtype == RDST_F32 || dtype == RDST_BF16, "rdst_last_conv_fwd: bad dtype %d", dtype);
  if (B == 0) return RDST_OK;
  const int64_t T = (int64_t)B * H * W;
  const unsigned grid = (unsigned)((T + 7) / 8);
  const size_t smem = (size_t)9 * Cin * sizeof(float);
  if (dtype == RDST_F32)
    last_conv_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>((const float*)x, ldx, w, bias, out_scale, out_bias, img, B, H, W, Cin);
  else
    last_conv_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, w, bias, out_scale,
                                                                               out_bias, img, B, H, W, Cin);
  RDST_CHECK_LAUNCH("rdst_last_conv_fwd");
  return RDST_OK;
}
