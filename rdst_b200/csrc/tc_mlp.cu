// Fused Swin MLP on tcgen05:   Y = X + fc2( GELU( fc1( LNhat(X) ) ) )          (bf16 operands, fp32 accumulate)
//
// Replaces norm2 + Mlp + residual of SwinTransformerBlock.forward (reference swin_transformer_sr.py:272, :23-29).
// One persistent CTA per SM keeps both weight matrices resident in shared memory as ready-made UMMA operand
// images and walks over 128-token tiles:
//   P1  coalesced load of the tile (8 rows x 64 B per warp instruction), LayerNorm statistics by 2 shuffles,
//       normalised bf16 rows written straight into the K-major A-operand image (conflict-free 128 B core matrices)
//   P2  fc1 as two N-halves of tcgen05.mma (accumulators in TMEM), each committed to its own mbarrier
//   P3  per half: tcgen05.ld -> +bias -> GELU -> bf16 -> second A-operand image; fc2's K-half is issued as soon as
//       its hidden half is staged, so the tensor pipe runs under the GELU of the other half
//   P5  tcgen05.ld of the fc2 accumulator -> +bias -> bf16 staging -> coalesced residual add and store
// Hidden activations never leave the SM.  LayerNorm gamma/beta are folded into fc1 by the host.
#include "common.cuh"
#include "umma.cuh"

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

// GELU(x) = x*Phi(x).  Fast form: Phi(x) ~ 0.5*(1+tanh(x*(a+b x^2+c x^4))) fitted to the exact erf form
// (max abs deviation 2.6e-5 on |x|<=8, i.e. far below bf16 resolution); EXACT uses erff.
template <bool EXACT>
__device__ __forceinline__ float gelu_fn(float x) {
  if (EXACT) return gelu_erf(x);
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float u = xc * xc;
  const float inner = xc * fmaf(u, fmaf(u, -3.53076214e-04f, 3.70152568e-02f), 7.97497252e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

template <int CP, int HP>
struct MlpCfg {
  static constexpr int NCH = CP / 8;                      // 16-byte chunks per activation row
  static constexpr int H0 = (HP / 2 + 15) / 16 * 16;      // fc1 N-half 0 (== fc2 K-half 0)
  static constexpr int H1 = HP - H0;
  static constexpr int W1_BYTES = HP * CP * 2;
  static constexpr int W2_BYTES = CP * HP * 2;
  static constexpr int A1_BYTES = 128 * CP * 2;
  static constexpr int A2_BYTES = 128 * HP * 2;
  static constexpr int PITCH = CP * 2 + 16;               // staging row pitch (bytes), conflict-free for 16 B accesses
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
  static constexpr int OFF_A1 = OFF_W2 + W2_BYTES;
  static constexpr int OFF_A2 = OFF_A1 + A1_BYTES;
  static constexpr int OFF_B1 = OFF_A2 + A2_BYTES;
  static constexpr int OFF_B2 = OFF_B1 + HP * 4;
  static constexpr int OFF_WT = OFF_B2 + CP * 4;           // fused DenseSTLayer tail: Linear(C -> 30, padded 32) image
  static constexpr int WT_BYTES = 32 * CP * 2;
  static constexpr int OFF_BT = OFF_WT + WT_BYTES;
  static constexpr int SMEM_PLAIN = OFF_WT;
  static constexpr int SMEM_TAIL = OFF_BT + 32 * 4;
  static constexpr int SMEM = SMEM_PLAIN;
  static constexpr int TM_FC1 = 0;                        // TMEM columns: fc1 accumulator [0,HP), fc2 at 256
  static constexpr int TM_FC2 = 256;
  static_assert(HP % 16 == 0 && CP % 32 == 0 && H1 % 16 == 0 && H1 > 0, "tile shape");
  static_assert(128 * PITCH <= A2_BYTES, "staging must fit in the dead A2 image");
  static_assert(HP <= 256 && CP <= 128, "TMEM budget");
};

// epilogue helper: 16 accumulator columns -> +bias -> (GELU) -> 16 bf16 = two 16-byte stores
template <bool GELU, bool EXACT>
__device__ __forceinline__ void epi16(uint32_t taddr, const float* __restrict__ bias, uint8_t* dst0, uint8_t* dst1) {
  uint32_t v[16];
  tmem_ld_x16(taddr, v);
  wait_ld();
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = __uint_as_float(v[2 * j]) + bias[2 * j];
    float b = __uint_as_float(v[2 * j + 1]) + bias[2 * j + 1];
    if (GELU) { a = gelu_fn<EXACT>(a); b = gelu_fn<EXACT>(b); }
    o[j] = pack_bf16x2(a, b);
  }
  *reinterpret_cast<uint4*>(dst0) = make_uint4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<uint4*>(dst1) = make_uint4(o[4], o[5], o[6], o[7]);
}

struct TailArgs {
  const uint8_t* wtimg;      // [CP/8][32][8] bf16, LN gamma folded
  const float* bt;           // [32]
  __nv_bfloat16* dense;      // &D[0][64 + 32 j]
  int64_t ldd;
  float scale;               // dense_scale
};

template <int CP, int HP, bool EXACT, bool TAIL>
__global__ void __launch_bounds__(256, 1)
stl_mlp_kernel(const __nv_bfloat16* __restrict__ X, int64_t ldx, __nv_bfloat16* __restrict__ Y, int64_t ldy,
               const uint8_t* __restrict__ w1img, const uint8_t* __restrict__ w2img,
               const float* __restrict__ b1, const float* __restrict__ b2, int64_t T, int creal, TailArgs ta) {
  using C = MlpCfg<CP, HP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[5];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sW1 = smem + C::OFF_W1;
  uint8_t* sW2 = smem + C::OFF_W2;
  uint8_t* sA1 = smem + C::OFF_A1;
  uint8_t* sA2 = smem + C::OFF_A2;
  float* sB1 = reinterpret_cast<float*>(smem + C::OFF_B1);
  float* sB2 = reinterpret_cast<float*>(smem + C::OFF_B2);

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  // resident weights (ready-made operand images) + biases
  for (int i = tid; i < (C::W1_BYTES + C::W2_BYTES) / 16; i += 256) {
    const uint8_t* src = i < C::W1_BYTES / 16 ? w1img + (size_t)i * 16 : w2img + (size_t)(i - C::W1_BYTES / 16) * 16;
    *reinterpret_cast<uint4*>(smem + (size_t)i * 16) = __ldg(reinterpret_cast<const uint4*>(src));
  }
  for (int i = tid; i < HP; i += 256) sB1[i] = b1[i];
  for (int i = tid; i < CP; i += 256) sB2[i] = b2[i];
  if (TAIL) {
    for (int i = tid; i < C::WT_BYTES / 16; i += 256)
      *reinterpret_cast<uint4*>(smem + C::OFF_WT + (size_t)i * 16) = __ldg(reinterpret_cast<const uint4*>(ta.wtimg) + i);
    for (int i = tid; i < 32; i += 256) reinterpret_cast<float*>(smem + C::OFF_BT)[i] = ta.bt[i];
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  const uint32_t aA1 = smem_u32(sA1), aA2 = smem_u32(sA2), aW1 = smem_u32(sW1), aW2 = smem_u32(sW2);
  const int row = tid & 127, half = tid >> 7;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float inv_c = 1.0f / (float)creal;
  const int64_t ntiles = (T + 127) / 128;
  uint32_t parity = 0;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, parity ^= 1) {
    const int64_t t0 = tile * 128;
    // ---------------- P1: load + LayerNorm -> A1 ----------------
#pragma unroll 1
    for (int g = warp; g < 16; g += 8) {
      const int r = g * 8 + (lane & 7);
      const int64_t t = t0 + r;
      uint4 raw[C::NCH / 4];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < C::NCH / 4; ++j) {
        const int c = (lane >> 3) + 4 * j;
        raw[j] = t < T ? __ldg(reinterpret_cast<const uint4*>(X + t * ldx) + c) : make_uint4(0, 0, 0, 0);
        const float2 f0 = unpack_bf16x2(raw[j].x), f1 = unpack_bf16x2(raw[j].y), f2 = unpack_bf16x2(raw[j].z),
                     f3 = unpack_bf16x2(raw[j].w);
        s += (f0.x + f0.y) + (f1.x + f1.y) + (f2.x + f2.y) + (f3.x + f3.y);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      const float mean = s * inv_c;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < C::NCH / 4; ++j) {
        const uint32_t w4[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w4[q]);
          ss += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
      }
      ss += __shfl_xor_sync(0xffffffffu, ss, 8);
      ss += __shfl_xor_sync(0xffffffffu, ss, 16);
      ss -= (float)(CP - creal) * mean * mean;                 // zero pads contributed mean^2 each
      const float rstd = rsqrtf(fmaxf(ss, 0.f) * inv_c + 1e-5f);
#pragma unroll
      for (int j = 0; j < C::NCH / 4; ++j) {
        const int c = (lane >> 3) + 4 * j;
        const uint32_t w4[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w4[q]);
          o[q] = pack_bf16x2((f.x - mean) * rstd, (f.y - mean) * rstd);
        }
        *reinterpret_cast<uint4*>(sA1 + c * 2048 + r * 16) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    // ---------------- P2: fc1 (two N-halves) ----------------
    if (tid == 0) {
      fence_after_sync();
      constexpr uint32_t id0 = make_idesc_bf16(128, C::H0, false, false);
      constexpr uint32_t id1 = make_idesc_bf16(128, C::H1, false, false);
#pragma unroll
      for (int ks = 0; ks < CP / 16; ++ks)
        mma_bf16_ss(tmem + C::TM_FC1, make_smem_desc(aA1 + ks * 4096, 2048, 128),
                    make_smem_desc(aW1 + ks * 2 * (HP * 16), HP * 16, 128), id0, ks > 0);
      commit(&bars[0]);
#pragma unroll
      for (int ks = 0; ks < CP / 16; ++ks)
        mma_bf16_ss(tmem + C::TM_FC1 + C::H0, make_smem_desc(aA1 + ks * 4096, 2048, 128),
                    make_smem_desc(aW1 + C::H0 * 16 + ks * 2 * (HP * 16), HP * 16, 128), id1, ks > 0);
      commit(&bars[1]);
    }
    // ---------------- P3: GELU epilogue per half, fc2 K-half issued behind it ----------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int hbase = h == 0 ? 0 : C::H0;
      const int hw = h == 0 ? C::H0 : C::H1;
      const int w0 = (hw / 2 + 15) / 16 * 16;                  // columns of this half taken by warpgroup 0
      const int cbeg = hbase + (half == 0 ? 0 : w0);
      const int cend = hbase + (half == 0 ? w0 : hw);
      mbar_wait(&bars[h], parity);
      fence_after_sync();
      for (int c0 = cbeg; c0 < cend; c0 += 16)
        epi16<true, EXACT>(lane_addr + C::TM_FC1 + c0, sB1 + c0, sA2 + (c0 / 8) * 2048 + row * 16,
                           sA2 + (c0 / 8 + 1) * 2048 + row * 16);
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        constexpr uint32_t id2 = make_idesc_bf16(128, CP, false, false);
        const int ks0 = hbase / 16, ks1 = (hbase + hw) / 16;
        for (int ks = ks0; ks < ks1; ++ks)
          mma_bf16_ss(tmem + C::TM_FC2, make_smem_desc(aA2 + ks * 4096, 2048, 128),
                      make_smem_desc(aW2 + ks * 2 * (CP * 16), CP * 16, 128), id2, ks > 0);
        commit(&bars[2 + h]);
      }
    }
    // ---------------- P5: fc2 epilogue -> staging -> coalesced residual add + store ----------------
    mbar_wait(&bars[2], parity);
    mbar_wait(&bars[3], parity);
    fence_after_sync();
    uint8_t* stg = sA2;                                       // A2 is dead once fc2 has completed
    {
      const int cbeg = half * (CP / 2), cend = cbeg + CP / 2;
      for (int c0 = cbeg; c0 < cend; c0 += 16)
        epi16<false, false>(lane_addr + C::TM_FC2 + c0, sB2 + c0, stg + row * C::PITCH + c0 * 2,
                            stg + row * C::PITCH + c0 * 2 + 16);
    }
    fence_before_sync();
    __syncthreads();
#pragma unroll 1
    for (int g = warp; g < 16; g += 8) {
      const int r = g * 8 + (lane & 7);
      const int64_t t = t0 + r;
      float yv[C::NCH / 4][8];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < C::NCH / 4; ++j) {
        const int c = (lane >> 3) + 4 * j;
        const uint4 m = *reinterpret_cast<const uint4*>(stg + r * C::PITCH + c * 16);
        const uint4 x = t < T ? __ldg(reinterpret_cast<const uint4*>(X + t * ldx) + c) : make_uint4(0, 0, 0, 0);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w}, xw[4] = {x.x, x.y, x.z, x.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a = unpack_bf16x2(mw[q]), b = unpack_bf16x2(xw[q]);
          yv[j][2 * q] = a.x + b.x;
          yv[j][2 * q + 1] = a.y + b.y;
          s += yv[j][2 * q] + yv[j][2 * q + 1];
          o[q] = pack_bf16x2(yv[j][2 * q], yv[j][2 * q + 1]);
        }
        if (!TAIL && t < T) *(reinterpret_cast<uint4*>(Y + t * ldy) + c) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (TAIL) {
        // LayerNorm of the block output (fp32, never rounded) -> A image for the growth projection
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        const float mean = s * inv_c;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < C::NCH / 4; ++j)
#pragma unroll
          for (int q = 0; q < 8; ++q) ss += (yv[j][q] - mean) * (yv[j][q] - mean);
        ss += __shfl_xor_sync(0xffffffffu, ss, 8);
        ss += __shfl_xor_sync(0xffffffffu, ss, 16);
        ss -= (float)(CP - creal) * mean * mean;
        const float rstd = rsqrtf(fmaxf(ss, 0.f) * inv_c + 1e-5f);
#pragma unroll
        for (int j = 0; j < C::NCH / 4; ++j) {
          const int c = (lane >> 3) + 4 * j;
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            o[q] = pack_bf16x2((yv[j][2 * q] - mean) * rstd, (yv[j][2 * q + 1] - mean) * rstd);
          *reinterpret_cast<uint4*>(sA1 + c * 2048 + r * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    if (TAIL) {
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        constexpr uint32_t idt = make_idesc_bf16(128, 32, false, false);
        const uint32_t aWT = smem_u32(smem + C::OFF_WT);
#pragma unroll
        for (int ks = 0; ks < CP / 16; ++ks)
          mma_bf16_ss(tmem + C::TM_FC1, make_smem_desc(aA1 + ks * 4096, 2048, 128),
                      make_smem_desc(aWT + ks * 2 * (32 * 16), 32 * 16, 128), idt, ks > 0);
        commit(&bars[4]);
      }
      mbar_wait(&bars[4], parity);
      fence_after_sync();
      if (half == 0) {
        const float* sBT = reinterpret_cast<const float*>(smem + C::OFF_BT);
        uint32_t v[32];
        tmem_ld_x32(lane_addr + C::TM_FC1, v);
        wait_ld();
        const int64_t t = t0 + row;
        if (t < T) {
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            o[j] = pack_bf16x2((__uint_as_float(v[2 * j]) + sBT[2 * j]) * ta.scale,
                               (__uint_as_float(v[2 * j + 1]) + sBT[2 * j + 1]) * ta.scale);
          uint4* dp = reinterpret_cast<uint4*>(ta.dense + t * ta.ldd);
#pragma unroll
          for (int q = 0; q < 4; ++q) dp[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
      fence_before_sync();
    }
    __syncthreads();        // staging (A2) and TMEM are reused by the next tile
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int CP, int HP>
static int launch_mlp(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w1, const void* w2, const float* b1,
                      const float* b2, int64_t T, int creal, int exact_gelu, const TailArgs* tail, int sms, cudaStream_t st) {
  using C = MlpCfg<CP, HP>;
  const int64_t ntiles = (T + 127) / 128;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  TailArgs ta{};
  int smem = C::SMEM_PLAIN;
  void (*k)(const __nv_bfloat16*, int64_t, __nv_bfloat16*, int64_t, const uint8_t*, const uint8_t*, const float*,
            const float*, int64_t, int, TailArgs);
  if (tail) {
    ta = *tail;
    smem = C::SMEM_TAIL;
    k = exact_gelu ? stl_mlp_kernel<CP, HP, true, true> : stl_mlp_kernel<CP, HP, false, true>;
  } else {
    k = exact_gelu ? stl_mlp_kernel<CP, HP, true, false> : stl_mlp_kernel<CP, HP, false, false>;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("rdst_stl_mlp_fwd_bf16: smem attr (%d B): %s", smem, cudaGetErrorString(e)); return RDST_E_CUDA; }
  k<<<grid, 256, smem, st>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, (const uint8_t*)w1,
                             (const uint8_t*)w2, b1, b2, T, creal, ta);
  return RDST_OK;
}

static int mlp_dispatch(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w1img, const void* w2img,
                        const float* b1, const float* b2, int64_t T, int C, int exact_gelu, const TailArgs* tail,
                        void* stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 60:  return launch_mlp<64, 128>(x, ldx, y, ldy, w1img, w2img, b1, b2, T, 60, exact_gelu, tail, sms, st);
    case 90:  return launch_mlp<96, 192>(x, ldx, y, ldy, w1img, w2img, b1, b2, T, 90, exact_gelu, tail, sms, st);
    case 120: return launch_mlp<128, 240>(x, ldx, y, ldy, w1img, w2img, b1, b2, T, 120, exact_gelu, tail, sms, st);
    default: set_error("rdst_stl_mlp_fwd_bf16: C=%d unsupported (60, 90, 120 with mlp_ratio 2)", C); return RDST_E_UNSUPPORTED;
  }
}

}  // namespace rdst

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
