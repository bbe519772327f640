// Fused shifted-window attention on tcgen05:
//     Y = X + proj( softmax( q k^T + rel_pos_bias + shift_mask ) v ),   [q|k|v] = qkv( LNhat(X) )
// (bf16 operands, fp32 accumulate in TMEM).  Replaces norm1 + roll + window_partition + WindowAttention +
// window_reverse + roll + residual of SwinTransformerBlock.forward (reference swin_transformer_sr.py:239-271,
// :110-141, :211-232).
//
// One persistent CTA per SM keeps the qkv / proj weights resident in shared memory as UMMA operand images and
// walks over tiles of 128 tokens = two 8x8 windows stacked on the 128 TMEM lanes:
//   P1   gather of the two windows by index arithmetic (cyclic shift and window partition never materialise),
//        LayerNorm statistics by 2 shuffles, normalised bf16 rows -> K-major A image
//   per head h (6):
//        qkv_h = A . Wqkv_h^T            tcgen05.mma M128 N{32,48,64} K=Cp        (issued two heads ahead)
//        drain: +bias -> bf16 -> Q (A image), K (B image, K-major), V (B image, MN-major)
//        S = Q K^T                        M128 N128 K{16,32}; the two windows sit on the diagonal 64x64 blocks
//        softmax on the diagonal block: +bias(table) +mask(closed form) -> exp2 -> bf16 P (row sums kept in regs)
//        O[:, h] = P V                    M128 N{16,32} K128 (off-diagonal P blocks are permanent zeros)
//   O / rowsum -> bf16 A image, proj = O . Wproj^T (M128 N=Cp), + bias -> staging -> coalesced residual add/store
// The (nW,6,64,64) score tensor of the reference never exists outside TMEM.
#include "common.cuh"
#include "umma.cuh"

namespace rdst {
using namespace umma;

__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 up2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int C_>
struct AttnCfg {
  static constexpr int C = C_;
  static constexpr int CP = C == 60 ? 64 : (C == 90 ? 96 : 128);
  static constexpr int HD = C / 6;
  static constexpr int NH = (3 * HD + 15) / 16 * 16;        // qkv columns of one head (padded): 32 / 48 / 64
  static constexpr int HDP = HD <= 16 ? 16 : 32;            // K of S = Q K^T
  static constexpr int HDV = HD <= 16 ? 16 : 32;            // N of O = P V
  static constexpr int HDO = HD == 20 ? 20 : 16;            // stride of one head inside the proj A operand
  static constexpr int KPROJ = (6 * HDO + 15) / 16 * 16;    // 96 / 96 / 128
  static constexpr int NCH = CP / 8;
  static constexpr int WQKV_BYTES = 6 * NH * CP * 2;
  static constexpr int WPROJ_BYTES = CP * KPROJ * 2;
  static constexpr int A1_BYTES = 128 * CP * 2;
  static constexpr int AO_BYTES = 128 * KPROJ * 2;
  static constexpr bool SWZ = (CP == 128);                  // staging: XOR-swizzled 256 B rows, else padded rows
  static constexpr int PITCH = SWZ ? CP * 2 : CP * 2 + 16;
  static constexpr int STG_BYTES = 128 * PITCH;
  static constexpr int AREG0 = A1_BYTES > AO_BYTES ? A1_BYTES : AO_BYTES;
  static constexpr int AREG_BYTES = AREG0 > STG_BYTES ? AREG0 : STG_BYTES;       // A1, later Ao, later staging
  static constexpr int P_BYTES = 128 * 128 * 2;
  static constexpr int AQ_BYTES = 128 * HDP * 2;
  static constexpr int BV_BYTES = 128 * HDV * 2;
  static constexpr int OFF_WQKV = 0;
  static constexpr int OFF_WPROJ = OFF_WQKV + WQKV_BYTES;
  static constexpr int OFF_A = OFF_WPROJ + WPROJ_BYTES;
  static constexpr int OFF_P = OFF_A + AREG_BYTES;
  static constexpr int OFF_AQ = OFF_P + P_BYTES;
  static constexpr int OFF_BK = OFF_AQ + AQ_BYTES;
  static constexpr int OFF_BV = OFF_BK + AQ_BYTES;
  static constexpr int OFF_TAB = OFF_BV + BV_BYTES;
  static constexpr int OFF_BQKV = OFF_TAB + 6 * 225 * 4;
  static constexpr int OFF_BPROJ = OFF_BQKV + 6 * NH * 4;
  static constexpr int OFF_SREG = OFF_BPROJ + CP * 4;
  static constexpr int OFF_SRED = OFF_SREG + 128 * 4;
  static constexpr int SMEM = OFF_SRED + 2 * 128 * 4;
  // TMEM columns
  static constexpr int TM_QKV0 = 0, TM_QKV1 = 64, TM_S = 128, TM_O = 256, TM_PROJ = 128;
  static_assert(2 * AQ_BYTES >= 2 * 6 * 128 * 4, "row-sum exchange must fit in the dead Q/K images");
  static_assert(HDP == 16 || 2 * 6 * 128 * 4 <= 3 * 2048, "row-sum exchange must not touch the zero K-pad chunk of Q");
  static_assert(SMEM <= 232448, "shared memory budget");
};

// load NCOL (multiple of 8) consecutive accumulator columns of this thread's TMEM lane
template <int NCOL>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, float (&f)[NCOL]) {
#pragma unroll
  for (int c = 0; c < NCOL; c += 8) {
    uint32_t v[8];
    tmem_ld_x8(taddr + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[c + j] = __uint_as_float(v[j]);
  }
  wait_ld();
}

struct WinGeom {
  int H, W, shift, nwx, nw_img;
  int64_t nwt;
};

// token index (in the un-shifted, raster-ordered activation) of position (iy,ix) of window `win`; region id for the mask
__device__ __forceinline__ int64_t win_token(const WinGeom& g, int64_t win, int iy, int ix, int& region, bool& edge) {
  const int b = (int)(win / g.nw_img);
  const int wl = (int)(win - (int64_t)b * g.nw_img);
  const int wy = wl / g.nwx, wx = wl - wy * g.nwx;
  const int hs = wy * 8 + iy, ws = wx * 8 + ix;             // coordinates on the shifted frame
  int hh = hs + g.shift; if (hh >= g.H) hh -= g.H;          // shifted[h'] = x[(h'+s) mod H]   (:245)
  int ww = ws + g.shift; if (ww >= g.W) ww -= g.W;
  const int rh = hs < g.H - 8 ? 0 : (hs < g.H - g.shift ? 1 : 2);
  const int rw = ws < g.W - 8 ? 0 : (ws < g.W - g.shift ? 1 : 2);
  region = rh * 3 + rw;
  edge = g.shift > 0 && (wy == g.H / 8 - 1 || wx == g.nwx - 1);
  return ((int64_t)b * g.H + hh) * g.W + ww;
}

template <int C_>
__global__ void __launch_bounds__(256, 1)
stl_attn_kernel(const __nv_bfloat16* __restrict__ X, int64_t ldx, __nv_bfloat16* __restrict__ Y, int64_t ldy,
                const uint8_t* __restrict__ wqkv_img, const uint8_t* __restrict__ wproj_img,
                const float* __restrict__ bqkv, const float* __restrict__ bproj, const float* __restrict__ table,
                WinGeom geo, float mask_val) {
  using K = AttnCfg<C_>;
  constexpr int CP = K::CP, HD = K::HD, NH = K::NH;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[5];          // 0,1: qkv ping/pong  2: S  3: PV  4: proj
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem + K::OFF_A;
  uint8_t* sP = smem + K::OFF_P;
  uint8_t* sAq = smem + K::OFF_AQ;
  uint8_t* sBk = smem + K::OFF_BK;
  uint8_t* sBv = smem + K::OFF_BV;
  float* sTab = reinterpret_cast<float*>(smem + K::OFF_TAB);
  float* sBqkv = reinterpret_cast<float*>(smem + K::OFF_BQKV);
  float* sBproj = reinterpret_cast<float*>(smem + K::OFF_BPROJ);
  int* sReg = reinterpret_cast<int*>(smem + K::OFF_SREG);
  float* sRed = reinterpret_cast<float*>(smem + K::OFF_SRED);
  float* sSum = reinterpret_cast<float*>(sAq);              // [2][6][128] exchange of row sums (Aq/Bk dead by then)

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < (K::WQKV_BYTES + K::WPROJ_BYTES) / 16; i += 256) {
    const uint8_t* src = i < K::WQKV_BYTES / 16 ? wqkv_img + (size_t)i * 16
                                                : wproj_img + (size_t)(i - K::WQKV_BYTES / 16) * 16;
    *reinterpret_cast<uint4*>(smem + (size_t)i * 16) = __ldg(reinterpret_cast<const uint4*>(src));
  }
  // zero P (off-diagonal blocks stay zero forever), Q/K/V images (K / N pads stay zero forever)
  for (int i = tid; i < (K::P_BYTES + 2 * K::AQ_BYTES + K::BV_BYTES) / 16; i += 256)
    *reinterpret_cast<uint4*>(sP + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 6 * 225; i += 256) sTab[i] = table[i];
  for (int i = tid; i < 6 * NH; i += 256) sBqkv[i] = bqkv[i];
  for (int i = tid; i < CP; i += 256) sBproj[i] = bproj[i];
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t aA = smem_u32(sA), aP = smem_u32(sP), aAq = smem_u32(sAq), aBk = smem_u32(sBk), aBv = smem_u32(sBv);
  const uint32_t aWqkv = smem_u32(smem + K::OFF_WQKV), aWproj = smem_u32(smem + K::OFF_WPROJ);

  const int row = tid & 127, part = tid >> 7;
  const int wsel = row >> 6, irow = row & 63, iy = irow >> 3, ix = irow & 7;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float inv_c = 1.0f / (float)C_;
  const int64_t ntiles = (geo.nwt + 1) / 2;
  uint32_t ph_q0 = 0, ph_q1 = 0, ph_s = 0, ph_o = 0, ph_p = 0;

  auto issue_qkv = [&](int h) {      // thread 0 only
    constexpr uint32_t idq = make_idesc_bf16(128, NH, false, false);
    const uint32_t d = tmem + ((h & 1) ? K::TM_QKV1 : K::TM_QKV0);
    const uint32_t wb = aWqkv + h * (NH * CP * 2);
#pragma unroll
    for (int ks = 0; ks < CP / 16; ++ks)
      mma_bf16_ss(d, make_smem_desc(aA + ks * 4096, 2048, 128), make_smem_desc(wb + ks * 2 * (NH * 16), NH * 16, 128),
                  idq, ks > 0);
    commit(&bars[h & 1]);
  };

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---------------- P1: gather two windows + LayerNorm -> A image ----------------
#pragma unroll 1
    for (int g = warp; g < 16; g += 8) {
      const int r = g * 8 + (lane & 7);
      const int64_t win = tile * 2 + (g >> 3);
      int region = 0; bool edge = false;
      int64_t t = -1;
      if (win < geo.nwt) t = win_token(geo, win, g & 7, lane & 7, region, edge);
      if ((lane >> 3) == 0) sReg[r] = edge ? region : -1;
      uint4 raw[K::NCH / 4];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < K::NCH / 4; ++j) {
        const int c = (lane >> 3) + 4 * j;
        raw[j] = t >= 0 ? __ldg(reinterpret_cast<const uint4*>(X + t * ldx) + c) : make_uint4(0, 0, 0, 0);
        const float2 f0 = up2(raw[j].x), f1 = up2(raw[j].y), f2 = up2(raw[j].z), f3 = up2(raw[j].w);
        s += (f0.x + f0.y) + (f1.x + f1.y) + (f2.x + f2.y) + (f3.x + f3.y);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      const float mean = s * inv_c;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < K::NCH / 4; ++j) {
        const uint32_t w4[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = up2(w4[q]);
          ss += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
      }
      ss += __shfl_xor_sync(0xffffffffu, ss, 8);
      ss += __shfl_xor_sync(0xffffffffu, ss, 16);
      ss -= (float)(CP - C_) * mean * mean;
      const float rstd = rsqrtf(fmaxf(ss, 0.f) * inv_c + 1e-5f);
#pragma unroll
      for (int j = 0; j < K::NCH / 4; ++j) {
        const int c = (lane >> 3) + 4 * j;
        const uint32_t w4[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = up2(w4[q]);
          o[q] = pk2((f.x - mean) * rstd, (f.y - mean) * rstd);
        }
        *reinterpret_cast<uint4*>(sA + c * 2048 + r * 16) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_qkv(0);
      issue_qkv(1);
    }
    const int myreg = sReg[row];
    const bool wmask = myreg >= 0;                 // warp-uniform: a warp covers 32 rows of one window
    float psum[6];

    // ---------------- heads ----------------
#pragma unroll
    for (int h = 0; h < 6; ++h) {
      // drain qkv_h : part 0 takes q,k ; part 1 takes v
      if (h & 1) { mbar_wait(&bars[1], ph_q1 & 1); ph_q1++; } else { mbar_wait(&bars[0], ph_q0 & 1); ph_q0++; }
      fence_after_sync();
      const uint32_t tq = lane_addr + ((h & 1) ? K::TM_QKV1 : K::TM_QKV0);
      const float* bq = sBqkv + h * NH;
      if (part == 0) {
        constexpr int NC = (2 * HD + 7) / 8 * 8;
        float f[NC];
        tmem_load_cols<NC>(tq, f);
#pragma unroll
        for (int sel = 0; sel < 2; ++sel) {          // 0: q -> Aq, 1: k -> Bk
          uint8_t* dst = (sel == 0 ? sAq : sBk) + row * 16;
#pragma unroll
          for (int c8 = 0; c8 < (HD + 7) / 8; ++c8) {
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int d0 = c8 * 8 + 2 * q, d1 = d0 + 1;
              const float a = d0 < HD ? f[sel * HD + d0] + bq[sel * HD + d0] : 0.f;
              const float b = d1 < HD ? f[sel * HD + d1] + bq[sel * HD + d1] : 0.f;
              o[q] = pk2(a, b);
            }
            *reinterpret_cast<uint4*>(dst + c8 * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      } else {
        // v columns start at 2*HD (not 8-aligned in general): load an aligned superset
        constexpr int C0 = (2 * HD) / 8 * 8;
        constexpr int NC = (3 * HD - C0 + 7) / 8 * 8;
        float f[NC];
        tmem_load_cols<NC>(tq + C0, f);
        if (h > 0) { mbar_wait(&bars[3], (ph_o - 1) & 1); }     // PV of head h-1 must have finished reading V
#pragma unroll
        for (int c8 = 0; c8 < (HD + 7) / 8; ++c8) {
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int d0 = c8 * 8 + 2 * q, d1 = d0 + 1;
            const float a = d0 < HD ? f[2 * HD - C0 + d0] + bq[2 * HD + d0] : 0.f;
            const float b = d1 < HD ? f[2 * HD - C0 + d1] + bq[2 * HD + d1] : 0.f;
            o[q] = pk2(a, b);
          }
          // MN-major V image: (token k=row, n=d) at (row/8)*128 + (d/8)*2048 + (row%8)*16 + (d%8)*2  == row*16 + c8*2048
          *reinterpret_cast<uint4*>(sBv + row * 16 + c8 * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        constexpr uint32_t ids = make_idesc_bf16(128, 128, false, false);
#pragma unroll
        for (int ks = 0; ks < K::HDP / 16; ++ks)
          mma_bf16_ss(tmem + K::TM_S, make_smem_desc(aAq + ks * 4096, 2048, 128), make_smem_desc(aBk + ks * 4096, 2048, 128),
                      ids, ks > 0);
        commit(&bars[2]);
        if (h + 2 < 6) issue_qkv(h + 2);
      }
      // ---- softmax on this thread's half row (32 of the 64 keys of its own window) ----
      mbar_wait(&bars[2], ph_s & 1); ph_s++;
      fence_after_sync();
      {
        uint32_t v[32];
        tmem_ld_x32(lane_addr + K::TM_S + 64 * wsel + 32 * part, v);
        wait_ld();
        const float* tb = sTab + h * 225 + ((iy + 7) * 15 + ix + 7) - 60 * part;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = __uint_as_float(v[j]) + tb[-((j >> 3) * 15 + (j & 7))];
          v[j] = __float_as_uint(t);
        }
        if (wmask) {
          const int* rg = sReg + 64 * wsel + 32 * part;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (rg[j] != myreg) v[j] = __float_as_uint(__uint_as_float(v[j]) + mask_val);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        sRed[part * 128 + row] = mx;
        pair_barrier(1 + (warp & 3));
        mx = fmaxf(mx, sRed[(1 - part) * 128 + row]);
        float sum = 0.f;
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2(__uint_as_float(v[2 * j]) - mx), p1 = ex2(__uint_as_float(v[2 * j + 1]) - mx);
          sum += p0 + p1;
          o[j] = pk2(p0, p1);
        }
        psum[h] = sum;
        uint8_t* dst = sP + (8 * wsel + 4 * part) * 2048 + row * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dst + q * 2048) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        constexpr uint32_t idv = make_idesc_bf16(128, K::HDV, false, true);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_bf16_ss(tmem + K::TM_O + h * K::HDV, make_smem_desc(aP + ks * 4096, 2048, 128),
                      make_smem_desc(aBv + ks * 256, 128, 2048), idv, ks > 0);
        commit(&bars[3]);
      }
      ph_o++;
    }
    // ---------------- O / rowsum -> A image for proj ----------------
    mbar_wait(&bars[3], (ph_o - 1) & 1);
    fence_after_sync();
#pragma unroll
    for (int h = 0; h < 6; ++h) sSum[(part * 6 + h) * 128 + row] = psum[h];
    __syncthreads();
    {
      constexpr int NC = (HD + 7) / 8 * 8;
#pragma unroll
      for (int hh = 0; hh < 3; ++hh) {
        const int h = part * 3 + hh;
        float f[NC];
        tmem_load_cols<NC>(lane_addr + K::TM_O + h * K::HDV, f);
        const float own = part ? psum[3 + hh] : psum[hh];
        const float inv = 1.0f / (own + sSum[((1 - part) * 6 + h) * 128 + row]);
        if (K::HDO == 16) {
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int d0 = c8 * 8 + 2 * q, d1 = d0 + 1;
              o[q] = pk2(d0 < HD ? f[d0] * inv : 0.f, d1 < HD ? f[d1] * inv : 0.f);
            }
            *reinterpret_cast<uint4*>(sA + (2 * h + c8) * 2048 + row * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        } else {   // HDO == 20: element k = 20h + d, written as 5 groups of 4 bf16 (8 bytes)
#pragma unroll
          for (int m = 0; m < 5; ++m) {
            const int k = 20 * h + 4 * m;
            const uint2 val = make_uint2(pk2(f[4 * m] * inv, f[4 * m + 1] * inv), pk2(f[4 * m + 2] * inv, f[4 * m + 3] * inv));
            *reinterpret_cast<uint2*>(sA + (k >> 3) * 2048 + row * 16 + (k & 7) * 2) = val;
          }
        }
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      constexpr uint32_t idp = make_idesc_bf16(128, CP, false, false);
#pragma unroll
      for (int ks = 0; ks < K::KPROJ / 16; ++ks)
        mma_bf16_ss(tmem + K::TM_PROJ, make_smem_desc(aA + ks * 4096, 2048, 128),
                    make_smem_desc(aWproj + ks * 2 * (CP * 16), CP * 16, 128), idp, ks > 0);
      commit(&bars[4]);
    }
    mbar_wait(&bars[4], ph_p & 1); ph_p++;
    fence_after_sync();
    // ---------------- proj epilogue -> swizzled staging (in the dead A region) -> coalesced residual store ----------------
    {
      constexpr int NC = CP / 2;
      const int cbeg = part * NC;
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(lane_addr + K::TM_PROJ + cbeg + c0, v);
        wait_ld();
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = pk2(__uint_as_float(v[2 * j]) + sBproj[cbeg + c0 + 2 * j], __uint_as_float(v[2 * j + 1]) + sBproj[cbeg + c0 + 2 * j + 1]);
        const int ch = (cbeg + c0) >> 3;
        const int sw = K::SWZ ? (row & 7) : 0;
        *reinterpret_cast<uint4*>(sA + row * K::PITCH + ((ch ^ sw) * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(sA + row * K::PITCH + (((ch + 1) ^ sw) * 16)) = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    fence_before_sync();
    __syncthreads();
#pragma unroll 1
    for (int g = warp; g < 16; g += 8) {
      const int r = g * 8 + (lane & 7);
      const int64_t win = tile * 2 + (g >> 3);
      if (win < geo.nwt) {
        int region; bool edge;
        const int64_t t = win_token(geo, win, g & 7, lane & 7, region, edge);
#pragma unroll
        for (int j = 0; j < K::NCH / 4; ++j) {
          const int c = (lane >> 3) + 4 * j;
          const uint4 m = *reinterpret_cast<const uint4*>(sA + r * K::PITCH + ((c ^ (K::SWZ ? (r & 7) : 0)) * 16));
          const uint4 x = __ldg(reinterpret_cast<const uint4*>(X + t * ldx) + c);
          const uint32_t mw[4] = {m.x, m.y, m.z, m.w}, xw[4] = {x.x, x.y, x.z, x.w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 a = up2(mw[q]), b = up2(xw[q]);
            o[q] = pk2(a.x + b.x, a.y + b.y);
          }
          *(reinterpret_cast<uint4*>(Y + t * ldy) + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int C_>
static int launch_attn(const void* x, int64_t ldx, void* y, int64_t ldy, const void* wqkv, const void* wproj,
                       const float* bqkv, const float* bproj, const float* table, int B, int H, int W, int shift,
                       int sms, cudaStream_t st) {
  using K = AttnCfg<C_>;
  WinGeom g;
  g.H = H; g.W = W; g.shift = shift; g.nwx = W / 8; g.nw_img = (H / 8) * (W / 8);
  g.nwt = (int64_t)B * g.nw_img;
  const int64_t ntiles = (g.nwt + 1) / 2;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  auto k = stl_attn_kernel<C_>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
  if (e != cudaSuccess) { set_error("rdst_stl_attn_fwd_bf16: smem attr (%d B): %s", K::SMEM, cudaGetErrorString(e)); return RDST_E_CUDA; }
  k<<<grid, 256, K::SMEM, st>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, (const uint8_t*)wqkv,
                                (const uint8_t*)wproj, bqkv, bproj, table, g, -100.0f * 1.4426950408889634f);
  return RDST_OK;
}

}  // namespace rdst

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
  RDST_REQUIRE(x != y, "rdst_stl_attn_fwd_bf16: in-place operation is not supported (windows read shifted neighbours)");
  if (B <= 0) return RDST_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  switch (C) {
    case 60:  RDST_REQUIRE(ldx >= 64 && ldy >= 64, "ld too small");   rc = launch_attn<60>(x, ldx, y, ldy, wqkv_img, wproj_img, bqkv, bproj, table, B, H, W, shift, sms, st); break;
    case 90:  RDST_REQUIRE(ldx >= 96 && ldy >= 96, "ld too small");   rc = launch_attn<90>(x, ldx, y, ldy, wqkv_img, wproj_img, bqkv, bproj, table, B, H, W, shift, sms, st); break;
    case 120: RDST_REQUIRE(ldx >= 128 && ldy >= 128, "ld too small"); rc = launch_attn<120>(x, ldx, y, ldy, wqkv_img, wproj_img, bqkv, bproj, table, B, H, W, shift, sms, st); break;
    default: set_error("rdst_stl_attn_fwd_bf16: C=%d unsupported (60, 90, 120 with 6 heads)", C); return RDST_E_UNSUPPORTED;
  }
  if (rc) return rc;
  RDST_CHECK_LAUNCH("rdst_stl_attn_fwd_bf16");
  return RDST_OK;
}
