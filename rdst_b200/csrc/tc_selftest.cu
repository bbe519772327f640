// UMMA plumbing self-test: one CTA computes D[128][N] = A[128][K] . B[N][K]^T on tcgen05 with the same
// shared-memory images, descriptors, commit/mbarrier protocol and TMEM read-back that the fused kernels use.
// tests/test_umma_selftest.py runs it over the (N, K, layout) combinations the network needs, so a descriptor
// mistake shows up here and not as a silent numerical error deep inside the fused kernels.
#include "common.cuh"
#include "umma.cuh"

namespace rdst {
using namespace umma;

// A: [128][K] bf16 row-major, B: [N][K] bf16 row-major, D: [128][N] fp32 (all 128 TMEM lanes are dumped even
// when m64 != 0 so the M=64 accumulator lane mapping can be inspected).
__global__ void __launch_bounds__(128) umma_selftest_kernel(const __nv_bfloat16* __restrict__ A,
                                                            const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                                                            int N, int K, int b_mn_major, int m64) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem;                         // K-major: (k/8)*2048 + r*16 + (k%8)*2
  uint8_t* sB = smem + (size_t)128 * K * 2;   // K-major: (k/8)*(N*16) + n*16 + (k%8)*2 ; MN-major: see below

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }

  // stage A (zero rows >= 64 when m64 so stale smem cannot leak in)
  for (int idx = tid; idx < 128 * (K / 8); idx += 128) {
    const int r = idx % 128, kc = idx / 128;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m64 != 1 || r < 64) v = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kc * 8);
    *reinterpret_cast<uint4*>(sA + (size_t)kc * 2048 + r * 16) = v;
  }
  if (m64 == 2) {
    const int N2 = 2 * N;
    for (int idx = tid; idx < N2 * (K / 8); idx += 128) {
      const int n = idx % N2, kc = idx / N2;
      *reinterpret_cast<uint4*>(sB + (size_t)kc * (N2 * 16) + n * 16) =
          *reinterpret_cast<const uint4*>(B + (size_t)n * K + kc * 8);
    }
  } else if (!b_mn_major) {
    for (int idx = tid; idx < N * (K / 8); idx += 128) {
      const int n = idx % N, kc = idx / N;
      *reinterpret_cast<uint4*>(sB + (size_t)kc * (N * 16) + n * 16) =
          *reinterpret_cast<const uint4*>(B + (size_t)n * K + kc * 8);
    }
  } else {
    // MN-major image: element (k, n) at (k/8)*128 + (n/8)*(K/8*128) + (k%8)*16 + (n%8)*2
    for (int idx = tid; idx < N * K; idx += 128) {
      const int n = idx % N, k = idx / N;
      *reinterpret_cast<__nv_bfloat16*>(sB + (size_t)(k / 8) * 128 + (size_t)(n / 8) * ((K / 8) * 128) + (k % 8) * 16 +
                                        (n % 8) * 2) = B[(size_t)n * K + k];
    }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (m64 == 3) {
    // stage A into TMEM columns [256, 256 + K/2): thread = row, packed bf16 pairs (k even in the low half)
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + 2 * (c0 + j));
      tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  if (tid == 0 && m64 == 2) {
    // lane-mask probe: rows 0-63 use B[0:N), rows 64-127 use B[N:2N), both into the same N columns
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    for (int half = 0; half < 2; ++half)
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2 * 2048, 2048, 128);
        const uint64_t db = make_smem_desc(smem_u32(sB) + half * N * 16 + ks * 2 * (2 * N * 16), 2 * N * 16, 128);
        mma_bf16_ss_masked(tmem, da, db, idesc, ks > 0 ? 1u : 0u, half ? 0xFFFFFFFFu : 0u, half ? 0xFFFFFFFFu : 0u,
                           half ? 0u : 0xFFFFFFFFu, half ? 0u : 0xFFFFFFFFu);
      }
    commit(&bar);
  } else if (tid == 0 && m64 == 3) {
    const uint32_t idesc = make_idesc_bf16(128, N, false, b_mn_major != 0);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t db;
      if (!b_mn_major) db = make_smem_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      else db = make_smem_desc(smem_u32(sB) + ks * 2 * 128, 128, (K / 8) * 128);
      mma_bf16_ts_masked(tmem, tmem + 256 + ks * 8, db, idesc, ks > 0 ? 1u : 0u, 0, 0, 0, 0);
    }
    commit(&bar);
  } else if (tid == 0) {
    // m64 == 4: fp16 accumulators (c_format = F16): the raw TMEM columns are dumped, two fp16 values per column
    const uint32_t idesc = m64 == 4 ? (make_idesc_f16(128, N, false, b_mn_major != 0) & ~(3u << 4))     // A/B are fp16 bit patterns
                                    : make_idesc_bf16(m64 ? 64 : 128, N, false, b_mn_major != 0);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2 * 2048, 2048, 128);
      uint64_t db;
      if (!b_mn_major) db = make_smem_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      else db = make_smem_desc(smem_u32(sB) + ks * 2 * 128, 128, (K / 8) * 128);
      mma_bf16_ss(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();

  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tmem_ld_x8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace rdst

extern "C" int rdst_umma_selftest(const void* a_bf16, const void* b_bf16, float* d, int N, int K, int b_mn_major,
                                  int m64, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(a_bf16 && b_bf16 && d, "rdst_umma_selftest: null pointer");
  RDST_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 16 == 0 && K <= 256,
               "rdst_umma_selftest: need 16<=N<=256 (N%%16==0), 16<=K<=256 (K%%16==0); got N=%d K=%d", N, K);
  const size_t smem = (size_t)128 * K * 2 + (size_t)(m64 == 2 ? 2 : 1) * N * K * 2;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("rdst_umma_selftest: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16,
                                                              d, N, K, b_mn_major, m64);
  RDST_CHECK_LAUNCH("rdst_umma_selftest");
  return RDST_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05 issue-pattern microbenchmark (timing only; operands are zeros).  One CTA, one issuing lane: `count` MMAs of
// M=128, N, K=16 are issued round-robin over `chains` different accumulators (chain c at TMEM column c * N), then one
// commit; out[0] = cycles from the first issue to the completion barrier, out[1] = cycles spent issuing.
//   a_tmem != 0: A operand from TMEM (TS mode), else from shared memory (SS);  masked != 0: disable-output-lane form.
// Answers "what does a short dependent accumulate chain cost" for the window-attention kernel.
namespace rdst {
using namespace umma;
template <int N, int CHAINS, bool TS, bool MASKED>
__global__ void __launch_bounds__(128) umma_bench_kernel(int reps, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 65536 / 16; i += 128) *reinterpret_cast<uint4*>(smem + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  {
    uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 512; c += 8) tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + c, z);
    wait_st();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp == 0) {
    // warp-uniform loop, constant descriptors, only the MMAs under elect.sync (the issue pattern of the fused kernels)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 32768;
    const unsigned long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t d = tm + (i % CHAINS) * N;
          const uint64_t db = make_smem_desc(sb + (i & 3) * 4096, N * 16, 128);
          if (TS) {
            if (MASKED) mma_bf16_ts_masked(d, tm + 448 + (i & 3) * 8, db, idesc, 1u, (i & 1) ? 0xFFFFFFFFu : 0u, (i & 1) ? 0xFFFFFFFFu : 0u,
                                           (i & 1) ? 0u : 0xFFFFFFFFu, (i & 1) ? 0u : 0xFFFFFFFFu);
            else mma_ts(d, tm + 448 + (i & 3) * 8, db, idesc, 1u);
          } else {
            mma_bf16_ss(d, make_smem_desc(sa + (i & 3) * 4096, 2048, 128), db, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    const unsigned long long t1 = clock64();
    if (elect_one()) commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const unsigned long long t2 = clock64();
    if ((tid & 31) == 0) { out[0] = t2 - t0; out[1] = t1 - t0; }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace rdst

extern "C" int rdst_umma_bench(int N, int chains, int count, int a_tmem, int masked, void* out_2_u64, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(out_2_u64 && chains >= 1 && chains * N <= 448 && count >= 16 && count % 16 == 0,
               "rdst_umma_bench: need chains*N<=448, count a multiple of 16");
  void (*k)(int, unsigned long long*) = nullptr;
#define RDST_UB(NN, CC, TT, MM) if (N == NN && chains == CC && (a_tmem != 0) == TT && (masked != 0) == MM) k = umma_bench_kernel<NN, CC, TT, MM>;
#define RDST_UB_N(NN) RDST_UB(NN, 1, true, false) RDST_UB(NN, 2, true, false) RDST_UB(NN, 1, true, true) RDST_UB(NN, 2, true, true) \
                      RDST_UB(NN, 1, false, false) RDST_UB(NN, 2, false, false)
  RDST_UB_N(16) RDST_UB_N(32) RDST_UB_N(64) RDST_UB_N(128)
  RDST_UB(32, 4, true, false) RDST_UB(64, 4, true, false) RDST_UB(32, 4, true, true) RDST_UB(64, 4, true, true)
  RDST_UB(256, 1, true, false) RDST_UB(256, 1, false, false) RDST_UB(192, 1, true, false) RDST_UB(192, 2, true, false)
#undef RDST_UB_N
#undef RDST_UB
  RDST_REQUIRE(k != nullptr, "rdst_umma_bench: combination N=%d chains=%d not instantiated", N, chains);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (e != cudaSuccess) { set_error("rdst_umma_bench: smem attr: %s", cudaGetErrorString(e)); return RDST_E_CUDA; }
  k<<<1, 128, 65536, (cudaStream_t)stream>>>(count / 16, (unsigned long long*)out_2_u64);
  RDST_CHECK_LAUNCH("rdst_umma_bench");
  return RDST_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// TMEM <-> register bandwidth probe: `nwarps` warps (1..16; warp w reads lane quadrant w % 4) each run `reps` rounds of
// four tcgen05.ld.32x32b.x32 (128 columns = 16 KB per warp and round) -- or tcgen05.st when `store` != 0 -- and wait.
// out[0] = cycles of the slowest warp, out[1] = bytes moved in total.
namespace rdst {
using namespace umma;
__global__ void __launch_bounds__(512) tmem_bw_kernel(int nwarps, int reps, int store, unsigned long long* __restrict__ out) {
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned long long tmax;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) tmax = 0;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i;
  tmem_st_x32(base, v); tmem_st_x32(base + 32, v); tmem_st_x32(base + 64, v); tmem_st_x32(base + 96, v);
  wait_st();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  uint32_t acc = 0;
  const unsigned long long t0 = clock64();
  if (warp < nwarps) {
    for (int r = 0; r < reps; ++r) {
      if (store) {
        tmem_st_x32(base, v); tmem_st_x32(base + 32, v); tmem_st_x32(base + 64, v); tmem_st_x32(base + 96, v);
        wait_st();
      } else {
        uint32_t a[32], b[32], c[32], d[32];
        tmem_ld_x32(base, a); tmem_ld_x32(base + 32, b); tmem_ld_x32(base + 64, c); tmem_ld_x32(base + 96, d);
        wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += a[i] ^ b[i] ^ c[i] ^ d[i];
      }
    }
  }
  const unsigned long long t1 = clock64();
  if (warp < nwarps && (tid & 31) == 0) atomicMax(&tmax, t1 - t0);
  __syncthreads();
  if (tid == 0) { out[0] = tmax; out[1] = (unsigned long long)nwarps * reps * 16384ull; out[2] = acc; }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}
}  // namespace rdst

extern "C" int rdst_tmem_bw_bench(int nwarps, int reps, int store, void* out_3_u64, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(out_3_u64 && nwarps >= 1 && nwarps <= 16 && reps >= 1, "rdst_tmem_bw_bench: 1 <= nwarps <= 16");
  tmem_bw_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(nwarps, reps, store, (unsigned long long*)out_3_u64);
  RDST_CHECK_LAUNCH("rdst_tmem_bw_bench");
  return RDST_OK;
}
