// Host side of the TMA path (tensor-map encoding + cache) and a self-test of the 4-D box copies the fused kernels use.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include <mutex>
#include <vector>

namespace rdst {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct MapKey {
  const void* base; int64_t ld; int B, H, W, C, bw, bh, bc;
  bool operator==(const MapKey& o) const {
    return base == o.base && ld == o.ld && B == o.B && H == o.H && W == o.W && C == o.C && bw == o.bw && bh == o.bh && bc == o.bc;
  }
};
struct MapEntry { MapKey key; CUtensorMap map; };
std::mutex g_mu;
std::vector<MapEntry*> g_maps;       // entries are never freed or moved: callers keep the returned pointers
}  // namespace

const CUtensorMap* get_act_tmap(const void* base, int64_t ld, int B, int H, int W, int C, int bw, int bh, int bc) {
  if (bc != 64 && bc != 32) { set_error("get_act_tmap: box channel width must be 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)"); return nullptr; }
  const MapKey key{base, ld, B, H, W, C, bw, bh, bc};
  std::lock_guard<std::mutex> lk(g_mu);
  for (MapEntry* e : g_maps)
    if (e->key == key) return &e->map;
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return nullptr; }
  MapEntry* e = new MapEntry();
  e->key = key;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  const cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&e->map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, bc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for base=%p ld=%lld B=%d H=%d W=%d C=%d", (int)r, base, (long long)ld, B, H, W, C);
    delete e;
    return nullptr;
  }
  if (g_maps.size() >= 4096) g_maps.clear();     // bounded; stale pointers stay valid (entries are leaked, 160 B each)
  g_maps.push_back(e);
  return &e->map;
}

// one 8x8 window (shifted frame origin hs0, ws0 of image b) of NP 64-channel panels: load, dump the raw shared-memory
// image, store to Y through its own map
__global__ void __launch_bounds__(128) tma_selftest_kernel(const __grid_constant__ CUtensorMap mx,
                                                           const __grid_constant__ CUtensorMap my, int np, int H, int W,
                                                           int shift, int b, int hs0, int ws0, uint8_t* __restrict__ dump) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    umma::mbar_arrive_expect_tx(&bar, np * 8192);
    for (int p = 0; p < np; ++p)
      for (int q = 0; q < 4; ++q) {
        int h = hs0 + 4 * (q >> 1) + shift; if (h >= H) h -= H;
        int w = ws0 + 4 * (q & 1) + shift; if (w >= W) w -= W;
        tma::load_4d(smem + p * 8192 + q * 2048, &mx, p * 64, w, h, b, &bar);
      }
  }
  umma::mbar_wait(&bar, 0);
  for (int i = tid; i < np * 8192 / 16; i += 128)
    reinterpret_cast<uint4*>(dump)[i] = reinterpret_cast<const uint4*>(smem)[i];
  __syncthreads();
  if (tid == 0) {
    umma::fence_proxy_async();
    for (int p = 0; p < np; ++p)
      for (int q = 0; q < 4; ++q) {
        int h = hs0 + 4 * (q >> 1) + shift; if (h >= H) h -= H;
        int w = ws0 + 4 * (q & 1) + shift; if (w >= W) w -= W;
        tma::store_4d(&my, p * 64, w, h, b, smem + p * 8192 + q * 2048);
      }
    umma::bulk_commit();
    umma::bulk_wait_read();
  }
}

}  // namespace rdst

extern "C" int rdst_tma_selftest(const void* x, int64_t ldx, void* y, int64_t ldy, int B, int H, int W, int C, int shift,
                                 int b, int hs0, int ws0, void* dump, void* stream) {
  using namespace rdst;
  RDST_REQUIRE(x && y && dump, "rdst_tma_selftest: null pointer");
  RDST_REQUIRE(C > 0 && C <= 128 && C % 8 == 0 && H % 8 == 0 && W % 8 == 0, "rdst_tma_selftest: bad shape");
  const CUtensorMap* mx = get_act_tmap(x, ldx, B, H, W, C, 4, 4);
  const CUtensorMap* my = get_act_tmap(y, ldy, B, H, W, C, 4, 4);
  if (!mx || !my) return RDST_E_CUDA;
  const int np = (C + 63) / 64;
  tma_selftest_kernel<<<1, 128, np * 8192 + 1024, (cudaStream_t)stream>>>(*mx, *my, np, H, W, shift, b, hs0, ws0, (uint8_t*)dump);
  RDST_CHECK_LAUNCH("rdst_tma_selftest");
  return RDST_OK;
}
