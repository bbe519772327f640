// TMA tensor maps over token-major activations ([B][H][W][ld] bf16) and the 4-D tile copies that use them.
//
// Window tiles are moved as 4x4-token boxes of 64 channels (128 bytes per token, SWIZZLE_128B): every (shifted)
// 8x8 window is exactly four such boxes -- the cyclic shift of ws/2 = 4 never splits a box -- so one descriptor
// serves shifted and unshifted blocks, and out-of-range coordinates zero-fill (missing window of an odd tile,
// channel pad of a 96-wide buffer).  Shared-memory image of one 64-channel panel: [token row][128 B], the 16-byte
// chunk c of row r stored at chunk position c ^ (r & 7)  (rows are 128 B, panels 1024-byte aligned).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rdst {

// host: encode (and cache) the map of a [B][H][W][C<=ld] bf16 activation with box {bc ch, bw, bh, 1}; bc = 64 ->
// SWIZZLE_128B (128-byte rows), bc = 32 -> SWIZZLE_64B (64-byte rows).  Returns nullptr (and sets the error string) on failure
const CUtensorMap* get_act_tmap(const void* base, int64_t ld, int B, int H, int W, int C, int bw, int bh, int bc = 64);

namespace tma {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// global -> shared box load; completes on `bar` with complete_tx(box bytes, out-of-range parts included)
__device__ __forceinline__ void load_4d(void* smem_dst, const CUtensorMap* map, int c, int w, int h, int b, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          s32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(w), "r"(h), "r"(b), "r"(s32(bar))
      : "memory");
}

// shared -> global box store (bulk-group completion); out-of-range parts are dropped
__device__ __forceinline__ void store_4d(const CUtensorMap* map, int c, int w, int h, int b, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c), "r"(w), "r"(h), "r"(b), "r"(s32(smem_src))
               : "memory");
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

}  // namespace tma
}  // namespace rdst
