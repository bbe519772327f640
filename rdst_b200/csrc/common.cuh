// Shared helpers for librdst_b200 (error reporting, storage-type load/store).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/rdst_b200.h"

namespace rdst {

void set_error(const char* fmt, ...);

#define RDST_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      rdst::set_error(__VA_ARGS__);                               \
      return RDST_E_INVALID;                                      \
    }                                                             \
  } while (0)

#define RDST_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) {                                                      \
      rdst::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));     \
      return RDST_E_CUDA;                                                         \
    }                                                                             \
  } while (0)

// Launch with programmatic stream serialization: the kernel may begin (prologue: barrier init, TMEM allocation, bulk
// weight loads) while the previous kernel of the stream drains; it calls griddepcontrol.wait before touching
// anything the previous kernel produced.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float ld_act(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_act(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_act(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_act(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// exact (erf) GELU, as nn.GELU() default  [reference swin_transformer_sr.py:14,25]
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace rdst
