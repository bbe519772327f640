// ABI bookkeeping of librdst_b200: version, thread-local error string, capability probe.
#include "common.cuh"

namespace rdst {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rdst

extern "C" int rdst_abi_version(void) { return RDST_ABI_VERSION; }

extern "C" const char* rdst_last_error(void) { return rdst::g_err; }

extern "C" int rdst_has_tcgen05(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return major == 10 ? 1 : 0;
}
