// Error plumbing and device probing for the C ABI.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace zsg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace zsg

extern "C" const char* zsg_last_error_string(void) { return zsg::g_err; }
extern "C" int zsg_abi_version(void) { return 8; }
extern "C" int zsg_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10;
}
