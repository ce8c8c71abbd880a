// Error reporting and device queries shared by the C-ABI entry points.
#include "common.h"

#include <cstring>

namespace csmoe {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return -1;
  }
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      set_error("cudaDeviceGetAttribute(MultiProcessorCount) failed");
      return -1;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace csmoe

extern "C" int csmoe_abi_version(void) { return CSMOE_ABI_VERSION; }

extern "C" const char* csmoe_last_error(void) { return csmoe::g_err; }

extern "C" int csmoe_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    csmoe::set_error("cudaGetDevice failed (no CUDA device?)");
    return CSMOE_ERR_CUDA;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    csmoe::set_error("cudaDeviceGetAttribute failed");
    return CSMOE_ERR_CUDA;
  }
  return major == 10 ? 1 : 0;
}
