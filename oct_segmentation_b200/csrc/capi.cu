// Error reporting and device queries of the octseg C-ABI.
#include <cstring>
#include "common.h"

namespace octseg {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace octseg

extern "C" const char* octseg_last_error(void) { return octseg::last_error_buf(); }

extern "C" int octseg_abi_version(void) { return OCTSEG_ABI_VERSION; }

extern "C" int octseg_sm_count(void) {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return octseg::fail(OCTSEG_ENODEV, "cudaGetDevice failed");
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return octseg::fail(OCTSEG_ENODEV, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return octseg::fail(OCTSEG_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                        prop.major, prop.minor);
  if (dev >= 0 && dev < 64) cached[dev] = prop.multiProcessorCount;
  return prop.multiProcessorCount;
}
