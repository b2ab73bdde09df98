// Library-level entry points: error reporting, version, device info.
#include "common.cuh"
#include <stdarg.h>

namespace pt {

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

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace pt

extern "C" const char* pt_last_error(void) { return pt::last_error_buf(); }

extern "C" int pt_version(void) { return 100; /* 0.1.0 */ }

extern "C" int pt_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  PT_REQUIRE(sm_count && cc_major && cc_minor, "pt_device_info: null output pointer");
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  PT_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  PT_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  PT_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return PT_OK;
}
