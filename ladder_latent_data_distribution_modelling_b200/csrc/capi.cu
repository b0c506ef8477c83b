// Library-level entry points: version, error text, device check.
#include "common.cuh"
#include "ladder_sm100.h"
#include <cstring>

namespace ladder {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

unsigned long long& launch_counter() {
  static unsigned long long n = 0;
  return n;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;   // B200; also lets the planner run on a host without a device
  }
  return cached;
}

}  // namespace ladder

extern "C" {

int ladder_version(void) { return 100; }

const char* ladder_last_error(void) { return ladder::error_buffer(); }

unsigned long long ladder_launch_count(void) { return ladder::launch_counter(); }

int ladder_device_check(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return ladder::fail(LADDER_ERR_CUDA, "device %d: %s", device, cudaGetErrorString(e));
  if (p.major != 10)
    return ladder::fail(LADDER_ERR_ARCH, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
  return LADDER_OK;
}

}
