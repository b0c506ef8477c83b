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

/* CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), the checksum of the TF tensor-bundle files the reference's savers write
 * (codes/base.py:37-48): host function, slicing-by-8.  `crc` chains calls (0 for the first). */
unsigned int ladder_crc32c(unsigned int crc, const void* data, size_t n) {
  static unsigned int table[8][256];
  static bool ready = false;
  if (!ready) {
    for (unsigned int i = 0; i < 256; ++i) {
      unsigned int c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][i] = c;
    }
    for (unsigned int i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xff];
    ready = true;
  }
  const unsigned char* p = static_cast<const unsigned char*>(data);
  unsigned int c = ~crc;
  while (n >= 8) {
    unsigned int lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xff] ^ table[6][(lo >> 8) & 0xff] ^ table[5][(lo >> 16) & 0xff] ^ table[4][lo >> 24] ^
        table[3][hi & 0xff] ^ table[2][(hi >> 8) & 0xff] ^ table[1][(hi >> 16) & 0xff] ^ table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}

int ladder_device_check(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return ladder::fail(LADDER_ERR_CUDA, "device %d: %s", device, cudaGetErrorString(e));
  if (p.major != 10)
    return ladder::fail(LADDER_ERR_ARCH, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
  return LADDER_OK;
}

}
