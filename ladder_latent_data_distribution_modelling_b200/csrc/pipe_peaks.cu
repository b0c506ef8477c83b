// Diagnostic kernels that measure the FP32-FMA and SFU (MUFU.EX2) issue peaks of the device.
// BASELINE.md asks for these: MEASURED_PEAKS.json only holds HBM and bf16 tensor peaks, and
// the mixture kernel's relevant pipes are FMA and SFU.  Timed from bench.py with CUDA events.
#include "common.cuh"
#include "ladder_sm100.h"

namespace ladder {

template <int KIND>
__global__ void __launch_bounds__(256) pipe_peak_kernel(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  const float m = 0.999f, c = 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) {
          a[i] = fmaf(a[i], m, c);
        } else {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
          a[i] = y;
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678f) out[0] = s;   // keep the chain alive without memory traffic
}

}  // namespace ladder

extern "C" int ladder_pipe_peak_launch(int kind, int blocks, int iters, float* out, cudaStream_t stream) {
  LADDER_REQUIRE(kind == 0 || kind == 1, "pipe_peak: kind 0 (FFMA) or 1 (MUFU.EX2)");
  LADDER_REQUIRE(blocks > 0 && iters > 0 && out != nullptr, "pipe_peak: bad arguments");
  if (kind == 0) ladder::pipe_peak_kernel<0><<<blocks, 256, 0, stream>>>(out, iters, 0.5f);
  else ladder::pipe_peak_kernel<1><<<blocks, 256, 0, stream>>>(out, iters, -0.5f);
  return ladder::check_launch("pipe_peak");
}
