// Shared helpers for the LaDDer sm_100a kernels (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#define LADDER_OK 0
#define LADDER_ERR_ARG (-1)
#define LADDER_ERR_CUDA (-2)
#define LADDER_ERR_ARCH (-3)
#define LADDER_ERR_WORKSPACE (-4)

namespace ladder {

// Thread-local last-error message, exposed through ladder_last_error().
char* error_buffer();
int fail(int code, const char* fmt, ...);

// number of kernels this library has enqueued in this process (bench.py reports it as gpu_launches)
unsigned long long& launch_counter();

inline int check_launch(const char* what) {
  ++launch_counter();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return LADDER_OK;
}

#define LADDER_REQUIRE(cond, ...) \
  do { if (!(cond)) return ::ladder::fail(LADDER_ERR_ARG, __VA_ARGS__); } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int num_sms();

// Activation codes shared by every op (reference uses leaky_relu(0.2), relu, tanh, none).
enum Act : int { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_TANH = 3 };

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case ACT_LEAKY: return v > 0.f ? v : 0.2f * v;
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_TANH: return tanhf(v);
    default: return v;
  }
}
// derivative expressed through the saved OUTPUT y (sign(y) == sign(pre) for leaky/relu)
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  switch (act) {
    case ACT_LEAKY: return y > 0.f ? 1.f : 0.2f;
    case ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}


// Output permutations fused into the GEMM epilogues (r = block size, 0 = none):
//  * forward:  y is written directly in depth_to_space (NHWC, DCR) layout -- row (b,h,w), col (i*r+j)*C'+c  ->
//              element (b, h*r+i, w*r+j, c) of [B, GH*r, GW*r, C']
//  * backward: the gradient w.r.t. a depth_to_space OUTPUT (row (b,y2,x2) of the [B,GH,GW,C] grid, col c) is
//              written at the position of the matching depth_to_space INPUT element
//              (b, y2/r, x2/r, ((y2%r)*r + x2%r)*C + c) of [B, GH/r, GW/r, C*r*r]
__device__ __forceinline__ long long d2s_dest(long long m, int n, int GH, int GW, int Ncols, int r) {
  const int Cp = Ncols / (r * r);
  const int w = (int)(m % GW);
  const long long q = m / GW;
  const int h = (int)(q % GH);
  const long long b = q / GH;
  const int ij = n / Cp, c = n % Cp, i = ij / r, j = ij % r;
  return ((b * GH * r + h * r + i) * ((long long)GW * r) + w * r + j) * Cp + c;
}
__device__ __forceinline__ long long s2d_dest(long long m, int n, int GH, int GW, int Ncols, int r) {
  const int x2 = (int)(m % GW);
  const long long q = m / GW;
  const int y2 = (int)(q % GH);
  const long long b = q / GH;
  return ((b * (GH / r) + y2 / r) * (GW / r) + x2 / r) * ((long long)Ncols * r * r) + ((y2 % r) * r + x2 % r) * Ncols + n;
}

}  // namespace ladder
