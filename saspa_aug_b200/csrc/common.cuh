// Shared helpers for the sm_100a kernels behind the saspa_b200 C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define SASPA_OK 0
#define SASPA_ERR_ARG (-1)
#define SASPA_ERR_WORKSPACE (-2)
#define SASPA_ERR_UNSUPPORTED (-3)
#define SASPA_ERR_DRIVER (-4)

void saspa_set_error(const char* fmt, ...);

#define SASPA_CHECK_ARG(cond, ...)                 \
  do {                                             \
    if (!(cond)) {                                 \
      saspa_set_error(__VA_ARGS__);                \
      return SASPA_ERR_ARG;                        \
    }                                              \
  } while (0)

#define SASPA_CUDA(call)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      saspa_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                         \
    }                                                                                         \
  } while (0)

#define SASPA_LAUNCH_CHECK()                                                                  \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      saspa_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                         \
    }                                                                                         \
  } while (0)

static inline int saspa_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// x * sigmoid(s * x) as FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL.  The flush-to-zero PTX forms are spelled out: __expf / __fdividef wrap the same
// two MUFU ops in denormal-range fix-ups (an extra compare and two predicated multiplies each) that matter only beyond |s * x| > 87, where
// this form already returns the right limit (t = inf -> rcp = 0 -> -0; t = 0 -> x).
__device__ __forceinline__ float x_sigmoid_f(float x, float neg_s_log2e) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * neg_s_log2e));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return x * r;
}
__device__ __forceinline__ float silu_f(float x) { return x_sigmoid_f(x, -1.4426950408889634f); }
__device__ __forceinline__ float quick_gelu_f(float x) { return x_sigmoid_f(x, -1.702f * 1.4426950408889634f); }
// erf-GELU on the FMA pipe only (GEGLU epilogue of the 64x64-level feed-forward GEMM: 16K gate values per 128 x 256 tile made the two
// MUFU ops of gelu_erf_f a 2048-clk-per-tile XU bill next to a 2560-clk main loop).  erf(u / sqrt 2) = u * Q(u^2), Q = degree-7
// near-minimax fit on |u| <= 4 (|erf error| <= 4.3e-5, i.e. |Phi error| <= 2.6e-5 incl. the clamp: erf(4 / sqrt 2) = 0.99994);
// |gelu error| <= 5.3e-5 * |x|, two orders below the bf16 rounding of the product it feeds.  13 instructions, no MUFU.
__device__ __forceinline__ float gelu_erf_poly_f(float x) {
  const float u = fminf(fmaxf(x, -4.0f), 4.0f);
  const float t = u * u;
  float q = fmaf(t, -3.161570339e-09f, 2.434221074e-07f);
  q = fmaf(q, t, -8.201730452e-06f);
  q = fmaf(q, t, 1.613347704e-04f);
  q = fmaf(q, t, -2.096408745e-03f);
  q = fmaf(q, t, 1.932974905e-02f);
  q = fmaf(q, t, -1.323507577e-01f);
  q = fmaf(q, t, 7.976950407e-01f);
  return x * fmaf(0.5f * u, q, 0.5f);
}
// Exact-erf GELU (F.gelu default, diffusers GEGLU / GELU) as x * Phi(x) with Phi from the Abramowitz-Stegun 7.1.26
// erfc approximation (|error| <= 1.5e-7 in Phi, far below bf16 / fp32-epilogue needs): two MUFU ops (rcp, ex2)
// and ~10 FMA-pipe instructions, branch-free; the negative tail is computed without cancellation.
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float h = 0.5f * poly * e;  // Phi(-|x|)
  return x * (x < 0.0f ? h : 1.0f - h);
}
