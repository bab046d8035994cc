// Issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a: does the packed form retire two fp32 FMAs per issue slot?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench tools/microbench/ffma2_bench.cu && /tmp/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 0.001f + i;
  const float m = 1.0001f, c = 0.5f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(m), "f"(c));
    } else {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long v, mm, cc;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(mm), "l"(cc));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(v));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sms * 2 * 1024 * 4);
  const int iters = 8192;
  k<MODE><<<sms * 2, 1024>>>(out, 16, 1.f);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<sms * 2, 1024>>>(out, iters, 1.f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double fmas = (double)sms * 2 * 1024 * iters * 16;
  printf("%-24s %8.3f ms  %8.1f GFMA/s = %6.1f fp32 FMA/clk/SM at %d MHz nominal\n", name, ms, fmas / ms * 1e-6, fmas / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  cudaFree(out);
}

int main() {
  run<0>("fma.rn.f32");
  run<1>("fma.rn.f32x2");
  return 0;
}
