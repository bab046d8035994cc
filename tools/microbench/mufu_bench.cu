// MUFU.EX2 throughput per SM on sm_100a for the f32 / f16 / bf16 flavours (the packed PTX forms ex2.approx.{f16x2,bf16x2} compile to
// two MUFU.EX2.{F16,BF16}; the question is whether those issue at the f32 rate).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench tools/microbench/mufu_bench.cu && /tmp/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned* out, int iters, unsigned seed) {
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        float f = __uint_as_float(x[i]), y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f));
        x[i] = __float_as_uint(y);
      } else if (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      } else if (MODE == 2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      } else {  // one 16-bit lane only
        unsigned short h = (unsigned short)x[i];
        asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h));
        x[i] = h;
      }
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int elems_per_instr) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out;
  cudaMalloc(&out, sms * 2 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<sms * 2, 1024>>>(out, 16, 1);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<sms * 2, 1024>>>(out, iters, 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double instr = (double)sms * 2 * 1024 * iters * 8;  // PTX instructions
  double exps = instr * elems_per_instr;
  printf("%-28s %8.3f ms  %7.2f Gexp/s  = %5.2f exp/clk/SM at %d MHz nominal\n", name, ms, exps / ms * 1e-6, exps / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("ex2.approx.f16 (scalar)", 1);
  return 0;
}
