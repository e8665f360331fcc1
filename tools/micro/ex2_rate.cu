// Development microbenchmark: MUFU.EX2 throughput per SM for f32 / f16x2 / bf16x2 operands.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/ex2_rate tools/micro/ex2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xBC00BC00u + i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int elems_per_op) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 1024>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 2 * 1024 * iters * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-10s %.3f ms  %.2f instr-lanes/ns  -> %.2f ex2 results / clk / SM (at %d MHz nominal)  %s\n", name, ms,
         ops / ms * 1e-6, ops * elems_per_op / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("f32", 1);
  run<1>("f16x2", 2);
  run<2>("bf16x2", 2);
  return 0;
}
