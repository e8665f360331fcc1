// Development microbenchmark: tensor-pipe time of one tcgen05.mma (M = 128, K = 16, bf16) as a function of N,
// with the A operand in shared memory (SS) or in tensor memory (TS).  Operands are zeros.  Not part of the product.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/umma_rate tools/micro/umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) k(int N, int ts, int iters, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tb = tslot;
  if (warp == 1 && lane == 0) {
    const uint32_t sb = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
    const uint64_t ad = make_desc(sb), bd = make_desc(sb + 16384);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (ts)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tb + 256), "r"(tb), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb + 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    cyc[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(512));
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {32, 64, 128, 256}) {
      k<<<148, 128, 64 * 1024>>>(N, ts, iters, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("M=128 N=%3d K=16 bf16, A in %s: %.1f cycles per tcgen05.mma   %s\n", N, ts ? "TMEM" : "smem", (double)h / iters, cudaGetErrorString(e));
    }
  return 0;
}
