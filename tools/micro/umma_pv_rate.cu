// Development microbenchmark (not part of the product): tensor-pipe cost of the tcgen05.mma shapes the attention
// kernel issues — M = 128, K = 16, bf16 — as a function of N, the B operand's major-ness / swizzle (V is consumed
// MN-major exactly as TMA wrote it), the A operand's home (shared memory / tensor memory), and the issue pattern:
// and of how often an instruction OVERWRITES the accumulator (scale-d = 0) instead of accumulating into it.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/umma_pv_rate tools/micro/umma_pv_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Cfg {
  int N, ts, b_mn, sw64, fresh_every, zero_st;
};

__global__ void __launch_bounds__(128, 1) k(Cfg c, int iters, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
    const uint32_t layout = c.sw64 ? 4u : 2u;
    const uint32_t sbo = c.sw64 ? 512u : 1024u;
    const uint64_t ad = make_desc(sb, 1024, 2u);
    const uint64_t bd0 = make_desc(sb + 32768, sbo, layout);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.b_mn << 16) | (((uint32_t)c.N >> 3) << 17) | ((128u >> 4) << 24);
    // MN-major B: one K = 16 step is two 8-k swizzle atoms = 2 * sbo bytes; K-major B: +32 B inside the swizzle row
    const uint32_t bstep = c.b_mn ? (2 * sbo) >> 4 : 2;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int j = i & 3;
      const uint32_t acc = (c.fresh_every > 0 && i % c.fresh_every == 0) ? 0u : 1u;
      const uint32_t d = tb + 256;
      const uint64_t bd = bd0 + (uint64_t)(bstep * j);
      if (c.ts) {
        const uint32_t a = tb + 8u * j;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad + 2 * j), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
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
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  printf("cycles per tcgen05.mma (M=128, K=16, bf16); fresh_every = k: every k-th instruction overwrites the accumulator (scale-d = 0)\n");
  printf("%4s %5s %6s %5s %12s %8s\n", "N", "A", "Bmajor", "swz", "fresh_every", "cycles");
  for (int ts = 0; ts < 2; ++ts)
    for (int b_mn = 0; b_mn < 2; ++b_mn)
      for (int N : {32, 64, 128, 256}) {
        if (b_mn && N != 32) continue;
        const int sw64 = b_mn;  // the attention kernel's V operand: MN-major, SWIZZLE_64B, N = 32
        for (int fe : {0, 1, 2, 4, 16}) {
          Cfg c{N, ts, b_mn, sw64, fe, 0};
          k<<<148, 128, 100 * 1024>>>(c, iters, cyc);
          cudaError_t e = cudaDeviceSynchronize();
          long long h;
          cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
          printf("%4d %5s %6s %5s %12d %8.1f   %s\n", N, ts ? "TMEM" : "smem", b_mn ? "MN" : "K", sw64 ? "64B" : "128B", fe,
                 (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
          if (e != cudaSuccess) return 1;
        }
      }
  return 0;
}
