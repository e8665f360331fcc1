// Development microbenchmark (not part of the product): what bounds the P V step of the attention kernel.
// One thread issues a fully unrolled body of sixteen tcgen05.mma (M = 128, N = 32, K = 16, bf16; A = P from tensor
// memory, B = V MN-major SWIZZLE_64B exactly as TMA wrote it) per round, under different accumulator patterns, with
// and without the softmax warps' tcgen05.ld / tcgen05.st traffic running beside it, and with the Q K^T instructions
// (N = 256) of the real kernel interleaved.  No runtime division or modulo in the issuing thread: an earlier version
// measured its own `i % k`.
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
}

// ROT: number of accumulators the sixteen P V instructions rotate over (1 = one dependent chain)
// FRESH: every 4th instruction overwrites (scale-d = 0) — the per-chunk accumulators of the v1 kernel
// WITH_S: two N = 256 Q K^T instructions (A, B from shared memory) in front of the sixteen
// TRAFFIC: warps 4-7 run tcgen05.ld.x64 + tcgen05.st.x32 loops on other columns meanwhile
// SS: A of the P V instructions from shared memory (v1) instead of tensor memory
template <int ROT, bool FRESH, bool WITH_S, bool TRAFFIC, bool SS>
__global__ void __launch_bounds__(256, 1) k(int rounds, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) stop = 0;
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
    const uint64_t q = make_desc(sb, 512, 4u), kk = make_desc(sb + 8192, 512, 4u);  // Q, K: K-major SWIZZLE_64B
    const uint64_t p_sm = make_desc(sb + 49152, 1024, 2u);                          // P in smem (v1): K-major SWIZZLE_128B
    const uint64_t v = make_desc(sb + 32768, 512, 4u);                              // V: MN-major SWIZZLE_64B
    constexpr uint32_t id_s = idesc_bf16(256, 0), id_pv = idesc_bf16(32, 1);
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (WITH_S) {
        mma_ss(tb + 256, q, kk, id_s, 0u);
        mma_ss(tb + 256, q + 2, kk + 2, id_s, 1u);
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const uint32_t d = tb + 32 + 64 * (uint32_t)(t % ROT);       // O tiles in columns [32,64), [96,128), ...
        const uint32_t a = tb + 64 * (uint32_t)(t / 4) + 8 * (uint32_t)(t & 3);
        const uint64_t bd = v + (uint64_t)(t * 64);                   // + t * 1024 B
        const uint32_t acc = (FRESH && (t & 3) == 0) ? 0u : 1u;
        if (SS) mma_ss(d, p_sm + (uint64_t)((t / 4) * 512 + 2 * (t & 3)), bd, id_pv, acc);
        else mma_ts(d, a, bd, id_pv, acc);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    cyc[blockIdx.x] = clock64() - t0;
    stop = 1;
  } else if (TRAFFIC && warp >= 4) {
    // the softmax warps' TMEM traffic: 64-column loads and 32-column stores on this warp's lane quarter, columns
    // [256, 512) when the Q K^T tile is not in use there, else [128, 256)
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (WITH_S ? 128u : 256u);
    uint32_t rg[64];
    uint32_t sink = 0;
    while (!stop) {
#pragma unroll 1
      for (int c = 0; c < (WITH_S ? 2 : 4); ++c) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
            "tcgen05.wait::ld.sync.aligned;\n"
            : "=r"(rg[0]), "=r"(rg[1]), "=r"(rg[2]), "=r"(rg[3]), "=r"(rg[4]), "=r"(rg[5]), "=r"(rg[6]), "=r"(rg[7]), "=r"(rg[8]), "=r"(rg[9]), "=r"(rg[10]), "=r"(rg[11]), "=r"(rg[12]), "=r"(rg[13]), "=r"(rg[14]), "=r"(rg[15]), "=r"(rg[16]), "=r"(rg[17]), "=r"(rg[18]), "=r"(rg[19]), "=r"(rg[20]), "=r"(rg[21]), "=r"(rg[22]), "=r"(rg[23]), "=r"(rg[24]), "=r"(rg[25]), "=r"(rg[26]), "=r"(rg[27]), "=r"(rg[28]), "=r"(rg[29]), "=r"(rg[30]), "=r"(rg[31]), "=r"(rg[32]), "=r"(rg[33]), "=r"(rg[34]), "=r"(rg[35]), "=r"(rg[36]), "=r"(rg[37]), "=r"(rg[38]), "=r"(rg[39]), "=r"(rg[40]), "=r"(rg[41]), "=r"(rg[42]), "=r"(rg[43]), "=r"(rg[44]), "=r"(rg[45]), "=r"(rg[46]), "=r"(rg[47]), "=r"(rg[48]), "=r"(rg[49]), "=r"(rg[50]), "=r"(rg[51]), "=r"(rg[52]), "=r"(rg[53]), "=r"(rg[54]), "=r"(rg[55]), "=r"(rg[56]), "=r"(rg[57]), "=r"(rg[58]), "=r"(rg[59]), "=r"(rg[60]), "=r"(rg[61]), "=r"(rg[62]), "=r"(rg[63])
            : "r"(base + 64u * c)
            : "memory");
#pragma unroll
        for (int i = 0; i < 64; ++i) sink ^= rg[i];
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\t"
            "tcgen05.wait::st.sync.aligned;\n" ::"r"(base + 64u * c),
            "r"(rg[0]), "r"(rg[1]), "r"(rg[2]), "r"(rg[3]), "r"(rg[4]), "r"(rg[5]), "r"(rg[6]), "r"(rg[7]), "r"(rg[8]), "r"(rg[9]),
            "r"(rg[10]), "r"(rg[11]), "r"(rg[12]), "r"(rg[13]), "r"(rg[14]), "r"(rg[15]), "r"(rg[16]), "r"(rg[17]), "r"(rg[18]),
            "r"(rg[19]), "r"(rg[20]), "r"(rg[21]), "r"(rg[22]), "r"(rg[23]), "r"(rg[24]), "r"(rg[25]), "r"(rg[26]), "r"(rg[27]),
            "r"(rg[28]), "r"(rg[29]), "r"(rg[30]), "r"(rg[31])
            : "memory");
      }
    }
    if (sink == 0x12345678u) cyc[200] = sink;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(512));
}

template <int ROT, bool FRESH, bool WITH_S, bool TRAFFIC, bool SS>
static void run(const char* what, long long* cyc) {
  auto kern = k<ROT, FRESH, WITH_S, TRAFFIC, SS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int rounds = 512;
  kern<<<148, 256, 100 * 1024>>>(rounds, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-92s %8.0f cycles per round (%5.1f per P V instruction)  %s\n", what, (double)h / rounds,
         (double)h / rounds / 16.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 256 * 8);
  printf("one round = sixteen tcgen05.mma M=128 N=32 K=16 (P V of one attention item) [+ two N=256 (Q K^T)]\n");
  run<1, false, false, false, false>("one accumulator (chain), P in TMEM", cyc);
  run<2, false, false, false, false>("2 accumulators in rotation, P in TMEM", cyc);
  run<4, false, false, false, false>("4 accumulators in rotation, P in TMEM", cyc);
  run<1, true, false, false, false>("one accumulator, every 4th instruction overwrites, P in TMEM", cyc);
  run<1, false, false, false, true>("one accumulator (chain), P in smem", cyc);
  run<4, false, false, false, true>("4 accumulators in rotation, P in smem", cyc);
  run<1, true, false, false, true>("one accumulator, every 4th overwrites, P in smem (the v1 kernel's pattern)", cyc);
  run<1, false, true, false, false>("Q K^T + chain, P in TMEM", cyc);
  run<4, false, true, false, false>("Q K^T + 4 accumulators in rotation, P in TMEM", cyc);
  run<1, false, false, true, false>("chain, P in TMEM, with tcgen05.ld/st traffic from 4 warps", cyc);
  run<4, false, false, true, false>("4 accumulators in rotation, P in TMEM, with tcgen05.ld/st traffic", cyc);
  run<1, false, true, true, false>("Q K^T + chain, P in TMEM, with traffic", cyc);
  run<4, false, true, true, false>("Q K^T + 4 accumulators in rotation, P in TMEM, with traffic", cyc);
  run<1, true, true, true, true>("Q K^T + per-chunk overwrite, P in smem, with traffic (v1 in situ)", cyc);
  return 0;
}
