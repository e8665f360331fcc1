// Development microbenchmark: tcgen05.ld throughput per SM for the 32x32b / 16x256b / 16x128b shapes, with 4 and 8
// warps, alone and while the tensor pipe runs 128x256x16 UMMAs on garbage operands.  Not part of the product.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/tmem_ld_rate tools/micro/tmem_ld_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define R64 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}"
#define O64(r) "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),"=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),"=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63])

template <int SHAPE>
__device__ __forceinline__ void ld8k(uint32_t taddr, uint32_t (&r)[64]) {
  if (SHAPE == 0)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 " R64 ", [%64];\ntcgen05.wait::ld.sync.aligned;\n" : O64(r) : "r"(taddr) : "memory");
  if (SHAPE == 1)
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x16.b32 " R64 ", [%64];\ntcgen05.wait::ld.sync.aligned;\n" : O64(r) : "r"(taddr) : "memory");
  if (SHAPE == 2)
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x32.b32 " R64 ", [%64];\ntcgen05.wait::ld.sync.aligned;\n" : O64(r) : "r"(taddr) : "memory");
  if (SHAPE == 3)  // the same 8 KB as four x16 loads in flight before one wait
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%64];\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%65];\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47}, [%66];\n"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%67];\n"
        "tcgen05.wait::ld.sync.aligned;\n" : O64(r) : "r"(taddr), "r"(taddr + 16), "r"(taddr + 32), "r"(taddr + 48) : "memory");
  if (SHAPE == 4)  // one x128 load: 16 KB per warp per wait (half as many waits per byte)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 " R64 ", [%64];\n" : O64(r) : "r"(taddr) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int SHAPE>
__global__ void __launch_bounds__(320, 1) k(int nwarps, int with_mma, int iters, long long* cyc, uint32_t* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tb = tslot;
  long long t0 = clock64();
  if (warp == 1 && lane == 0 && with_mma) {
    // keep the tensor pipe busy: 128x256x16 UMMAs into columns [256, 512) from zeroed shared memory
    const uint32_t sb = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
    const uint64_t ad = make_desc(sb), bd = make_desc(sb + 16384);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    for (int i = 0; i < iters * with_mma; ++i) {
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb + 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    cyc[blockIdx.x * 2 + 1] = clock64() - t0;
  }
  if (warp >= 2 && warp < 2 + nwarps) {
    uint32_t r[64], acc = 0;
    const int q = warp & 3;
    // 32x32b: 32 lanes x 64 columns.  16xNb: 16 lanes x 128 columns (lane base = first 16 lanes of the quarter)
    const uint32_t taddr = tb + ((uint32_t)(q * 32) << 16) + ((warp - 2) >> 2) * 128;
    for (int i = 0; i < iters; ++i) {
      ld8k<SHAPE>(taddr + ((SHAPE == 0 || SHAPE >= 3) ? (i & 1) * 64 : 0), r);
      if (SHAPE == 4) {  // second 8 KB in flight before the wait
        uint32_t r2[64];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 " R64 ", [%64];\ntcgen05.wait::ld.sync.aligned;\n" : O64(r2) : "r"(taddr + 64 - (i & 1) * 64) : "memory");
        acc ^= r2[0] ^ r2[31] ^ r2[63];
      }
      acc ^= r[0] ^ r[31] ^ r[63];  // static indices: a dynamic index would push the 64 registers to local memory
    }
    if (lane == 0 && warp == 2) cyc[blockIdx.x * 2] = clock64() - t0;
    sink[blockIdx.x * 320 + threadIdx.x] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(512));
}

template <int SHAPE>
void run(const char* name) {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 2 * 8); cudaMalloc(&sink, 148 * 320 * 4);
  cudaFuncSetAttribute(k<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int nw : {4, 8}) {
    for (int mma : {0, 1, 4}) {
      cudaMemset(cyc, 0, 148 * 2 * 8);
      k<SHAPE><<<148, 320, 64 * 1024>>>(nw, mma, iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("%-10s warps=%d  UMMAs per ld-iteration=%d : ld %.1f B/clk/SM", name, nw, mma, (double)nw * iters * (SHAPE == 4 ? 16384 : 8192) / (double)h[0]);
      if (mma) printf("   |  UMMA %.0f cycles each (128 alone)", (double)h[1] / (iters * mma));
      printf("   %s\n", cudaGetErrorString(e));
    }
  }
}
int main() {
  run<0>("32x32b.x64");
  run<1>("16x256b.x16");
  run<2>("16x128b.x32");
  run<3>("4 x (32x32b.x16) in flight");
  run<4>("2 x (32x32b.x64) in flight");
  return 0;
}
