// Development microbenchmark: TMA store throughput vs box shape (how fast can 148 persistent CTAs write a
// [136544 x 768] bf16 matrix as 128x256 tiles?).  Not part of the product.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/tma_store_rate tools/micro/tma_store_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// Each of `nwarps` warps owns a (box_rows x box_cols) staging buffer and walks its share of the boxes of each tile.
template <int KEEP>
__global__ void __launch_bounds__(256, 1)
k(const __grid_constant__ CUtensorMap tm, int M, int N, int box_rows, int box_cols, int box_bytes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const int m_tiles = (M + 127) / 128, n_tiles = N / 256;
  const int bpr = 256 / box_cols, bpc = 128 / box_rows, nbox = bpr * bpc;
  for (int t = blockIdx.x; t < m_tiles * n_tiles; t += gridDim.x) {
    const int m0 = (t / n_tiles) * 128, n0 = (t % n_tiles) * 256;
    for (int b = warp; b < nbox; b += 8) {
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(KEEP) : "memory");
        tma_store_2d(&tm, base + (warp * (KEEP + 1)) * box_bytes % (200 * 1024 / 8 * 8), n0 + (b % bpr) * box_cols, m0 + (b / bpr) * box_rows);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int M = 136544, N = 768;
  void* out;
  cudaMalloc(&out, (size_t)M * N * 2);
  struct Cfg { int rows, cols; CUtensorMapSwizzle sw; const char* name; } cfgs[] = {
      {32, 64, CU_TENSOR_MAP_SWIZZLE_128B, "32x128B  sw128 (current)"},
      {128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128x128B sw128"},
      {32, 256, CU_TENSOR_MAP_SWIZZLE_NONE, "32x512B  none"},
      {64, 256, CU_TENSOR_MAP_SWIZZLE_NONE, "64x512B  none"},
      {16, 256, CU_TENSOR_MAP_SWIZZLE_NONE, "16x512B  none"},
      {32, 128, CU_TENSOR_MAP_SWIZZLE_NONE, "32x256B  none"},
  };
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t gstride[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {(cuuint32_t)c.cols, (cuuint32_t)c.rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int box_bytes = c.rows * c.cols * 2;
    for (int keep = 0; keep < 2; ++keep) {
      if ((size_t)8 * (keep + 1) * box_bytes > 200 * 1024) continue;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int it = 0; it < 3; ++it) {
        if (it == 1) cudaEventRecord(e0);
        if (keep == 0) k<0><<<148, 256, 220 * 1024>>>(tm, M, N, c.rows, c.cols, box_bytes);
        else k<1><<<148, 256, 220 * 1024>>>(tm, M, N, c.rows, c.cols, box_bytes);
      }
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      ms /= 2;
      printf("%-26s bufs/warp=%d  %.1f us  %.0f GB/s  (%s)\n", c.name, keep + 1, ms * 1e3, (double)M * N * 2 / (ms * 1e-3) * 1e-9,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
