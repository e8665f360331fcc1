// gemm_simt.cu — true-fp32 FFMA GEMM for the parity mode (CSE_FP32).
//
// C[M,N] = A[M,K] * W[N,K]^T (+ bias_scale*bias) (relu) (+ residual).  Both operands are
// K-contiguous (activations row-major, nn.Linear weights [out,in]).  The fp32 mode exists to
// meet the 1e-4 relative-L2 bar against the reference's fp32 path (tensor-core TF32/BF16 inputs
// would not); the performance mode is gemm_tc.cu.
// Tile 128x128x16, 256 threads, 8x8 register micro-tile per thread (split 4+4 in both
// directions so shared-memory reads are conflict-free float4).
#include "common.cuh"

namespace cse {

constexpr int kBM = 128, kBN = 128, kBK = 16;
constexpr int kPad = 4;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int lda,
                                                        const float* __restrict__ W,
                                                        const float* __restrict__ bias,
                                                        float bias_scale,
                                                        const float* residual, float* C, int ldc,
                                                        int M, int N, int K, int relu) {
  __shared__ __align__(16) float As[2][kBK][kBM + kPad];
  __shared__ __align__(16) float Bs[2][kBK][kBN + kPad];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 thread grid
  // global->smem mapping: each thread moves 2 float4 of A and 2 of W per k-tile
  const int lrow = tid >> 2;          // 0..63
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  float4 ra[2], rb[2];

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = m0 + lrow + h * 64;
      if (row < M)
        ra[h] = *reinterpret_cast<const float4*>(A + (size_t)row * lda + k0 + lk);
      else
        ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int col = n0 + lrow + h * 64;  // N is a multiple of kBN (checked by the launcher)
      rb[h] = *reinterpret_cast<const float4*>(W + (size_t)col * K + k0 + lk);
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x;
      As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z;
      As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x;
      Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z;
      Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  const int nk = K / kBK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j}
#pragma unroll
  for (int ih = 0; ih < 2; ++ih) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + ih * 64 + ty * 4 + i;
      if (row >= M) continue;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int col = n0 + jh * 64 + tx * 4;
        float4 v = make_float4(acc[ih * 4 + i][jh * 4 + 0], acc[ih * 4 + i][jh * 4 + 1],
                               acc[ih * 4 + i][jh * 4 + 2], acc[ih * 4 + i][jh * 4 + 3]);
        if (bias != nullptr) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + col);
          v.x = fmaf(bias_scale, bb.x, v.x);
          v.y = fmaf(bias_scale, bb.y, v.y);
          v.z = fmaf(bias_scale, bb.z, v.z);
          v.w = fmaf(bias_scale, bb.w, v.w);
        }
        if (relu) {
          v.x = fmaxf(v.x, 0.f);
          v.y = fmaxf(v.y, 0.f);
          v.z = fmaxf(v.z, 0.f);
          v.w = fmaxf(v.w, 0.f);
        }
        if (residual != nullptr) {
          const float4 r = *reinterpret_cast<const float4*>(residual + (size_t)row * ldc + col);
          v.x += r.x;
          v.y += r.y;
          v.z += r.z;
          v.w += r.w;
        }
        *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = v;
      }
    }
  }
}

int launch_gemm_simt(const float* A, int lda, const float* W, const float* bias, float bias_scale,
                     const float* residual, float* C, int ldc, int M, int N, int K, int relu,
                     cudaStream_t st) {
  if (N % kBN != 0 || K % kBK != 0 || lda % 4 != 0 || ldc % 4 != 0) {
    set_error("gemm_simt: need N %% %d == 0, K %% %d == 0, lda/ldc %% 4 == 0 (N=%d K=%d lda=%d ldc=%d)",
              kBN, kBK, N, K, lda, ldc);
    return 1;
  }
  if (M <= 0) return 0;
  dim3 grid(N / kBN, ceil_div(M, kBM));
  KernelScope prof(kClsGemmSimt, st);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(A, lda, W, bias, bias_scale, residual, C, ldc, M, N, K,
                                         relu);
  return check_launch("gemm_simt_kernel");
}

}  // namespace cse
