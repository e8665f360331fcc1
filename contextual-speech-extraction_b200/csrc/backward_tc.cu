// backward_tc.cu — nn.Linear backward on the tcgen05 GEMM (bf16 operands, fp32 accumulate).
//
// STATUS: compiled and argument-checked only — written after this round's GPU budget was spent, NOT yet run
// on hardware.  Nothing on the default path calls it; tests/test_backward_tc_gpu.py is skipped unless
// CSE_EXPERIMENTAL=1.  It is the first step of the performance-mode backward (DESIGN.md §7 "Next"): the
// fp32 SIMT dgrad + wgrad it replaces are 72 % of the measured training step.
//
// Both gradients are expressed as the forward GEMM C = A W^T that gemm_tc.cu already implements:
//   dgrad  dA[M,K]  = dC[M,N] (W^T)[K,N]^T          A := bf16(dC),      W := bf16(W^T)
//   wgrad  dW[N,K] += (dC^T)[N,M] (A^T)[K,M]^T      A := bf16(dC)^T,    W := bf16(A)^T,  K_gemm = M (padded to 64)
// so the only new device code is a transposing cast.  The wgrad output has (N/128)*(K/256) <= 8 tiles; a
// split over M (several launches accumulating through the TMA reduce-add epilogue, or native MN-major UMMA
// operands) is the follow-up once this is measured.
// Reference: autograd of nn.Linear (CSE_transformer.py:335-340,468-477; ContSep.py:229,247,255,258).
#include "common.cuh"

namespace cse {

// X [M, N] (row stride ld, float or bf16) -> XT [N, Mpad] bf16, zero-filled for m >= M.
template <typename T>
__global__ void __launch_bounds__(256) transpose_cast_kernel(const T* __restrict__ X, int ld, int M, int N,
                                                             int Mpad, bf16* __restrict__ XT) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, m0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int m = m0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (m < M && n < N) ? to_f(X[(size_t)m * ld + n]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int n = n0 + i, m = m0 + threadIdx.x;
    if (n < N && m < Mpad) XT[(size_t)n * Mpad + m] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

static int launch_transpose_cast(const void* X, int x_is_bf16, int ld, int M, int N, int Mpad, bf16* XT,
                                 cudaStream_t st) {
  dim3 grid(ceil_div(Mpad, 32), ceil_div(N, 32));
  if (x_is_bf16)
    transpose_cast_kernel<bf16><<<grid, dim3(32, 8), 0, st>>>((const bf16*)X, ld, M, N, Mpad, XT);
  else
    transpose_cast_kernel<float><<<grid, dim3(32, 8), 0, st>>>((const float*)X, ld, M, N, Mpad, XT);
  return check_launch("transpose_cast_kernel");
}

struct TcBwdScratch {
  bf16 *dC16, *WT16, *XT, *YT;
  float* WT32;
  size_t total;
};

static TcBwdScratch carve_tc_bwd(char* base, size_t M, size_t N, size_t K) {
  TcBwdScratch s;
  const size_t Mpad = align_up(M, 64);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off = align_up(off + bytes, 256);
    return p;
  };
  s.dC16 = (bf16*)take(M * N * 2);
  s.WT32 = (float*)take(N * K * 4);
  s.WT16 = (bf16*)take(N * K * 2);
  s.XT = (bf16*)take(N * Mpad * 2);
  s.YT = (bf16*)take(K * Mpad * 2);
  s.total = off;
  return s;
}

}  // namespace cse

using namespace cse;

extern "C" {

size_t cse_linear_bwd_tc_scratch_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return carve_tc_bwd(nullptr, (size_t)M, (size_t)N, (size_t)K).total;
}

int cse_linear_bwd_tc(const void* A, int a_is_bf16, int lda, const float* W, const float* dC, int M, int N,
                      int K, void* dA, int dA_fp32, int ldda, float* dW, float* dbias, void* scratch,
                      size_t scratch_bytes, void* stream) {
  CSE_REQUIRE(A && W && dC && scratch && M > 0, "linear_bwd_tc: bad argument");
  CSE_REQUIRE(N % 128 == 0 && K % 128 == 0, "linear_bwd_tc: N and K must be multiples of 128 (N=%d K=%d)", N, K);
  CSE_REQUIRE(dbias == nullptr || N % 256 == 0, "linear_bwd_tc: bias gradient needs N %% 256 == 0 (N=%d)", N);
  CSE_REQUIRE(((uintptr_t)scratch & 255) == 0, "linear_bwd_tc: scratch must be 256-byte aligned");
  const TcBwdScratch s = carve_tc_bwd((char*)scratch, (size_t)M, (size_t)N, (size_t)K);
  CSE_REQUIRE(scratch_bytes >= s.total, "linear_bwd_tc: scratch too small (%zu < %zu bytes)", scratch_bytes, s.total);
  cudaStream_t st = (cudaStream_t)stream;
  const int Mpad = (int)align_up((size_t)M, 64);
  if (dbias != nullptr) {
    if (launch_colsum(dC, N, M, N, dbias, st)) return 1;
  }
  if (dA != nullptr) {
    if (launch_f32_to_bf16(dC, s.dC16, (size_t)M * N, st)) return 1;
    if (launch_transpose(W, N, K, s.WT32, st)) return 1;                     // W^T [K,N] fp32
    if (launch_f32_to_bf16(s.WT32, s.WT16, (size_t)N * K, st)) return 1;
    if (launch_gemm_tc(s.dC16, N, s.WT16, nullptr, 0.f, nullptr, dA, ldda, M, K, N, 0, dA_fp32 ? 1 : 0, st)) return 1;
  }
  if (dW != nullptr) {
    if (launch_transpose_cast(dC, 0, N, M, N, Mpad, s.XT, st)) return 1;     // (dC^T) [N, Mpad]
    if (launch_transpose_cast(A, a_is_bf16, lda, M, K, Mpad, s.YT, st)) return 1;  // (A^T) [K, Mpad]
    // dW[N,K] += XT YT^T : in-place accumulation through the TMA reduce-add epilogue
    if (launch_gemm_tc(s.XT, Mpad, s.YT, nullptr, 0.f, dW, dW, K, N, K, Mpad, 0, 1, st)) return 1;
  }
  return 0;
}

}  // extern "C"
