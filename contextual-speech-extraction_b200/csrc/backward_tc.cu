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
#include <cstdlib>

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

// CSE_WGRAD_TRANSPOSE=1: the first version, which fed the forward (K-major) GEMM transposed bf16 copies of W, dC and A
// (A/B aid)
static bool wgrad_transposed_copies() {
  static const bool on = []() {
    const char* e = getenv("CSE_WGRAD_TRANSPOSE");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

// W16: optional bf16 copy of W [N,K] (the forward's packed weights); NULL = cast W here.
// relu: optional saved bf16 output [M,N] of a ReLU that followed this Linear: dC is masked by it on the fly
// (needs dbias, i.e. the fused bias-gradient / cast pass).
static int linear_bwd_tc(const void* A, int a_is_bf16, int lda, const float* W, const bf16* W16, const float* dC,
                         const bf16* relu, int M, int N, int K, void* dA, int dA_fp32, int ldda, float* dW,
                         float* dbias, void* scratch, size_t scratch_bytes, void* stream) {
  CSE_REQUIRE(relu == nullptr || dbias != nullptr, "linear_bwd_tc: the ReLU mask rides on the bias-gradient pass");
  CSE_REQUIRE(A && W && dC && scratch && M > 0, "linear_bwd_tc: bad argument");
  CSE_REQUIRE(N % 128 == 0 && K % 128 == 0, "linear_bwd_tc: N and K must be multiples of 128 (N=%d K=%d)", N, K);
  CSE_REQUIRE(dbias == nullptr || N % 256 == 0, "linear_bwd_tc: bias gradient needs N %% 256 == 0 (N=%d)", N);
  CSE_REQUIRE(((uintptr_t)scratch & 255) == 0, "linear_bwd_tc: scratch must be 256-byte aligned");
  const TcBwdScratch s = carve_tc_bwd((char*)scratch, (size_t)M, (size_t)N, (size_t)K);
  CSE_REQUIRE(scratch_bytes >= s.total, "linear_bwd_tc: scratch too small (%zu < %zu bytes)", scratch_bytes, s.total);
  cudaStream_t st = (cudaStream_t)stream;
  const int Mpad = (int)align_up((size_t)M, 64);
  const bool transposed_copies = wgrad_transposed_copies();
  CSE_REQUIRE(relu == nullptr || !transposed_copies, "linear_bwd_tc: no ReLU mask in the transposed-copies mode");
  bool have_dc16 = false;  // the bf16 copy of dC feeds both the dgrad and the wgrad
  if (dbias != nullptr) {  // bias gradient and the bf16 copy in one pass over dC
    if (launch_colsum_cast(dC, relu, M, N, dbias, s.dC16, st)) return 1;
    have_dc16 = true;
  }
  if (dA != nullptr) {
    if (!have_dc16 && launch_f32_to_bf16(dC, s.dC16, (size_t)M * N, st)) return 1;
    have_dc16 = true;
    if (!transposed_copies) {
      // dA = dC W with W as stored ([N,K] row-major = the MN-major B operand): no transposed weight copy
      if (W16 == nullptr) {
        if (launch_f32_to_bf16(W, s.WT16, (size_t)N * K, st)) return 1;
        W16 = s.WT16;
      }
      if (launch_gemm_tc_dgrad(s.dC16, N, W16, K, dA, ldda, dA_fp32 ? 1 : 0, M, K, N, st)) return 1;
    } else {
      if (launch_transpose(W, N, K, s.WT32, st)) return 1;                     // W^T [K,N] fp32
      if (launch_f32_to_bf16(s.WT32, s.WT16, (size_t)N * K, st)) return 1;
      if (launch_gemm_tc(s.dC16, N, s.WT16, nullptr, 0.f, nullptr, dA, ldda, M, K, N, 0, dA_fp32 ? 1 : 0, st)) return 1;
    }
  }
  if (dW != nullptr) {
    if (!transposed_copies) {
      // dW[N,K] += dC^T A straight from the row-major bf16 dC [M,N] and A [M,K]: MN-major UMMA operands, the token
      // dimension split over the CTA pairs, TMA reduce-add into dW
      if (!have_dc16 && launch_f32_to_bf16(dC, s.dC16, (size_t)M * N, st)) return 1;
      const bf16* A16 = (const bf16*)A;
      int lda16 = lda;
      if (!a_is_bf16) {  // (only the stack's first linear sees an fp32 input)
        CSE_REQUIRE(lda == K, "linear_bwd_tc: an fp32 A must be contiguous (lda=%d K=%d)", lda, K);
        if (launch_f32_to_bf16((const float*)A, s.YT, (size_t)M * K, st)) return 1;
        A16 = s.YT;
        lda16 = K;
      }
      if (launch_gemm_tc_wgrad(s.dC16, N, A16, lda16, dW, K, N, K, M, st)) return 1;
    } else {
      if (launch_transpose_cast(dC, 0, N, M, N, Mpad, s.XT, st)) return 1;     // (dC^T) [N, Mpad]
      if (launch_transpose_cast(A, a_is_bf16, lda, M, K, Mpad, s.YT, st)) return 1;  // (A^T) [K, Mpad]
      // dW[N,K] += XT YT^T : in-place accumulation through the TMA reduce-add epilogue, the token dimension split
      // over the CTA pairs (2-8 output tiles alone would leave 140 SMs idle)
      if (launch_gemm_tc_splitk(s.XT, Mpad, s.YT, Mpad, dW, K, N, K, Mpad, st)) return 1;
    }
  }
  return 0;
}

int cse_linear_bwd_tc(const void* A, int a_is_bf16, int lda, const float* W, const float* dC, int M, int N,
                      int K, void* dA, int dA_fp32, int ldda, float* dW, float* dbias, void* scratch,
                      size_t scratch_bytes, void* stream) {
  return linear_bwd_tc(A, a_is_bf16, lda, W, nullptr, dC, nullptr, M, N, K, dA, dA_fp32, ldda, dW, dbias, scratch, scratch_bytes,
                       stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// EXPERIMENTAL, same status as above: one transformer layer backward in the performance mode —
// bf16 recompute with the forward's own kernels (unfused FFN so that the hidden activation exists),
// tensor-core dgrad / wgrad through cse_linear_bwd_tc, fp32 LayerNorm / attention / residual gradients.
// ---------------------------------------------------------------------------------------------
namespace cse {

__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst,
                                                          size_t n8) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256)
    st8(dst + i * 8, ld8(src + i * 8));
}

static int launch_bf16_to_f32(const bf16* src, float* dst, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  const size_t n8 = n / 8;  // callers pass multiples of 256
  const int grid = (int)min((size_t)148 * 8, (n8 + 255) / 256);
  bf16_to_f32_kernel<<<grid, 256, 0, st>>>(src, dst, n8);
  return check_launch("bf16_to_f32_kernel");
}

// d[i] = F[i] > 0 ? d[i] : 0 with the saved activation in bf16 and the gradient in fp32
__global__ void __launch_bounds__(256) relu_bwd_mixed_kernel(const bf16* __restrict__ F, float* __restrict__ d,
                                                             size_t n8) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
    const f8 f = ld8(F + i * 8);
    f8 g = ld8(d + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = f.v[k] > 0.f ? g.v[k] : 0.f;
    st8(d + i * 8, g);
  }
}

static int launch_relu_bwd_mixed(const bf16* F, float* d, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  const size_t n8 = n / 8;
  const int grid = (int)min((size_t)148 * 8, (n8 + 255) / 256);
  relu_bwd_mixed_kernel<<<grid, 256, 0, st>>>(F, d, n8);
  return check_launch("relu_bwd_mixed_kernel");
}

struct LayerWs16 {
  bf16 *H, *H1, *QKV, *AO, *F1;   // H: norm2 output, H1: norm1 output (kept for in_proj's weight gradient)
  float *Rmid, *dBig, *dH, *QKV32, *AO32;
  char* lin;          // scratch of cse_linear_bwd_tc, sized for the largest of the four linears
  size_t lin_bytes, total;
};

static LayerWs16 carve_layer_ws16(char* ws, size_t M) {
  LayerWs16 w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = ws ? ws + off : nullptr;
    off = align_up(off + bytes, 256);
    return p;
  };
  w.H = (bf16*)take(M * kN * 2);
  w.H1 = (bf16*)take(M * kN * 2);
  w.QKV = (bf16*)take(M * 3 * kN * 2);
  w.AO = (bf16*)take(M * kN * 2);
  w.F1 = (bf16*)take(M * kFfn * 2);
  w.Rmid = (float*)take(M * kN * 4);
  w.dBig = (float*)take(M * kFfn * 4);
  w.dH = (float*)take(M * kN * 4);
  w.QKV32 = (float*)take(M * 3 * kN * 4);
  w.AO32 = (float*)take(M * kN * 4);
  size_t lin = 0;
  const size_t shapes[4][2] = {{(size_t)kN, (size_t)kFfn}, {(size_t)kFfn, (size_t)kN}, {(size_t)kN, (size_t)kN},
                               {(size_t)3 * kN, (size_t)kN}};
  for (auto& s : shapes) {
    const size_t b = carve_tc_bwd(nullptr, M, s[0], s[1]).total;
    if (b > lin) lin = b;
  }
  w.lin_bytes = lin;
  w.lin = take(lin);
  w.total = off;
  return w;
}

}  // namespace cse

extern "C" {

size_t cse_layer_bwd_bf16_workspace_bytes(int nseq, int n) {
  if (nseq <= 0 || n <= 0) return 0;
  return carve_layer_ws16(nullptr, (size_t)nseq * n).total;
}

int cse_layer_bwd_bf16(const cse_layer_params* p, const cse_layer_grads* g, const float* R_in, float* dR,
                       int nseq, int n, void* workspace, size_t workspace_bytes, void* stream) {
  CSE_REQUIRE(p && g && R_in && dR && workspace && nseq > 0 && n > 0, "layer_bwd_bf16: bad argument");
  CSE_REQUIRE(p->in_proj_w_bf16 && p->out_proj_w_bf16 && p->ffn1_w_bf16 && p->ffn2_w_bf16,
              "layer_bwd_bf16: bf16 weights missing (cse_pack_bf16)");
  CSE_REQUIRE(((uintptr_t)workspace & 255) == 0, "layer_bwd_bf16: workspace must be 256-byte aligned");
  CSE_REQUIRE((long long)nseq * n <= 2147483647LL / kFfn, "layer_bwd_bf16: %d x %d rows overflow the 32-bit tile index", nseq, n);
  const int M = nseq * n;
  const LayerWs16 w = carve_layer_ws16((char*)workspace, (size_t)M);
  CSE_REQUIRE(workspace_bytes >= w.total, "layer_bwd_bf16: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t row_bytes = (size_t)M * kN * sizeof(float);

  // ---- recompute in the performance mode (the launch sequence of abi.cu:run_stack with the unfused FFN) ----
  if (launch_layernorm(R_in, p->ln1_g, p->ln1_b, M, 1e-6f, CSE_BF16, w.H1, st)) return 1;
  if (launch_gemm_tc(w.H1, kN, (const bf16*)p->in_proj_w_bf16, p->in_proj_b, 1.f, nullptr, w.QKV, 3 * kN, M, 3 * kN, kN, 0, 0, st)) return 1;
  if (launch_attention(w.QKV, nseq, n, CSE_BF16, w.AO, st)) return 1;
  CSE_CUDA(cudaMemcpyAsync(w.Rmid, R_in, row_bytes, cudaMemcpyDeviceToDevice, st));
  if (launch_gemm_tc(w.AO, kN, (const bf16*)p->out_proj_w_bf16, p->out_proj_b, 1.f, w.Rmid, w.Rmid, kN, M, kN, kN, 0, 1, st)) return 1;
  if (launch_layernorm(w.Rmid, p->ln2_g, p->ln2_b, M, 1e-6f, CSE_BF16, w.H, st)) return 1;
  if (launch_gemm_tc(w.H, kN, (const bf16*)p->ffn1_w_bf16, p->ffn1_b, 1.f, nullptr, w.F1, kFfn, M, kFfn, kN, 1, 0, st)) return 1;

  // ---- FFN sub-block ----
  if (linear_bwd_tc(w.F1, 1, kFfn, p->ffn2_w, (const bf16*)p->ffn2_w_bf16, dR, nullptr, M, kN, kFfn, w.dBig, 1, kFfn, g->ffn2_w, g->ffn2_b, w.lin,
                        w.lin_bytes, stream)) return 1;
  // (the ReLU between ffn.0 and ffn.3 is differentiated inside the next call's bias-gradient / cast pass)
  const bf16* relu_mask = w.F1;
  if (wgrad_transposed_copies()) {
    if (launch_relu_bwd_mixed(w.F1, w.dBig, (size_t)M * kFfn, st)) return 1;
    relu_mask = nullptr;
  }
  if (linear_bwd_tc(w.H, 1, kN, p->ffn1_w, (const bf16*)p->ffn1_w_bf16, w.dBig, relu_mask, M, kFfn, kN, w.dH, 1, kN, g->ffn1_w, g->ffn1_b, w.lin,
                        w.lin_bytes, stream)) return 1;
  if (launch_layernorm_bwd(w.Rmid, p->ln2_g, w.dH, M, 1e-6f, dR, 1, g->ln2_g, g->ln2_b, st)) return 1;

  // ---- attention sub-block ----
  if (linear_bwd_tc(w.AO, 1, kN, p->out_proj_w, (const bf16*)p->out_proj_w_bf16, dR, nullptr, M, kN, kN, w.dH, 1, kN, g->out_proj_w, g->out_proj_b, w.lin,
                        w.lin_bytes, stream)) return 1;
  float* dQKV = w.dBig;  // [M,768] fp32
  if (n <= 256) {  // tensor-core attention backward straight from the bf16 recompute (attention_bwd_mma.cu)
    if (launch_attention_bwd_bf16(w.QKV, w.AO, w.dH, nseq, n, dQKV, st)) return 1;
  } else {         // longer sequences (inter stack beyond ~30 s of audio): the fp32 kernel
    if (launch_bf16_to_f32(w.QKV, w.QKV32, (size_t)M * 3 * kN, st)) return 1;
    if (launch_bf16_to_f32(w.AO, w.AO32, (size_t)M * kN, st)) return 1;
    if (launch_attention_bwd(w.QKV32, w.AO32, w.dH, nseq, n, dQKV, st)) return 1;
  }
  // (norm1's output is still in H1 from the recompute above)
  if (linear_bwd_tc(w.H1, 1, kN, p->in_proj_w, (const bf16*)p->in_proj_w_bf16, dQKV, nullptr, M, 3 * kN, kN, w.dH, 1, kN, g->in_proj_w, g->in_proj_b, w.lin,
                        w.lin_bytes, stream)) return 1;
  return launch_layernorm_bwd(R_in, p->ln1_g, w.dH, M, 1e-6f, dR, 1, g->ln1_g, g->ln1_b, st);
}

}  // extern "C"
