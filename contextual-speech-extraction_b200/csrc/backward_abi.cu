// backward_abi.cu — extern "C" entry points of the training path (include/cse_b200.h, "backward").
//
// cse_layer_bwd differentiates TransformerEncoderLayer.forward (CSE_transformer.py:385-416) the way
// autograd does for the reference's `loss.backward()` (train_ContSep.py:402-419), with activation
// checkpointing at layer granularity: the caller keeps only the layer's input residual stream, the
// layer is recomputed in fp32 and then walked backwards.  Parameter gradients accumulate (+=).
#include "common.cuh"

namespace cse {

struct LayerWs {
  float *H, *QKV, *AO, *Rmid, *F1, *dBig, *dH, *WT;
  size_t total;
};

static LayerWs carve_layer_ws(char* ws, size_t M) {
  LayerWs w;
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = ws ? (float*)(ws + off) : nullptr;
    off = align_up(off + floats * sizeof(float), 256);
    return p;
  };
  w.H = take(M * kN);
  w.QKV = take(M * 3 * kN);
  w.AO = take(M * kN);
  w.Rmid = take(M * kN);
  w.F1 = take(M * kFfn);
  w.dBig = take(M * kFfn);
  w.dH = take(M * kN);
  w.WT = take((size_t)kFfn * kN);
  w.total = off;
  return w;
}

// fp32 forward of one layer, keeping every intermediate the backward pass needs.
// R_in is left untouched; the layer output would be Rmid + F1 W2^T + b2 (written to R_out if given).
static int layer_forward_f32(const cse_layer_params& lp, const float* R_in, float* R_out, int nseq, int n,
                             const LayerWs& w, bool need_hidden, cudaStream_t st) {
  const int M = nseq * n;
  if (launch_layernorm(R_in, lp.ln1_g, lp.ln1_b, M, 1e-6f, CSE_FP32, w.H, st)) return 1;
  if (launch_gemm_simt(w.H, kN, lp.in_proj_w, lp.in_proj_b, 1.f, nullptr, w.QKV, 3 * kN, M, 3 * kN, kN, 0, st)) return 1;
  if (launch_attention(w.QKV, nseq, n, CSE_FP32, w.AO, st)) return 1;
  if (launch_gemm_simt(w.AO, kN, lp.out_proj_w, lp.out_proj_b, 1.f, R_in, w.Rmid, kN, M, kN, kN, 0, st)) return 1;
  if (launch_layernorm(w.Rmid, lp.ln2_g, lp.ln2_b, M, 1e-6f, CSE_FP32, w.H, st)) return 1;
  if (need_hidden || R_out != nullptr) {
    if (launch_gemm_simt(w.H, kN, lp.ffn1_w, lp.ffn1_b, 1.f, nullptr, w.F1, kFfn, M, kFfn, kN, 1, st)) return 1;
  }
  if (R_out != nullptr) {
    if (launch_gemm_simt(w.F1, kFfn, lp.ffn2_w, lp.ffn2_b, 1.f, w.Rmid, R_out, kN, M, kN, kFfn, 0, st)) return 1;
  }
  return 0;
}

// nn.Linear backward, see cse_linear_bwd.
static int linear_bwd(const float* A, int lda, const float* W, const float* dC, int lddc, int M, int N,
                      int K, float* dA, int ldda, float* dW, float* dbias, float* scratch_wt,
                      cudaStream_t st) {
  if (dW != nullptr) {
    if (launch_wgrad(dC, lddc, A, lda, M, N, K, dW, st)) return 1;
  }
  if (dbias != nullptr) {
    if (launch_colsum(dC, lddc, M, N, dbias, st)) return 1;
  }
  if (dA != nullptr) {
    CSE_REQUIRE(scratch_wt != nullptr, "linear_bwd: dA requested but scratch_wt is NULL");
    if (launch_transpose(W, N, K, scratch_wt, st)) return 1;           // W^T [K,N]
    // dA[M,K] = dC[M,N] (W^T)[K,N]^T : the forward GEMM with W^T as the weight
    if (launch_gemm_simt(dC, lddc, scratch_wt, nullptr, 0.f, nullptr, dA, ldda, M, K, N, 0, st)) return 1;
  }
  return 0;
}

}  // namespace cse

using namespace cse;

extern "C" {

int cse_si_snr_bwd(const float* source, const float* estimate, const float* g_out, int B, int T, int C,
                   float* d_source, float* d_estimate, void* stream) {
  CSE_REQUIRE(source && estimate && g_out && B > 0 && T > 0, "si_snr_bwd: bad argument");
  return launch_si_snr_bwd(source, estimate, B, T, C, 0, g_out, nullptr, d_source, d_estimate,
                           (cudaStream_t)stream);
}

int cse_pit_si_snr_bwd(const float* source, const float* estimate_source, const float* g_loss,
                       const int* perm, int B, int T, int C, float* d_source, float* d_estimate_source,
                       void* stream) {
  CSE_REQUIRE(source && estimate_source && g_loss && perm && B > 0 && T > 0, "pit_si_snr_bwd: bad argument");
  return launch_si_snr_bwd(source, estimate_source, B, T, C, 1, g_loss, perm, d_source, d_estimate_source,
                           (cudaStream_t)stream);
}

int cse_tm_si_snr_bwd(const float* preds, const float* target, const float* g_out, int B, int T,
                      float* d_preds, float* d_target, void* stream) {
  CSE_REQUIRE(preds && target && g_out && B > 0 && T > 0, "tm_si_snr_bwd: bad argument");
  return launch_tm_si_snr_bwd(preds, target, B, T, g_out, d_preds, d_target, (cudaStream_t)stream);
}

int cse_linear_bwd(const float* A, int lda, const float* W, const float* dC, int lddc, int M, int N, int K,
                   float* dA, int ldda, float* dW, float* dbias, float* scratch_wt, void* stream) {
  CSE_REQUIRE(A && W && dC, "linear_bwd: NULL argument");
  CSE_REQUIRE(N % 128 == 0 && K % 128 == 0, "linear_bwd: N and K must be multiples of 128 (N=%d K=%d)", N, K);
  return linear_bwd(A, lda, W, dC, lddc, M, N, K, dA, ldda, dW, dbias, scratch_wt, (cudaStream_t)stream);
}

int cse_layernorm_bwd(const float* x, const float* g, const float* dy, int M, float eps, int accumulate,
                      float* dx, float* dg, float* db, void* stream) {
  CSE_REQUIRE(x && g && dy && dx, "layernorm_bwd: NULL argument");
  return launch_layernorm_bwd(x, g, dy, M, eps, dx, accumulate, dg, db, (cudaStream_t)stream);
}

int cse_attention_bwd(const float* qkv, const float* out, const float* d_out, int nseq, int n, float* d_qkv,
                      void* stream) {
  CSE_REQUIRE(qkv && out && d_out && d_qkv, "attention_bwd: NULL argument");
  return launch_attention_bwd(qkv, out, d_out, nseq, n, d_qkv, (cudaStream_t)stream);
}

int cse_attention_bwd_bf16(const void* qkv_bf16, const void* out_bf16, const float* d_out, int nseq, int n,
                           float* d_qkv, void* stream) {
  CSE_REQUIRE(qkv_bf16 && out_bf16 && d_out && d_qkv, "attention_bwd_bf16: NULL argument");
  return launch_attention_bwd_bf16((const bf16*)qkv_bf16, (const bf16*)out_bf16, d_out, nseq, n, d_qkv,
                                   (cudaStream_t)stream);
}

size_t cse_layer_workspace_bytes(int nseq, int n) {
  if (nseq <= 0 || n <= 0) return 0;
  return carve_layer_ws(nullptr, (size_t)nseq * n).total;
}

int cse_layer_fwd(const cse_layer_params* p, float* R, int nseq, int n, int precision, void* workspace,
                  size_t workspace_bytes, void* stream) {
  CSE_REQUIRE(p && R && workspace && nseq > 0 && n > 0, "layer_fwd: bad argument");
  CSE_REQUIRE(((uintptr_t)workspace & 255) == 0, "layer_fwd: workspace must be 256-byte aligned");
  const size_t M = (size_t)nseq * n;
  const LayerWs w = carve_layer_ws((char*)workspace, M);
  CSE_REQUIRE(workspace_bytes >= w.total, "layer_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == CSE_FP32) return layer_forward_f32(*p, R, R, nseq, n, w, true, st);
  CSE_REQUIRE(precision == CSE_BF16, "layer_fwd: unknown precision %d", precision);
  CSE_REQUIRE(p->in_proj_w_bf16 && p->out_proj_w_bf16 && p->ffn1_w_bf16 && p->ffn2_w_bf16,
              "layer_fwd: bf16 weights missing");
  // the default performance-mode launch sequence of abi.cu:run_stack; bf16 buffers alias the fp32 carve
  bf16* H = (bf16*)w.H;
  bf16* QKV = (bf16*)w.QKV;
  bf16* AO = (bf16*)w.AO;
  const int Mi = (int)M;
  if (launch_layernorm(R, p->ln1_g, p->ln1_b, Mi, 1e-6f, CSE_BF16, H, st)) return 1;
  if (launch_gemm_tc(H, kN, (const bf16*)p->in_proj_w_bf16, p->in_proj_b, 1.f, nullptr, QKV, 3 * kN, Mi, 3 * kN, kN, 0, 0, st)) return 1;
  if (launch_attention(QKV, nseq, n, CSE_BF16, AO, st)) return 1;
  if (launch_gemm_tc(AO, kN, (const bf16*)p->out_proj_w_bf16, p->out_proj_b, 1.f, R, R, kN, Mi, kN, kN, 0, 1, st)) return 1;
  // norm2 rides along in the feed-forward kernel's LayerNorm warps (ffn_tc.cu; H receives norm2(R)) — same results as
  // layernorm_kernel + launch_ffn_tc
  return launch_ffn_tc_ln(R, p->ln2_g, p->ln2_b, 1e-6f, H, (const bf16*)p->ffn1_w_bf16, p->ffn1_b,
                          (const bf16*)p->ffn2_w_bf16, p->ffn2_b, nullptr, nullptr, nullptr, Mi, st);
}

int cse_layer_bwd(const cse_layer_params* p, const cse_layer_grads* g, const float* R_in, float* dR,
                  int nseq, int n, void* workspace, size_t workspace_bytes, void* stream) {
  CSE_REQUIRE(p && g && R_in && dR && workspace && nseq > 0 && n > 0, "layer_bwd: bad argument");
  CSE_REQUIRE(((uintptr_t)workspace & 255) == 0, "layer_bwd: workspace must be 256-byte aligned");
  CSE_REQUIRE((long long)nseq * n <= 2147483647LL / kFfn, "layer_bwd: %d x %d rows overflow the 32-bit tile index", nseq, n);
  const int M = nseq * n;
  const LayerWs w = carve_layer_ws((char*)workspace, (size_t)M);
  CSE_REQUIRE(workspace_bytes >= w.total, "layer_bwd: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;

  // recompute: H = LN2(Rmid), QKV, AO, Rmid, F1 = relu(H W1^T + b1)
  if (layer_forward_f32(*p, R_in, nullptr, nseq, n, w, true, st)) return 1;

  // ---- FFN sub-block: R_out = Rmid + F1 W2^T + b2 ----
  // dF1 = dR W2 ; dW2 += dR^T F1 ; db2 += colsum(dR)
  if (linear_bwd(w.F1, kFfn, p->ffn2_w, dR, kN, M, kN, kFfn, w.dBig, kFfn, g->ffn2_w, g->ffn2_b, w.WT, st)) return 1;
  if (launch_relu_bwd(w.F1, w.dBig, (size_t)M * kFfn, st)) return 1;
  // dH2 = dF1 W1 ; dW1 += dF1^T H2 ; db1 += colsum(dF1)
  if (linear_bwd(w.H, kN, p->ffn1_w, w.dBig, kFfn, M, kFfn, kN, w.dH, kN, g->ffn1_w, g->ffn1_b, w.WT, st)) return 1;
  // dR (now dL/dRmid) += LN2'(Rmid) dH2
  if (launch_layernorm_bwd(w.Rmid, p->ln2_g, w.dH, M, 1e-6f, dR, 1, g->ln2_g, g->ln2_b, st)) return 1;

  // ---- attention sub-block: Rmid = R_in + AO Wo^T + bo ----
  if (linear_bwd(w.AO, kN, p->out_proj_w, dR, kN, M, kN, kN, w.dH, kN, g->out_proj_w, g->out_proj_b, w.WT, st)) return 1;
  float* dQKV = w.dBig;  // [M,768]
  if (launch_attention_bwd(w.QKV, w.AO, w.dH, nseq, n, dQKV, st)) return 1;
  // H1 = LN1(R_in) again (H was overwritten by LN2's output)
  if (launch_layernorm(R_in, p->ln1_g, p->ln1_b, M, 1e-6f, CSE_FP32, w.H, st)) return 1;
  if (linear_bwd(w.H, kN, p->in_proj_w, dQKV, 3 * kN, M, 3 * kN, kN, w.dH, kN, g->in_proj_w, g->in_proj_b, w.WT, st)) return 1;
  // dR (now dL/dR_in) += LN1'(R_in) dH1
  return launch_layernorm_bwd(R_in, p->ln1_g, w.dH, M, 1e-6f, dR, 1, g->ln1_g, g->ln1_b, st);
}

}  // extern "C"
