/* cse_b200.h — C ABI of the B200-native Sepformer hot path.
 *
 * The reference (miraodasilva/contextual-speech-extraction) is pure Python and exposes no FFI;
 * its seam for this path is Python class composition.  This header is the
 * boundary a maintainer would bind instead: plain pointers and sizes, caller-owned DEVICE
 * memory, no hidden allocation, kernels launched on the caller's stream.  Each entry point
 * names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; cse_last_error() gives the text
 *     (thread-local).  Nothing aborts the process.
 *   - all pointers are device pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - activations are CHANNELS-LAST: a [B,N,L] reference tensor is stored [B,L,N] (N = 256).
 *   - `precision`: CSE_FP32 = true-fp32 arithmetic (parity mode, <=1e-4 rel-L2 vs the fp32
 *     reference); CSE_BF16 = bf16 tensor-core operands, fp32 accumulate / residual / norms
 *     (the reference's autocast policy; performance mode).
 */
#ifndef CSE_B200_H
#define CSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSE_API __attribute__((visibility("default")))

#define CSE_FP32 0
#define CSE_BF16 1

#define CSE_N 256        /* channels            ContSep.py:10,13-14 */
#define CSE_K 250        /* chunk length        ContSep.py:16       */
#define CSE_LAYERS 8     /* layers per stack    ContSep.py:18       */
#define CSE_BLOCKS 2     /* dual-path blocks    ContSep.py:15       */
#define CSE_FFN 1024     /* FFN width           ContSep.py:21       */
#define CSE_HEADS 8      /* attention heads     ContSep.py:20       */
#define CSE_CTX 4096     /* context width       ContSep.py:8        */

/* One pre-norm transformer layer (CSE_transformer.py:253-421).  fp32 master pointers are the
 * nn.Parameter storages; *_bf16 are the packed copies written by cse_pack_bf16 (CSE_BF16 only). */
typedef struct {
  const float* in_proj_w;   /* [768,256]  self_att.att.in_proj_weight */
  const float* in_proj_b;   /* [768] */
  const float* out_proj_w;  /* [256,256]  self_att.att.out_proj.weight */
  const float* out_proj_b;  /* [256] */
  const float* ffn1_w;      /* [1024,256] pos_ffn.ffn.0.weight */
  const float* ffn1_b;      /* [1024] */
  const float* ffn2_w;      /* [256,1024] pos_ffn.ffn.3.weight */
  const float* ffn2_b;      /* [256] */
  const float* ln1_g;       /* norm1.norm.weight */
  const float* ln1_b;
  const float* ln2_g;       /* norm2.norm.weight */
  const float* ln2_b;
  const void* in_proj_w_bf16;
  const void* out_proj_w_bf16;
  const void* ffn1_w_bf16;
  const void* ffn2_w_bf16;
} cse_layer_params;

/* SBTransformerBlock_CSE (CSE_transformer.py:11-106): 8 layers + final LayerNorm + PE table. */
typedef struct {
  cse_layer_params layer[CSE_LAYERS];
  const float* final_g;     /* mdl.norm.norm.weight */
  const float* final_b;
  const float* pe;          /* pos_enc.pe [2500,256] */
} cse_stack_params;

/* Dual_Computation_Block_CSE (ContSep.py:372-533). */
typedef struct {
  cse_stack_params intra;
  cse_stack_params inter;
  const float* intra_norm_g;  /* GroupNorm(1,256) affine */
  const float* intra_norm_b;
  const float* inter_norm_g;
  const float* inter_norm_b;
  const float* intra_map_w;   /* intra_context_mapper.weight [256,4096] or NULL (c == 0) */
  const float* intra_map_b;
  const float* inter_map_w;
  const float* inter_map_b;
} cse_block_params;

/* Encoder + Dual_Path_Model_CSE + Decoder (ContSep.py:7-100, 103-370). */
typedef struct {
  const float* enc_w;       /* encoder.conv1d.weight [256,16] */
  const float* norm_g;      /* masknet.norm */
  const float* norm_b;
  const float* conv1d_w;    /* masknet.conv1d.weight [256,256] */
  cse_block_params block[CSE_BLOCKS];
  const float* prelu;       /* masknet.prelu.weight [1] */
  const float* conv2d_w;    /* masknet.conv2d.weight [spk*256,256] */
  const float* conv2d_b;    /* [spk*256] */
  const float* out_w;       /* masknet.output.0 [256,256] */
  const float* out_b;
  const float* gate_w;      /* masknet.output_gate.0 */
  const float* gate_b;
  const float* end_w;       /* masknet.end_conv1x1.weight [256,256] */
  const float* dec_w;       /* decoder.weight [256,16] */
  const void* conv1d_w_bf16;
  const void* conv2d_w_bf16;
  const void* out_w_bf16;
  const void* gate_w_bf16;
  const void* end_w_bf16;
} cse_params;

/* ---- library ---- */
CSE_API int cse_version(void);
CSE_API const char* cse_last_error(void);

/* Kernel launches issued by this library since load (bench.py's gpu_launches). */
CSE_API long long cse_launch_count(void);
/* Test hook: 0 = automatic choice, 1 = force the mma.sync online-softmax bf16 attention kernel,
 * 2 = force the tcgen05 kernel v1 (n <= 256), 4 = force the tcgen05 kernel v4 — so every kernel can be checked at
 * the same shapes. */
CSE_API int cse_debug_force_mma_attention(int on);
/* Debug: device buffer of 64 x 16 int64 that CTA 0 of the tcgen05 attention kernel fills with clock64 stamps of its
 * pipeline events (tools/attn_trace.py); NULL turns it off. */
CSE_API int cse_debug_attention_trace(long long* device_buffer);
/* Optional device timing per kernel class (0 tcgen05 GEMM, 1 attention, 2 LayerNorm, 3 fp32 SIMT
 * GEMM, 4 fused feed-forward kernel — counted in class 0 when n_classes == 4): while enabled, each launch of those classes is bracketed by a CUDA event pair on its
 * stream; cse_profile_collect sums and clears them (synchronises on the recorded events). */
CSE_API int cse_profile_enable(int on);
CSE_API int cse_profile_collect(double* ms_by_class, long long* launches_by_class, int n_classes);

/* ---- shape algebra (ContSep.py:270-335 `_padding/_Segmentation`; encoder/decoder lengths) ---- */
typedef struct {
  int B, T, c, spk;
  int L;      /* encoder frames (T-16)/8+1 */
  int gap;    /* ContSep.py:287 */
  int S;      /* chunks */
  int T_est;  /* 8(L-1)+16 */
} cse_shape;
CSE_API int cse_path_shape(int B, int T, int c, int spk, cse_shape* out);

/* ---- whole path ---- */
/* Bytes of scratch cse_forward needs for this call (caller allocates, 256-byte aligned). */
CSE_API size_t cse_workspace_bytes(int B, int T, int c, int n_masks, int precision);

/* Number of bf16 elements cse_pack_bf16 writes; then fills the *_bf16 members of `p`
 * (host struct) with pointers into `packed`.  Replaces autocast's per-call weight casts. */
CSE_API size_t cse_pack_bf16_elems(int n_masks);
CSE_API int cse_pack_bf16(cse_params* p_host, int n_masks, void* packed, size_t packed_elems,
                          void* stream);

/* Sepformer.forward minus the two host-side Linear layers (ContSep.py:53-100, ContExt.py:54-129,
 * sepformer.py:42-81): encoder -> masknet(ctx) -> mask * mix_w -> decoder -> pad/trim.
 *   mix  [B,T] fp32;  ctx [B,c,4096] fp32 or NULL when c == 0
 *   n_masks: masks to estimate AND decode (spk for Sepformer/ContSep; 1 for ContExt, which uses
 *            only est_mask[0], ContExt.py:113-119 — conv2d rows [0,256) suffice)
 *   est  [B,T,n_masks] fp32 out
 *   pred_head [B,256] fp32 out or NULL (ContSep.py:516-517, last block's inter token 0 mean) */
CSE_API int cse_forward(const cse_params* p_host, const float* mix, const float* ctx,
                        int B, int T, int c, int n_masks, int precision,
                        float* est, float* pred_head,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers (pinned recommended): H2D of mix/ctx, forward, D2H of est and
 * pred_head, stream-ordered; returns after the stream is synchronised.  The reference-facing
 * end-to-end entry (`model(mix.cuda(), ctx).cpu()` in test.py:231-245). */
CSE_API int cse_forward_host(const cse_params* p_host, const float* mix_host, const float* ctx_host,
                             int B, int T, int c, int n_masks, int precision,
                             float* est_host, float* pred_head_host,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Pipelined host entry: `depth` forwards in flight (serving / eval loops, test.py:231-245 per batch).
 * Each slot owns one cse_workspace_bytes() region of the caller's workspace (device staging of its inputs and
 * outputs included) and a CUDA graph of the forward captured at create time; H2D, forward and D2H run on three
 * internal streams chained by events, so the copy-in of step i+1 and the copy-out of step i-1 overlap the forward
 * of step i.  The graph bakes in the pointers of `p_host` (bf16 pack included): keep them alive and unchanged, and
 * re-create the pipeline after a parameter re-allocation.  Host buffers must be pinned and stay valid until the
 * slot is waited on.  submit() on a slot that is still in flight waits for it first.
 *   cse_pipeline_submit  -> *slot: pass it to cse_pipeline_wait to collect est_host / pred_head_host. */
typedef struct cse_pipeline cse_pipeline;
CSE_API size_t cse_pipeline_workspace_bytes(int B, int T, int c, int n_masks, int precision, int depth);
CSE_API int cse_pipeline_create(const cse_params* p_host, int B, int T, int c, int n_masks, int precision,
                                int depth, void* workspace, size_t workspace_bytes, cse_pipeline** out);
CSE_API int cse_pipeline_submit(cse_pipeline* pipe, const float* mix_host, const float* ctx_host,
                                float* est_host, float* pred_head_host, int* slot);
CSE_API int cse_pipeline_wait(cse_pipeline* pipe, int slot);
CSE_API int cse_pipeline_destroy(cse_pipeline* pipe);

/* Dual_Path_Model_CSE.forward on its own (ContSep.py:205-268; speechbrain Dual_Path_Model when
 * c == 0): E = mix_w channels-last [B,L,256] in the activation dtype of `precision` ->
 * mask [B,L,n_masks,256] fp32 (post-ReLU; reference layout mask[s,b,n,l]) and pred_head. */
CSE_API int cse_masknet_fwd(const cse_params* p_host, const void* E, const float* ctx,
                            int B, int L, int c, int n_masks, int precision,
                            float* mask, float* pred_head,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- per-stage entry points (unit-testable; same kernels cse_forward launches) ---- */
/* act_dtype: CSE_FP32 -> float activations, CSE_BF16 -> __nv_bfloat16 activations. */

/* speechbrain Encoder: relu(conv1d(k=16,s=8,no bias)) (ContSep.py:10,69).  mix [B,T] ->
 * out [B,L,256] act; also per-CTA (sum,sumsq) partials for masknet.norm -> gn_part [B,parts,2]. */
CSE_API int cse_encoder_fwd(const float* mix, const float* w, int B, int T, int act_dtype,
                            void* out, float* gn_part, int* n_parts, void* stream);

/* GroupNorm(1,256,eps=1e-8) statistics finalise: partials -> (mean, rstd) [B,2] (ContSep.py:164). */
CSE_API int cse_gn_finalize(const float* gn_part, int B, int n_parts, double count, float eps,
                            float* stat, void* stream);

/* masknet.norm apply (ContSep.py:226): out = (x-mean)*rstd*g + b, x [B,L,256] act -> act. */
CSE_API int cse_gn_apply(const void* x, const float* stat, const float* g, const float* b,
                         int B, int L, int act_dtype, void* out, void* stream);

/* nn.Linear / 1x1 conv: C[M,N] = A[M,K] W[N,K]^T (+bias)(relu)(+residual).
 * CSE_FP32: fp32 SIMT FFMA.  CSE_BF16: tcgen05 + TMA, bf16 operands (W must be bf16), fp32
 * accumulate in TMEM.  out_fp32 != 0 -> C is float (residual, if given, is float [M,N] and
 * may alias C); else C has the activation dtype.  bias_scale multiplies the bias (2 for the
 * overlap-add / conv2d commutation). */
CSE_API int cse_linear(const void* A, int lda, const void* W, const float* bias, float bias_scale,
                       const float* residual, void* C, int ldc, int M, int N, int K,
                       int relu, int out_fp32, int precision, void* stream);

/* Position-wise feed-forward sub-block with its residual add, fused (CSE_transformer.py:407-411,
 * PositionalwiseFeedForward :547-566), CSE_BF16 only, d_model 256 / d_ffn 1024:
 *   R[M,256] (fp32, in place) += relu(A[M,256] W1[1024,256]^T + b1) W2[256,1024]^T + b2
 * A = norm2 output (bf16), W1/W2 bf16.  The hidden activation stays on chip (TMEM -> smem). */
CSE_API int cse_ffn_fused(const void* A_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16,
                          const float* b2, float* R, int M, void* stream);

/* The feed-forward sub-block WITH the LayerNorms either side of it (CSE_transformer.py:406-411 of this layer, :387 of
 * the next one), CSE_BF16 only:
 *   A = LayerNorm(R; ln2_g, ln2_b, eps) (bf16, left in `scratch_bf16` [M,256]);
 *   R[M,256] (fp32, in place) += relu(A W1^T + b1) W2^T + b2;
 *   H1[M,256] (bf16) = LayerNorm(R_new; next_ln1_g, next_ln1_b, eps)       (H1_bf16 == NULL: not produced)
 * norm2 runs in dedicated warps ahead of the tensor pipe, the next layer's norm1 in the output warps on the rows they
 * have just updated: a transformer layer is four launches (in_proj, attention, out_proj, this). */
CSE_API int cse_ffn_ln_fused(float* R, const float* ln2_g, const float* ln2_b, float eps, void* scratch_bf16,
                             const void* W1_bf16, const float* b1, const void* W2_bf16, const float* b2,
                             const float* next_ln1_g, const float* next_ln1_b, void* H1_bf16, int M, void* stream);

/* Attention output projection + residual add + pre-FFN LayerNorm in one tcgen05 kernel (CSE_transformer.py:399-408:
 * `src = src + self_att(...)` then `norm2(src)`), CSE_BF16 only:
 *   R[M,256] (fp32, in place) += A[M,K] W[256,K]^T + bias;   H[M,256] (bf16) = LayerNorm(R; gamma, beta, eps)
 * A, W bf16 (K % 64 == 0).  The GEMM epilogue reads and writes the residual row straight from / to global memory
 * and normalises it while it is on chip, so the stream is not re-read by a separate LayerNorm kernel. */
CSE_API int cse_linear_residual_ln(const void* A_bf16, int lda, const void* W_bf16, const float* bias, float* R,
                                   const float* gamma, const float* beta, float eps, void* H_bf16, int M, int K,
                                   void* stream);

/* Pre-norm sub-block head `norm(src)` -> Linear fused (CSE_transformer.py:385-390 norm1 -> in_proj,
 * :407-411 norm2 -> ffn.0 + ReLU), CSE_BF16 only: C[M,N] bf16 = act(LN(R[M,256]) W[N,256]^T + bias),
 * N % 256 == 0.  The fp32 residual row is read once; LayerNorm runs in the GEMM's A-operand producer warps. */
CSE_API int cse_ln_linear(const float* R, const float* gamma, const float* beta, float eps,
                          const void* W_bf16, const float* bias, void* C, int ldc, int M, int N,
                          int relu, void* stream);

/* nn.LayerNorm(256, eps) over rows (CSE_transformer.py:358-359,386,408): x [M,256] fp32 -> act. */
CSE_API int cse_layernorm_fwd(const float* x, const float* g, const float* b, int M, float eps,
                              int act_dtype, void* out, void* stream);

/* nn.MultiheadAttention core (CSE_transformer.py:535-557 -> F.sdpa): qkv [nseq*n,768] act with
 * q|k|v column blocks, 8 heads of 32 -> out [nseq*n,256] act.  No mask, no dropout. */
CSE_API int cse_attention_fwd(const void* qkv, int nseq, int n, int act_dtype, void* out,
                              void* stream);

/* _padding + _Segmentation (ContSep.py:270-335): x0 [B,L,256] fp32 -> X [B,S,K,256] fp32. */
CSE_API int cse_segment(const float* x0, int B, int L, int S, float* X, void* stream);

/* Context-token prompt + positional encoding (ContSep.py:474-482 / :506-513 and
 * CSE_transformer.py:102-104): builds the residual stream of a stack.
 *   inter == 0: R[(b*S+s), c+k, :] = X[b,s,k,:] + pe[c+k];  inter != 0: R[(b*K+k), c+s, :] = ...
 *   rows j < c: ctok[b,j,:] + pe[j].  X [B,S,K,256] fp32, ctok [B,c,256] fp32 or NULL. */
CSE_API int cse_build_sequences(const float* X, const float* ctok, const float* pe,
                                int B, int S, int c, int inter, float* R, void* stream);

/* {intra,inter}_context_mapper = nn.Linear(4096,256) (ContSep.py:450-451,480,511). */
CSE_API int cse_context_map(const float* ctx, const float* w, const float* b, int rows, int in_dim,
                            float* out, void* stream);

/* Tail of a stack inside Dual_Computation_Block_CSE.forward (ContSep.py:487-502 / :518-531):
 * final LayerNorm(eps 1e-6) of R, drop the c context rows, GroupNorm(1,256,eps 1e-8) over the
 * sample, + skip.  out [B,S,K,256] fp32.  skip [B,S,K,256] fp32.  Uses scratch gn_part/stat. */
CSE_API int cse_stack_finish(const float* R, const float* ln_g, const float* ln_b,
                             const float* gn_g, const float* gn_b, const float* skip,
                             int B, int S, int c, int inter, float* out,
                             float* gn_part, float* stat, void* stream);

/* pred_head = mean_k LN(R_inter[(b,k), 0, :]) (ContSep.py:516-517). */
CSE_API int cse_pred_head(const float* R_inter, const float* ln_g, const float* ln_b,
                          int B, int S, int c, float* pred_head, void* stream);

/* PReLU then _over_add (ContSep.py:244,337-370), commuted in front of conv2d:
 * X [B,S,K,256] fp32 -> U [B,L,256] act. */
CSE_API int cse_prelu_overlap_add(const float* X, const float* prelu, int B, int S, int L,
                                  int act_dtype, void* U, void* stream);

/* tanh(o) * sigmoid(g) (ContSep.py:255). */
CSE_API int cse_gate(const void* o, const void* g, size_t n, int act_dtype, void* out, void* stream);

/* ReLU mask, mask*mix_w, ConvTranspose1d(256,1,16,stride 8) and length fix
 * (ContSep.py:263,79-95): mask_pre [B*L*n_masks,256] act (row = (b,l,s)), E [B,L,256] act,
 * dec_w [256,16] -> est [B,T,n_masks] fp32.  frames scratch [B*L*n_masks,16] fp32. */
/* E == NULL: plain Decoder.forward on an already-masked [rows,256] input (no ReLU/multiply). */
CSE_API int cse_mask_decode(const void* mask_pre, const void* E, const float* dec_w,
                            int B, int L, int T, int n_masks, int act_dtype,
                            float* frames, float* est, void* stream);

/* ---- losses ---- */
/* speechbrain cal_si_snr (train_ContSep.py:352,386): NEGATIVE SI-SNR of `estimate` against
 * `source`; both [B,T,C] fp32 (batch-major; the reference passes [T,B,C]) -> out [B,C]. */
CSE_API int cse_si_snr(const float* source, const float* estimate, int B, int T, int C,
                       float* out, void* stream);

/* get_si_snr_with_pitwrapper(source, estimate_source) (train_ContSep.py:346,391-393):
 * [B,T,C] x2 -> loss [B], perm [B,C] (perm[b,i] = column of `source` paired with
 * estimate_source[:, :, i]).  C <= 4. */
CSE_API int cse_pit_si_snr(const float* source, const float* estimate_source, int B, int T, int C,
                           float* loss, int* perm, void* stream);

/* torchmetrics ScaleInvariantSignalNoiseRatio (train_ContExt.py:339,367): [B,T] x2 -> dB [B]. */
CSE_API int cse_tm_si_snr(const float* preds, const float* target, int B, int T,
                          float* out, void* stream);

/* ---- ContSep selection tail (estimate tensors are [B,T,n_streams] as the model returns them) ---- */
/* Training (train_ContSep.py:386-388): sisnr[b,s] = -cal_si_snr(source = gt[b], estimate = est[b,:,s]) (estimate
 * detached), label[b] = argmax_s (first maximum, int64), loss[0] = CrossEntropyLoss(logits [B,n_streams], label)
 * when ce != 0, else BCEWithLogitsLoss(logits [B] (the single-logit head, 2 streams), label.float()); both batch
 * means.  dlogits (same shape as logits) = d loss / d logits, so the backward pass is a scale by the incoming
 * gradient.  item_loss: B floats of scratch.  One launch per item + a fixed-order mean: no host round trip. */
CSE_API int cse_selection_loss(const float* gt, const float* est, const float* logits, int B, int T,
                               int n_streams, int ce, float* sisnr, long long* label, float* loss,
                               float* dlogits, float* item_loss, void* stream);
/* Eval (test.py:234-239): pick[b] = argmax softmax(logits[b]) (ce) or sigmoid(logit[b]) > 0.5, out[b,:] =
 * est[b,:,pick[b]] — the reference moves the logits to the host for this (`ctx_pred.cpu()`, test.py:236). */
CSE_API int cse_select_stream(const float* est, const float* logits, int B, int T, int n_streams, int ce,
                              float* out, long long* pick, void* stream);
/* Eval (test.py:248-255): sisnr[b,j] = -cal_si_snr(source = sources[b,:,j], estimate = enhanced[b]);
 * acc[b] = 1 iff sisnr[b,0] >= sisnr[b,j] for every interferer j >= 1 (column 0 = the target speaker). */
CSE_API int cse_selection_accuracy(const float* enhanced, const float* sources, int B, int T, int n_sources,
                                   float* sisnr, int* acc, void* stream);

/* ---- backward (training step, BASELINE configs[2]) ----
 * The reference trains through autograd (`loss.backward()`, train_ContSep.py:402-419,
 * train_ContExt.py:372-389); these entry points are the hand-written gradients of the same
 * operators, in true-fp32 arithmetic (parity mode).  Parameter gradients are ACCUMULATED (+=) into
 * caller-owned buffers, like autograd's .grad; zero them before the first micro-step. */

/* d cal_si_snr / d(source, estimate): g_out [B,C] = dL/d out of cse_si_snr -> d_source, d_estimate
 * [B,T,C] (either may be NULL).  Overwrites. */
CSE_API int cse_si_snr_bwd(const float* source, const float* estimate, const float* g_out,
                           int B, int T, int C, float* d_source, float* d_estimate, void* stream);

/* d get_si_snr_with_pitwrapper: g_loss [B], perm [B,C] as returned by cse_pit_si_snr (the chosen
 * permutation is a constant of the backward pass, as in PitWrapper) -> d_source, d_estimate_source. */
CSE_API int cse_pit_si_snr_bwd(const float* source, const float* estimate_source, const float* g_loss,
                               const int* perm, int B, int T, int C, float* d_source,
                               float* d_estimate_source, void* stream);

/* d torchmetrics SI-SNR: g_out [B] -> d_preds, d_target [B,T] (either may be NULL). */
CSE_API int cse_tm_si_snr_bwd(const float* preds, const float* target, const float* g_out, int B, int T,
                              float* d_preds, float* d_target, void* stream);

/* nn.Linear backward for C = A W^T + bias (fp32): dC [M,N] ->
 *   dA [M,K] = dC W            (overwritten; NULL to skip; needs scratch_wt [N*K] floats)
 *   dW [N,K] += dC^T A         (NULL to skip)      dbias [N] += column sums of dC (NULL to skip)
 * N, K multiples of 128 (N of 256 when dbias is given). */
CSE_API int cse_linear_bwd(const float* A, int lda, const float* W, const float* dC, int lddc,
                           int M, int N, int K, float* dA, int ldda, float* dW, float* dbias,
                           float* scratch_wt, void* stream);

/* nn.LayerNorm(256) backward: x, dy [M,256] -> dx (+= when accumulate != 0), dg/db [256] += . */
CSE_API int cse_layernorm_bwd(const float* x, const float* g, const float* dy, int M, float eps,
                              int accumulate, float* dx, float* dg, float* db, void* stream);

/* Attention core backward: qkv [nseq*n,768], out [nseq*n,256] (forward result), d_out -> d_qkv
 * [nseq*n,768] (overwritten).  n <= 390. */
CSE_API int cse_attention_bwd(const float* qkv, const float* out, const float* d_out, int nseq, int n,
                              float* d_qkv, void* stream);

/* The same gradient on the tensor cores (mma.sync bf16, fp32 softmax algebra and accumulation), performance mode:
 * qkv [M,768] and out [M,256] are the bf16 tensors of the autocast forward, d_out [M,256] and d_qkv [M,768] fp32.
 * n <= 256.  Used by cse_layer_bwd_bf16; P and dS are rounded to bf16 for their second contraction, as autocast
 * does (train_ContSep.py:383). */
CSE_API int cse_attention_bwd_bf16(const void* qkv_bf16, const void* out_bf16, const float* d_out, int nseq, int n,
                                   float* d_qkv, void* stream);

/* Gradient buffers of one transformer layer (same shapes as cse_layer_params' fp32 members). */
typedef struct {
  float* in_proj_w;
  float* in_proj_b;
  float* out_proj_w;
  float* out_proj_b;
  float* ffn1_w;
  float* ffn1_b;
  float* ffn2_w;
  float* ffn2_b;
  float* ln1_g;
  float* ln1_b;
  float* ln2_g;
  float* ln2_b;
} cse_layer_grads;

/* TransformerEncoderLayer.forward (CSE_transformer.py:385-416) on the fp32 residual stream, in place:
 * R [nseq*n,256].  precision as cse_forward.  Workspace: cse_layer_workspace_bytes. */
CSE_API size_t cse_layer_workspace_bytes(int nseq, int n);
CSE_API int cse_layer_fwd(const cse_layer_params* p_host, float* R, int nseq, int n, int precision,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the same layer from its INPUT residual stream (activation checkpointing at layer
 * granularity: the layer is recomputed in fp32, then differentiated):
 *   R_in [M,256] (layer input, unchanged), dR [M,256]: in = dL/dR_out, out = dL/dR_in;
 *   grads_host: parameter gradients, accumulated. */
CSE_API int cse_layer_bwd(const cse_layer_params* p_host, const cse_layer_grads* grads_host,
                          const float* R_in, float* dR, int nseq, int n,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- training path, non-transformer stages (fp32; csrc/train_ops.cu) ---- */

/* select_norm('ln') = nn.GroupNorm(1,256,eps) over x [B,rows,256] per sample (ContSep.py:164,226,
 * 423-424), optionally fused with the block skip connection (ContSep.py:498-502,527-531):
 * out = (x-mean)*rstd*g + b (+ skip).  stat [B,2] = (mean, rstd) is kept for the backward pass;
 * part_scratch holds B*64*2 floats. */
CSE_API int cse_groupnorm_fwd(const float* x, const float* g, const float* b, const float* skip,
                              int B, int rows, float eps, float* out, float* stat,
                              float* part_scratch, void* stream);
/* dx overwritten, dg/db [256] accumulated (NULL to skip).  d skip = dy (identity; left to the caller). */
CSE_API int cse_groupnorm_bwd(const float* x, const float* stat, const float* g, const float* dy,
                              int B, int rows, float* dx, float* dg, float* db,
                              float* part_scratch, void* stream);

/* Inverse relayout of cse_build_sequences (ContSep.py:487-489 / :518-521 `[:, c:]` + permutes, and the
 * adjoint of the forward relayout): R [nseq*n,256] -> X [B,S,K,256] without the c context rows (X may
 * be NULL); ctok_sum [B,c,256] = per-sample sum of the context rows over all sequences (the gradient
 * of the broadcast prompt token; NULL to skip). */
CSE_API int cse_sequences_to_chunks(const float* R, int B, int S, int c, int inter, float* X,
                                    float* ctok_sum, void* stream);

/* Backward of cse_prelu_overlap_add (fp32): dU [B,L,256] -> dX [B,S,K,256] (overwritten),
 * dprelu [1] accumulated (NULL to skip).  With prelu = 1 the forward is exactly the adjoint of
 * cse_segment, so segmentation needs no backward kernel of its own. */
CSE_API int cse_prelu_overlap_add_bwd(const float* X, const float* prelu, const float* dU,
                                      int B, int S, int L, float* dX, float* dprelu, void* stream);

/* Backward of cse_gate: y = tanh(o) * sigmoid(g); d_o, d_g overwritten. */
CSE_API int cse_gate_bwd(const float* o, const float* g, const float* d_out, size_t n,
                         float* d_o, float* d_g, void* stream);

/* Backward of cse_mask_decode (fp32, E != NULL): d_est [B,T,n_masks] -> d_mask_pre [B*L*n_masks,256]
 * and dE [B,L,256] (both overwritten), d_dec_w [256,16] accumulated (NULL to skip). */
CSE_API int cse_mask_decode_bwd(const float* mask_pre, const float* E, const float* dec_w,
                                const float* d_est, int B, int L, int T, int n_masks,
                                float* d_mask_pre, float* dE, float* d_dec_w, void* stream);

/* Backward of cse_encoder_fwd (fp32): d_w [256,16] += sum dE * (E > 0) * frames(mix). */
CSE_API int cse_encoder_bwd(const float* mix, const float* E, const float* dE, int B, int T,
                            float* d_w, void* stream);

/* ---- EXPERIMENTAL (compiled, not yet run on hardware — see csrc/backward_tc.cu) ----
 * nn.Linear backward on the tcgen05 GEMM: bf16 operands, fp32 accumulate.  A [M,K] float or bf16
 * (a_is_bf16), W [N,K] fp32 master, dC [M,N] fp32 contiguous ->
 *   dA [M,K] (fp32 if dA_fp32 else bf16; NULL to skip), dW [N,K] fp32 += , dbias [N] += .
 * scratch: cse_linear_bwd_tc_scratch_bytes(M,N,K), 256-byte aligned. */
CSE_API size_t cse_linear_bwd_tc_scratch_bytes(int M, int N, int K);
CSE_API int cse_linear_bwd_tc(const void* A, int a_is_bf16, int lda, const float* W, const float* dC,
                              int M, int N, int K, void* dA, int dA_fp32, int ldda, float* dW,
                              float* dbias, void* scratch, size_t scratch_bytes, void* stream);

/* EXPERIMENTAL (same status): cse_layer_bwd in the performance mode — bf16 recompute with the forward's own
 * tcgen05 kernels, tensor-core dgrad / wgrad, fp32 LayerNorm / attention / residual gradients.  p_host needs
 * the *_bf16 members (cse_pack_bf16 layout: any bf16 copies of the four weight matrices). */
CSE_API size_t cse_layer_bwd_bf16_workspace_bytes(int nseq, int n);
CSE_API int cse_layer_bwd_bf16(const cse_layer_params* p_host, const cse_layer_grads* grads_host,
                               const float* R_in, float* dR, int nseq, int n,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused optimiser step ----
 * Replaces, per update (train_ContSep.py:233,402-419; train_ContExt.py:372-389):
 *   scaler.unscale_(optimizer); grad_norm = clip_grad_norm_(model.parameters(), max_norm=5.0);
 *   [skip when the norm is not finite] optimizer.step() with optim.AdamW(amsgrad=True); scaler.update()
 * by three launches over a table of <= 16384-element chunks of the parameter tensors, without the reference's host
 * synchronisation on the norm.
 *   cse_optim_chunk_count   chunks a parameter list needs (host arithmetic; -1 on bad input)
 *   cse_optim_table_fill    writes the 48-byte chunk descriptors {param, grad, exp_avg, exp_avg_sq,
 *                           max_exp_avg_sq, n, pad} for DEVICE pointers into a caller-owned HOST buffer, which
 *                           the caller copies to the device (max_exp_avg_sq may be NULL when amsgrad == 0)
 *   cse_optim_step          device_table: that copy; state: 64 B of zero-initialised device memory
 *                           {double step; float scale; int growth_tracker; float total_norm; float coef;
 *                            int found_inf; float step_size; float bc2_sqrt; float reserved[7]} (set `scale`
 *                           to the GradScaler's initial scale when use_scaler != 0); partial: n_chunks floats
 *                           of device scratch.  total_norm is the norm of the unscaled gradients
 *                           (clip_grad_norm_'s return value); found_inf = 1 means the update was skipped and,
 *                           with use_scaler, the scale backed off — GradScaler.step/update semantics.
 *                           write_back_grads != 0 also stores the unscaled, clipped gradients (what
 *                           clip_grad_norm_ leaves in .grad). */
CSE_API long long cse_optim_chunk_count(int n_tensors, const long long* numel);
CSE_API int cse_optim_table_fill(int n_tensors, const long long* numel, void* const* param, void* const* grad,
                                 void* const* exp_avg, void* const* exp_avg_sq, void* const* max_exp_avg_sq,
                                 void* host_table, size_t host_table_bytes);
/* Rewrites only the gradient pointers of a filled host table (autograd hands out new .grad storages every step). */
CSE_API int cse_optim_table_set_grads(int n_tensors, const long long* numel, void* const* grad, void* host_table,
                                      size_t host_table_bytes);
CSE_API int cse_optim_step(const void* device_table, long long n_chunks, float lr, float beta1, float beta2,
                           float eps, float weight_decay, int amsgrad, float max_norm, int use_scaler,
                           float growth_factor, float backoff_factor, int growth_interval, int write_back_grads,
                           void* state, float* partial, void* stream);

/* ---- evaluation metrics on the device (test.py:198-201,241-245,291-301) ----
 * cse_sdr: torchmetrics.functional.audio.signal_distortion_ratio(preds, target, use_cg_iter=None,
 *   filter_length, zero_mean, load_diag) per item, float64 inside: preds, target [B,T] fp32 -> out [B] (dB).
 *   filter_length <= 1024 (the reference uses the default 512); has_load_diag = 0 means load_diag=None.
 *   workspace: cse_sdr_workspace_bytes(B, T, filter_length) bytes, 8-byte aligned.
 * cse_metric_update: the running state of a torchmetrics metric object — acc[0] += sum(values[0..n)),
 *   acc[1] += n (two doubles in device memory, zero-initialised by the caller); compute() = acc[0] / acc[1].
 *   SI-SNR values come from cse_tm_si_snr, SDR values from cse_sdr. */
CSE_API size_t cse_sdr_workspace_bytes(int B, int T, int filter_length);
CSE_API int cse_sdr(const float* preds, const float* target, int B, int T, int filter_length, int zero_mean,
                    int has_load_diag, double load_diag, float* out, void* workspace, size_t workspace_bytes,
                    void* stream);
CSE_API int cse_metric_update(const float* values, int n, double* acc, void* stream);

/* ---- loader-side mixture synthesis + collation on the device ----
 * Ragged clips come as one flat fp32 buffer per role plus B+1 int64 element offsets (device); outputs are the
 * collated, zero right-padded [B, T_out] tensors collate_fn builds (dataset_train_CSE.py:507-601).
 * cse_mix_audio: n_noise = 1 -> mix_audio(signal, noise1, snr1, pad)            (dataset_train_CSE.py:417-456)
 *                n_noise = 2 -> mix_audio_3spk(signal, noise1, noise2, snr1, snr2, pad)       (:458-505)
 *   snr1 / snr2: B float64 values on the device (the dataset draws numpy float64 scalars, :257-263).  out_len[b]
 *   (optional) = samples item b occupies (len(signal), or the longest clip for 3 speakers); T_out must cover it.
 *   Same dtype promotion as numpy in the reference: float32 energies, float64 gains / mixture / peak scale.
 * cse_peak_normalize: out[b] = x_b / max|x_b| * peak in float32 (dataset_train_CSE.py:237,274), zero-padded.
 * cse_decimate: rows of [B, T_in] (valid lengths len_in or NULL) low-pass filtered with taps[n_taps] (device; centre
 *   tap aligned with the kept samples) and decimated by `down` -> [B, T_out], len_out[b] = ceil(len / down): the
 *   16 kHz -> 8 kHz step (dataset_train_CSE.py:393-398).  The reference calls librosa.resample, whose soxr back end
 *   is an absent dependency; the taps are an argument (default of the host mirror: scipy.signal.resample_poly's). */
CSE_API int cse_mix_audio(const float* signal, const long long* signal_off, const float* noise1,
                          const long long* noise1_off, const float* noise2, const long long* noise2_off,
                          const double* snr1, const double* snr2, int B, int n_noise, int pad, long long T_out,
                          float* mixed, float* signal_out, float* noise1_out, float* noise2_out, int* out_len,
                          void* stream);
CSE_API int cse_peak_normalize(const float* x, const long long* off, int B, float peak, long long T_out,
                               float* out, void* stream);
CSE_API int cse_decimate(const float* x, const int* len_in, int B, long long T_in, int down, const float* taps,
                         int n_taps, long long T_out, float* y, int* len_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSE_B200_H */
