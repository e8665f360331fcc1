// abi.cu — extern "C" boundary (include/cse_b200.h) and the host-side orchestration of the path.
//
// cse_forward is Sepformer.forward of the reference (ContSep.py:53-100, ContExt.py:54-129,
// sepformer.py:42-81) with Dual_Path_Model_CSE.forward (ContSep.py:205-268) and
// Dual_Computation_Block_CSE.forward (ContSep.py:453-533) unrolled into kernel launches on the
// caller's stream.  No allocation, no synchronisation (except cse_forward_host), no global state
// besides the TMA descriptor cache in gemm_tc.cu.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace cse {

static thread_local char g_err[768] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};

// ---- per-kernel-class timing ----
struct ProfRec { int cls; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;       // records of the current collection window
static std::vector<cudaEvent_t> g_ev_pool;
static std::mutex g_prof_mu;

static cudaEvent_t prof_event() {
  if (!g_ev_pool.empty()) {
    cudaEvent_t e = g_ev_pool.back();
    g_ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

KernelScope::KernelScope(int cls, cudaStream_t st) : slot_(-1), st_(st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> g(g_prof_mu);
  ProfRec r{cls, prof_event(), prof_event()};
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  slot_ = (int)g_prof.size() - 1;
}
KernelScope::~KernelScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> g(g_prof_mu);
  cudaEventRecord(g_prof[slot_].e1, st_);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    dst[i] = __float2bfloat16_rn(src[i]);
}

int launch_f32_to_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  const int grid = (int)min((size_t)148 * 8, (n + 255) / 256);
  f32_to_bf16_kernel<<<grid, 256, 0, st>>>(src, dst, n);
  return check_launch("f32_to_bf16_kernel");
}

// ---------------------------------------------------------------------------------------------
// workspace plan
// ---------------------------------------------------------------------------------------------
struct Plan {
  cse_shape sh;
  int n_masks, precision;
  size_t esz;        // activation element size
  size_t m_max;      // rows of the larger stack
  int enc_parts;
  // byte offsets into the workspace
  size_t E, En, x0, XA, XB, Ra, Rb, H, QKV, AO, F1, ctok, part, stat, U, V, O, G, MP, frames;
  size_t h_mix, h_ctx, h_est, h_pred;  // device staging for cse_forward_host
  size_t total;
};

static int make_shape(int B, int T, int c, int spk, cse_shape* s) {
  CSE_REQUIRE(B >= 1 && c >= 0 && spk >= 1, "path_shape: need B >= 1, c >= 0, spk >= 1 (B=%d c=%d spk=%d)", B, c, spk);
  CSE_REQUIRE(T >= kEncK, "path_shape: mixture of %d samples is shorter than the %d-tap encoder kernel", T, kEncK);
  s->B = B; s->T = T; s->c = c; s->spk = spk;
  s->L = (T - kEncK) / kEncS + 1;
  s->gap = kK - (kP + s->L % kK) % kK;
  s->S = 2 * (s->L + s->gap + kP) / kK;
  s->T_est = kEncS * (s->L - 1) + kEncK;
  CSE_REQUIRE(s->S + c <= 2500 && kK + c <= 2500,
              "path_shape: %d tokens exceed the positional table (2500)", s->S + c);
  return 0;
}

static int make_plan(int B, int T, int c, int n_masks, int precision, Plan* p) {
  if (make_shape(B, T, c, n_masks, &p->sh)) return 1;
  CSE_REQUIRE(precision == CSE_FP32 || precision == CSE_BF16, "unknown precision %d", precision);
  p->n_masks = n_masks;
  p->precision = precision;
  p->esz = precision == CSE_BF16 ? 2 : 4;
  const cse_shape& s = p->sh;
  const size_t rows_i = (size_t)B * s.S * (kK + c), rows_e = (size_t)B * kK * (s.S + c);
  p->m_max = rows_i > rows_e ? rows_i : rows_e;
  p->enc_parts = encoder_parts(s.L);
  const size_t BL = (size_t)B * s.L, chunk = (size_t)B * s.S * kK, Mh = BL * n_masks;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  p->E = take(BL * kN * p->esz);
  p->En = take(BL * kN * p->esz);
  p->x0 = take(BL * kN * 4);
  p->XA = take(chunk * kN * 4);
  p->XB = take(chunk * kN * 4);
  p->Ra = take(rows_i * kN * 4);
  p->Rb = take(rows_e * kN * 4);
  p->H = take(p->m_max * kN * p->esz);
  p->QKV = take(p->m_max * 3 * kN * p->esz);   // QKV and AO are contiguous: together they hold F1
  p->AO = take(p->m_max * kN * p->esz);
  p->F1 = p->QKV;
  p->ctok = take((size_t)4 * B * (c > 0 ? c : 1) * kN * 4);
  const int parts = p->enc_parts > kFinishParts ? p->enc_parts : kFinishParts;
  p->part = take((size_t)B * parts * 2 * 4);
  p->stat = take((size_t)B * 2 * 4);
  p->U = take(BL * kN * p->esz);
  p->V = take(Mh * kN * p->esz);
  p->O = take(Mh * kN * p->esz);
  p->G = take(Mh * kN * p->esz);
  p->MP = take(Mh * kN * p->esz);
  p->frames = take(Mh * kEncK * 4);
  p->h_mix = take((size_t)B * T * 4);
  p->h_ctx = take((size_t)B * (c > 0 ? c : 1) * CSE_CTX * 4);
  p->h_est = take((size_t)B * T * n_masks * 4);
  p->h_pred = take((size_t)B * kN * 4);
  p->total = off;
  // F1 [m_max,1024] must fit in QKV+AO: the two takes are adjacent only if no padding was inserted
  if (p->AO != p->QKV + align_up(p->m_max * 3 * kN * p->esz, 256) ||
      (p->m_max * 3 * kN * p->esz) % 256 != 0) {
    // fall back to a dedicated F1 buffer
    p->F1 = take(p->m_max * kFfn * p->esz);
    p->total = off;
  }
  return 0;
}

static int linear(const Plan& pl, const void* A, int lda, const float* W32, const void* W16,
                  const float* bias, float bias_scale, const float* residual, void* C, int ldc,
                  int M, int N, int K, int relu, int out_fp32, cudaStream_t st) {
  if (pl.precision == CSE_BF16) {
    CSE_REQUIRE(W16 != nullptr, "bf16 weights missing: call cse_pack_bf16 first");
    return launch_gemm_tc((const bf16*)A, lda, (const bf16*)W16, bias, bias_scale, residual, C, ldc,
                          M, N, K, relu, out_fp32, st);
  }
  return launch_gemm_simt((const float*)A, lda, W32, bias, bias_scale, residual, (float*)C, ldc, M,
                          N, K, relu, st);
}

bool pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("CSE_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

// CSE_FFN_FUSED=0 keeps the two-GEMM feed-forward path (A/B aid; the fused kernel is the default)
static bool ffn_fused_enabled() {
  static const bool on = []() {
    const char* e = getenv("CSE_FFN_FUSED");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

// CSE_OUTPROJ_LN=1 runs out-proj (+R) and norm2 as ONE kernel (cse_linear_residual_ln: LayerNorm in the GEMM
// epilogue).  Off by default: correct and 32 launches / 4.5 GB of HBM reads per forward lighter, but measured 95 us
// against 58 + 38 us for the two kernels it replaces — its epilogue is a chain of TMEM round trips
// (profiles/r02_experiments.md section 2).
static bool outproj_ln_fused_enabled() {
  static const bool on = []() {
    const char* e = getenv("CSE_OUTPROJ_LN");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

// The two LayerNorms of a layer ride along in the feed-forward kernel (ffn_tc.cu, LNF variant): norm2 in its own
// LayerNorm warps ahead of the tensor pipe, norm1 of the NEXT layer in its output warps — four launches per layer.
// CSE_FFN_LN=0 keeps the separate layernorm_kernel launches (A/B aid).
static int ffn_ln_mode() {   // 0: off, 1: norm2 + next norm1, 2: norm2 only (norm1 stays a launch), 3: auto (default)
  static const int mode = []() {
    const char* e = getenv("CSE_FFN_LN");
    return e == nullptr ? 3 : (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1));
  }();
  return mode;
}
// auto: the LayerNorms ride in the feed-forward kernel when its CTAs have several row tiles each to hide them behind
// (>= 16 k rows: +7-10 % on the forward); with one tile per CTA they are latency in front of and behind the tile and
// the separate launches are 5-10 % faster (profiles/r02_sweep_cfg4_cfg5.md)
constexpr size_t kFfnLnMinRows = 16000;

// CSE_LN_FUSED=1 runs norm1 -> in_proj as one kernel (gemm_ln_tc.cu).  Off by default: measured 123-132 us
// against 36 + 67 us for the two kernels (its LayerNorm warps cannot keep enough loads in flight inside the
// 96 registers a 576-thread CTA leaves them), profiles/r01_experiments.md.
static bool ln_qkv_fused_enabled() {
  static const bool on = []() {
    const char* e = getenv("CSE_LN_FUSED");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

static bool ffn_ln_path(const Plan& pl) {
  const int mode = ffn_ln_mode();
  return pl.precision == CSE_BF16 && ffn_fused_enabled() && mode != 0 && (mode != 3 || pl.m_max >= kFfnLnMinRows) &&
         !ln_qkv_fused_enabled() && !outproj_ln_fused_enabled();
}
static bool ffn_ln_next_norm1() { return ffn_ln_mode() != 2; }   // (mode 2: A/B aid, norm1 stays a launch)

// SBTransformerBlock_CSE body after the PE add: 8 pre-norm layers on the fp32 residual stream R
// (TransformerEncoderLayer.forward, CSE_transformer.py:385-416).  The final LayerNorm belongs to
// the stack tail (stack_finish / pred_head).
// h_ready: norm1 of the first layer is already in ws+H (written by the kernel that built R: build_seq / stack tail)
static int run_stack(const Plan& pl, const cse_stack_params& sp, float* R, int nseq, int n,
                     char* ws, cudaStream_t st, bool h_ready = false) {
  const int M = nseq * n;
  void* H = ws + pl.H;
  void* QKV = ws + pl.QKV;
  void* AO = ws + pl.AO;
  void* F1 = ws + pl.F1;
  const int act = pl.precision;
  const bool ffn_ln = ffn_ln_path(pl);
  for (int l = 0; l < CSE_LAYERS; ++l) {
    const cse_layer_params& lp = sp.layer[l];
    if (ffn_ln && ffn_ln_next_norm1() && (l > 0 || h_ready)) {
      // norm1(R) is already in H: the previous layer's feed-forward kernel wrote it
      if (linear(pl, H, kN, lp.in_proj_w, lp.in_proj_w_bf16, lp.in_proj_b, 1.f, nullptr, QKV, 3 * kN, M,
                 3 * kN, kN, 0, 0, st)) return 1;
    } else if (pl.precision == CSE_BF16 && ln_qkv_fused_enabled()) {
      // norm1 -> in_proj in one kernel: the residual row is read once and normalised in the GEMM's A producer
      CSE_REQUIRE(lp.in_proj_w_bf16, "bf16 weights missing: call cse_pack_bf16 first");
      if (launch_gemm_ln_tc(R, lp.ln1_g, lp.ln1_b, 1e-6f, (const bf16*)lp.in_proj_w_bf16, lp.in_proj_b,
                            (bf16*)QKV, 3 * kN, M, 3 * kN, 0, st)) return 1;
    } else {
      if (launch_layernorm(R, lp.ln1_g, lp.ln1_b, M, 1e-6f, act, H, st)) return 1;
      if (linear(pl, H, kN, lp.in_proj_w, lp.in_proj_w_bf16, lp.in_proj_b, 1.f, nullptr, QKV, 3 * kN, M,
                 3 * kN, kN, 0, 0, st)) return 1;
    }
    if (launch_attention(QKV, nseq, n, act, AO, st)) return 1;
    if (pl.precision == CSE_BF16 && outproj_ln_fused_enabled()) {
      // x = x + out_proj(attention); norm2(x) in the same kernel: the row is normalised while it is on chip
      CSE_REQUIRE(lp.out_proj_w_bf16, "bf16 weights missing: call cse_pack_bf16 first");
      if (launch_gemm_tc_residual_ln((const bf16*)AO, kN, (const bf16*)lp.out_proj_w_bf16, lp.out_proj_b, R,
                                     lp.ln2_g, lp.ln2_b, 1e-6f, (bf16*)H, M, kN, st)) return 1;
    } else {
      if (linear(pl, AO, kN, lp.out_proj_w, lp.out_proj_w_bf16, lp.out_proj_b, 1.f, R, R, kN, M, kN, kN,
                 0, 1, st)) return 1;
      if (!ffn_ln && launch_layernorm(R, lp.ln2_g, lp.ln2_b, M, 1e-6f, act, H, st)) return 1;
    }
    if (ffn_ln) {
      // norm2 -> Linear -> ReLU -> Linear -> +R -> next layer's norm1, one kernel; AO (consumed by out_proj) is the
      // scratch that receives norm2(R)
      CSE_REQUIRE(lp.ffn1_w_bf16 && lp.ffn2_w_bf16, "bf16 weights missing: call cse_pack_bf16 first");
      const cse_layer_params* nx = l + 1 < CSE_LAYERS && ffn_ln_next_norm1() ? &sp.layer[l + 1] : nullptr;
      if (launch_ffn_tc_ln(R, lp.ln2_g, lp.ln2_b, 1e-6f, (bf16*)AO, (const bf16*)lp.ffn1_w_bf16, lp.ffn1_b,
                           (const bf16*)lp.ffn2_w_bf16, lp.ffn2_b, nx ? nx->ln1_g : nullptr,
                           nx ? nx->ln1_b : nullptr, nx ? (bf16*)H : nullptr, M, st)) return 1;
      continue;
    }
    if (pl.precision == CSE_BF16 && ffn_fused_enabled()) {
      // Linear -> ReLU -> Linear -> residual add in one kernel: the [M,1024] hidden never leaves the SM
      CSE_REQUIRE(lp.ffn1_w_bf16 && lp.ffn2_w_bf16, "bf16 weights missing: call cse_pack_bf16 first");
      if (launch_ffn_tc((const bf16*)H, (const bf16*)lp.ffn1_w_bf16, lp.ffn1_b, (const bf16*)lp.ffn2_w_bf16,
                        lp.ffn2_b, R, M, st)) return 1;
      continue;
    }
    if (linear(pl, H, kN, lp.ffn1_w, lp.ffn1_w_bf16, lp.ffn1_b, 1.f, nullptr, F1, kFfn, M, kFfn, kN, 1,
               0, st)) return 1;
    if (linear(pl, F1, kFfn, lp.ffn2_w, lp.ffn2_w_bf16, lp.ffn2_b, 1.f, R, R, kN, M, kN, kFfn, 0, 1,
               st)) return 1;
  }
  return 0;
}

// Dual_Path_Model_CSE.forward (ContSep.py:205-268) on channels-last mix_w `E` whose masknet.norm
// partial statistics are already in ws+part (n_parts per sample).  Leaves the pre-ReLU mask
// [B*L*n_masks, 256] (row = (b,l,s)) in ws+MP.
static int masknet_impl(const cse_params* p, const void* E, int n_parts, const float* ctx,
                        const Plan& pl, float* pred_head, char* ws, cudaStream_t st) {
  const cse_shape& s = pl.sh;
  const int B = s.B, c = s.c, L = s.L, S = s.S, n_masks = pl.n_masks, act = pl.precision;
  void* En = ws + pl.En;
  float* x0 = (float*)(ws + pl.x0);
  float* XA = (float*)(ws + pl.XA);
  float* XB = (float*)(ws + pl.XB);
  float* Ra = (float*)(ws + pl.Ra);
  float* Rb = (float*)(ws + pl.Rb);
  float* part = (float*)(ws + pl.part);
  float* stat = (float*)(ws + pl.stat);
  float* ctok = (float*)(ws + pl.ctok);
  const size_t ctok_stride = (size_t)B * (c > 0 ? c : 1) * kN;
  auto tok = [&](int blk, int inter) -> float* {
    return c > 0 ? ctok + (size_t)(blk * 2 + inter) * ctok_stride : nullptr;
  };

  // masknet.norm + masknet.conv1d + segmentation (ContSep.py:226-236)
  if (launch_gn_finalize(part, B, n_parts, (double)L * kN, 1e-8f, stat, st)) return 1;
  if (launch_gn_apply(E, stat, p->norm_g, p->norm_b, B, L, act, En, st)) return 1;
  if (linear(pl, En, kN, p->conv1d_w, p->conv1d_w_bf16, nullptr, 0.f, nullptr, x0, kN, B * L, kN, kN, 0,
             1, st)) return 1;
  if (launch_segment(x0, B, L, S, XA, st)) return 1;

  // context prompt tokens of both blocks (ContSep.py:480,511)
  if (c > 0) {
    CSE_REQUIRE(ctx != nullptr, "forward: c=%d but ctx is NULL", c);
    for (int blk = 0; blk < CSE_BLOCKS; ++blk) {
      const cse_block_params& bp = p->block[blk];
      CSE_REQUIRE(bp.intra_map_w && bp.inter_map_w && bp.intra_map_b && bp.inter_map_b,
                  "forward: context mappers missing (model built without add_ctx()?)");
      if (launch_context_map(ctx, bp.intra_map_w, bp.intra_map_b, B * c, CSE_CTX, tok(blk, 0), st)) return 1;
      if (launch_context_map(ctx, bp.inter_map_w, bp.inter_map_b, B * c, CSE_CTX, tok(blk, 1), st)) return 1;
    }
  }

  // In the 4-launch layer mode (norm1 of layers 1-7 comes from the previous layer's feed-forward kernel) the kernel
  // that BUILDS a stack's residual stream also writes norm1 of its first layer: no standalone LayerNorm launch is left
  const bool h1 = ffn_ln_path(pl) && ffn_ln_next_norm1();
  bf16* Hn = h1 ? (bf16*)(ws + pl.H) : nullptr;
  if (launch_build_sequences(XA, tok(0, 0), p->block[0].intra.pe, B, S, c, 0, Ra, st,
                             p->block[0].intra.layer[0].ln1_g, p->block[0].intra.layer[0].ln1_b, Hn)) return 1;
  for (int blk = 0; blk < CSE_BLOCKS; ++blk) {
    const cse_block_params& bp = p->block[blk];
    // intra: sequences = chunks (ContSep.py:474-502)
    if (run_stack(pl, bp.intra, Ra, B * S, kK + c, ws, st, h1)) return 1;
    if (launch_stack_finish(Ra, bp.intra.final_g, bp.intra.final_b, bp.intra_norm_g, bp.intra_norm_b, XA,
                            B, S, c, 0, XB, Rb, bp.inter.pe, tok(blk, 1), part, stat, st,
                            bp.inter.layer[0].ln1_g, bp.inter.layer[0].ln1_b, Hn)) return 1;
    // inter: sequences = in-chunk positions (ContSep.py:506-531)
    if (run_stack(pl, bp.inter, Rb, B * kK, S + c, ws, st, h1)) return 1;
    const bool last = blk == CSE_BLOCKS - 1;
    if (last && pred_head != nullptr) {
      if (launch_pred_head(Rb, bp.inter.final_g, bp.inter.final_b, B, S, c, pred_head, st)) return 1;
    }
    if (launch_stack_finish(Rb, bp.inter.final_g, bp.inter.final_b, bp.inter_norm_g, bp.inter_norm_b, XB,
                            B, S, c, 1, XA, last ? nullptr : Ra,
                            last ? nullptr : p->block[blk + 1].intra.pe,
                            last ? nullptr : tok(blk + 1, 0), part, stat, st,
                            last ? nullptr : p->block[blk + 1].intra.layer[0].ln1_g,
                            last ? nullptr : p->block[blk + 1].intra.layer[0].ln1_b, last ? nullptr : Hn)) return 1;
  }

  // mask head (ContSep.py:244-266) with the overlap-add commuted in front of conv2d
  void* U = ws + pl.U;
  void* V = ws + pl.V;
  void* O = ws + pl.O;
  void* G = ws + pl.G;
  void* MP = ws + pl.MP;
  const int Mh = B * L * n_masks;
  if (launch_prelu_ola(XA, p->prelu, B, S, L, act, U, st)) return 1;
  if (linear(pl, U, kN, p->conv2d_w, p->conv2d_w_bf16, p->conv2d_b, 2.f, nullptr, V, n_masks * kN, B * L,
             n_masks * kN, kN, 0, 0, st)) return 1;
  if (linear(pl, V, kN, p->out_w, p->out_w_bf16, p->out_b, 1.f, nullptr, O, kN, Mh, kN, kN, 0, 0, st)) return 1;
  if (linear(pl, V, kN, p->gate_w, p->gate_w_bf16, p->gate_b, 1.f, nullptr, G, kN, Mh, kN, kN, 0, 0, st)) return 1;
  if (launch_gate(O, G, (size_t)Mh * kN, act, O, st)) return 1;
  if (linear(pl, O, kN, p->end_w, p->end_w_bf16, nullptr, 0.f, nullptr, MP, kN, Mh, kN, kN, 0, 0, st)) return 1;
  return 0;
}

static int forward_impl(const cse_params* p, const float* mix, const float* ctx, const Plan& pl,
                        float* est, float* pred_head, char* ws, cudaStream_t st) {
  const cse_shape& s = pl.sh;
  void* E = ws + pl.E;
  // encoder (ContSep.py:69) with the masknet.norm partial sums fused in
  if (launch_encoder(mix, p->enc_w, s.B, s.T, s.L, pl.precision, E, (float*)(ws + pl.part), pl.enc_parts, st))
    return 1;
  if (masknet_impl(p, E, pl.enc_parts, ctx, pl, pred_head, ws, st)) return 1;
  // relu(mask) * mix_w -> ConvTranspose1d -> pad / trim (ContSep.py:263,79-95)
  return launch_mask_decode(ws + pl.MP, E, p->dec_w, s.B, s.L, s.T, pl.n_masks, pl.precision,
                            (float*)(ws + pl.frames), est, st);
}

}  // namespace cse

using namespace cse;

extern "C" {

int cse_version(void) { return 100; }

const char* cse_last_error(void) { return g_err; }

long long cse_launch_count(void) { return g_launches.load(); }

int cse_debug_force_mma_attention(int on) {
  g_attention_mode = on;  // 0 auto, 1 force mma.sync, 2 force tcgen05 v1 (n <= 256), 4 force tcgen05 v4
  return 0;
}

int cse_debug_attention_trace(long long* device_buffer) {
  g_attention_trace = device_buffer;  // [64 items][16 slots] of clock64 stamps written by CTA 0 of attention_tc4_kernel; NULL = off
  return 0;
}

int cse_profile_enable(int on) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  g_prof_on = on != 0;
  return 0;
}

int cse_profile_collect(double* ms_by_class, long long* launches_by_class, int n_classes) {
  CSE_REQUIRE(ms_by_class && launches_by_class && n_classes >= 4, "profile_collect: need at least 4 classes");
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (int i = 0; i < n_classes; ++i) { ms_by_class[i] = 0.0; launches_by_class[i] = 0; }
  for (ProfRec& r : g_prof) {
    CSE_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CSE_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    const int cls = r.cls < n_classes ? r.cls : (int)kClsGemmTc;   // a caller with 4 slots gets the feed-forward kernel in the GEMM class
    ms_by_class[cls] += ms;
    launches_by_class[cls] += 1;
    g_ev_pool.push_back(r.e0);
    g_ev_pool.push_back(r.e1);
  }
  g_prof.clear();
  return 0;
}

int cse_path_shape(int B, int T, int c, int spk, cse_shape* out) {
  CSE_REQUIRE(out != nullptr, "path_shape: out is NULL");
  return make_shape(B, T, c, spk, out);
}

size_t cse_workspace_bytes(int B, int T, int c, int n_masks, int precision) {
  Plan pl;
  if (make_plan(B, T, c, n_masks, precision, &pl)) return 0;
  return pl.total;
}

size_t cse_pack_bf16_elems(int n_masks) {
  const size_t per_layer = (size_t)3 * kN * kN + kN * kN + 2 * (size_t)kFfn * kN;
  return per_layer * CSE_LAYERS * 2 * CSE_BLOCKS + (size_t)kN * kN * 4 + (size_t)n_masks * kN * kN;
}

int cse_pack_bf16(cse_params* p, int n_masks, void* packed, size_t packed_elems, void* stream) {
  CSE_REQUIRE(p != nullptr && packed != nullptr, "pack_bf16: NULL argument");
  CSE_REQUIRE(packed_elems >= cse_pack_bf16_elems(n_masks), "pack_bf16: buffer too small (%zu < %zu)",
              packed_elems, cse_pack_bf16_elems(n_masks));
  cudaStream_t st = (cudaStream_t)stream;
  bf16* dst = (bf16*)packed;
  auto conv = [&](const float* src, size_t n, const void** out) -> int {
    if (src == nullptr) { set_error("pack_bf16: a weight pointer is NULL"); return 1; }
    if (launch_f32_to_bf16(src, dst, n, st)) return 1;
    *out = dst;
    dst += n;
    return 0;
  };
  for (int b = 0; b < CSE_BLOCKS; ++b) {
    for (int path = 0; path < 2; ++path) {
      cse_stack_params& sp = path ? p->block[b].inter : p->block[b].intra;
      for (int l = 0; l < CSE_LAYERS; ++l) {
        cse_layer_params& lp = sp.layer[l];
        if (conv(lp.in_proj_w, (size_t)3 * kN * kN, &lp.in_proj_w_bf16)) return 1;
        if (conv(lp.out_proj_w, (size_t)kN * kN, &lp.out_proj_w_bf16)) return 1;
        if (conv(lp.ffn1_w, (size_t)kFfn * kN, &lp.ffn1_w_bf16)) return 1;
        if (conv(lp.ffn2_w, (size_t)kFfn * kN, &lp.ffn2_w_bf16)) return 1;
      }
    }
  }
  if (conv(p->conv1d_w, (size_t)kN * kN, &p->conv1d_w_bf16)) return 1;
  if (conv(p->conv2d_w, (size_t)n_masks * kN * kN, &p->conv2d_w_bf16)) return 1;
  if (conv(p->out_w, (size_t)kN * kN, &p->out_w_bf16)) return 1;
  if (conv(p->gate_w, (size_t)kN * kN, &p->gate_w_bf16)) return 1;
  if (conv(p->end_w, (size_t)kN * kN, &p->end_w_bf16)) return 1;
  return 0;
}

int cse_forward(const cse_params* p, const float* mix, const float* ctx, int B, int T, int c,
                int n_masks, int precision, float* est, float* pred_head, void* workspace,
                size_t workspace_bytes, void* stream) {
  CSE_REQUIRE(p && mix && est && workspace, "forward: NULL argument");
  Plan pl;
  if (make_plan(B, T, c, n_masks, precision, &pl)) return 1;
  CSE_REQUIRE(workspace_bytes >= pl.total, "forward: workspace too small (%zu < %zu bytes)",
              workspace_bytes, pl.total);
  CSE_REQUIRE(((uintptr_t)workspace & 255) == 0, "forward: workspace must be 256-byte aligned");
  return forward_impl(p, mix, ctx, pl, est, pred_head, (char*)workspace, (cudaStream_t)stream);
}

int cse_forward_host(const cse_params* p, const float* mix_host, const float* ctx_host, int B, int T,
                     int c, int n_masks, int precision, float* est_host, float* pred_head_host,
                     void* workspace, size_t workspace_bytes, void* stream) {
  CSE_REQUIRE(p && mix_host && est_host && workspace, "forward_host: NULL argument");
  Plan pl;
  if (make_plan(B, T, c, n_masks, precision, &pl)) return 1;
  CSE_REQUIRE(workspace_bytes >= pl.total, "forward_host: workspace too small (%zu < %zu bytes)",
              workspace_bytes, pl.total);
  CSE_REQUIRE(((uintptr_t)workspace & 255) == 0, "forward_host: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* d_mix = (float*)(ws + pl.h_mix);
  float* d_ctx = (float*)(ws + pl.h_ctx);
  float* d_est = (float*)(ws + pl.h_est);
  float* d_pred = (float*)(ws + pl.h_pred);
  CSE_CUDA(cudaMemcpyAsync(d_mix, mix_host, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
  if (c > 0) {
    CSE_REQUIRE(ctx_host != nullptr, "forward_host: c=%d but ctx is NULL", c);
    CSE_CUDA(cudaMemcpyAsync(d_ctx, ctx_host, (size_t)B * c * CSE_CTX * 4, cudaMemcpyHostToDevice, st));
  }
  if (forward_impl(p, d_mix, c > 0 ? d_ctx : nullptr, pl, d_est, pred_head_host ? d_pred : nullptr, ws, st))
    return 1;
  CSE_CUDA(cudaMemcpyAsync(est_host, d_est, (size_t)B * T * n_masks * 4, cudaMemcpyDeviceToHost, st));
  if (pred_head_host)
    CSE_CUDA(cudaMemcpyAsync(pred_head_host, d_pred, (size_t)B * kN * 4, cudaMemcpyDeviceToHost, st));
  CSE_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Pipelined host entry: `depth` forwards in flight.  Slot i owns a full workspace (incl. the device
// staging of its inputs / outputs) and a CUDA graph of forward_impl captured on it; three streams
// (H2D, forward, D2H) chained by events, so the copy-in of step i+1 and the copy-out of step i-1
// overlap the forward of step i and the ~220 launches of a forward replay as one graph launch.
// Reference call site: the eval loop `model(mix.cuda(), ctx)` ... `.cpu()` per batch (test.py:231-245).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPipeDepth = 8;

struct PipeSlot {
  char* ws = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t h2d = nullptr, fwd = nullptr, d2h = nullptr;
  bool in_flight = false;
};

}  // extern "C"

struct cse_pipeline {
  cse::Plan pl;
  int depth = 0, next = 0;
  bool has_pred = false;
  cudaStream_t s_in = nullptr, s_fw = nullptr, s_out = nullptr;
  PipeSlot slot[kMaxPipeDepth];
};

static void pipeline_free(cse_pipeline* q) {
  if (q == nullptr) return;
  for (int i = 0; i < q->depth; ++i) {
    PipeSlot& s = q->slot[i];
    if (s.d2h && s.in_flight) cudaEventSynchronize(s.d2h);
    if (s.exec) cudaGraphExecDestroy(s.exec);
    if (s.h2d) cudaEventDestroy(s.h2d);
    if (s.fwd) cudaEventDestroy(s.fwd);
    if (s.d2h) cudaEventDestroy(s.d2h);
  }
  if (q->s_in) cudaStreamDestroy(q->s_in);
  if (q->s_fw) cudaStreamDestroy(q->s_fw);
  if (q->s_out) cudaStreamDestroy(q->s_out);
  delete q;
}

extern "C" {

size_t cse_pipeline_workspace_bytes(int B, int T, int c, int n_masks, int precision, int depth) {
  Plan pl;
  if (depth < 1 || depth > kMaxPipeDepth || make_plan(B, T, c, n_masks, precision, &pl)) return 0;
  return align_up(pl.total, 256) * (size_t)depth;
}

int cse_pipeline_create(const cse_params* p, int B, int T, int c, int n_masks, int precision, int depth,
                        void* workspace, size_t workspace_bytes, cse_pipeline** out) {
  CSE_REQUIRE(p && workspace && out, "pipeline_create: NULL argument");
  CSE_REQUIRE(depth >= 1 && depth <= kMaxPipeDepth, "pipeline_create: depth %d outside [1,%d]", depth, kMaxPipeDepth);
  CSE_REQUIRE(((uintptr_t)workspace & 255) == 0, "pipeline_create: workspace must be 256-byte aligned");
  cse_pipeline* q = new cse_pipeline();
  if (make_plan(B, T, c, n_masks, precision, &q->pl)) { delete q; return 1; }
  const size_t per_slot = align_up(q->pl.total, 256);
  if (workspace_bytes < per_slot * depth) {
    set_error("pipeline_create: workspace too small (%zu < %zu bytes)", workspace_bytes, per_slot * depth);
    delete q;
    return 1;
  }
  q->depth = depth;
  q->has_pred = c > 0;
  auto fail = [&](const char* what, cudaError_t e) {
    set_error("pipeline_create: %s failed: %s", what, cudaGetErrorString(e));
    pipeline_free(q);
    return 1;
  };
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&q->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&q->s_fw, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&q->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
  const Plan& pl = q->pl;
  for (int i = 0; i < depth; ++i) {
    PipeSlot& s = q->slot[i];
    s.ws = (char*)workspace + per_slot * i;
    if ((e = cudaEventCreateWithFlags(&s.h2d, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreateWithFlags(&s.fwd, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreateWithFlags(&s.d2h, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    float* d_mix = (float*)(s.ws + pl.h_mix);
    float* d_ctx = c > 0 ? (float*)(s.ws + pl.h_ctx) : nullptr;
    float* d_est = (float*)(s.ws + pl.h_est);
    float* d_pred = q->has_pred ? (float*)(s.ws + pl.h_pred) : nullptr;
    // inputs of the warm-up run must be finite: zero them
    if ((e = cudaMemsetAsync(d_mix, 0, (size_t)B * T * 4, q->s_fw)) != cudaSuccess) return fail("memset", e);
    if (c > 0 && (e = cudaMemsetAsync(d_ctx, 0, (size_t)B * c * CSE_CTX * 4, q->s_fw)) != cudaSuccess)
      return fail("memset", e);
    // warm-up outside capture: one-time function attributes and tensor-map encodes happen here
    if (forward_impl(p, d_mix, d_ctx, pl, d_est, d_pred, s.ws, q->s_fw)) { pipeline_free(q); return 1; }
    if ((e = cudaStreamSynchronize(q->s_fw)) != cudaSuccess) return fail("warm-up forward", e);
    if ((e = cudaStreamBeginCapture(q->s_fw, cudaStreamCaptureModeThreadLocal)) != cudaSuccess)
      return fail("cudaStreamBeginCapture", e);
    const int rc = forward_impl(p, d_mix, d_ctx, pl, d_est, d_pred, s.ws, q->s_fw);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(q->s_fw, &graph);
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      pipeline_free(q);
      return 1;
    }
    if (e != cudaSuccess) return fail("cudaStreamEndCapture", e);
    e = cudaGraphInstantiate(&s.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail("cudaGraphInstantiate", e);
  }
  *out = q;
  return 0;
}

int cse_pipeline_submit(cse_pipeline* q, const float* mix_host, const float* ctx_host, float* est_host,
                        float* pred_head_host, int* slot_out) {
  CSE_REQUIRE(q && mix_host && est_host, "pipeline_submit: NULL argument");
  const Plan& pl = q->pl;
  const cse_shape& sh = pl.sh;
  CSE_REQUIRE(sh.c == 0 || ctx_host != nullptr, "pipeline_submit: c=%d but ctx is NULL", sh.c);
  CSE_REQUIRE(pred_head_host == nullptr || q->has_pred, "pipeline_submit: pred_head needs c > 0");
  const int i = q->next;
  PipeSlot& s = q->slot[i];
  if (s.in_flight) {  // the slot's previous result has not been collected: wait for it (keeps the depth bounded)
    CSE_CUDA(cudaEventSynchronize(s.d2h));
    s.in_flight = false;
  }
  CSE_CUDA(cudaMemcpyAsync(s.ws + pl.h_mix, mix_host, (size_t)sh.B * sh.T * 4, cudaMemcpyHostToDevice, q->s_in));
  if (sh.c > 0)
    CSE_CUDA(cudaMemcpyAsync(s.ws + pl.h_ctx, ctx_host, (size_t)sh.B * sh.c * CSE_CTX * 4, cudaMemcpyHostToDevice, q->s_in));
  CSE_CUDA(cudaEventRecord(s.h2d, q->s_in));
  CSE_CUDA(cudaStreamWaitEvent(q->s_fw, s.h2d, 0));
  CSE_CUDA(cudaGraphLaunch(s.exec, q->s_fw));
  CSE_CUDA(cudaEventRecord(s.fwd, q->s_fw));
  CSE_CUDA(cudaStreamWaitEvent(q->s_out, s.fwd, 0));
  CSE_CUDA(cudaMemcpyAsync(est_host, s.ws + pl.h_est, (size_t)sh.B * sh.T * pl.n_masks * 4, cudaMemcpyDeviceToHost, q->s_out));
  if (pred_head_host)
    CSE_CUDA(cudaMemcpyAsync(pred_head_host, s.ws + pl.h_pred, (size_t)sh.B * kN * 4, cudaMemcpyDeviceToHost, q->s_out));
  CSE_CUDA(cudaEventRecord(s.d2h, q->s_out));
  s.in_flight = true;
  q->next = (i + 1) % q->depth;
  if (slot_out) *slot_out = i;
  return 0;
}

int cse_pipeline_wait(cse_pipeline* q, int slot) {
  CSE_REQUIRE(q && slot >= 0 && slot < q->depth, "pipeline_wait: bad slot");
  PipeSlot& s = q->slot[slot];
  if (!s.in_flight) return 0;
  CSE_CUDA(cudaEventSynchronize(s.d2h));
  s.in_flight = false;
  return 0;
}

int cse_pipeline_destroy(cse_pipeline* q) {
  pipeline_free(q);
  return 0;
}

int cse_masknet_fwd(const cse_params* p, const void* E, const float* ctx, int B, int L, int c,
                    int n_masks, int precision, float* mask, float* pred_head, void* workspace,
                    size_t workspace_bytes, void* stream) {
  CSE_REQUIRE(p && E && mask && workspace, "masknet_fwd: NULL argument");
  CSE_REQUIRE(L >= 1, "masknet_fwd: L must be >= 1");
  Plan pl;
  if (make_plan(B, kEncS * (L - 1) + kEncK, c, n_masks, precision, &pl)) return 1;
  CSE_REQUIRE(workspace_bytes >= pl.total, "masknet_fwd: workspace too small (%zu < %zu bytes)",
              workspace_bytes, pl.total);
  CSE_REQUIRE(((uintptr_t)workspace & 255) == 0, "masknet_fwd: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  if (launch_gn_stats(E, B, L, precision, (float*)(ws + pl.part), kFinishParts, st)) return 1;
  if (masknet_impl(p, E, kFinishParts, ctx, pl, pred_head, ws, st)) return 1;
  return launch_relu_f32(ws + pl.MP, (size_t)B * L * n_masks * kN, precision, mask, st);
}

// ---- per-stage entry points ----
int cse_encoder_fwd(const float* mix, const float* w, int B, int T, int act_dtype, void* out,
                    float* gn_part, int* n_parts, void* stream) {
  CSE_REQUIRE(mix && w && out && gn_part, "encoder_fwd: NULL argument");
  cse_shape s;
  if (make_shape(B, T, 0, 1, &s)) return 1;
  const int parts = encoder_parts(s.L);
  if (n_parts) *n_parts = parts;
  return launch_encoder(mix, w, B, T, s.L, act_dtype, out, gn_part, parts, (cudaStream_t)stream);
}

int cse_gn_finalize(const float* gn_part, int B, int n_parts, double count, float eps, float* stat,
                    void* stream) {
  CSE_REQUIRE(gn_part && stat && B > 0 && n_parts > 0, "gn_finalize: bad argument");
  return launch_gn_finalize(gn_part, B, n_parts, count, eps, stat, (cudaStream_t)stream);
}

int cse_gn_apply(const void* x, const float* stat, const float* g, const float* b, int B, int L,
                 int act_dtype, void* out, void* stream) {
  CSE_REQUIRE(x && stat && g && b && out, "gn_apply: NULL argument");
  return launch_gn_apply(x, stat, g, b, B, L, act_dtype, out, (cudaStream_t)stream);
}

int cse_linear(const void* A, int lda, const void* W, const float* bias, float bias_scale,
               const float* residual, void* C, int ldc, int M, int N, int K, int relu, int out_fp32,
               int precision, void* stream) {
  CSE_REQUIRE(A && W && C, "linear: NULL argument");
  if (precision == CSE_BF16)
    return launch_gemm_tc((const bf16*)A, lda, (const bf16*)W, bias, bias_scale, residual, C, ldc, M,
                          N, K, relu, out_fp32, (cudaStream_t)stream);
  CSE_REQUIRE(precision == CSE_FP32, "linear: unknown precision %d", precision);
  return launch_gemm_simt((const float*)A, lda, (const float*)W, bias, bias_scale, residual, (float*)C,
                          ldc, M, N, K, relu, (cudaStream_t)stream);
}

int cse_ln_linear(const float* R, const float* gamma, const float* beta, float eps, const void* W_bf16,
                  const float* bias, void* C, int ldc, int M, int N, int relu, void* stream) {
  CSE_REQUIRE(R && gamma && beta && W_bf16 && C, "ln_linear: NULL argument");
  return launch_gemm_ln_tc(R, gamma, beta, eps, (const bf16*)W_bf16, bias, (bf16*)C, ldc, M, N, relu,
                           (cudaStream_t)stream);
}

int cse_linear_residual_ln(const void* A_bf16, int lda, const void* W_bf16, const float* bias, float* R,
                           const float* gamma, const float* beta, float eps, void* H_bf16, int M, int K,
                           void* stream) {
  CSE_REQUIRE(A_bf16 && W_bf16 && bias && R && gamma && beta && H_bf16, "linear_residual_ln: NULL argument");
  return launch_gemm_tc_residual_ln((const bf16*)A_bf16, lda, (const bf16*)W_bf16, bias, R, gamma, beta, eps,
                                    (bf16*)H_bf16, M, K, (cudaStream_t)stream);
}

int cse_ffn_fused(const void* A_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16,
                  const float* b2, float* R, int M, void* stream) {
  CSE_REQUIRE(A_bf16 && W1_bf16 && b1 && W2_bf16 && b2 && R, "ffn_fused: NULL argument");
  return launch_ffn_tc((const bf16*)A_bf16, (const bf16*)W1_bf16, b1, (const bf16*)W2_bf16, b2, R, M,
                       (cudaStream_t)stream);
}

int cse_ffn_ln_fused(float* R, const float* ln2_g, const float* ln2_b, float eps, void* scratch_bf16,
                     const void* W1_bf16, const float* b1, const void* W2_bf16, const float* b2,
                     const float* next_ln1_g, const float* next_ln1_b, void* H1_bf16, int M, void* stream) {
  CSE_REQUIRE(R && ln2_g && ln2_b && scratch_bf16 && W1_bf16 && b1 && W2_bf16 && b2, "ffn_ln_fused: NULL argument");
  return launch_ffn_tc_ln(R, ln2_g, ln2_b, eps, (bf16*)scratch_bf16, (const bf16*)W1_bf16, b1, (const bf16*)W2_bf16,
                          b2, next_ln1_g, next_ln1_b, (bf16*)H1_bf16, M, (cudaStream_t)stream);
}

int cse_layernorm_fwd(const float* x, const float* g, const float* b, int M, float eps, int act_dtype,
                      void* out, void* stream) {
  CSE_REQUIRE(x && g && b && out, "layernorm_fwd: NULL argument");
  return launch_layernorm(x, g, b, M, eps, act_dtype, out, (cudaStream_t)stream);
}

int cse_attention_fwd(const void* qkv, int nseq, int n, int act_dtype, void* out, void* stream) {
  CSE_REQUIRE(qkv && out, "attention_fwd: NULL argument");
  return launch_attention(qkv, nseq, n, act_dtype, out, (cudaStream_t)stream);
}

int cse_segment(const float* x0, int B, int L, int S, float* X, void* stream) {
  CSE_REQUIRE(x0 && X, "segment: NULL argument");
  return launch_segment(x0, B, L, S, X, (cudaStream_t)stream);
}

int cse_build_sequences(const float* X, const float* ctok, const float* pe, int B, int S, int c,
                        int inter, float* R, void* stream) {
  CSE_REQUIRE(X && pe && R && (c == 0 || ctok), "build_sequences: NULL argument");
  return launch_build_sequences(X, ctok, pe, B, S, c, inter, R, (cudaStream_t)stream);
}

int cse_context_map(const float* ctx, const float* w, const float* b, int rows, int in_dim, float* out,
                    void* stream) {
  CSE_REQUIRE(ctx && w && b && out, "context_map: NULL argument");
  return launch_context_map(ctx, w, b, rows, in_dim, out, (cudaStream_t)stream);
}

int cse_stack_finish(const float* R, const float* ln_g, const float* ln_b, const float* gn_g,
                     const float* gn_b, const float* skip, int B, int S, int c, int inter, float* out,
                     float* gn_part, float* stat, void* stream) {
  CSE_REQUIRE(R && ln_g && ln_b && gn_g && gn_b && skip && out && gn_part && stat, "stack_finish: NULL argument");
  return launch_stack_finish(R, ln_g, ln_b, gn_g, gn_b, skip, B, S, c, inter, out, nullptr, nullptr,
                             nullptr, gn_part, stat, (cudaStream_t)stream);
}

int cse_pred_head(const float* R_inter, const float* ln_g, const float* ln_b, int B, int S, int c,
                  float* pred_head, void* stream) {
  CSE_REQUIRE(R_inter && ln_g && ln_b && pred_head, "pred_head: NULL argument");
  return launch_pred_head(R_inter, ln_g, ln_b, B, S, c, pred_head, (cudaStream_t)stream);
}

int cse_prelu_overlap_add(const float* X, const float* prelu, int B, int S, int L, int act_dtype, void* U,
                          void* stream) {
  CSE_REQUIRE(X && prelu && U, "prelu_overlap_add: NULL argument");
  return launch_prelu_ola(X, prelu, B, S, L, act_dtype, U, (cudaStream_t)stream);
}

int cse_gate(const void* o, const void* g, size_t n, int act_dtype, void* out, void* stream) {
  CSE_REQUIRE(o && g && out, "gate: NULL argument");
  return launch_gate(o, g, n, act_dtype, out, (cudaStream_t)stream);
}

int cse_mask_decode(const void* mask_pre, const void* E, const float* dec_w, int B, int L, int T,
                    int n_masks, int act_dtype, float* frames, float* est, void* stream) {
  CSE_REQUIRE(mask_pre && dec_w && frames && est, "mask_decode: NULL argument");  // E may be NULL
  return launch_mask_decode(mask_pre, E, dec_w, B, L, T, n_masks, act_dtype, frames, est,
                            (cudaStream_t)stream);
}

int cse_si_snr(const float* source, const float* estimate, int B, int T, int C, float* out, void* stream) {
  CSE_REQUIRE(source && estimate && out && B > 0 && T > 0, "si_snr: bad argument");
  return launch_si_snr(source, estimate, B, T, C, out, (cudaStream_t)stream);
}

int cse_pit_si_snr(const float* source, const float* estimate_source, int B, int T, int C, float* loss,
                   int* perm, void* stream) {
  CSE_REQUIRE(source && estimate_source && loss && perm && B > 0 && T > 0, "pit_si_snr: bad argument");
  return launch_pit(source, estimate_source, B, T, C, loss, perm, (cudaStream_t)stream);
}

int cse_tm_si_snr(const float* preds, const float* target, int B, int T, float* out, void* stream) {
  CSE_REQUIRE(preds && target && out && B > 0 && T > 0, "tm_si_snr: bad argument");
  return launch_tm_si_snr(preds, target, B, T, out, (cudaStream_t)stream);
}

// ---- ContSep selection tail ----
int cse_selection_loss(const float* gt, const float* est, const float* logits, int B, int T, int n_streams,
                       int ce, float* sisnr, long long* label, float* loss, float* dlogits, float* item_loss,
                       void* stream) {
  CSE_REQUIRE(gt && est && logits && sisnr && label && loss && dlogits && item_loss && B > 0 && T > 0,
              "selection_loss: bad argument");
  CSE_REQUIRE(ce || n_streams == 2, "selection_loss: the single-logit BCE head exists for 2 streams only (ContSep.py:46-51)");
  return launch_selection_loss(gt, est, logits, B, T, n_streams, ce, sisnr, label, item_loss, loss, dlogits,
                               (cudaStream_t)stream);
}

int cse_select_stream(const float* est, const float* logits, int B, int T, int n_streams, int ce, float* out,
                      long long* pick, void* stream) {
  CSE_REQUIRE(est && logits && out && pick && B > 0 && T > 0 && n_streams >= 1, "select_stream: bad argument");
  CSE_REQUIRE(ce || n_streams == 2, "select_stream: the single-logit head exists for 2 streams only");
  return launch_select_stream(est, logits, B, T, n_streams, ce, out, pick, (cudaStream_t)stream);
}

int cse_selection_accuracy(const float* enhanced, const float* sources, int B, int T, int n_sources, float* sisnr,
                           int* acc, void* stream) {
  CSE_REQUIRE(enhanced && sources && sisnr && acc && B > 0 && T > 0, "selection_accuracy: bad argument");
  return launch_selection_accuracy(enhanced, sources, B, T, n_sources, sisnr, acc, (cudaStream_t)stream);
}

}  // extern "C"
