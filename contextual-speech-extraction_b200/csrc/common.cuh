// common.cuh — shared device/host helpers for the sm_100a Sepformer kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cse_b200.h"

namespace cse {

constexpr int kN = CSE_N;          // 256 channels
constexpr int kK = CSE_K;          // 250 frames per chunk
constexpr int kP = CSE_K / 2;      // hop
constexpr int kHeads = CSE_HEADS;  // 8
constexpr int kDh = CSE_N / CSE_HEADS;  // 32
constexpr int kFfn = CSE_FFN;
constexpr int kEncK = 16;
constexpr int kEncS = 8;

typedef __nv_bfloat16 bf16;

// ---- error plumbing (abi.cu owns the thread-local buffer) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define CSE_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      cse::set_error(__VA_ARGS__);      \
      return 1;                         \
    }                                   \
  } while (0)

#define CSE_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      cse::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                           \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

// ---- activation load/store: 8 consecutive channels per lane (256 = 32 lanes x 8) ----
struct f8 {
  float v[8];
};

__device__ __forceinline__ f8 ld8(const float* p) {
  f8 r;
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ f8 ld8(const bf16* p) {
  f8 r;
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void st8(float* p, const f8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const f8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Row access for one-warp-per-row kernels over 256 channels: lane L owns channels [4L, 4L + 4) and [128 + 4L, +4), so
// each of the two 16-byte accesses of a row is one contiguous 512-byte request (with 8 consecutive channels per lane
// both accesses touch all 32 sectors of the row and use half of each).  `row` points at the row's first channel.
__device__ __forceinline__ f8 ld_row8(const float* row, int lane) {
  const float4 a = reinterpret_cast<const float4*>(row)[lane];
  const float4 b = reinterpret_cast<const float4*>(row)[32 + lane];
  f8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ f8 ld_row8_cg(const float* row, int lane) {  // L2 only (rows another proxy rewrites)
  const float4 a = __ldcg(reinterpret_cast<const float4*>(row) + lane);
  const float4 b = __ldcg(reinterpret_cast<const float4*>(row) + 32 + lane);
  f8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_row8(float* row, int lane, const f8& r) {
  reinterpret_cast<float4*>(row)[lane] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  reinterpret_cast<float4*>(row)[32 + lane] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void st_row8(bf16* row, int lane, const f8& r) {
  __nv_bfloat162 lo[2] = {__floats2bfloat162_rn(r.v[0], r.v[1]), __floats2bfloat162_rn(r.v[2], r.v[3])};
  __nv_bfloat162 hi[2] = {__floats2bfloat162_rn(r.v[4], r.v[5]), __floats2bfloat162_rn(r.v[6], r.v[7])};
  reinterpret_cast<uint2*>(row)[lane] = *reinterpret_cast<const uint2*>(lo);
  reinterpret_cast<uint2*>(row)[32 + lane] = *reinterpret_cast<const uint2*>(hi);
}

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }
// value after a store/load round trip through the activation type
template <typename T> __device__ __forceinline__ float round_act(float x) { return to_f(from_f<T>(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Row LayerNorm of 256 channels held as 8 per lane; two-pass (mean, then centred variance).  NR rows at once: the
// per-row arithmetic is the same for every NR (same results), but the NR butterfly reductions advance together —
// NR independent shuffles per step instead of NR dependent chains of five.
template <int NR>
__device__ __forceinline__ void ln_rows(f8 (&x)[NR], const f8& g, const f8& b, float eps) {
  float s[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[r] += x[r].v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < NR; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  }
  float q[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const float mean = s[r] * (1.0f / kN);
    q[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[r].v[i] -= mean;
      q[r] += x[r].v[i] * x[r].v[i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < NR; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const float rstd = rsqrtf(q[r] * (1.0f / kN) + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[r].v[i] = x[r].v[i] * rstd * g.v[i] + b.v[i];
  }
}
__device__ __forceinline__ void ln_row(f8& x, const f8& g, const f8& b, float eps) {
  f8 (&one)[1] = reinterpret_cast<f8 (&)[1]>(x);
  ln_rows<1>(one, g, b, eps);
}

// ---- optional per-kernel-class device timing (bench.py roofline leg; off by default) ----
enum KernelClass { kClsGemmTc = 0, kClsAttention = 1, kClsLayerNorm = 2, kClsGemmSimt = 3, kClsFfn = 4, kClsCount = 5 };
struct KernelScope {  // records a CUDA event pair around the launches made in its lifetime
  KernelScope(int cls, cudaStream_t st);
  ~KernelScope();
  int slot_;
  cudaStream_t st_;
};

// ---- programmatic dependent launch (PDL) ----
// The kernels of a transformer layer run back to back on one stream.  Launched with the programmatic-
// serialization attribute, kernel N+1 may be scheduled onto an SM as soon as kernel N's CTA there has
// retired: its prologue (barrier init, TMEM allocation, descriptor fetch) then overlaps N's tail, and it
// blocks in pdl_wait() until ALL of N's memory operations are complete and visible.  Every kernel calls
// pdl_launch_dependents() first and pdl_wait() before its first global read OR write; both are no-ops in a
// launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
bool pdl_enabled();  // abi.cu: CSE_PDL=0 turns the launch attribute off (A/B aid)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- per-device one-time setup ----
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per device: a process that touches a second
// GPU (cuda:1 after cuda:0) must repeat the opt-in there.  `configured_on_this_device()` returns true once the caller
// has called `mark_configured()` on the CURRENT device.
struct DeviceOnce {
  unsigned long long done = 0;  // one bit per device ordinal (ordinals >= 64 are simply re-configured every call)
  bool configured_on_this_device() const {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < 64 && ((done >> dev) & 1ull);
  }
  void mark_configured() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64) done |= 1ull << dev;
  }
};
int sm_count();  // gemm_tc.cu: multiprocessors of the CURRENT device

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- kernel launchers shared between translation units (defined in the named .cu) ----
// frontend.cu
int launch_encoder(const float* mix, const float* w, int B, int T, int L, int act, void* out,
                   float* gn_part, int n_parts, cudaStream_t st);
int encoder_parts(int L);
int launch_gn_finalize(const float* part, int B, int n_parts, double count, float eps, float* stat,
                       cudaStream_t st);
int launch_gn_stats(const void* x, int B, int rows, int act, float* gn_part, int n_parts,
                    cudaStream_t st);
int launch_gn_apply(const void* x, const float* stat, const float* g, const float* b, int B, int L,
                    int act, void* out, cudaStream_t st);
int launch_segment(const float* x0, int B, int L, int S, float* X, cudaStream_t st);
// ln_g / ln_b / H (optional): also write norm1 of the stack's first layer (bf16) for every row
int launch_build_sequences(const float* X, const float* ctok, const float* pe, int B, int S, int c,
                           int inter, float* R, cudaStream_t st, const float* ln_g = nullptr,
                           const float* ln_b = nullptr, bf16* H = nullptr);
int launch_context_map(const float* ctx, const float* w, const float* b, int rows, int in_dim,
                       float* out, cudaStream_t st);
// norm.cu
int launch_layernorm(const float* x, const float* g, const float* b, int M, float eps, int act,
                     void* out, cudaStream_t st);
constexpr int kFinishParts = 64;  // GroupNorm partials per sample in stack_finish
int launch_stack_finish(const float* R, const float* ln_g, const float* ln_b, const float* gn_g,
                        const float* gn_b, const float* skip, int B, int S, int c, int inter,
                        float* out, float* next_R, const float* next_pe, const float* next_ctok,
                        float* gn_part, float* stat, cudaStream_t st, const float* next_ln_g = nullptr,
                        const float* next_ln_b = nullptr, bf16* next_H = nullptr);
int launch_pred_head(const float* R, const float* ln_g, const float* ln_b, int B, int S, int c,
                     float* out, cudaStream_t st);
// gemm_simt.cu / gemm_tc.cu
int launch_gemm_simt(const float* A, int lda, const float* W, const float* bias, float bias_scale,
                     const float* residual, float* C, int ldc, int M, int N, int K, int relu,
                     cudaStream_t st);
int launch_gemm_tc(const bf16* A, int lda, const bf16* W, const float* bias, float bias_scale,
                   const float* residual, void* C, int ldc, int M, int N, int K, int relu,
                   int out_fp32, cudaStream_t st);
int launch_attention_bwd_bf16(const bf16* qkv, const bf16* out, const float* d_out, int nseq, int n, float* d_qkv,
                              cudaStream_t st);  // tensor-core attention backward, n <= 256
int launch_gemm_tc_residual_ln(const bf16* A, int lda, const bf16* W, const float* bias, float* R, const float* gamma,
                               const float* beta, float eps, bf16* H, int M, int K, cudaStream_t st);
int launch_gemm_tc_dgrad(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int out_fp32, int M, int N,
                         int K, cudaStream_t st);  // C = A W for the row-major W [K,N] (B operand MN-major)
int launch_gemm_tc_wgrad(const bf16* X, int ldx, const bf16* Y, int ldy, float* C, int ldc, int M, int N, int T,
                         cudaStream_t st);  // C (fp32) += X^T Y over the T tokens (MN-major operands, split over the CTA pairs)
int launch_gemm_tc_splitk(const bf16* A, int lda, const bf16* W, int ldw, float* C, int ldc, int M, int N, int K,
                          cudaStream_t st);  // C (fp32) += A W^T, K split over the CTA pairs (wgrad)
// gemm_ln_tc.cu: C(bf16) = act(LN(R) W^T + bias), K = 256
int launch_ffn_tc(const bf16* A, const bf16* W1, const float* b1, const bf16* W2, const float* b2, float* R,
                  int M, cudaStream_t st);
// R += FFN(norm2(R)); H1 = norm1_next(R) (nullptr: not produced); scratch [M,256] bf16 receives norm2(R)
int launch_ffn_tc_ln(float* R, const float* ln2_g, const float* ln2_b, float eps, bf16* scratch, const bf16* W1,
                     const float* b1, const bf16* W2, const float* b2, const float* ln1n_g, const float* ln1n_b,
                     bf16* H1, int M, cudaStream_t st);
int launch_gemm_ln_tc(const float* R, const float* gamma, const float* beta, float eps, const bf16* W,
                      const float* bias, bf16* C, int ldc, int M, int N, int relu, cudaStream_t st);
// attention.cu
int launch_attention(const void* qkv, int nseq, int n, int act, void* out, cudaStream_t st);
int launch_attention_tc(const bf16* qkv, int nseq, int n, bf16* out, cudaStream_t st);  // attention_tc.cu
extern int g_attention_mode;             // 0 auto, 1 force mma.sync, 2 force tcgen05 v1, 4 force tcgen05 v4
extern long long* g_attention_trace;    // attention_tc.cu: optional clock64 trace buffer of CTA 0 (debug)
// head.cu
int launch_prelu_ola(const float* X, const float* prelu, int B, int S, int L, int act, void* U,
                     cudaStream_t st);
int launch_gate(const void* o, const void* g, size_t n, int act, void* out, cudaStream_t st);
int launch_relu_f32(const void* x, size_t n, int act, float* out, cudaStream_t st);
int launch_mask_decode(const void* mask_pre, const void* E, const float* dec_w, int B, int L, int T,
                       int n_masks, int act, float* frames, float* est, cudaStream_t st);
// loss.cu
int launch_si_snr(const float* source, const float* estimate, int B, int T, int C, float* out,
                  cudaStream_t st);
int launch_pit(const float* source, const float* est, int B, int T, int C, float* loss, int* perm,
               cudaStream_t st);
int launch_tm_si_snr(const float* preds, const float* target, int B, int T, float* out,
                     cudaStream_t st);
int launch_si_snr_bwd(const float* source, const float* estimate, int B, int T, int C, int mode,
                      const float* gout, const int* perm, float* d_source, float* d_estimate,
                      cudaStream_t st);
int launch_tm_si_snr_bwd(const float* preds, const float* target, int B, int T, const float* gout,
                         float* d_preds, float* d_target, cudaStream_t st);
int launch_selection_loss(const float* gt, const float* est, const float* logits, int B, int T, int C, int ce,
                          float* sisnr, long long* label, float* item_loss, float* loss, float* dlogits,
                          cudaStream_t st);
int launch_select_stream(const float* est, const float* logits, int B, int T, int C, int ce, float* out,
                         long long* pick, cudaStream_t st);
int launch_selection_accuracy(const float* enhanced, const float* sources, int B, int T, int C, float* sisnr,
                              int* acc, cudaStream_t st);
// backward.cu (fp32 training path)
int launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t st);
int launch_wgrad(const float* X, int ldx, const float* Y, int ldy, int M, int N1, int N2, float* dW,
                 cudaStream_t st);
int launch_colsum(const float* dC, int ld, int M, int N, float* db, cudaStream_t st);
int launch_colsum_cast(const float* dC, const bf16* relu, int M, int N, float* db, bf16* dC16,
                       cudaStream_t st);  // + bf16 copy of dC, optionally through the ReLU mask of the saved activation
int launch_relu_bwd(const float* F, float* d, size_t n, cudaStream_t st);
int launch_layernorm_bwd(const float* x, const float* g, const float* dy, int M, float eps, float* dx,
                         int accumulate, float* dg, float* db, cudaStream_t st);
int launch_attention_bwd(const float* qkv, const float* out, const float* d_out, int nseq, int n,
                         float* d_qkv, cudaStream_t st);
// pack.cu (in abi.cu)
int launch_f32_to_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st);

}  // namespace cse
