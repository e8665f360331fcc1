// optim.cu — the optimiser step of the training loop as three launches over a chunk table.
//
// Reference (train_ContSep.py:233,402-419; train_ContExt.py:372-389):
//     scaler.unscale_(optimizer)                                          (fp16 only)
//     grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
//     if not isfinite(grad_norm): skip          /   scaler.step(optimizer); scaler.update()
//     optimizer.step()        with optim.AdamW(params, lr, weight_decay, amsgrad=True)
// which in eager PyTorch is ~30 M parameters x (unscale, norm, clip, 5 AdamW passes) over ~420 tensors: several
// hundred small launches and a host synchronisation on the norm.  Here:
//   optim_sumsq_kernel    one CTA per chunk: sum of squares of the gradient chunk -> partial[chunk]
//   optim_finalize_kernel one CTA: deterministic sum of the partials in double -> total norm of the UNSCALED
//                         gradients, clip coefficient, found_inf, step counter and bias corrections, and the
//                         GradScaler.update() rule — all kept in a small caller-owned device state, no host sync
//   optim_adamw_kernel    one CTA per chunk: AdamW(amsgrad) update with the clip coefficient and the inverse loss
//                         scale folded into the gradient load; skipped on device when found_inf
// HBM-bound: 4 B read for the norm + 20 B read / 16 B written per parameter for the update.
#include "common.cuh"

namespace cse {

struct OptimChunk {  // one contiguous run of <= kOptChunk elements of one parameter tensor
  float* p;
  float* g;
  float* m;
  float* v;
  float* vmax;
  int n;
  int pad;
};
static_assert(sizeof(OptimChunk) == 48, "chunk descriptor layout is part of the ABI");

constexpr int kOptChunk = 16384;
constexpr int kOptThreads = 256;

// device state, 16 floats (cse_optim_state in the header)
struct OptimState {
  double step;          // number of updates applied so far (torch keeps it as a float tensor per parameter)
  float scale;          // GradScaler scale (1 when no scaler is used)
  int growth_tracker;   // GradScaler: consecutive finite steps
  float total_norm;     // norm of the unscaled gradients of the last call (clip_grad_norm_'s return value)
  float coef;           // inv_scale * min(1, max_norm / (total_norm + 1e-6)) — what every gradient is multiplied by
  int found_inf;        // 1: the last call skipped the update
  float step_size;      // lr / (1 - beta1^step)
  float bc2_sqrt;       // sqrt(1 - beta2^step)
  float reserved[7];
};
static_assert(sizeof(OptimState) == 64, "optimiser state layout is part of the ABI");

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < kOptThreads / 32) ? red[threadIdx.x] : 0.f;
  if (warp == 0) t = warp_sum(t);
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(kOptThreads) optim_sumsq_kernel(const OptimChunk* __restrict__ chunks,
                                                                 float* __restrict__ partial) {
  __shared__ float red[kOptThreads / 32];
  const OptimChunk c = chunks[blockIdx.x];
  float s = 0.f;
  const int n4 = ((reinterpret_cast<uintptr_t>(c.g) & 15) == 0) ? (c.n >> 2) : 0;
  const float4* g4 = reinterpret_cast<const float4*>(c.g);
  for (int i = threadIdx.x; i < n4; i += kOptThreads) {
    const float4 g = g4[i];
    s += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
  }
  for (int i = 4 * n4 + threadIdx.x; i < c.n; i += kOptThreads) s += c.g[i] * c.g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

struct OptimHyper {
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  float growth_factor, backoff_factor;
  int growth_interval, use_scaler;
};

__global__ void __launch_bounds__(kOptThreads) optim_finalize_kernel(const float* __restrict__ partial, int n_chunks,
                                                                    OptimHyper h, OptimState* __restrict__ st) {
  __shared__ double red[kOptThreads];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_chunks; i += kOptThreads) s += (double)partial[i];  // fixed order: deterministic
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = kOptThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const float scale = h.use_scaler ? st->scale : 1.f;
  const double inv_scale = 1.0 / (double)scale;  // GradScaler.unscale_: grads *= 1/scale
  const double total = sqrt(red[0]) * inv_scale;
  const bool finite = isfinite(total) && isfinite(red[0]);
  st->total_norm = (float)total;
  st->found_inf = finite ? 0 : 1;
  if (finite) {
    // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1 (max_norm <= 0: no clipping)
    double coef = 1.0;
    if (h.max_norm > 0.f) coef = fmin(1.0, (double)h.max_norm / (total + 1e-6));
    st->coef = (float)(coef * inv_scale);
    const double step = st->step + 1.0;
    st->step = step;
    st->step_size = (float)((double)h.lr / (1.0 - pow((double)h.beta1, step)));
    st->bc2_sqrt = (float)sqrt(1.0 - pow((double)h.beta2, step));
  } else {
    st->coef = 0.f;
  }
  if (h.use_scaler) {  // GradScaler.update()
    if (!finite) {
      st->scale = scale * h.backoff_factor;
      st->growth_tracker = 0;
    } else if (++st->growth_tracker == h.growth_interval) {
      st->scale = scale * h.growth_factor;
      st->growth_tracker = 0;
    }
  }
}

// torch.optim.AdamW single-tensor rule, amsgrad (torch/optim/adamw.py -> adam.py:_single_tensor_adam):
//   p *= 1 - lr*wd;  m = lerp(m, g, 1-beta1);  v = v*beta2 + (1-beta2) g^2;  vmax = max(vmax, v);
//   p -= step_size * m / (sqrt(vmax) / sqrt(bc2) + eps)
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float& vmax, float coef,
                                          float decay, float w1, float beta2, float w2, float step_size,
                                          float bc2_sqrt, float eps, bool amsgrad) {
  g *= coef;
  p *= decay;
  m = m + w1 * (g - m);
  v = v * beta2;
  v = v + w2 * g * g;
  float d = v;
  if (amsgrad) {
    vmax = fmaxf(vmax, v);
    d = vmax;
  }
  const float denom = sqrtf(d) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kOptThreads) optim_adamw_kernel(const OptimChunk* __restrict__ chunks,
                                                                 const OptimState* __restrict__ st, OptimHyper h,
                                                                 int amsgrad, int write_back_grads) {
  if (st->found_inf) return;  // non-finite gradients: the reference skips the update (and the scaler backs off)
  const OptimChunk c = chunks[blockIdx.x];
  const float coef = st->coef, step_size = st->step_size, bc2s = st->bc2_sqrt;
  const float decay = 1.f - h.lr * h.weight_decay, w1 = 1.f - h.beta1, w2 = 1.f - h.beta2;
  const bool ams = amsgrad != 0;
  const bool al = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) |
                    reinterpret_cast<uintptr_t>(c.m) | reinterpret_cast<uintptr_t>(c.v) |
                    (ams ? reinterpret_cast<uintptr_t>(c.vmax) : 0)) & 15) == 0;
  const int n4 = al ? (c.n >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += kOptThreads) {
    float4 p = reinterpret_cast<float4*>(c.p)[i];
    float4 g = reinterpret_cast<const float4*>(c.g)[i];
    float4 m = reinterpret_cast<float4*>(c.m)[i];
    float4 v = reinterpret_cast<float4*>(c.v)[i];
    float4 x = ams ? reinterpret_cast<float4*>(c.vmax)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    adamw_one(p.x, g.x, m.x, v.x, x.x, coef, decay, w1, h.beta2, w2, step_size, bc2s, h.eps, ams);
    adamw_one(p.y, g.y, m.y, v.y, x.y, coef, decay, w1, h.beta2, w2, step_size, bc2s, h.eps, ams);
    adamw_one(p.z, g.z, m.z, v.z, x.z, coef, decay, w1, h.beta2, w2, step_size, bc2s, h.eps, ams);
    adamw_one(p.w, g.w, m.w, v.w, x.w, coef, decay, w1, h.beta2, w2, step_size, bc2s, h.eps, ams);
    reinterpret_cast<float4*>(c.p)[i] = p;
    reinterpret_cast<float4*>(c.m)[i] = m;
    reinterpret_cast<float4*>(c.v)[i] = v;
    if (ams) reinterpret_cast<float4*>(c.vmax)[i] = x;
    if (write_back_grads)
      reinterpret_cast<float4*>(c.g)[i] = make_float4(g.x * coef, g.y * coef, g.z * coef, g.w * coef);
  }
  for (int i = 4 * n4 + threadIdx.x; i < c.n; i += kOptThreads) {
    float p = c.p[i], m = c.m[i], v = c.v[i], x = ams ? c.vmax[i] : 0.f;
    const float g = c.g[i];
    adamw_one(p, g, m, v, x, coef, decay, w1, h.beta2, w2, step_size, bc2s, h.eps, ams);
    c.p[i] = p;
    c.m[i] = m;
    c.v[i] = v;
    if (ams) c.vmax[i] = x;
    if (write_back_grads) c.g[i] = g * coef;
  }
}

}  // namespace cse

using namespace cse;

extern "C" {

long long cse_optim_chunk_count(int n_tensors, const long long* numel) {
  if (n_tensors < 0 || (n_tensors > 0 && numel == nullptr)) return -1;
  long long n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (numel[i] < 0) return -1;
    n += (numel[i] + kOptChunk - 1) / kOptChunk;
  }
  return n;
}

int cse_optim_table_fill(int n_tensors, const long long* numel, void* const* param, void* const* grad,
                         void* const* exp_avg, void* const* exp_avg_sq, void* const* max_exp_avg_sq,
                         void* host_table, size_t host_table_bytes) {
  CSE_REQUIRE(n_tensors > 0 && numel && param && grad && exp_avg && exp_avg_sq && host_table,
              "optim_table_fill: NULL argument");
  const long long n_chunks = cse_optim_chunk_count(n_tensors, numel);
  CSE_REQUIRE(n_chunks > 0, "optim_table_fill: empty parameter list");
  CSE_REQUIRE(host_table_bytes >= (size_t)n_chunks * sizeof(OptimChunk), "optim_table_fill: table too small (%zu < %zu)",
              host_table_bytes, (size_t)n_chunks * sizeof(OptimChunk));
  OptimChunk* t = reinterpret_cast<OptimChunk*>(host_table);
  long long k = 0;
  for (int i = 0; i < n_tensors; ++i) {
    CSE_REQUIRE(numel[i] == 0 || (param[i] && grad[i] && exp_avg[i] && exp_avg_sq[i]),
                "optim_table_fill: tensor %d has a NULL buffer", i);
    for (long long off = 0; off < numel[i]; off += kOptChunk, ++k) {
      OptimChunk& c = t[k];
      c.p = (float*)param[i] + off;
      c.g = (float*)grad[i] + off;
      c.m = (float*)exp_avg[i] + off;
      c.v = (float*)exp_avg_sq[i] + off;
      c.vmax = (max_exp_avg_sq && max_exp_avg_sq[i]) ? (float*)max_exp_avg_sq[i] + off : nullptr;
      c.n = (int)((numel[i] - off) < kOptChunk ? (numel[i] - off) : kOptChunk);
      c.pad = 0;
    }
  }
  return 0;
}

int cse_optim_table_set_grads(int n_tensors, const long long* numel, void* const* grad, void* host_table,
                              size_t host_table_bytes) {
  CSE_REQUIRE(n_tensors > 0 && numel && grad && host_table, "optim_table_set_grads: NULL argument");
  const long long n_chunks = cse_optim_chunk_count(n_tensors, numel);
  CSE_REQUIRE(n_chunks > 0 && host_table_bytes >= (size_t)n_chunks * sizeof(OptimChunk),
              "optim_table_set_grads: table too small");
  OptimChunk* t = reinterpret_cast<OptimChunk*>(host_table);
  long long k = 0;
  for (int i = 0; i < n_tensors; ++i) {
    CSE_REQUIRE(numel[i] == 0 || grad[i], "optim_table_set_grads: tensor %d has a NULL gradient", i);
    for (long long off = 0; off < numel[i]; off += kOptChunk, ++k) t[k].g = (float*)grad[i] + off;
  }
  return 0;
}

int cse_optim_step(const void* device_table, long long n_chunks, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int amsgrad, float max_norm, int use_scaler, float growth_factor,
                   float backoff_factor, int growth_interval, int write_back_grads, void* state, float* partial,
                   void* stream) {
  CSE_REQUIRE(device_table && state && partial, "optim_step: NULL argument");
  CSE_REQUIRE(n_chunks > 0 && n_chunks < 2147483647LL, "optim_step: bad chunk count %lld", n_chunks);
  CSE_REQUIRE(((uintptr_t)state & 7) == 0, "optim_step: state must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  OptimHyper h{lr, beta1, beta2, eps, weight_decay, max_norm, growth_factor, backoff_factor, growth_interval,
               use_scaler};
  const OptimChunk* chunks = reinterpret_cast<const OptimChunk*>(device_table);
  optim_sumsq_kernel<<<(int)n_chunks, kOptThreads, 0, st>>>(chunks, partial);
  if (check_launch("optim_sumsq_kernel")) return 1;
  optim_finalize_kernel<<<1, kOptThreads, 0, st>>>(partial, (int)n_chunks, h, reinterpret_cast<OptimState*>(state));
  if (check_launch("optim_finalize_kernel")) return 1;
  optim_adamw_kernel<<<(int)n_chunks, kOptThreads, 0, st>>>(chunks, reinterpret_cast<const OptimState*>(state), h,
                                                            amsgrad, write_back_grads);
  return check_launch("optim_adamw_kernel");
}

}  // extern "C"
