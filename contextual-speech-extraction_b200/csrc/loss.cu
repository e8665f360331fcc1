// loss.cu — SI-SNR family as one-pass warp-shuffle reductions.
//
// Reference: speechbrain cal_si_snr / PitWrapper / get_si_snr_with_pitwrapper
// (train_ContSep.py:346,352,386,391-393; test.py:248-252) and torchmetrics
// ScaleInvariantSignalNoiseRatio (train_ContExt.py:339,367).  The reference builds the C x C
// pairwise matrix by repeating tensors and loops over batch items and permutations in Python
// (~20 tiny reductions per item); here one CTA per item accumulates the five sufficient
// statistics {sum a, sum a^2, sum b, sum b^2, sum a_i b_j} in double in a single pass
// (4*T*C*2 algorithmic bytes per item) and the closed forms below finish the job.
#include "common.cuh"

namespace cse {

constexpr int kMaxC = 4;
constexpr int kLossThreads = 512;

struct PairStats {
  double sa[kMaxC], saa[kMaxC], sb[kMaxC], sbb[kMaxC], sab[kMaxC][kMaxC];
};

// a, b: [T, C] slices of one item (row stride C).  Result valid in thread 0.
template <int C>
__device__ void pair_stats(const float* __restrict__ a, const float* __restrict__ b, int T,
                           PairStats& out) {
  constexpr int NV = 4 * C + C * C;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  for (int t = threadIdx.x; t < T; t += kLossThreads) {
    float av[C], bv[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
      av[i] = a[(size_t)t * C + i];
      bv[i] = b[(size_t)t * C + i];
    }
#pragma unroll
    for (int i = 0; i < C; ++i) {
      acc[i] += (double)av[i];
      acc[C + i] += (double)av[i] * (double)av[i];
      acc[2 * C + i] += (double)bv[i];
      acc[3 * C + i] += (double)bv[i] * (double)bv[i];
#pragma unroll
      for (int j = 0; j < C; ++j) acc[4 * C + i * C + j] += (double)av[i] * (double)bv[j];
    }
  }
  __shared__ double s_red[kLossThreads / 32][NV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double v = 0.0;
      for (int w = 0; w < kLossThreads / 32; ++w) v += s_red[w][i];
      acc[i] = v;
    }
#pragma unroll
    for (int i = 0; i < C; ++i) {
      out.sa[i] = acc[i];
      out.saa[i] = acc[C + i];
      out.sb[i] = acc[2 * C + i];
      out.sbb[i] = acc[3 * C + i];
#pragma unroll
      for (int j = 0; j < C; ++j) out.sab[i][j] = acc[4 * C + i * C + j];
    }
  }
}

// speechbrain cal_si_snr for the pair (source column i of a, estimate column j of b):
//   s = a_i - mean, e = b_j - mean, dot = <e,s>, energy = |s|^2 + eps, proj = dot*s/energy,
//   returns -10 log10(|proj|^2 / (|e - proj|^2 + eps) + eps),  eps = 1e-8.
__device__ double sb_neg_si_snr(const PairStats& st, int i, int j, int T) {
  const double eps = 1e-8;
  const double n = (double)T;
  const double ss = st.saa[i] - st.sa[i] * st.sa[i] / n;
  const double ee = st.sbb[j] - st.sb[j] * st.sb[j] / n;
  const double dot = st.sab[i][j] - st.sa[i] * st.sb[j] / n;
  const double energy = ss + eps;
  const double proj2 = dot * dot * ss / (energy * energy);
  double noise2 = ee - 2.0 * dot * dot / energy + proj2;
  if (noise2 < 0.0) noise2 = 0.0;
  const double ratio = proj2 / (noise2 + eps);
  return -10.0 * log10(ratio + eps);
}

template <int C>
__global__ void __launch_bounds__(kLossThreads) si_snr_kernel(const float* __restrict__ source,
                                                              const float* __restrict__ estimate,
                                                              int T, float* __restrict__ out) {
  __shared__ PairStats st;
  const size_t b = blockIdx.x;
  pair_stats<C>(source + b * T * C, estimate + b * T * C, T, st);
  if (threadIdx.x == 0) {
    for (int i = 0; i < C; ++i) out[b * C + i] = (float)sb_neg_si_snr(st, i, i, T);
  }
}

// PitWrapper: loss_mat[i][j] = cal_si_snr(source = source[:, j], estimate = estimate_source[:, i]);
// permutations in itertools (lexicographic) order, strict '>' so the first minimum wins.
template <int C>
__global__ void __launch_bounds__(kLossThreads) pit_kernel(const float* __restrict__ source,
                                                           const float* __restrict__ est_src, int T,
                                                           float* __restrict__ loss,
                                                           int* __restrict__ perm) {
  __shared__ PairStats st;
  const size_t b = blockIdx.x;
  pair_stats<C>(source + b * T * C, est_src + b * T * C, T, st);
  if (threadIdx.x == 0) {
    float mat[C][C];
    for (int i = 0; i < C; ++i)
      for (int j = 0; j < C; ++j) mat[i][j] = (float)sb_neg_si_snr(st, j, i, T);
    int p[C], best_p[C];
    for (int i = 0; i < C; ++i) p[i] = i;
    float best = 0.f;
    bool have = false;
    while (true) {
      float v = 0.f;
      for (int i = 0; i < C; ++i) v += mat[i][p[i]];
      v /= (float)C;
      if (!have || best > v) {
        best = v;
        have = true;
        for (int i = 0; i < C; ++i) best_p[i] = p[i];
      }
      // next lexicographic permutation
      int k = C - 2;
      while (k >= 0 && p[k] > p[k + 1]) --k;
      if (k < 0) break;
      int l = C - 1;
      while (p[l] < p[k]) --l;
      int tmp = p[k]; p[k] = p[l]; p[l] = tmp;
      for (int x = k + 1, y = C - 1; x < y; ++x, --y) { tmp = p[x]; p[x] = p[y]; p[y] = tmp; }
    }
    loss[b] = best;
    for (int i = 0; i < C; ++i) perm[b * C + i] = best_p[i];
  }
}

// torchmetrics SI-SNR: alpha = (<p,t>+eps)/(|t|^2+eps); 10 log10((|alpha t|^2+eps)/(|alpha t - p|^2+eps)),
// zero-mean inputs, eps = finfo(float32).eps.
__global__ void __launch_bounds__(kLossThreads) tm_si_snr_kernel(const float* __restrict__ preds,
                                                                 const float* __restrict__ target,
                                                                 int T, float* __restrict__ out) {
  __shared__ PairStats st;
  const size_t b = blockIdx.x;
  pair_stats<1>(target + b * T, preds + b * T, T, st);
  if (threadIdx.x == 0) {
    const double eps = 1.1920928955078125e-07;
    const double n = (double)T;
    const double tt = st.saa[0] - st.sa[0] * st.sa[0] / n;
    const double pp = st.sbb[0] - st.sb[0] * st.sb[0] / n;
    const double pt = st.sab[0][0] - st.sa[0] * st.sb[0] / n;
    const double alpha = (pt + eps) / (tt + eps);
    const double ts2 = alpha * alpha * tt;
    double noise2 = ts2 - 2.0 * alpha * pt + pp;
    if (noise2 < 0.0) noise2 = 0.0;
    out[b] = (float)(10.0 * log10((ts2 + eps) / (noise2 + eps)));
  }
}

// ---------------------------------------------------------------------------------------------
// Backward.  Every loss above is a function of the five centred sufficient statistics of a pair
// (ss = |s~|^2, ee = |e~|^2, d = <e~, s~>), so its gradient with respect to the raw signals is
//   dL/ds[t] = 2 G_ss s~[t] + G_d e~[t],   dL/de[t] = 2 G_ee e~[t] + G_d s~[t]
// (centring is a symmetric projection and both right-hand sides are already zero-mean).  One CTA
// per item: pass 1 re-reduces the statistics, thread 0 turns them into the coefficient table,
// pass 2 streams the signals once more (L2-resident) and writes the gradients.
// Algorithmic bytes per item: 4*T*C*2 read twice + 4*T*C*2 written.
// ---------------------------------------------------------------------------------------------
struct PairGrad { double g_ss, g_ee, g_d; };

// gradient of sb_neg_si_snr(st, i, j) with respect to (ss_i, ee_j, d_ij)
__device__ PairGrad sb_neg_si_snr_grad(const PairStats& st, int i, int j, int T) {
  const double eps = 1e-8;
  const double n = (double)T;
  const double ss = st.saa[i] - st.sa[i] * st.sa[i] / n;
  const double ee = st.sbb[j] - st.sb[j] * st.sb[j] / n;
  const double dot = st.sab[i][j] - st.sa[i] * st.sb[j] / n;
  const double energy = ss + eps;
  const double proj2 = dot * dot * ss / (energy * energy);
  double noise2 = ee - 2.0 * dot * dot / energy + proj2;
  if (noise2 < 0.0) noise2 = 0.0;
  const double ratio = proj2 / (noise2 + eps);
  const double dl_dr = -(10.0 / 2.302585092994046) / (ratio + eps);
  const double dr_dp = 1.0 / (noise2 + eps);
  const double dr_dn = -proj2 / ((noise2 + eps) * (noise2 + eps));
  const double dp_dd = 2.0 * dot * ss / (energy * energy);
  const double dp_dss = dot * dot * (energy - 2.0 * ss) / (energy * energy * energy);
  const double dn_dd = -4.0 * dot / energy + dp_dd;
  const double dn_dss = 2.0 * dot * dot / (energy * energy) + dp_dss;
  PairGrad g;
  g.g_d = dl_dr * (dr_dp * dp_dd + dr_dn * dn_dd);
  g.g_ss = dl_dr * (dr_dp * dp_dss + dr_dn * dn_dss);
  g.g_ee = dl_dr * dr_dn;
  return g;
}

// mode 0: cal_si_snr, gout [B,C] (pair i <-> i).  mode 1: PIT, gout [B], perm [B,C] from the forward
// (estimate_source column i is paired with source column perm[i], weight gout/C).
template <int C>
__global__ void __launch_bounds__(kLossThreads) si_snr_bwd_kernel(const float* __restrict__ source,
                                                                  const float* __restrict__ estimate,
                                                                  int T, int mode,
                                                                  const float* __restrict__ gout,
                                                                  const int* __restrict__ perm,
                                                                  float* __restrict__ d_source,
                                                                  float* __restrict__ d_estimate) {
  __shared__ PairStats st;
  __shared__ float As[C], Ae[C], Bm[C][C], mean_s[C], mean_e[C];
  const size_t b = blockIdx.x;
  const float* a = source + b * T * C;
  const float* e = estimate + b * T * C;
  pair_stats<C>(a, e, T, st);
  if (threadIdx.x == 0) {
    double w[C][C];
    for (int i = 0; i < C; ++i)
      for (int j = 0; j < C; ++j) w[i][j] = 0.0;
    if (mode == 0) {
      for (int i = 0; i < C; ++i) w[i][i] = (double)gout[b * C + i];
    } else {
      for (int i = 0; i < C; ++i) w[perm[b * C + i]][i] += (double)gout[b] / (double)C;
    }
    double as[C], ae[C];
    for (int i = 0; i < C; ++i) as[i] = ae[i] = 0.0;
    for (int i = 0; i < C; ++i)
      for (int j = 0; j < C; ++j) {
        float bm = 0.f;
        if (w[i][j] != 0.0) {
          const PairGrad g = sb_neg_si_snr_grad(st, i, j, T);
          as[i] += 2.0 * w[i][j] * g.g_ss;
          ae[j] += 2.0 * w[i][j] * g.g_ee;
          bm = (float)(w[i][j] * g.g_d);
        }
        Bm[i][j] = bm;
      }
    for (int i = 0; i < C; ++i) {
      As[i] = (float)as[i];
      Ae[i] = (float)ae[i];
      mean_s[i] = (float)(st.sa[i] / (double)T);
      mean_e[i] = (float)(st.sb[i] / (double)T);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += kLossThreads) {
    float sv[C], ev[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
      sv[i] = a[(size_t)t * C + i] - mean_s[i];
      ev[i] = e[(size_t)t * C + i] - mean_e[i];
    }
#pragma unroll
    for (int i = 0; i < C; ++i) {
      float gs = As[i] * sv[i], ge = Ae[i] * ev[i];
#pragma unroll
      for (int j = 0; j < C; ++j) {
        gs = fmaf(Bm[i][j], ev[j], gs);
        ge = fmaf(Bm[j][i], sv[j], ge);
      }
      if (d_source != nullptr) d_source[(b * T + t) * C + i] = gs;
      if (d_estimate != nullptr) d_estimate[(b * T + t) * C + i] = ge;
    }
  }
}

// torchmetrics SI-SNR backward: out = 10 log10((ts2+eps)/(noise2+eps)), see tm_si_snr_kernel.
__global__ void __launch_bounds__(kLossThreads) tm_si_snr_bwd_kernel(const float* __restrict__ preds,
                                                                     const float* __restrict__ target,
                                                                     int T, const float* __restrict__ gout,
                                                                     float* __restrict__ d_preds,
                                                                     float* __restrict__ d_target) {
  __shared__ PairStats st;
  __shared__ float coef[6];  // G_pt, 2 G_pp, 2 G_tt, mean_t, mean_p
  const size_t b = blockIdx.x;
  const float* tg = target + b * T;
  const float* pr = preds + b * T;
  pair_stats<1>(tg, pr, T, st);
  if (threadIdx.x == 0) {
    const double eps = 1.1920928955078125e-07;
    const double n = (double)T;
    const double tt = st.saa[0] - st.sa[0] * st.sa[0] / n;
    const double pp = st.sbb[0] - st.sb[0] * st.sb[0] / n;
    const double pt = st.sab[0][0] - st.sa[0] * st.sb[0] / n;
    const double alpha = (pt + eps) / (tt + eps);
    const double ts2 = alpha * alpha * tt;
    double noise2 = ts2 - 2.0 * alpha * pt + pp;
    if (noise2 < 0.0) noise2 = 0.0;
    const double k10 = 10.0 / 2.302585092994046;
    const double g_ts2 = k10 / (ts2 + eps), g_n = -k10 / (noise2 + eps);
    const double g_alpha = g_ts2 * 2.0 * alpha * tt + g_n * (2.0 * alpha * tt - 2.0 * pt);
    const double g_pt = g_alpha / (tt + eps) - 2.0 * alpha * g_n;
    const double g_tt = -g_alpha * alpha / (tt + eps) + (g_ts2 + g_n) * alpha * alpha;
    const double go = (double)gout[b];
    coef[0] = (float)(go * g_pt);
    coef[1] = (float)(go * 2.0 * g_n);
    coef[2] = (float)(go * 2.0 * g_tt);
    coef[3] = (float)(st.sa[0] / n);
    coef[4] = (float)(st.sb[0] / n);
  }
  __syncthreads();
  const float g_pt = coef[0], g_pp2 = coef[1], g_tt2 = coef[2], mt = coef[3], mp = coef[4];
  for (int t = threadIdx.x; t < T; t += kLossThreads) {
    const float tv = tg[t] - mt, pv = pr[t] - mp;
    if (d_preds != nullptr) d_preds[b * T + t] = fmaf(g_pt, tv, g_pp2 * pv);
    if (d_target != nullptr) d_target[b * T + t] = fmaf(g_pt, pv, g_tt2 * tv);
  }
}

int launch_si_snr_bwd(const float* source, const float* estimate, int B, int T, int C, int mode,
                      const float* gout, const int* perm, float* d_source, float* d_estimate,
                      cudaStream_t st) {
  switch (C) {
    case 1: si_snr_bwd_kernel<1><<<B, kLossThreads, 0, st>>>(source, estimate, T, mode, gout, perm, d_source, d_estimate); break;
    case 2: si_snr_bwd_kernel<2><<<B, kLossThreads, 0, st>>>(source, estimate, T, mode, gout, perm, d_source, d_estimate); break;
    case 3: si_snr_bwd_kernel<3><<<B, kLossThreads, 0, st>>>(source, estimate, T, mode, gout, perm, d_source, d_estimate); break;
    case 4: si_snr_bwd_kernel<4><<<B, kLossThreads, 0, st>>>(source, estimate, T, mode, gout, perm, d_source, d_estimate); break;
    default: set_error("si_snr_bwd: C=%d unsupported (1..%d)", C, kMaxC); return 1;
  }
  return check_launch("si_snr_bwd_kernel");
}

int launch_tm_si_snr_bwd(const float* preds, const float* target, int B, int T, const float* gout,
                         float* d_preds, float* d_target, cudaStream_t st) {
  tm_si_snr_bwd_kernel<<<B, kLossThreads, 0, st>>>(preds, target, T, gout, d_preds, d_target);
  return check_launch("tm_si_snr_bwd_kernel");
}

int launch_si_snr(const float* source, const float* estimate, int B, int T, int C, float* out,
                  cudaStream_t st) {
  switch (C) {
    case 1: si_snr_kernel<1><<<B, kLossThreads, 0, st>>>(source, estimate, T, out); break;
    case 2: si_snr_kernel<2><<<B, kLossThreads, 0, st>>>(source, estimate, T, out); break;
    case 3: si_snr_kernel<3><<<B, kLossThreads, 0, st>>>(source, estimate, T, out); break;
    case 4: si_snr_kernel<4><<<B, kLossThreads, 0, st>>>(source, estimate, T, out); break;
    default: set_error("si_snr: C=%d unsupported (1..%d)", C, kMaxC); return 1;
  }
  return check_launch("si_snr_kernel");
}

int launch_pit(const float* source, const float* est, int B, int T, int C, float* loss, int* perm,
               cudaStream_t st) {
  switch (C) {
    case 1: pit_kernel<1><<<B, kLossThreads, 0, st>>>(source, est, T, loss, perm); break;
    case 2: pit_kernel<2><<<B, kLossThreads, 0, st>>>(source, est, T, loss, perm); break;
    case 3: pit_kernel<3><<<B, kLossThreads, 0, st>>>(source, est, T, loss, perm); break;
    case 4: pit_kernel<4><<<B, kLossThreads, 0, st>>>(source, est, T, loss, perm); break;
    default: set_error("pit: C=%d unsupported (1..%d)", C, kMaxC); return 1;
  }
  return check_launch("pit_kernel");
}

int launch_tm_si_snr(const float* preds, const float* target, int B, int T, float* out,
                     cudaStream_t st) {
  tm_si_snr_kernel<<<B, kLossThreads, 0, st>>>(preds, target, T, out);
  return check_launch("tm_si_snr_kernel");
}

// ---------------------------------------------------------------------------------------------
// ContSep selection tail: the immediate consumer of `est` / `context_pred`.
//   training (train_ContSep.py:386-388): per-stream SI-SNR of the estimate against the ground truth ->
//     label = argmax -> CrossEntropy / BCEWithLogits of the selector logits (estimate detached);
//   eval (test.py:234-239, 248-255): pick the stream the selector chose, and the "was it the right one"
//     accuracy bit (SI-SNR against the target >= SI-SNR against every interferer).
// The reference does these with ~25 tiny ATen launches and three host round trips (`.cpu()` at test.py:236);
// here each is one CTA per item over the same one-pass sufficient statistics as the losses above.
// ---------------------------------------------------------------------------------------------

// x [T] against the C columns of y [T,C]: sums {x, x^2, y_j, y_j^2, x y_j} in double.  Valid in thread 0.
struct OneManyStats {
  double sx, sxx, sy[kMaxC], syy[kMaxC], sxy[kMaxC];
};

template <int C>
__device__ void one_vs_many_stats(const float* __restrict__ x, const float* __restrict__ y, int T, OneManyStats& out) {
  constexpr int NV = 2 + 3 * C;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  for (int t = threadIdx.x; t < T; t += kLossThreads) {
    const double xv = (double)x[t];
    acc[0] += xv;
    acc[1] += xv * xv;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const double yv = (double)y[(size_t)t * C + j];
      acc[2 + j] += yv;
      acc[2 + C + j] += yv * yv;
      acc[2 + 2 * C + j] += xv * yv;
    }
  }
  __shared__ double s_red[kLossThreads / 32][NV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double v = 0.0;
      for (int w = 0; w < kLossThreads / 32; ++w) v += s_red[w][i];
      acc[i] = v;
    }
    out.sx = acc[0];
    out.sxx = acc[1];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      out.sy[j] = acc[2 + j];
      out.syy[j] = acc[2 + C + j];
      out.sxy[j] = acc[2 + 2 * C + j];
    }
  }
}

// speechbrain cal_si_snr from centred statistics: ss = |source~|^2, ee = |estimate~|^2, dot = <e~, s~>
__device__ double sb_neg_si_snr_centred(double ss, double ee, double dot) {
  const double eps = 1e-8;
  const double energy = ss + eps;
  const double proj2 = dot * dot * ss / (energy * energy);
  double noise2 = ee - 2.0 * dot * dot / energy + proj2;
  if (noise2 < 0.0) noise2 = 0.0;
  return -10.0 * log10(proj2 / (noise2 + eps) + eps);
}

// torch.argmax semantics: first maximum; a NaN counts as the maximum
__device__ int argmax_first(const float* v, int n) {
  int best = 0;
  for (int i = 1; i < n; ++i) {
    const bool better = (v[i] > v[best]) || (isnan(v[i]) && !isnan(v[best]));
    if (better) best = i;
  }
  return best;
}

template <int C>
__global__ void __launch_bounds__(kLossThreads) selection_loss_kernel(
    const float* __restrict__ gt, const float* __restrict__ est, const float* __restrict__ logits, int T, int ce,
    int B, float* __restrict__ sisnr, long long* __restrict__ label, float* __restrict__ item_loss,
    float* __restrict__ dlogits) {
  __shared__ OneManyStats st;
  const size_t b = blockIdx.x;
  one_vs_many_stats<C>(gt + b * T, est + b * T * C, T, st);
  if (threadIdx.x != 0) return;
  const double n = (double)T;
  const double ss = st.sxx - st.sx * st.sx / n;              // source = ground truth
  float v[C];
  for (int j = 0; j < C; ++j) {
    const double ee = st.syy[j] - st.sy[j] * st.sy[j] / n;
    const double dot = st.sxy[j] - st.sx * st.sy[j] / n;
    v[j] = (float)(-sb_neg_si_snr_centred(ss, ee, dot));     // sisnrs = -1 * cal_si_snr(...)
    sisnr[b * C + j] = v[j];
  }
  const int y = argmax_first(v, C);
  label[b] = y;
  const float inv_b = 1.0f / (float)B;
  if (ce) {                                                  // nn.CrossEntropyLoss (mean over the batch)
    const float* z = logits + b * C;
    float m = z[0];
    for (int j = 1; j < C; ++j) m = fmaxf(m, z[j]);
    float se = 0.f;
    for (int j = 0; j < C; ++j) se += expf(z[j] - m);
    const float lse = m + logf(se);
    item_loss[b] = lse - z[y];
    for (int j = 0; j < C; ++j) dlogits[b * C + j] = (expf(z[j] - lse) - (j == y ? 1.f : 0.f)) * inv_b;
  } else {                                                   // nn.BCEWithLogitsLoss on the single logit, target = label
    const float z = logits[b], t = (float)y;
    item_loss[b] = fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
    dlogits[b] = (1.f / (1.f + expf(-z)) - t) * inv_b;
  }
}

__global__ void mean_in_order_kernel(const float* __restrict__ item, int B, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += item[b];                // fixed order: deterministic
    out[0] = s / (float)B;
  }
}

// pick = argmax softmax(logits) (== argmax logits) or sigmoid(logit) > 0.5 (== logit > 0); out[b,t] = est[b,t,pick[b]]
__global__ void __launch_bounds__(256) select_stream_kernel(const float* __restrict__ est,
                                                            const float* __restrict__ logits, int T, int C, int ce,
                                                            float* __restrict__ out, long long* __restrict__ pick) {
  const size_t b = blockIdx.y;
  int p;
  if (ce) {
    p = argmax_first(logits + b * C, C);
  } else {
    p = logits[b] > 0.f ? 1 : 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) pick[b] = p;
  const float* e = est + b * T * C + p;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < T; t += gridDim.x * 256) out[b * T + t] = e[(size_t)t * C];
}

// enhanced [T] is the ESTIMATE, the columns of sources [T,C] are the sources (column 0 = target, others = interferers)
template <int C>
__global__ void __launch_bounds__(kLossThreads) selection_accuracy_kernel(const float* __restrict__ enhanced,
                                                                          const float* __restrict__ sources, int T,
                                                                          float* __restrict__ sisnr,
                                                                          int* __restrict__ acc) {
  __shared__ OneManyStats st;
  const size_t b = blockIdx.x;
  one_vs_many_stats<C>(enhanced + b * T, sources + b * T * C, T, st);
  if (threadIdx.x != 0) return;
  const double n = (double)T;
  const double ee = st.sxx - st.sx * st.sx / n;
  float v[C];
  for (int j = 0; j < C; ++j) {
    const double ss = st.syy[j] - st.sy[j] * st.sy[j] / n;
    const double dot = st.sxy[j] - st.sx * st.sy[j] / n;
    v[j] = (float)(-sb_neg_si_snr_centred(ss, ee, dot));
    sisnr[b * C + j] = v[j];
  }
  int ok = 1;
  for (int j = 1; j < C; ++j) ok *= (v[0] >= v[j]) ? 1 : 0;   // test.py:251-255
  acc[b] = ok;
}

int launch_selection_loss(const float* gt, const float* est, const float* logits, int B, int T, int C, int ce,
                          float* sisnr, long long* label, float* item_loss, float* loss, float* dlogits,
                          cudaStream_t st) {
  switch (C) {
    case 2: selection_loss_kernel<2><<<B, kLossThreads, 0, st>>>(gt, est, logits, T, ce, B, sisnr, label, item_loss, dlogits); break;
    case 3: selection_loss_kernel<3><<<B, kLossThreads, 0, st>>>(gt, est, logits, T, ce, B, sisnr, label, item_loss, dlogits); break;
    case 4: selection_loss_kernel<4><<<B, kLossThreads, 0, st>>>(gt, est, logits, T, ce, B, sisnr, label, item_loss, dlogits); break;
    default: set_error("selection_loss: %d streams unsupported (2..%d)", C, kMaxC); return 1;
  }
  if (check_launch("selection_loss_kernel")) return 1;
  mean_in_order_kernel<<<1, 32, 0, st>>>(item_loss, B, loss);
  return check_launch("mean_in_order_kernel");
}

int launch_select_stream(const float* est, const float* logits, int B, int T, int C, int ce, float* out,
                         long long* pick, cudaStream_t st) {
  dim3 grid((unsigned)min(ceil_div(T, 256), 148), (unsigned)B);
  select_stream_kernel<<<grid, 256, 0, st>>>(est, logits, T, C, ce, out, pick);
  return check_launch("select_stream_kernel");
}

int launch_selection_accuracy(const float* enhanced, const float* sources, int B, int T, int C, float* sisnr,
                              int* acc, cudaStream_t st) {
  switch (C) {
    case 1: selection_accuracy_kernel<1><<<B, kLossThreads, 0, st>>>(enhanced, sources, T, sisnr, acc); break;
    case 2: selection_accuracy_kernel<2><<<B, kLossThreads, 0, st>>>(enhanced, sources, T, sisnr, acc); break;
    case 3: selection_accuracy_kernel<3><<<B, kLossThreads, 0, st>>>(enhanced, sources, T, sisnr, acc); break;
    case 4: selection_accuracy_kernel<4><<<B, kLossThreads, 0, st>>>(enhanced, sources, T, sisnr, acc); break;
    default: set_error("selection_accuracy: %d sources unsupported (1..%d)", C, kMaxC); return 1;
  }
  return check_launch("selection_accuracy_kernel");
}

}  // namespace cse
