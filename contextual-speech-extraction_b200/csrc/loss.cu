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

}  // namespace cse
