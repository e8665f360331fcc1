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
