// metrics.cu — evaluation metrics of the reference's test loop on the device.
//
// Reference (test.py:198-201,241-245,291-301): torchmetrics.audio.ScaleInvariantSignalNoiseRatio and
// torchmetrics.audio.SignalDistortionRatio objects, `.update(preds, target)` per batch and `.compute()` at the end
// (running mean over every item seen), for the enhanced and for the unprocessed mixture (SI-SNRi / SDRi).
// SI-SNR per item is cse_tm_si_snr (loss.cu).  This file adds
//   * SDR per item — torchmetrics.functional.audio.signal_distortion_ratio(filter_length = 512, zero_mean = False,
//     load_diag = None, use_cg_iter = None): in float64, normalise both signals to unit norm, r = first 512 lags
//     of the target's autocorrelation, b = first 512 lags of the target/preds cross-correlation, solve the
//     symmetric Toeplitz system R(r) sol = b, coh = b.sol, SDR = 10 log10(coh / (1 - coh)).
//     torchmetrics gets r and b from an FFT of length >= 2T - 1 (i.e. LINEAR correlations) and calls a dense LU
//     solver; here the same 2 x 512 lags are direct float64 sums (sdr_corr_kernel, partials reduced in a fixed
//     order) and the system is solved by the Levinson recursion in one CTA per item (sdr_solve_kernel) — same
//     quantities, agreement ~1e-12 dB on CPU prototypes;
//   * the running (sum, count) state of a metric object as two doubles in device memory (metric_update_kernel),
//     so a whole evaluation epoch needs no host synchronisation until compute().
#include "common.cuh"

namespace cse {

constexpr int kSdrMaxLen = 1024;  // filter_length limit of the one-CTA solver (the reference uses 512)
constexpr int kCorrLags = 64;     // lags per CTA
constexpr int kCorrThreads = 256;
constexpr int kCorrTime = 2048;   // samples of the target per CTA

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of two doubles; result valid in every thread; red: 2 * 32 doubles
__device__ __forceinline__ void block_sum2_d(double& a, double& b, double* red) {
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();  // red may still be read by the previous call
  if (lane == 0) {
    red[warp] = a;
    red[32 + warp] = b;
  }
  __syncthreads();
  double x = 0.0, y = 0.0;
  for (int w = 0; w < nw; ++w) {  // fixed order: deterministic
    x += red[w];
    y += red[32 + w];
  }
  a = x;
  b = y;
}

// stats[b] = {sum t, sum p, sum t^2, sum p^2} (the sums of squares are of the mean-removed signals when zero_mean)
__global__ void __launch_bounds__(1024) sdr_stats_kernel(const float* __restrict__ preds,
                                                         const float* __restrict__ target, int T, int zero_mean,
                                                         double* __restrict__ stats) {
  __shared__ double red[64];
  const float* p = preds + (size_t)blockIdx.x * T;
  const float* t = target + (size_t)blockIdx.x * T;
  double mt = 0.0, mp = 0.0;
  if (zero_mean) {
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      mt += (double)t[i];
      mp += (double)p[i];
    }
    block_sum2_d(mt, mp, red);
    mt /= T;
    mp /= T;
  }
  double st = 0.0, sp = 0.0;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const double a = (double)t[i] - mt, c = (double)p[i] - mp;
    st += a * a;
    sp += c * c;
  }
  block_sum2_d(st, sp, red);
  if (threadIdx.x == 0) {
    double* s = stats + 4 * blockIdx.x;
    s[0] = mt;
    s[1] = mp;
    s[2] = st;
    s[3] = sp;
  }
}

// part[b][split][0][k] = sum_{t in split} t[t] t[t+k],  part[b][split][1][k] = sum t[t] p[t+k]   (raw, un-normalised)
__global__ void __launch_bounds__(kCorrThreads) sdr_corr_kernel(const float* __restrict__ preds,
                                                                const float* __restrict__ target, int T, int L,
                                                                int n_split, const double* __restrict__ stats,
                                                                double* __restrict__ part) {
  __shared__ double s_t[kCorrTime + kCorrLags];
  __shared__ double s_p[kCorrTime + kCorrLags];
  __shared__ double red[2][kCorrThreads];
  const int b = blockIdx.z, split = blockIdx.y, k0 = blockIdx.x * kCorrLags;
  const float* p = preds + (size_t)b * T;
  const float* t = target + (size_t)b * T;
  const double mt = stats[4 * b], mp = stats[4 * b + 1];
  const int t0 = split * kCorrTime;
  for (int i = threadIdx.x; i < kCorrTime + kCorrLags; i += kCorrThreads) {
    const int g = t0 + k0 + i;  // the shifted operands start k0 samples later
    s_t[i] = g < T ? (double)t[g] - mt : 0.0;
    s_p[i] = g < T ? (double)p[g] - mp : 0.0;
  }
  __syncthreads();
  // thread = (lag, time slice): the un-shifted target sample is a broadcast across the 64 lag threads of a slice
  const int lag = threadIdx.x & (kCorrLags - 1), slice = threadIdx.x / kCorrLags;
  constexpr int kSlices = kCorrThreads / kCorrLags;
  double r = 0.0, c = 0.0;
  const int i1 = min(kCorrTime, T - t0);
  for (int i = slice; i < i1; i += kSlices) {
    const int g = t0 + i;
    const double a = (double)t[g] - mt;  // L1-resident, same address across the lag threads
    r += a * s_t[i + lag];
    c += a * s_p[i + lag];
  }
  red[0][threadIdx.x] = r;
  red[1][threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x < kCorrLags && k0 + threadIdx.x < L) {
    double rr = 0.0, cc = 0.0;
#pragma unroll
    for (int s = 0; s < kSlices; ++s) {
      rr += red[0][s * kCorrLags + threadIdx.x];
      cc += red[1][s * kCorrLags + threadIdx.x];
    }
    double* dst = part + ((size_t)(b * n_split + split) * 2) * L;
    dst[k0 + threadIdx.x] = rr;
    dst[L + k0 + threadIdx.x] = cc;
  }
}

// One CTA per item: reduce the partials, normalise, Levinson recursion for the symmetric Toeplitz system, SDR in dB.
__global__ void __launch_bounds__(kSdrMaxLen) sdr_solve_kernel(const double* __restrict__ part,
                                                               const double* __restrict__ stats, int L, int n_split,
                                                               int has_load_diag, double load_diag,
                                                               float* __restrict__ out) {
  __shared__ double r[kSdrMaxLen], bb[kSdrMaxLen], f[kSdrMaxLen], x[kSdrMaxLen];
  __shared__ double red[64];
  const int b = blockIdx.x, i = threadIdx.x;
  // torchmetrics: x / clamp(norm(x), min = 1e-6) before the correlations
  const double nt = fmax(sqrt(stats[4 * b + 2]), 1e-6), np_ = fmax(sqrt(stats[4 * b + 3]), 1e-6);
  if (i < L) {
    double rr = 0.0, cc = 0.0;
    for (int s = 0; s < n_split; ++s) {  // fixed order
      const double* src = part + ((size_t)(b * n_split + s) * 2) * L;
      rr += src[i];
      cc += src[L + i];
    }
    r[i] = rr / (nt * nt);
    bb[i] = cc / (nt * np_);
    f[i] = 0.0;
    x[i] = 0.0;
  }
  __syncthreads();
  if (i == 0) {
    if (has_load_diag) r[0] += load_diag;
    f[0] = 1.0 / r[0];
    x[0] = bb[0] / r[0];
  }
  __syncthreads();
  for (int m = 1; m < L; ++m) {
    // ef = sum_{j<m} r[m-j] f[j],  ex = sum_{j<m} r[m-j] x[j]
    double ef = (i < m) ? r[m - i] * f[i] : 0.0;
    double ex = (i < m) ? r[m - i] * x[i] : 0.0;
    block_sum2_d(ef, ex, red);
    const double denom = 1.0 - ef * ef;
    // forward vector of order m+1: fn[j] = (f[j] - ef * f[m-j]) / denom with f[m] = 0 and f[-0 reversed] handled below
    double fnew = 0.0;
    if (i <= m) {
      const double fj = (i < m) ? f[i] : 0.0;
      const double bj = (i >= 1) ? f[m - i] : 0.0;
      fnew = (fj - ef * bj) / denom;
    }
    __syncthreads();
    if (i <= m) f[i] = fnew;
    __syncthreads();
    if (i <= m) x[i] += (bb[m] - ex) * f[m - i];
    __syncthreads();
  }
  double coh = (i < L) ? bb[i] * x[i] : 0.0, dummy = 0.0;
  block_sum2_d(coh, dummy, red);
  if (i == 0) out[b] = (float)(10.0 * log10(coh / (1.0 - coh)));
}

// acc[0] += sum(values), acc[1] += n   (torchmetrics: sum_<metric> / total state of a metric object)
__global__ void metric_update_kernel(const float* __restrict__ values, int n, double* __restrict__ acc) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += (double)values[i];
  s = warp_sum_d(s);
  if (threadIdx.x == 0) {
    acc[0] += s;
    acc[1] += (double)n;
  }
}

}  // namespace cse

using namespace cse;

extern "C" {

size_t cse_sdr_workspace_bytes(int B, int T, int filter_length) {
  if (B <= 0 || T <= 0 || filter_length <= 0) return 0;
  const size_t n_split = ((size_t)T + kCorrTime - 1) / kCorrTime;
  return ((size_t)B * 4 + (size_t)B * n_split * 2 * filter_length) * sizeof(double);
}

int cse_sdr(const float* preds, const float* target, int B, int T, int filter_length, int zero_mean,
            int has_load_diag, double load_diag, float* out, void* workspace, size_t workspace_bytes,
            void* stream) {
  CSE_REQUIRE(preds && target && out && workspace, "sdr: NULL argument");
  CSE_REQUIRE(B > 0 && T > 0, "sdr: bad shape B=%d T=%d", B, T);
  CSE_REQUIRE(filter_length >= 1 && filter_length <= kSdrMaxLen, "sdr: filter_length %d outside [1,%d]",
              filter_length, kSdrMaxLen);
  CSE_REQUIRE(workspace_bytes >= cse_sdr_workspace_bytes(B, T, filter_length), "sdr: workspace too small");
  CSE_REQUIRE(((uintptr_t)workspace & 7) == 0, "sdr: workspace must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  double* stats = (double*)workspace;
  double* part = stats + (size_t)B * 4;
  const int n_split = (T + kCorrTime - 1) / kCorrTime;
  sdr_stats_kernel<<<B, 1024, 0, st>>>(preds, target, T, zero_mean, stats);
  if (check_launch("sdr_stats_kernel")) return 1;
  dim3 grid((filter_length + kCorrLags - 1) / kCorrLags, n_split, B);
  sdr_corr_kernel<<<grid, kCorrThreads, 0, st>>>(preds, target, T, filter_length, n_split, stats, part);
  if (check_launch("sdr_corr_kernel")) return 1;
  const int threads = ((filter_length + 31) / 32) * 32;
  sdr_solve_kernel<<<B, threads, 0, st>>>(part, stats, filter_length, n_split, has_load_diag, load_diag, out);
  return check_launch("sdr_solve_kernel");
}

int cse_metric_update(const float* values, int n, double* acc, void* stream) {
  CSE_REQUIRE(values && acc && n > 0, "metric_update: bad argument");
  metric_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(values, n, acc);
  return check_launch("metric_update_kernel");
}

}  // extern "C"
