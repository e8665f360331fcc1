// mixture.cu — mixture synthesis and collation of the training / evaluation loader on the device.
//
// Reference, per dataset item on CPU workers (src/data/dataset_train_CSE.py, mix_aud.py):
//   :237,274   x / max|x| * 0.9                                    peak normalisation of every clip (float32)
//   :417-456   mix_audio(signal, noise, snr, pad)                  2-speaker mixture  (= mix_aud.py:58-96)
//   :458-505   mix_audio_3spk(signal, n1, n2, snr1, snr2, pad)     3-speaker mixture  (= mix_aud.py:3-55)
//   :393-398   librosa.resample(16 kHz -> 8 kHz)                   per output signal
//   :507-601   collate_fn: right-pad every item to the batch maximum
// At >= 4 k audio-seconds per second per GPU the CPU loader (6 workers, README.md:143) is two orders of magnitude too
// slow; here a whole batch of ragged clips (one flat buffer + offsets) is mixed, scaled, and written straight into
// the collated [B, T_out] tensors.  Arithmetic follows numpy's promotion in the reference: energies of float32
// data, then float64 gains / mixture / peak scale, float32 on the final store.
// One CTA per item, three passes over clips that live in L2 (a 16 s clip at 16 kHz is 1 MB): HBM-bound, small.
#include "common.cuh"

namespace cse {

constexpr int kMixThreads = 1024;

__device__ __forceinline__ double mx_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double mx_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum (op = 0) or max (op = 1) of up to three doubles; results valid in every thread
__device__ __forceinline__ void mx_block_reduce3(double& a, double& b, double& c, int op, double* red) {
  a = op ? mx_warp_max(a) : mx_warp_sum(a);
  b = op ? mx_warp_max(b) : mx_warp_sum(b);
  c = op ? mx_warp_max(c) : mx_warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) {
    red[warp] = a;
    red[32 + warp] = b;
    red[64 + warp] = c;
  }
  __syncthreads();
  double x = red[0], y = red[32], z = red[64];
  for (int w = 1; w < nw; ++w) {  // fixed order: deterministic
    x = op ? fmax(x, red[w]) : x + red[w];
    y = op ? fmax(y, red[32 + w]) : y + red[32 + w];
    z = op ? fmax(z, red[64 + w]) : z + red[64 + w];
  }
  a = x;
  b = y;
  c = z;
}

// A clip as the reference sees it after its repeat / cut / zero-pad rules: value at position t of the mixture
struct Clip {
  const float* p;
  long long len;   // samples that exist
  long long elen;  // samples the energy is taken over
  int repeat;      // play in a loop up to the mixture length
  __device__ __forceinline__ double at(long long t) const {
    if (repeat) return (double)p[t % len];
    return t < len ? (double)p[t] : 0.0;
  }
};

// mix_audio (n_noise = 1) and mix_audio_3spk (n_noise = 2).  out[j]: 0 mixed, 1 signal, 2 noise1, 3 noise2, each
// [B, T_out] fp32, right-padded with zeros (collate_fn).  out_len[b] = samples written for item b.
__global__ void __launch_bounds__(kMixThreads) mix_audio_kernel(
    const float* __restrict__ sig, const long long* __restrict__ sig_off, const float* __restrict__ n1,
    const long long* __restrict__ n1_off, const float* __restrict__ n2, const long long* __restrict__ n2_off,
    const double* __restrict__ snr1, const double* __restrict__ snr2, int n_noise, int pad, long long T_out,
    float* __restrict__ o_mix, float* __restrict__ o_sig, float* __restrict__ o_n1, float* __restrict__ o_n2,
    int* __restrict__ out_len) {
  __shared__ double red[96];
  const int b = blockIdx.x;
  Clip S{sig + sig_off[b], sig_off[b + 1] - sig_off[b], 0, 0};
  Clip A{n1 + n1_off[b], n1_off[b + 1] - n1_off[b], 0, 0};
  Clip C{nullptr, 0, 0, 0};
  long long T;  // mixture length
  if (n_noise == 1) {
    // mix_audio: the mixture has the signal's length; a shorter noise is looped (pad = False) or zero-padded
    // (pad = True, energy over its own length); a longer noise is cut
    T = S.len;
    S.elen = S.len;
    A.repeat = (!pad && S.len > A.len) ? 1 : 0;
    if (A.len > S.len) A.len = S.len;
    A.elen = A.repeat ? T : A.len;
  } else {
    // mix_audio_3spk: the mixture has the longest clip's length; shorter clips are looped (pad = False, energy over
    // the looped clip) or zero-padded (pad = True, energy over their own length)
    C = Clip{n2 + n2_off[b], n2_off[b + 1] - n2_off[b], 0, 0};
    T = max(S.len, max(A.len, C.len));
    S.repeat = (!pad && T > S.len) ? 1 : 0;
    A.repeat = (!pad && T > A.len) ? 1 : 0;
    C.repeat = (!pad && T > C.len) ? 1 : 0;
    S.elen = S.repeat ? T : S.len;
    A.elen = A.repeat ? T : A.len;
    C.elen = C.repeat ? T : C.len;
  }
  // ---- pass 1: energies (np.mean(x ** 2) of float32 data; the squares are float32 products) ----
  double es = 0.0, ea = 0.0, ec = 0.0;
  for (long long t = threadIdx.x; t < T; t += kMixThreads) {
    if (t < S.elen) { const float v = (float)S.at(t); es += (double)(v * v); }
    if (t < A.elen) { const float v = (float)A.at(t); ea += (double)(v * v); }
    if (n_noise == 2 && t < C.elen) { const float v = (float)C.at(t); ec += (double)(v * v); }
  }
  mx_block_reduce3(es, ea, ec, 0, red);
  const double e_s = (double)(float)(es / (double)S.elen);   // the reference's means are float32 scalars
  const double e_a = (double)(float)(ea / (double)A.elen);
  const double g1 = sqrt(pow(10.0, -snr1[b] / 10.0) * e_s / e_a);
  double ws, wa, wc = 0.0;
  if (n_noise == 1) {  // a * signal + b * noise keeps the signal's energy for uncorrelated clips
    ws = sqrt(1.0 / (1.0 + g1 * g1));
    wa = sqrt(g1 * g1 / (1.0 + g1 * g1));
  } else {
    const double e_c = (double)(float)(ec / (double)C.elen);
    ws = 1.0;
    wa = g1;
    wc = sqrt(pow(10.0, -snr2[b] / 10.0) * e_s / e_c);
  }
  // ---- pass 2: peak of the mixture ----
  double pk = 0.0, d0 = 0.0, d1 = 0.0;
  for (long long t = threadIdx.x; t < T; t += kMixThreads) {
    double m = ws * S.at(t) + wa * A.at(t);
    if (n_noise == 2) m += wc * C.at(t);
    pk = fmax(pk, fabs(m));
  }
  mx_block_reduce3(pk, d0, d1, 1, red);
  const double scale = 1.0 / pk * 0.9;
  // ---- pass 3: scaled outputs into the collated layout ----
  float* om = o_mix + (size_t)b * T_out;
  float* os = o_sig + (size_t)b * T_out;
  float* oa = o_n1 + (size_t)b * T_out;
  float* oc = n_noise == 2 ? o_n2 + (size_t)b * T_out : nullptr;
  for (long long t = threadIdx.x; t < T_out; t += kMixThreads) {
    if (t < T) {
      const double s = ws * S.at(t), a = wa * A.at(t);
      double m = s + a, c = 0.0;
      if (n_noise == 2) {
        c = wc * C.at(t);
        m += c;
      }
      om[t] = (float)(scale * m);
      os[t] = (float)(scale * s);
      oa[t] = (float)(scale * a);
      if (oc) oc[t] = (float)(scale * c);
    } else {
      om[t] = 0.f;
      os[t] = 0.f;
      oa[t] = 0.f;
      if (oc) oc[t] = 0.f;
    }
  }
  if (out_len && threadIdx.x == 0) out_len[b] = (int)T;
}

// out[b, t] = x_b[t] / max|x_b| * peak in float32 (two roundings, as numpy does it), zero-padded to T_out
__global__ void __launch_bounds__(kMixThreads) peak_normalize_kernel(const float* __restrict__ x,
                                                                    const long long* __restrict__ off, float peak,
                                                                    long long T_out, float* __restrict__ out) {
  __shared__ double red[96];
  const int b = blockIdx.x;
  const float* p = x + off[b];
  const long long len = off[b + 1] - off[b];
  double m = 0.0, d0 = 0.0, d1 = 0.0;
  for (long long t = threadIdx.x; t < len; t += kMixThreads) m = fmax(m, (double)fabsf(p[t]));
  mx_block_reduce3(m, d0, d1, 1, red);
  const float mf = (float)m;
  float* o = out + (size_t)b * T_out;
  for (long long t = threadIdx.x; t < T_out; t += kMixThreads)
    o[t] = t < len ? __fmul_rn(__fdiv_rn(p[t], mf), peak) : 0.f;
}

// Integer decimation by `down` with an FIR low-pass h[0..n_taps) whose centre tap (n_taps / 2) is aligned with the
// kept input samples: y[i] = sum_j h[j] x[i * down + n_taps / 2 - j], zeros outside the clip (scipy.signal.resample_poly
// with up = 1).  Rows of a [B, T_in] batch with per-row valid lengths; output rows [B, T_out], zero beyond
// ceil(len / down).
__global__ void __launch_bounds__(256) decimate_kernel(const float* __restrict__ x, const int* __restrict__ len_in,
                                                       long long T_in, int down, const float* __restrict__ h,
                                                       int n_taps, long long T_out, float* __restrict__ y,
                                                       int* __restrict__ len_out) {
  extern __shared__ float s_h[];
  for (int j = threadIdx.x; j < n_taps; j += blockDim.x) s_h[j] = h[j];
  __syncthreads();
  const int b = blockIdx.y;
  const long long n = len_in ? (long long)len_in[b] : T_in;
  const long long n_out = (n + down - 1) / down;
  const float* p = x + (size_t)b * T_in;
  const int half = n_taps / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < T_out; i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    if (i < n_out) {
      const long long c = i * down + half;
      for (int j = 0; j < n_taps; ++j) {
        const long long t = c - j;
        if (t >= 0 && t < n) acc += (double)s_h[j] * (double)p[t];
      }
    }
    y[(size_t)b * T_out + i] = (float)acc;
  }
  if (len_out && blockIdx.x == 0 && threadIdx.x == 0) len_out[b] = (int)n_out;
}

}  // namespace cse

using namespace cse;

extern "C" {

int cse_mix_audio(const float* signal, const long long* signal_off, const float* noise1, const long long* noise1_off,
                  const float* noise2, const long long* noise2_off, const double* snr1, const double* snr2,
                  int B, int n_noise, int pad, long long T_out, float* mixed, float* signal_out, float* noise1_out,
                  float* noise2_out, int* out_len, void* stream) {
  CSE_REQUIRE(B > 0 && (n_noise == 1 || n_noise == 2), "mix_audio: B=%d n_noise=%d", B, n_noise);
  CSE_REQUIRE(signal && signal_off && noise1 && noise1_off && snr1 && mixed && signal_out && noise1_out,
              "mix_audio: NULL argument");
  CSE_REQUIRE(n_noise == 1 || (noise2 && noise2_off && snr2 && noise2_out), "mix_audio: 3-speaker call needs noise2");
  CSE_REQUIRE(T_out > 0, "mix_audio: T_out=%lld", T_out);
  mix_audio_kernel<<<B, kMixThreads, 0, (cudaStream_t)stream>>>(signal, signal_off, noise1, noise1_off, noise2,
                                                              noise2_off, snr1, snr2, n_noise, pad, T_out, mixed,
                                                              signal_out, noise1_out, noise2_out, out_len);
  return check_launch("mix_audio_kernel");
}

int cse_peak_normalize(const float* x, const long long* off, int B, float peak, long long T_out, float* out,
                       void* stream) {
  CSE_REQUIRE(x && off && out && B > 0 && T_out > 0, "peak_normalize: bad argument");
  peak_normalize_kernel<<<B, kMixThreads, 0, (cudaStream_t)stream>>>(x, off, peak, T_out, out);
  return check_launch("peak_normalize_kernel");
}

int cse_decimate(const float* x, const int* len_in, int B, long long T_in, int down, const float* taps, int n_taps,
                 long long T_out, float* y, int* len_out, void* stream) {
  CSE_REQUIRE(x && taps && y && B > 0 && T_in > 0 && T_out > 0, "decimate: bad argument");
  CSE_REQUIRE(down >= 1 && n_taps >= 1 && n_taps <= 8192, "decimate: down=%d n_taps=%d", down, n_taps);
  const int blocks = (int)((T_out + 255) / 256 < 1024 ? (T_out + 255) / 256 : 1024);
  decimate_kernel<<<dim3(blocks, B), 256, n_taps * sizeof(float), (cudaStream_t)stream>>>(x, len_in, T_in, down, taps,
                                                                                          n_taps, T_out, y, len_out);
  return check_launch("decimate_kernel");
}

}  // extern "C"
