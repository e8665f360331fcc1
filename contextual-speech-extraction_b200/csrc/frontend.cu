// frontend.cu — waveform encoder, masknet.norm, chunking, context prompt (HBM-bound kernels).
//
// Reference: speechbrain Encoder (ContSep.py:10,69), select_norm('ln') = GroupNorm(1,256)
// (ContSep.py:164,226), _padding/_Segmentation (ContSep.py:270-335), context mappers and prompt
// concat (ContSep.py:474-482, 506-513), positional encoding add (CSE_transformer.py:102-104).
#include "common.cuh"

namespace cse {

// --------------------------------------------------------------------------------------------
// Encoder: out[b,l,n] = relu(sum_k w[n,k] * mix[b, 8l+k]).  One thread per channel n keeps its
// 16 taps in registers; the CTA stages a run of samples in shared memory (broadcast reads) and
// writes 256 contiguous channels per frame (coalesced).  Also emits the per-CTA (sum, sumsq)
// partials of masknet.norm so the GroupNorm statistics cost no extra pass over HBM.
// Algorithmic bytes: 4*T + e*N*L per mixture.
// --------------------------------------------------------------------------------------------
constexpr int kEncFrames = 64;  // frames per CTA

template <typename T>
__global__ void __launch_bounds__(kN) encoder_kernel(const float* __restrict__ mix,
                                                     const float* __restrict__ w, int Tlen, int L,
                                                     T* __restrict__ out,
                                                     float* __restrict__ gn_part) {
  __shared__ float s_mix[kEncFrames * kEncS + kEncS];
  __shared__ float s_red[2][kN / 32];
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * kEncFrames;
  const int n = threadIdx.x;
  const int nfr = min(kEncFrames, L - l0);
  const int nsamp = nfr * kEncS + kEncS;
  const float* src = mix + (size_t)b * Tlen + (size_t)l0 * kEncS;
  for (int i = threadIdx.x; i < nsamp; i += kN) s_mix[i] = src[i];
  float wk[kEncK];
#pragma unroll
  for (int k = 0; k < kEncK; ++k) wk[k] = w[n * kEncK + k];
  __syncthreads();
  float sum = 0.f, sq = 0.f;
  T* dst = out + ((size_t)b * L + l0) * kN + n;
  for (int l = 0; l < nfr; ++l) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kEncK; ++k) acc = fmaf(wk[k], s_mix[l * kEncS + k], acc);
    acc = fmaxf(acc, 0.f);
    const T y = from_f<T>(acc);
    dst[(size_t)l * kN] = y;
    const float yr = to_f(y);
    sum += yr;
    sq += yr * yr;
  }
  sum = warp_sum(sum);
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = sum;
    s_red[1][threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int i = 0; i < kN / 32; ++i) {
      a += s_red[0][i];
      c += s_red[1][i];
    }
    float* p = gn_part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
    p[0] = a;
    p[1] = c;
  }
}

int encoder_parts(int L) { return ceil_div(L, kEncFrames); }

int launch_encoder(const float* mix, const float* w, int B, int T, int L, int act, void* out,
                   float* gn_part, int n_parts, cudaStream_t st) {
  dim3 grid(n_parts, B);
  if (act == CSE_BF16)
    encoder_kernel<bf16><<<grid, kN, 0, st>>>(mix, w, T, L, (bf16*)out, gn_part);
  else
    encoder_kernel<float><<<grid, kN, 0, st>>>(mix, w, T, L, (float*)out, gn_part);
  return check_launch("encoder_kernel");
}

// --------------------------------------------------------------------------------------------
// GroupNorm(1, 256) statistics: partial (sum, sumsq) -> (mean, rstd), accumulated in double.
// --------------------------------------------------------------------------------------------
__global__ void gn_finalize_kernel(const float* __restrict__ part, int n_parts, double count,
                                   float eps, float* __restrict__ stat) {
  const int b = blockIdx.x;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < n_parts; i += 32) {
    s += (double)part[((size_t)b * n_parts + i) * 2];
    q += (double)part[((size_t)b * n_parts + i) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (threadIdx.x == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[b * 2] = (float)mean;
    stat[b * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

int launch_gn_finalize(const float* part, int B, int n_parts, double count, float eps, float* stat,
                       cudaStream_t st) {
  gn_finalize_kernel<<<B, 32, 0, st>>>(part, n_parts, count, eps, stat);
  return check_launch("gn_finalize_kernel");
}

// Stand-alone partial statistics of x [B, rows, 256] (used when the encoder did not run in this
// call, i.e. Dual_Path_Model.forward on its own): grid (parts, B), deterministic partials.
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, int rows,
                                                       float* __restrict__ gn_part) {
  __shared__ float s_red[2][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.y;
  float sum = 0.f, sq = 0.f;
  for (int r = blockIdx.x * 8 + wid; r < rows; r += gridDim.x * 8) {
    const f8 v = ld8(x + ((size_t)b * rows + r) * kN + lane * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sum += v.v[i];
      sq += v.v[i] * v.v[i];
    }
  }
  sum = warp_sum(sum);
  sq = warp_sum(sq);
  if (lane == 0) {
    s_red[0][wid] = sum;
    s_red[1][wid] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a += s_red[0][i];
      q += s_red[1][i];
    }
    float* p = gn_part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
    p[0] = a;
    p[1] = q;
  }
}

int launch_gn_stats(const void* x, int B, int rows, int act, float* gn_part, int n_parts,
                    cudaStream_t st) {
  dim3 grid(n_parts, B);
  if (act == CSE_BF16)
    gn_stats_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, rows, gn_part);
  else
    gn_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)x, rows, gn_part);
  return check_launch("gn_stats_kernel");
}

// masknet.norm apply: one warp per frame row.
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ x,
                                                       const float* __restrict__ stat,
                                                       const float* __restrict__ g,
                                                       const float* __restrict__ bta, int L,
                                                       size_t rows, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  const f8 gg = ld8(g + lane * 8), bb = ld8(bta + lane * 8);
  for (size_t r = warp; r < rows; r += nwarps) {
    const int b = (int)(r / L);
    const float mean = stat[b * 2], rstd = stat[b * 2 + 1];
    f8 v = ld8(x + r * kN + lane * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] = (v.v[i] - mean) * rstd * gg.v[i] + bb.v[i];
    st8(out + r * kN + lane * 8, v);
  }
}

int launch_gn_apply(const void* x, const float* stat, const float* g, const float* b, int B, int L,
                    int act, void* out, cudaStream_t st) {
  const size_t rows = (size_t)B * L;
  const int grid = (int)min((size_t)148 * 8, (rows + 7) / 8);
  if (act == CSE_BF16)
    gn_apply_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, stat, g, b, L, rows, (bf16*)out);
  else
    gn_apply_kernel<float><<<grid, 256, 0, st>>>((const float*)x, stat, g, b, L, rows, (float*)out);
  return check_launch("gn_apply_kernel");
}

// --------------------------------------------------------------------------------------------
// Segmentation: X[b,s,k,:] = padded[b, s*P+k, :], padded = [P zeros | x0 | gap+P zeros].
// One warp per destination row (1 KB contiguous), source row is contiguous too.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) segment_kernel(const float* __restrict__ x0, int L, int S,
                                                      size_t rows, float* __restrict__ X) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  for (size_t r = warp; r < rows; r += nwarps) {
    const int k = (int)(r % kK);
    const size_t bs = r / kK;
    const int s = (int)(bs % S);
    const int b = (int)(bs / S);
    const int l = s * kP + k - kP;
    f8 v;
    if (l >= 0 && l < L) {
      v = ld8(x0 + ((size_t)b * L + l) * kN + lane * 8);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = 0.f;
    }
    st8(X + r * kN + lane * 8, v);
  }
}

int launch_segment(const float* x0, int B, int L, int S, float* X, cudaStream_t st) {
  const size_t rows = (size_t)B * S * kK;
  const int grid = (int)min((size_t)148 * 8, (rows + 7) / 8);
  segment_kernel<<<grid, 256, 0, st>>>(x0, L, S, rows, X);
  return check_launch("segment_kernel");
}

// --------------------------------------------------------------------------------------------
// Residual-stream builder: prepend the mapped context tokens and add the sinusoid table.
//   intra: R[(b*S+s), c+k] = X[b,s,k] + pe[c+k]     inter: R[(b*K+k), c+s] = X[b,s,k] + pe[c+s]
// The context token sits at position 0, audio frames at c.. (ContSep.py:482 then
// CSE_transformer.py:104).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_seq_kernel(const float* __restrict__ X,
                                                        const float* __restrict__ ctok,
                                                        const float* __restrict__ pe, int B, int S,
                                                        int c, int inter, float* __restrict__ R,
                                                        const float* __restrict__ ln_g,
                                                        const float* __restrict__ ln_b, bf16* __restrict__ H) {
  // H != NULL: also norm1 of the stack's first layer on every row (it is in the warp's registers)
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  const int n = (inter ? S : kK) + c;
  const int per_b = inter ? kK : S;  // sequences per sample
  const size_t rows = (size_t)B * per_b * n;
  f8 lg, lb;
  if (H != nullptr) {
    lg = ld8(ln_g + lane * 8);
    lb = ld8(ln_b + lane * 8);
  }
  for (size_t r = warp; r < rows; r += nwarps) {
    const int pos = (int)(r % n);
    const size_t seq = r / n;
    const int q = (int)(seq % per_b);
    const int b = (int)(seq / per_b);
    f8 v;
    if (pos < c) {
      v = ld8(ctok + ((size_t)b * c + pos) * kN + lane * 8);
    } else {
      const int s = inter ? (pos - c) : q;
      const int k = inter ? q : (pos - c);
      v = ld8(X + (((size_t)b * S + s) * kK + k) * kN + lane * 8);
    }
    const f8 p = ld8(pe + (size_t)pos * kN + lane * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] += p.v[i];
    st8(R + r * kN + lane * 8, v);
    if (H != nullptr) {
      ln_row(v, lg, lb, 1e-6f);
      st8(H + r * kN + lane * 8, v);
    }
  }
}

int launch_build_sequences(const float* X, const float* ctok, const float* pe, int B, int S, int c,
                           int inter, float* R, cudaStream_t st, const float* ln_g, const float* ln_b, bf16* H) {
  const size_t rows = (size_t)B * (inter ? kK : S) * ((inter ? S : kK) + c);
  const int grid = (int)min((size_t)148 * 8, (rows + 7) / 8);
  build_seq_kernel<<<grid, 256, 0, st>>>(X, ctok, pe, B, S, c, inter, R, ln_g, ln_b, H);
  return check_launch("build_seq_kernel");
}

// --------------------------------------------------------------------------------------------
// Context mapper nn.Linear(4096 -> 256) on a handful of rows: one warp per (row, out) dot
// product, float4 loads.  4 MB of weights per mapper; negligible next to the stacks.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) context_map_kernel(const float* __restrict__ ctx,
                                                          const float* __restrict__ w,
                                                          const float* __restrict__ bias, int rows,
                                                          int in_dim, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (warp >= rows * kN) return;
  const int r = warp / kN, o = warp % kN;
  const float4* a = reinterpret_cast<const float4*>(ctx + (size_t)r * in_dim);
  const float4* ww = reinterpret_cast<const float4*>(w + (size_t)o * in_dim);
  float acc = 0.f;
  for (int i = lane; i < in_dim / 4; i += 32) {
    const float4 x = a[i], y = ww[i];
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
    acc = fmaf(x.z, y.z, acc);
    acc = fmaf(x.w, y.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)r * kN + o] = acc + bias[o];
}

int launch_context_map(const float* ctx, const float* w, const float* b, int rows, int in_dim,
                       float* out, cudaStream_t st) {
  if (in_dim % 4 != 0) {
    set_error("context_map: in_dim %d not a multiple of 4", in_dim);
    return 1;
  }
  const int grid = ceil_div(rows * kN, 8);
  context_map_kernel<<<grid, 256, 0, st>>>(ctx, w, b, rows, in_dim, out);
  return check_launch("context_map_kernel");
}

}  // namespace cse
