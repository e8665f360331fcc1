// train_ops.cu — fp32 forward/backward kernels of the non-transformer stages, training path.
//
// Together with backward.cu these are the gradients autograd derives for the reference's
// Sepformer.forward when the train scripts call `loss.backward()` (train_ContExt.py:366-389):
//   GroupNorm(1,256) (+ skip)              ContSep.py:164,226,423-424,498-502,527-531
//   chunk <-> sequence relayout, ctx token  ContSep.py:474-482,487-489,506-513,518-521
//   PReLU + _over_add                       ContSep.py:244,337-370
//   tanh * sigmoid gate                     ContSep.py:255
//   relu(mask) * mix_w -> ConvTranspose1d   ContSep.py:263,79-95
//   Conv1d(k16,s8) + ReLU encoder           ContSep.py:10,69
// All are HBM-bound streaming kernels: one warp per 256-channel row (8 channels per lane) or one
// thread per channel where a [256,16] filter gradient is accumulated in registers.
#include "common.cuh"

namespace cse {

constexpr int kGnParts = 64;

// ---------------------------------------------------------------------------------------------
// GroupNorm forward apply with optional skip: out = (x - mean) * rstd * g + b (+ skip)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_skip_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ stat,
                                                            const float* __restrict__ g,
                                                            const float* __restrict__ bta,
                                                            const float* __restrict__ skip, int rows_per_b,
                                                            size_t rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  const f8 gg = ld8(g + lane * 8), bb = ld8(bta + lane * 8);
  for (size_t r = warp; r < rows; r += nwarps) {
    const int b = (int)(r / rows_per_b);
    const float mean = stat[b * 2], rstd = stat[b * 2 + 1];
    f8 v = ld8(x + r * kN + lane * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] = (v.v[i] - mean) * rstd * gg.v[i] + bb.v[i];
    if (skip != nullptr) {
      const f8 s = ld8(skip + r * kN + lane * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] += s.v[i];
    }
    st8(out + r * kN + lane * 8, v);
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm backward.  Per sample (count = rows*256): dxhat = dy*g,
//   dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat)),  dg += sum dy*xhat,  db += sum dy.
// Pass 1: grid (kGnParts, B) -> deterministic partials of (sum dxhat, sum dxhat*xhat) + dg/db atomics.
// Pass 2: each warp re-sums the kGnParts partials of its sample (L2-resident) and writes dx.
// Algorithmic bytes: x and dy read twice, dx written once = 5 * 4 * rows * 256 per sample.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ stat,
                                                           const float* __restrict__ g,
                                                           const float* __restrict__ dy, int rows,
                                                           float* __restrict__ part,
                                                           float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float s_red[2][8];
  __shared__ float s_g[8][kN], s_b[8][kN];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const float mean = stat[b * 2], rstd = stat[b * 2 + 1];
  const f8 gg = ld8(g + lane * 8);
  float s1 = 0.f, s2 = 0.f, ag[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ag[i] = ab[i] = 0.f;
  for (int r = blockIdx.x * 8 + wid; r < rows; r += gridDim.x * 8) {
    const size_t off = ((size_t)b * rows + r) * kN + lane * 8;
    const f8 xv = ld8(x + off), dv = ld8(dy + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (xv.v[i] - mean) * rstd;
      const float dh = dv.v[i] * gg.v[i];
      s1 += dh;
      s2 += dh * xh;
      ag[i] += dv.v[i] * xh;
      ab[i] += dv.v[i];
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    s_red[0][wid] = s1;
    s_red[1][wid] = s2;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_g[wid][lane * 8 + i] = ag[i];
    s_b[wid][lane * 8 + i] = ab[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a += s_red[0][i];
      q += s_red[1][i];
    }
    float* p = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
    p[0] = a;
    p[1] = q;
  }
  float tg = 0.f, tb = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    tg += s_g[w][threadIdx.x];
    tb += s_b[w][threadIdx.x];
  }
  if (dg != nullptr) atomicAdd(dg + threadIdx.x, tg);
  if (db != nullptr) atomicAdd(db + threadIdx.x, tb);
}

__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ stat,
                                                           const float* __restrict__ g,
                                                           const float* __restrict__ dy,
                                                           const float* __restrict__ part, int n_parts,
                                                           int rows_per_b, size_t rows,
                                                           float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  const f8 gg = ld8(g + lane * 8);
  const float inv_count = 1.0f / ((float)rows_per_b * (float)kN);
  int cur_b = -1;
  float m1 = 0.f, m2 = 0.f, mean = 0.f, rstd = 0.f;
  for (size_t r = warp; r < rows; r += nwarps) {
    const int b = (int)(r / rows_per_b);
    if (b != cur_b) {  // warp-uniform
      float a = 0.f, q = 0.f;
      for (int i = lane; i < n_parts; i += 32) {
        a += part[((size_t)b * n_parts + i) * 2];
        q += part[((size_t)b * n_parts + i) * 2 + 1];
      }
      m1 = warp_sum(a) * inv_count;
      m2 = warp_sum(q) * inv_count;
      mean = stat[b * 2];
      rstd = stat[b * 2 + 1];
      cur_b = b;
    }
    const f8 xv = ld8(x + r * kN + lane * 8), dv = ld8(dy + r * kN + lane * 8);
    f8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (xv.v[i] - mean) * rstd;
      o.v[i] = rstd * (dv.v[i] * gg.v[i] - m1 - xh * m2);
    }
    st8(dx + r * kN + lane * 8, o);
  }
}

// ---------------------------------------------------------------------------------------------
// Sequence rows -> chunk tensor (inverse relayout of build_seq_kernel, context rows dropped):
//   X[b,s,k,:] = R[row(b,s,k), :];  and the column sums of the context rows per sample.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seq_to_chunks_kernel(const float* __restrict__ R, int S, int c,
                                                            int inter, size_t rows, float* __restrict__ X) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  for (size_t r = warp; r < rows; r += nwarps) {
    const int k = (int)(r % kK);
    const size_t bs = r / kK;
    const int s = (int)(bs % S);
    const size_t b = bs / S;
    const size_t src = inter ? (b * kK + k) * (size_t)(S + c) + c + s : (b * S + s) * (size_t)(kK + c) + c + k;
    st8(X + r * kN + lane * 8, ld8(R + src * kN + lane * 8));
  }
}

// grid (c, B), 256 threads = channels: ctok_sum[b,j,:] = sum over the sample's sequences of R[(b,q), j, :]
__global__ void __launch_bounds__(256) ctx_rows_sum_kernel(const float* __restrict__ R, int per_b, int n,
                                                           int c, float* __restrict__ out) {
  const int j = blockIdx.x, b = blockIdx.y;
  float acc = 0.f;
  for (int q = 0; q < per_b; ++q) acc += R[(((size_t)b * per_b + q) * n + j) * kN + threadIdx.x];
  out[((size_t)b * c + j) * kN + threadIdx.x] = acc;
}

// ---------------------------------------------------------------------------------------------
// PReLU + overlap-add backward: dX[b,s,k,:] = prelu'(X[b,s,k,:]) * dU[b, s*P+k-P, :] (0 outside
// [0,L)), dprelu += sum (X < 0) * X * dU.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prelu_ola_bwd_kernel(const float* __restrict__ X,
                                                            const float* __restrict__ prelu,
                                                            const float* __restrict__ dU, int S, int L,
                                                            size_t rows, float* __restrict__ dX,
                                                            float* __restrict__ dprelu) {
  __shared__ float s_red[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t warp = (size_t)blockIdx.x * 8 + wid;
  const size_t nwarps = (size_t)gridDim.x * 8;
  const float a = prelu[0];
  float da = 0.f;
  for (size_t r = warp; r < rows; r += nwarps) {
    const int k = (int)(r % kK);
    const size_t bs = r / kK;
    const int s = (int)(bs % S);
    const size_t b = bs / S;
    const int l = s * kP + k - kP;
    f8 o;
    if (l >= 0 && l < L) {
      const f8 xv = ld8(X + r * kN + lane * 8);
      const f8 du = ld8(dU + (b * L + l) * kN + lane * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool pos = xv.v[i] >= 0.f;
        o.v[i] = pos ? du.v[i] : a * du.v[i];
        da += pos ? 0.f : xv.v[i] * du.v[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
    }
    st8(dX + r * kN + lane * 8, o);
  }
  da = warp_sum(da);
  if (lane == 0) s_red[wid] = da;
  __syncthreads();
  if (threadIdx.x == 0 && dprelu != nullptr) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_red[i];
    atomicAdd(dprelu, t);
  }
}

// ---------------------------------------------------------------------------------------------
// gate backward: y = tanh(o) * sigmoid(g)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ o, const float* __restrict__ g,
                                                       const float* __restrict__ d, size_t n8,
                                                       float* __restrict__ d_o, float* __restrict__ d_g) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
    const f8 a = ld8(o + i * 8), b = ld8(g + i * 8), dd = ld8(d + i * 8);
    f8 ro, rg;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = tanhf(a.v[k]);
      const float s = 1.0f / (1.0f + expf(-b.v[k]));
      ro.v[k] = dd.v[k] * s * (1.0f - t * t);
      rg.v[k] = dd.v[k] * t * s * (1.0f - s);
    }
    st8(d_o + i * 8, ro);
    st8(d_g + i * 8, rg);
  }
}

// ---------------------------------------------------------------------------------------------
// relu(mask_pre) * E -> ConvTranspose1d(256,1,16,stride 8) -> pad/trim, backward.
// One thread per channel n, a CTA walks tiles of kDecTile (b,l) positions:
//   df[(b,l,s)][k] = d_est[b, 8l+k, s]  (0 for 8l+k >= T: trimmed samples carry no gradient)
//   dm = sum_k df[k] w[n,k];  d_mask_pre = (m > 0) ? dm * e : 0;  dE[b,l,n] = sum_s dm * relu(m)
//   d_w[n,k] += sum_rows relu(m) * e * df[k]   (16 register accumulators per thread, atomics at the end)
// ---------------------------------------------------------------------------------------------
constexpr int kDecTile = 16;
constexpr int kMaxMasks = 4;

__global__ void __launch_bounds__(256) mask_decode_bwd_kernel(const float* __restrict__ mask_pre,
                                                              const float* __restrict__ E,
                                                              const float* __restrict__ dec_w,
                                                              const float* __restrict__ d_est, int L, int T,
                                                              int n_masks, size_t BL,
                                                              float* __restrict__ d_mask_pre,
                                                              float* __restrict__ dE,
                                                              float* __restrict__ d_dec_w) {
  __shared__ float s_df[kDecTile * kMaxMasks][kEncK];
  const int n = threadIdx.x;
  float w[kEncK], acc[kEncK];
#pragma unroll
  for (int k = 0; k < kEncK; ++k) {
    w[k] = dec_w[n * kEncK + k];
    acc[k] = 0.f;
  }
  const size_t n_tiles = (BL + kDecTile - 1) / kDecTile;
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t p0 = tile * kDecTile;
    __syncthreads();  // previous tile's readers are done
    for (int idx = threadIdx.x; idx < kDecTile * n_masks * kEncK; idx += 256) {
      const int k = idx % kEncK;
      const int s = (idx / kEncK) % n_masks;
      const int pos = idx / (kEncK * n_masks);
      const size_t bl = p0 + pos;
      float v = 0.f;
      if (bl < BL) {
        const size_t b = bl / L;
        const int l = (int)(bl % L);
        const int t = l * kEncS + k;
        if (t < T) v = d_est[(b * T + t) * n_masks + s];
      }
      s_df[pos * n_masks + s][k] = v;
    }
    __syncthreads();
    for (int pos = 0; pos < kDecTile; ++pos) {
      const size_t bl = p0 + pos;
      if (bl >= BL) break;
      const float e = E[bl * kN + n];
      float de = 0.f;
      for (int s = 0; s < n_masks; ++s) {
        const size_t r = bl * n_masks + s;
        const float m = mask_pre[r * kN + n];
        const float rm = fmaxf(m, 0.f);
        const float masked = rm * e;
        float dm = 0.f;
#pragma unroll
        for (int k = 0; k < kEncK; ++k) {
          const float f = s_df[pos * n_masks + s][k];
          dm = fmaf(f, w[k], dm);
          acc[k] = fmaf(masked, f, acc[k]);
        }
        d_mask_pre[r * kN + n] = m > 0.f ? dm * e : 0.f;
        de = fmaf(dm, rm, de);
      }
      dE[bl * kN + n] = de;
    }
  }
  if (d_dec_w != nullptr) {
#pragma unroll
    for (int k = 0; k < kEncK; ++k) atomicAdd(d_dec_w + n * kEncK + k, acc[k]);
  }
}

// ---------------------------------------------------------------------------------------------
// Encoder backward: E = relu(frames W^T), frames[b,l,k] = mix[b, 8l+k]:
//   d_w[n,k] += sum_{b,l} dE[b,l,n] * (E[b,l,n] > 0) * mix[b, 8l+k]
// ---------------------------------------------------------------------------------------------
constexpr int kEncTile = 32;

__global__ void __launch_bounds__(256) encoder_bwd_kernel(const float* __restrict__ mix,
                                                          const float* __restrict__ E,
                                                          const float* __restrict__ dE, int L, int T,
                                                          size_t BL, float* __restrict__ d_w) {
  __shared__ float s_fr[kEncTile][kEncK];
  const int n = threadIdx.x;
  float acc[kEncK];
#pragma unroll
  for (int k = 0; k < kEncK; ++k) acc[k] = 0.f;
  const size_t n_tiles = (BL + kEncTile - 1) / kEncTile;
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t p0 = tile * kEncTile;
    __syncthreads();
    for (int idx = threadIdx.x; idx < kEncTile * kEncK; idx += 256) {
      const int k = idx % kEncK, pos = idx / kEncK;
      const size_t bl = p0 + pos;
      float v = 0.f;
      if (bl < BL) {
        const size_t b = bl / L;
        const int l = (int)(bl % L);
        v = mix[b * T + (size_t)l * kEncS + k];
      }
      s_fr[pos][k] = v;
    }
    __syncthreads();
    for (int pos = 0; pos < kEncTile; ++pos) {
      const size_t bl = p0 + pos;
      if (bl >= BL) break;
      const float gr = E[bl * kN + n] > 0.f ? dE[bl * kN + n] : 0.f;
#pragma unroll
      for (int k = 0; k < kEncK; ++k) acc[k] = fmaf(gr, s_fr[pos][k], acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < kEncK; ++k) atomicAdd(d_w + n * kEncK + k, acc[k]);
}

static int row_grid(size_t rows) { return (int)min((size_t)148 * 8, (rows + 7) / 8); }

}  // namespace cse

using namespace cse;

extern "C" {

int cse_groupnorm_fwd(const float* x, const float* g, const float* b, const float* skip, int B, int rows,
                      float eps, float* out, float* stat, float* part_scratch, void* stream) {
  CSE_REQUIRE(x && g && b && out && stat && part_scratch && B > 0 && rows > 0, "groupnorm_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (launch_gn_stats(x, B, rows, CSE_FP32, part_scratch, kGnParts, st)) return 1;
  if (launch_gn_finalize(part_scratch, B, kGnParts, (double)rows * kN, eps, stat, st)) return 1;
  const size_t total = (size_t)B * rows;
  gn_apply_skip_kernel<<<row_grid(total), 256, 0, st>>>(x, stat, g, b, skip, rows, total, out);
  return check_launch("gn_apply_skip_kernel");
}

int cse_groupnorm_bwd(const float* x, const float* stat, const float* g, const float* dy, int B, int rows,
                      float* dx, float* dg, float* db, float* part_scratch, void* stream) {
  CSE_REQUIRE(x && stat && g && dy && dx && part_scratch && B > 0 && rows > 0, "groupnorm_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  gn_bwd_stats_kernel<<<dim3(kGnParts, B), 256, 0, st>>>(x, stat, g, dy, rows, part_scratch, dg, db);
  if (check_launch("gn_bwd_stats_kernel")) return 1;
  const size_t total = (size_t)B * rows;
  gn_bwd_apply_kernel<<<row_grid(total), 256, 0, st>>>(x, stat, g, dy, part_scratch, kGnParts, rows, total, dx);
  return check_launch("gn_bwd_apply_kernel");
}

int cse_sequences_to_chunks(const float* R, int B, int S, int c, int inter, float* X, float* ctok_sum,
                            void* stream) {
  CSE_REQUIRE(R && B > 0 && S > 0 && c >= 0 && (X || ctok_sum), "sequences_to_chunks: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (X != nullptr) {
    const size_t rows = (size_t)B * S * kK;
    seq_to_chunks_kernel<<<row_grid(rows), 256, 0, st>>>(R, S, c, inter, rows, X);
    if (check_launch("seq_to_chunks_kernel")) return 1;
  }
  if (ctok_sum != nullptr && c > 0) {
    const int per_b = inter ? kK : S, n = (inter ? S : kK) + c;
    ctx_rows_sum_kernel<<<dim3(c, B), 256, 0, st>>>(R, per_b, n, c, ctok_sum);
    if (check_launch("ctx_rows_sum_kernel")) return 1;
  }
  return 0;
}

int cse_prelu_overlap_add_bwd(const float* X, const float* prelu, const float* dU, int B, int S, int L,
                              float* dX, float* dprelu, void* stream) {
  CSE_REQUIRE(X && prelu && dU && dX && B > 0 && S > 0 && L > 0, "prelu_overlap_add_bwd: bad argument");
  const size_t rows = (size_t)B * S * kK;
  prelu_ola_bwd_kernel<<<row_grid(rows), 256, 0, (cudaStream_t)stream>>>(X, prelu, dU, S, L, rows, dX, dprelu);
  return check_launch("prelu_ola_bwd_kernel");
}

int cse_gate_bwd(const float* o, const float* g, const float* d_out, size_t n, float* d_o, float* d_g,
                 void* stream) {
  CSE_REQUIRE(o && g && d_out && d_o && d_g, "gate_bwd: NULL argument");
  CSE_REQUIRE(n % 8 == 0, "gate_bwd: element count %zu not a multiple of 8", n);
  if (n == 0) return 0;
  const size_t n8 = n / 8;
  const int grid = (int)min((size_t)148 * 8, (n8 + 255) / 256);
  gate_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(o, g, d_out, n8, d_o, d_g);
  return check_launch("gate_bwd_kernel");
}

int cse_mask_decode_bwd(const float* mask_pre, const float* E, const float* dec_w, const float* d_est, int B,
                        int L, int T, int n_masks, float* d_mask_pre, float* dE, float* d_dec_w,
                        void* stream) {
  CSE_REQUIRE(mask_pre && E && dec_w && d_est && d_mask_pre && dE, "mask_decode_bwd: NULL argument");
  CSE_REQUIRE(n_masks >= 1 && n_masks <= kMaxMasks, "mask_decode_bwd: n_masks=%d unsupported (1..%d)", n_masks, kMaxMasks);
  CSE_REQUIRE(B > 0 && L > 0 && T > 0, "mask_decode_bwd: bad shape");
  const size_t BL = (size_t)B * L;
  const int grid = (int)min((size_t)148 * 4, (BL + kDecTile - 1) / kDecTile);
  mask_decode_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask_pre, E, dec_w, d_est, L, T, n_masks, BL,
                                                                 d_mask_pre, dE, d_dec_w);
  return check_launch("mask_decode_bwd_kernel");
}

int cse_encoder_bwd(const float* mix, const float* E, const float* dE, int B, int T, float* d_w, void* stream) {
  CSE_REQUIRE(mix && E && dE && d_w && B > 0, "encoder_bwd: bad argument");
  CSE_REQUIRE(T >= kEncK, "encoder_bwd: mixture of %d samples is shorter than the encoder kernel", T);
  const int L = (T - kEncK) / kEncS + 1;
  const size_t BL = (size_t)B * L;
  const int grid = (int)min((size_t)148 * 4, (BL + kEncTile - 1) / kEncTile);
  encoder_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mix, E, dE, L, T, BL, d_w);
  return check_launch("encoder_bwd_kernel");
}

}  // extern "C"
