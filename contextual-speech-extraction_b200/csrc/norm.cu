// norm.cu — warp-shuffle LayerNorm and the stack tails (final LN -> GroupNorm -> skip -> relayout).
//
// Reference: sb LayerNorm = nn.LayerNorm(256, eps 1e-6) (CSE_transformer.py:197,358-359,386,408);
// Dual_Computation_Block_CSE tails (ContSep.py:487-502 intra, :516-531 inter);
// pred_head (ContSep.py:516-517).
#include <cstdlib>

#include "common.cuh"

namespace cse {

// One warp per row of 256 fp32 channels (1 KB): 8 channels per lane (ld_row8's split ownership), two shuffle trees.
// Algorithmic bytes per row: 1024 in + e*256 out.
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x,
                                                        const float* __restrict__ g,
                                                        const float* __restrict__ b, size_t M,
                                                        float eps, T* __restrict__ out, int reverse) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  const f8 gg = ld_row8(g, lane), bb = ld_row8(b, lane);  // parameters: not produced by the previous kernel
  pdl_wait();
  for (size_t rr = warp; rr < M; rr += nwarps) {
    // reverse: sweep from the last row down, i.e. start with the rows the producer of x (a GEMM that walks
    // the rows upwards) touched last and that are still in L2
    const size_t r = reverse ? M - 1 - rr : rr;
    f8 v = ld_row8(x + r * kN, lane);
    ln_row(v, gg, bb, eps);
    st_row8(out + r * kN, lane, v);
  }
}

int launch_layernorm(const float* x, const float* g, const float* b, int M, float eps, int act,
                     void* out, cudaStream_t st) {
  const int grid = (int)min((size_t)148 * 8, ((size_t)M + 7) / 8);
  const int reverse = 1;  // descending sweep: see launch_attention (attention.cu)
  KernelScope prof(kClsLayerNorm, st);
  if (act == CSE_BF16)
    launch_pdl(layernorm_kernel<bf16>, dim3(grid), dim3(256), 0, st, 1, x, g, b, (size_t)M, eps, (bf16*)out, reverse);
  else
    launch_pdl(layernorm_kernel<float>, dim3(grid), dim3(256), 0, st, 1, x, g, b, (size_t)M, eps, (float*)out, reverse);
  return check_launch("layernorm_kernel");
}

// --------------------------------------------------------------------------------------------
// Stack tail, pass 1: GroupNorm statistics of y = LN_final(R) over the non-context rows of each
// sample.  grid (kFinishParts, B); each warp strides over the sample's rows; deterministic
// per-CTA partials (no atomics).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t stack_row(int b, int s, int k, int S, int c, int inter) {
  // row of R that holds chunk s / in-chunk position k of sample b
  return inter ? ((size_t)b * kK + k) * (size_t)(S + c) + c + s
               : ((size_t)b * S + s) * (size_t)(kK + c) + c + k;
}

__global__ void __launch_bounds__(256) finish_stats_kernel(const float* __restrict__ R,
                                                           const float* __restrict__ ln_g,
                                                           const float* __restrict__ ln_b, int S,
                                                           int c, int inter,
                                                           float* __restrict__ gn_part) {
  __shared__ float s_red[2][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = gridDim.y - 1 - blockIdx.y;  // CTAs are scheduled in blockIdx order: last sample first
  const f8 gg = ld8(ln_g + lane * 8), bb = ld8(ln_b + lane * 8);
  const int rows = S * kK;
  float sum = 0.f, sq = 0.f;
  // two rows per trip, both loads issued before either row's LayerNorm (a chain of two warp reductions): at one
  // row per trip the 40 resident warps kept 40 KB in flight per SM — 0.58 of the HBM roofline.  The accumulation
  // order is unchanged (row oo, then oo + step).
  const int step = gridDim.x * 8;
  for (int oo = blockIdx.x * 8 + wid; oo < rows; oo += 2 * step) {
    const int o = rows - 1 - oo;  // descending sweep (the last FFN walked the rows upwards; pass 2 walks up again)
    const bool two = oo + step < rows;
    const int o2 = two ? o - step : o;
    f8 v = ld8(R + stack_row(b, o / kK, o % kK, S, c, inter) * kN + lane * 8);
    f8 v2 = ld8(R + stack_row(b, o2 / kK, o2 % kK, S, c, inter) * kN + lane * 8);
    ln_row(v, gg, bb, 1e-6f);
    ln_row(v2, gg, bb, 1e-6f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sum += v.v[i];
      sq += v.v[i] * v.v[i];
    }
    if (two) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sum += v2.v[i];
        sq += v2.v[i] * v2.v[i];
      }
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

// Pass 2: out[b,s,k] = GN(LN(R row)) * gn_g + gn_b + skip[b,s,k]; optionally also seeds the NEXT
// stack's residual stream (other layout) with + pe and its context rows, so the chunk tensor is
// never re-read for the relayout.
__global__ void __launch_bounds__(256) finish_apply_kernel(
    const float* __restrict__ R, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
    const float* __restrict__ gn_g, const float* __restrict__ gn_b, const float* __restrict__ skip,
    const float* __restrict__ stat, int B, int S, int c, int inter, float* __restrict__ out,
    float* __restrict__ next_R, const float* __restrict__ next_pe,
    const float* __restrict__ next_ctok, const float* __restrict__ next_ln_g,
    const float* __restrict__ next_ln_b, bf16* __restrict__ next_H) {
  // next_H != NULL: also norm1 of the next stack's first layer on every row written to next_R (the row is in the
  // warp's registers: the stack's first layernorm_kernel launch and its 140 MB re-read disappear)
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  const f8 lg = ld8(ln_g + lane * 8), lb = ld8(ln_b + lane * 8);
  const f8 gg = ld8(gn_g + lane * 8), gb = ld8(gn_b + lane * 8);
  f8 ng, nb;
  if (next_H != nullptr) {
    ng = ld8(next_ln_g + lane * 8);
    nb = ld8(next_ln_b + lane * 8);
  }
  const size_t rows = (size_t)B * S * kK;
  const int next_inter = !inter;
  // two rows per trip: the four loads (R row and skip row of both) are in flight together — the kernel's 86
  // registers leave 16 warps per SM, and at one row per trip they kept 32 KB in flight (0.6 of the HBM roofline)
  for (size_t o0 = warp; o0 < rows; o0 += 2 * nwarps) {
    const bool two = o0 + nwarps < rows;
    size_t oq[2] = {o0, two ? o0 + nwarps : o0};
    int kq[2], sq_[2], bq[2];
    f8 vq[2], skq[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t o = (uint32_t)oq[u];   // rows < 2^31 (checked by the launcher)
      kq[u] = (int)(o % kK);
      const uint32_t bs = o / kK;
      sq_[u] = (int)(bs % (uint32_t)S);
      bq[u] = (int)(bs / (uint32_t)S);
      vq[u] = ld8(R + stack_row(bq[u], sq_[u], kq[u], S, c, inter) * kN + lane * 8);
      skq[u] = ld8(skip + oq[u] * kN + lane * 8);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const size_t o = oq[u];
      const int k = kq[u], s = sq_[u], b = bq[u];
      f8 v = vq[u];
      ln_row(v, lg, lb, 1e-6f);
      const float mean = stat[b * 2], rstd = stat[b * 2 + 1];
      const f8& sk = skq[u];
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = (v.v[i] - mean) * rstd * gg.v[i] + gb.v[i] + sk.v[i];
      st8(out + o * kN + lane * 8, v);
      if (next_R != nullptr) {
        const int pos = c + (next_inter ? s : k);
        const f8 p = ld8(next_pe + (size_t)pos * kN + lane * 8);
        f8 w;
#pragma unroll
        for (int i = 0; i < 8; ++i) w.v[i] = v.v[i] + p.v[i];
        st8(next_R + stack_row(b, s, k, S, c, next_inter) * kN + lane * 8, w);
        if (next_H != nullptr) {
          ln_row(w, ng, nb, 1e-6f);
          st8(next_H + stack_row(b, s, k, S, c, next_inter) * kN + lane * 8, w);
        }
        // the warp that owns the first audio row of a next-stack sequence also writes its prompt
        const bool first = next_inter ? (s == 0) : (k == 0);
        if (first) {
          const size_t base = stack_row(b, s, k, S, c, next_inter) - (size_t)pos;
          for (int j = 0; j < c; ++j) {
            f8 t = ld8(next_ctok + ((size_t)b * c + j) * kN + lane * 8);
            const f8 pj = ld8(next_pe + (size_t)j * kN + lane * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) t.v[i] += pj.v[i];
            st8(next_R + (base + j) * kN + lane * 8, t);
            if (next_H != nullptr) {
              ln_row(t, ng, nb, 1e-6f);
              st8(next_H + (base + j) * kN + lane * 8, t);
            }
          }
        }
      }
    }
  }
}

int launch_stack_finish(const float* R, const float* ln_g, const float* ln_b, const float* gn_g,
                        const float* gn_b, const float* skip, int B, int S, int c, int inter,
                        float* out, float* next_R, const float* next_pe, const float* next_ctok,
                        float* gn_part, float* stat, cudaStream_t st, const float* next_ln_g,
                        const float* next_ln_b, bf16* next_H) {
  dim3 g1(kFinishParts, B);
  finish_stats_kernel<<<g1, 256, 0, st>>>(R, ln_g, ln_b, S, c, inter, gn_part);
  if (check_launch("finish_stats_kernel")) return 1;
  if (launch_gn_finalize(gn_part, B, kFinishParts, (double)S * kK * kN, 1e-8f, stat, st)) return 1;
  const size_t rows = (size_t)B * S * kK;
  if (rows >= ((size_t)1 << 31)) {
    set_error("stack_finish: %zu rows exceed the 32-bit row index", rows);
    return 1;
  }
  const int grid = (int)min((size_t)148 * 8, (rows + 7) / 8);
  finish_apply_kernel<<<grid, 256, 0, st>>>(R, ln_g, ln_b, gn_g, gn_b, skip, stat, B, S, c, inter,
                                            out, next_R, next_pe, next_ctok, next_ln_g, next_ln_b,
                                            next_R != nullptr ? next_H : nullptr);
  return check_launch("finish_apply_kernel");
}

// --------------------------------------------------------------------------------------------
// pred_head[b,:] = mean over k of LN_final(R_inter[(b,k), 0, :])   (token 0 = first context row)
// One CTA per sample; warp w folds rows k = w, w+8, ... in a fixed order (deterministic).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pred_head_kernel(const float* __restrict__ R,
                                                        const float* __restrict__ ln_g,
                                                        const float* __restrict__ ln_b, int S,
                                                        int c, float* __restrict__ out) {
  __shared__ float s_acc[8][kN];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const f8 gg = ld8(ln_g + lane * 8), bb = ld8(ln_b + lane * 8);
  f8 acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
  for (int k = wid; k < kK; k += 8) {
    f8 v = ld8(R + ((size_t)b * kK + k) * (size_t)(S + c) * kN + lane * 8);
    ln_row(v, gg, bb, 1e-6f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] += v.v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_acc[wid][lane * 8 + i] = acc.v[i];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += s_acc[w][threadIdx.x];
  out[(size_t)b * kN + threadIdx.x] = t * (1.0f / kK);
}

int launch_pred_head(const float* R, const float* ln_g, const float* ln_b, int B, int S, int c,
                     float* out, cudaStream_t st) {
  pred_head_kernel<<<B, 256, 0, st>>>(R, ln_g, ln_b, S, c, out);
  return check_launch("pred_head_kernel");
}

}  // namespace cse
