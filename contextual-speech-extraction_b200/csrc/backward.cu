// backward.cu — fp32 backward kernels of one pre-norm transformer layer (training path, cfg 3).
//
// Reference: the gradients autograd derives for TransformerEncoderLayer.forward
// (CSE_transformer.py:385-416): LayerNorm (nn.LayerNorm(256, eps 1e-6), :358-359), the packed
// in_proj / out_proj / FFN nn.Linear layers (:335-340, :468-477) and softmax(q k^T / sqrt(32)) v
// (:535-557 -> F.scaled_dot_product_attention).  True-fp32 arithmetic like gemm_simt.cu: this is
// the parity-mode backward (<= 1e-4 rel-L2 against autograd over the fp32 reference); a tcgen05
// dgrad / wgrad pair is the performance-mode follow-up.
//
// Kernels
//   transpose_kernel       W[N,K] -> W^T[K,N], so a dgrad dA = dC W is the forward GEMM with W^T
//   wgrad_kernel           dW[N1,N2] += X[M,N1]^T Y[M,N2]  (split over M, fp32 atomics)
//   colsum_kernel          db[N] += sum_m dC[m,N]
//   relu_bwd_kernel        d *= (F > 0)
//   layernorm_bwd_kernel   warp per row, shuffle reductions; dgamma / dbeta per CTA then atomics
//   attention_bwd_kernel   one CTA per (sequence, head): softmax statistics and dq by query row,
//                          dk / dv by key row (no atomics), q/k/v/dO of the head in shared memory
#include "common.cuh"

namespace cse {

// ---------------------------------------------------------------------------------------------
// transpose: in [rows, cols] -> out [cols, rows]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int rows, int cols,
                                                        float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

int launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, rows, cols, out);
  return check_launch("transpose_kernel");
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[i, j] += sum_m X[m, i] * Y[m, j].  For nn.Linear y = a W^T: X = dY [M,N], Y = a [M,K].
// Both operands are read as stored ([m][col] rows): the 16-row slab of each lands in shared memory
// already in the [k][m]-style layout the FFMA micro-kernel wants, so no transposition is needed.
// 128x128 output tile, 256 threads, 8x8 register micro-tile (as gemm_simt.cu); grid.z splits M.
// ---------------------------------------------------------------------------------------------
constexpr int kWT = 128, kWK = 16, kWPad = 4;

__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ X, int ldx,
                                                    const float* __restrict__ Y, int ldy, int M,
                                                    int m_per_split, float* __restrict__ dW, int ldw) {
  __shared__ __align__(16) float Xs[2][kWK][kWT + kWPad];
  __shared__ __align__(16) float Ys[2][kWK][kWT + kWPad];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * kWT, j0 = blockIdx.x * kWT;
  const int m_begin = blockIdx.z * m_per_split;
  const int m_end = min(M, m_begin + m_per_split);
  if (m_begin >= m_end) return;
  const int ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 5;          // 0..7 (+8)
  const int lc = (tid & 31) * 4;    // 0..124
  float4 rx[2], ry[2];

  auto load_tile = [&](int m0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + lr + h * 8;
      if (m < m_end) {
        rx[h] = *reinterpret_cast<const float4*>(X + (size_t)m * ldx + i0 + lc);
        ry[h] = *reinterpret_cast<const float4*>(Y + (size_t)m * ldy + j0 + lc);
      } else {
        rx[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        ry[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(&Xs[buf][lr + h * 8][lc]) = rx[h];
      *reinterpret_cast<float4*>(&Ys[buf][lr + h * 8][lc]) = ry[h];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tile(m_begin);
  store_tile(0);
  __syncthreads();
  const int nk = (m_end - m_begin + kWK - 1) / kWK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile(m_begin + (kt + 1) * kWK);
#pragma unroll
    for (int k = 0; k < kWK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Xs[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Xs[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ys[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ys[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int ih = 0; ih < 2; ++ih)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i0 + ih * 64 + ty * 4 + i;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          atomicAdd(dW + (size_t)row * ldw + j0 + jh * 64 + tx * 4 + j, acc[ih * 4 + i][jh * 4 + j]);
    }
}

int launch_wgrad(const float* X, int ldx, const float* Y, int ldy, int M, int N1, int N2, float* dW,
                 cudaStream_t st) {
  if (N1 % kWT != 0 || N2 % kWT != 0 || ldx % 4 != 0 || ldy % 4 != 0) {
    set_error("wgrad: need N1, N2 %% %d == 0 and ldx, ldy %% 4 == 0 (N1=%d N2=%d ldx=%d ldy=%d)", kWT, N1,
              N2, ldx, ldy);
    return 1;
  }
  if (M <= 0) return 0;
  const int tiles = (N1 / kWT) * (N2 / kWT);
  int splits = ceil_div(148 * 4, tiles);                 // ~4 CTAs per SM in flight over the whole grid
  const int max_splits = ceil_div(M, 4 * kWK);           // at least 64 rows per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int m_per_split = ceil_div(ceil_div(M, splits), kWK) * kWK;
  splits = ceil_div(M, m_per_split);
  dim3 grid(N2 / kWT, N1 / kWT, splits);
  KernelScope prof(kClsGemmSimt, st);
  wgrad_kernel<<<grid, 256, 0, st>>>(X, ldx, Y, ldy, M, m_per_split, dW, N2);
  return check_launch("wgrad_kernel");
}

// ---------------------------------------------------------------------------------------------
// bias gradient: db[n] += sum_m dC[m, n]   (N % 256 == 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dC, int ld, int M,
                                                     int rows_per_cta, float* __restrict__ db) {
  const int col = blockIdx.y * 256 + threadIdx.x;
  const int m0 = blockIdx.x * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int m = m0;
  for (; m + 3 < m1; m += 4) {
    s0 += dC[(size_t)m * ld + col];
    s1 += dC[(size_t)(m + 1) * ld + col];
    s2 += dC[(size_t)(m + 2) * ld + col];
    s3 += dC[(size_t)(m + 3) * ld + col];
  }
  for (; m < m1; ++m) s0 += dC[(size_t)m * ld + col];
  if (m0 < m1) atomicAdd(db + col, (s0 + s1) + (s2 + s3));
}

// the same pass also writing the bf16 copy of dC the tensor-core dgrad / wgrad consume (one read of dC instead of
// two), optionally through a ReLU mask: relu != NULL -> dC[m,n] counts only where relu[m,n] > 0 (the saved bf16
// activation of Linear -> ReLU: torch's threshold_backward folded into this pass)
// Thread t: four consecutive columns ((t & 63) * 4 of the CTA's 256-column slab), rows m0 + (t >> 6), + 4, ...: 16-byte
// loads, 8-byte bf16 stores, four rows per thread in flight (the first version walked one column per thread with
// 4-byte loads and ran at a fifth of the HBM rate: 3.2 ms of a 16.5 ms training step).
template <bool RELU>
__global__ void __launch_bounds__(256) colsum_cast_kernel(const float* __restrict__ dC, const bf16* __restrict__ relu,
                                                          int ld, int M, int rows_per_cta, float* __restrict__ db,
                                                          bf16* __restrict__ dC16) {
  __shared__ float4 s_part[4][64];
  const int c4 = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int col = blockIdx.y * 256 + c4 * 4;
  const int m0 = blockIdx.x * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load = [&](int m) { return __ldg(reinterpret_cast<const float4*>(dC + (size_t)m * ld + col)); };
  auto mask = [&](int m, float4& v) {
    if (RELU) {
      const uint2 r = __ldg(reinterpret_cast<const uint2*>(relu + (size_t)m * ld + col));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
      const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
      v.x = a.x > 0.f ? v.x : 0.f;
      v.y = a.y > 0.f ? v.y : 0.f;
      v.z = b.x > 0.f ? v.z : 0.f;
      v.w = b.y > 0.f ? v.w : 0.f;
    }
  };
  auto put = [&](int m, const float4& v) {
    __nv_bfloat162 o[2] = {__floats2bfloat162_rn(v.x, v.y), __floats2bfloat162_rn(v.z, v.w)};
    *reinterpret_cast<uint2*>(dC16 + (size_t)m * ld + col) = *reinterpret_cast<const uint2*>(o);
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  };
  int m = m0 + rl;
  for (; m + 12 < m1; m += 16) {
    float4 v0 = load(m), v1 = load(m + 4), v2 = load(m + 8), v3 = load(m + 12);
    mask(m, v0);
    mask(m + 4, v1);
    mask(m + 8, v2);
    mask(m + 12, v3);
    put(m, v0);
    put(m + 4, v1);
    put(m + 8, v2);
    put(m + 12, v3);
  }
  for (; m < m1; m += 4) {
    float4 v = load(m);
    mask(m, v);
    put(m, v);
  }
  s_part[rl][c4] = acc;
  __syncthreads();
  if (threadIdx.x < 64 && m0 < m1) {
    const float4 a = s_part[0][c4], b = s_part[1][c4], c = s_part[2][c4], d = s_part[3][c4];
    // one 16-byte reduction per four columns: the CTAs' sums of a slab all land on the same eight 128-byte lines, and
    // with a scalar atomic per column (136 k per launch at N = 256) those lines were the kernel's bottleneck
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(db + blockIdx.y * 256 + c4 * 4),
                 "f"((a.x + b.x) + (c.x + d.x)), "f"((a.y + b.y) + (c.y + d.y)), "f"((a.z + b.z) + (c.z + d.z)),
                 "f"((a.w + b.w) + (c.w + d.w))
                 : "memory");
  }
}

int launch_colsum_cast(const float* dC, const bf16* relu, int M, int N, float* db, bf16* dC16, cudaStream_t st) {
  if (N % 256 != 0) {
    set_error("colsum_cast: need N %% 256 == 0 (N=%d)", N);
    return 1;
  }
  if (((uintptr_t)dC | (uintptr_t)dC16 | (uintptr_t)relu | (uintptr_t)db) & 15) {
    set_error("colsum_cast: operands must be 16-byte aligned");
    return 1;
  }
  if (M <= 0) return 0;
  int ctas = ceil_div(148 * 4, N / 256);
  int rows_per_cta = ceil_div(M, ctas);
  if (rows_per_cta < 64) rows_per_cta = 64;
  ctas = ceil_div(M, rows_per_cta);
  if (relu != nullptr)
    colsum_cast_kernel<true><<<dim3(ctas, N / 256), 256, 0, st>>>(dC, relu, N, M, rows_per_cta, db, dC16);
  else
    colsum_cast_kernel<false><<<dim3(ctas, N / 256), 256, 0, st>>>(dC, nullptr, N, M, rows_per_cta, db, dC16);
  return check_launch("colsum_cast_kernel");
}

int launch_colsum(const float* dC, int ld, int M, int N, float* db, cudaStream_t st) {
  if (N % 256 != 0) {
    set_error("colsum: need N %% 256 == 0 (N=%d)", N);
    return 1;
  }
  if (M <= 0) return 0;
  int ctas = ceil_div(148 * 8, N / 256);
  int rows_per_cta = ceil_div(M, ctas);
  if (rows_per_cta < 32) rows_per_cta = 32;
  ctas = ceil_div(M, rows_per_cta);
  colsum_kernel<<<dim3(ctas, N / 256), 256, 0, st>>>(dC, ld, M, rows_per_cta, db);
  return check_launch("colsum_kernel");
}

// ---------------------------------------------------------------------------------------------
// ReLU backward: d[i] = F[i] > 0 ? d[i] : 0   (torch threshold_backward on the saved output)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float4* __restrict__ F, float4* __restrict__ d,
                                                       size_t n4) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4 f = F[i];
    float4 g = d[i];
    g.x = f.x > 0.f ? g.x : 0.f;
    g.y = f.y > 0.f ? g.y : 0.f;
    g.z = f.z > 0.f ? g.z : 0.f;
    g.w = f.w > 0.f ? g.w : 0.f;
    d[i] = g;
  }
}

int launch_relu_bwd(const float* F, float* d, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  if (n % 4 != 0) {
    set_error("relu_bwd: element count %zu is not a multiple of 4", n);
    return 1;
  }
  const int grid = (int)min((size_t)148 * 8, (n / 4 + 255) / 256);
  relu_bwd_kernel<<<grid, 256, 0, st>>>((const float4*)F, (float4*)d, n / 4);
  return check_launch("relu_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  y = xhat * g + b, xhat = (x - mean) * rstd:
//   dxhat = dy * g;  dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
//   dg += sum_rows dy * xhat;  db += sum_rows dy
// One warp per row (8 channels per lane, as the forward kernel); statistics are recomputed from x.
// Algorithmic bytes per row: 1024 (x) + 1024 (dy) + 1024 (dx) [+ 1024 when accumulating into dx].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ g,
                                                            const float* __restrict__ dy, size_t M,
                                                            float eps, float* dx, int accumulate,
                                                            float* __restrict__ dg,
                                                            float* __restrict__ db) {
  __shared__ __align__(16) float s_g[8][kN], s_b[8][kN];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t warp = (size_t)blockIdx.x * 8 + wid;
  const size_t nwarps = (size_t)gridDim.x * 8;
  const f8 gg = ld8(g + lane * 8);
  float ag[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ag[i] = ab[i] = 0.f;
  for (size_t r = warp; r < M; r += nwarps) {
    f8 xv = ld8(x + r * kN + lane * 8);
    const f8 dv = ld8(dy + r * kN + lane * 8);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += xv.v[i];
    const float mean = warp_sum(s) * (1.0f / kN);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xv.v[i] -= mean;
      q += xv.v[i] * xv.v[i];
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / kN) + eps);
    float s1 = 0.f, s2 = 0.f;
    f8 dh;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xv.v[i] *= rstd;                       // xhat
      dh.v[i] = dv.v[i] * gg.v[i];           // dxhat
      s1 += dh.v[i];
      s2 += dh.v[i] * xv.v[i];
      ag[i] += dv.v[i] * xv.v[i];
      ab[i] += dv.v[i];
    }
    s1 = warp_sum(s1) * (1.0f / kN);
    s2 = warp_sum(s2) * (1.0f / kN);
    f8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = rstd * (dh.v[i] - s1 - xv.v[i] * s2);
    if (accumulate) {
      const f8 prev = ld8(dx + r * kN + lane * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] += prev.v[i];
    }
    st8(dx + r * kN + lane * 8, o);
  }
  if (dg == nullptr && db == nullptr) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_g[wid][lane * 8 + i] = ag[i];
    s_b[wid][lane * 8 + i] = ab[i];
  }
  __syncthreads();
  // threads 0-63: four columns of dg each, threads 64-127: of db; one 16-byte reduction per thread (the CTAs' sums all
  // land on the same sixteen 128-byte lines: 512 scalar atomics per CTA were a third of this kernel's time)
  if (threadIdx.x < 128) {
    const int c4 = threadIdx.x & 63;
    const bool is_g = threadIdx.x < 64;
    float* dst = is_g ? dg : db;
    if (dst != nullptr) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float4 v = *reinterpret_cast<const float4*>(is_g ? &s_g[w][c4 * 4] : &s_b[w][c4 * 4]);
        t.x += v.x;
        t.y += v.y;
        t.z += v.z;
        t.w += v.w;
      }
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst + c4 * 4), "f"(t.x), "f"(t.y), "f"(t.z),
                   "f"(t.w)
                   : "memory");
    }
  }
}

int launch_layernorm_bwd(const float* x, const float* g, const float* dy, int M, float eps, float* dx,
                         int accumulate, float* dg, float* db, cudaStream_t st) {
  if (M <= 0) return 0;
  if (((uintptr_t)dg | (uintptr_t)db) & 15) {
    set_error("layernorm_bwd: dg / db must be 16-byte aligned");
    return 1;
  }
  const int grid = (int)min((size_t)148 * 4, ((size_t)M + 7) / 8);
  KernelScope prof(kClsLayerNorm, st);
  layernorm_bwd_kernel<<<grid, 256, 0, st>>>(x, g, dy, (size_t)M, eps, dx, accumulate, dg, db);
  return check_launch("layernorm_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------
// Attention backward (no mask, no dropout), one CTA per (sequence, head).
//   P = softmax(scale * Q K^T), O = P V;  D_i = dO_i . O_i
//   dV_j = sum_i P_ij dO_i;  dS_ij = P_ij (dO_i . V_j - D_i);  dQ_i = scale * sum_j dS_ij K_j;
//   dK_j = scale * sum_i dS_ij Q_i
// Phase A (thread = query row): row max / sum (log-sum-exp) and D_i, then dQ_i.
// Phase B (thread = key row):   dK_j, dV_j accumulated in registers over all query rows.
// The probabilities are recomputed in both phases from the shared-memory copies of q/k/v, so nothing
// but qkv and the forward output is needed from the forward pass.  512*n + 8*n bytes of shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kAttnBwdThreads = 256;

__global__ void __launch_bounds__(kAttnBwdThreads) attention_bwd_kernel(const float* __restrict__ qkv,
                                                                        const float* __restrict__ out,
                                                                        const float* __restrict__ d_out,
                                                                        int n, float* __restrict__ d_qkv) {
  extern __shared__ __align__(16) float smem_b[];
  float* Qs = smem_b;                       // [n][32]
  float* Ks = Qs + (size_t)n * kDh;
  float* Vs = Ks + (size_t)n * kDh;
  float* Gs = Vs + (size_t)n * kDh;         // dO
  float* lse = Gs + (size_t)n * kDh;        // [n]
  float* Dd = lse + n;                      // [n]
  const int h = blockIdx.x % kHeads;
  const size_t seq = blockIdx.x / kHeads;
  const float* base = qkv + seq * n * (3 * kN) + h * kDh;
  const float* obase = out + seq * n * kN + h * kDh;
  const float* gbase = d_out + seq * n * kN + h * kDh;
  float* dbase = d_qkv + seq * n * (3 * kN) + h * kDh;
  for (int i = threadIdx.x; i < n * (kDh / 4); i += blockDim.x) {
    const int j = i / (kDh / 4), d4 = (i % (kDh / 4)) * 4;
    *reinterpret_cast<float4*>(Qs + j * kDh + d4) = *reinterpret_cast<const float4*>(base + (size_t)j * (3 * kN) + d4);
    *reinterpret_cast<float4*>(Ks + j * kDh + d4) = *reinterpret_cast<const float4*>(base + (size_t)j * (3 * kN) + kN + d4);
    *reinterpret_cast<float4*>(Vs + j * kDh + d4) = *reinterpret_cast<const float4*>(base + (size_t)j * (3 * kN) + 2 * kN + d4);
    *reinterpret_cast<float4*>(Gs + j * kDh + d4) = *reinterpret_cast<const float4*>(gbase + (size_t)j * kN + d4);
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)

  // ---- phase A: per query row ----
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    float q[kDh], go[kDh], dq[kDh];
    float D = 0.f;
#pragma unroll
    for (int d = 0; d < kDh; ++d) {
      q[d] = Qs[r * kDh + d] * scale;
      go[d] = Gs[r * kDh + d];
      D = fmaf(go[d], obase[(size_t)r * kN + d], D);
      dq[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kDh; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(Ks + j * kDh + d);
        s = fmaf(q[d], kk.x, s);
        s = fmaf(q[d + 1], kk.y, s);
        s = fmaf(q[d + 2], kk.z, s);
        s = fmaf(q[d + 3], kk.w, s);
      }
      const float mn = fmaxf(m, s);
      l = l * expf(m - mn) + expf(s - mn);   // m = -inf on the first key -> expf(-inf) = 0
      m = mn;
    }
    const float L = m + logf(l);
    for (int j = 0; j < n; ++j) {
      float s = 0.f, dp = 0.f;
      float kr[kDh];
#pragma unroll
      for (int d = 0; d < kDh; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(Ks + j * kDh + d);
        const float4 vv = *reinterpret_cast<const float4*>(Vs + j * kDh + d);
        kr[d] = kk.x; kr[d + 1] = kk.y; kr[d + 2] = kk.z; kr[d + 3] = kk.w;
        s = fmaf(q[d], kk.x, s);
        s = fmaf(q[d + 1], kk.y, s);
        s = fmaf(q[d + 2], kk.z, s);
        s = fmaf(q[d + 3], kk.w, s);
        dp = fmaf(go[d], vv.x, dp);
        dp = fmaf(go[d + 1], vv.y, dp);
        dp = fmaf(go[d + 2], vv.z, dp);
        dp = fmaf(go[d + 3], vv.w, dp);
      }
      const float ds = expf(s - L) * (dp - D);
#pragma unroll
      for (int d = 0; d < kDh; ++d) dq[d] = fmaf(ds, kr[d], dq[d]);
    }
    lse[r] = L;
    Dd[r] = D;
    float* dst = dbase + (size_t)r * (3 * kN);
#pragma unroll
    for (int d = 0; d < kDh; d += 4)
      *reinterpret_cast<float4*>(dst + d) =
          make_float4(dq[d] * scale, dq[d + 1] * scale, dq[d + 2] * scale, dq[d + 3] * scale);
  }
  __syncthreads();

  // ---- phase B: per key row ----
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float k[kDh], v[kDh], dk[kDh], dv[kDh];
#pragma unroll
    for (int d = 0; d < kDh; ++d) {
      k[d] = Ks[j * kDh + d] * scale;
      v[d] = Vs[j * kDh + d];
      dk[d] = 0.f;
      dv[d] = 0.f;
    }
    for (int i = 0; i < n; ++i) {
      float s = 0.f, dp = 0.f;
      float qr[kDh], gr[kDh];
#pragma unroll
      for (int d = 0; d < kDh; d += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(Qs + i * kDh + d);
        const float4 gg = *reinterpret_cast<const float4*>(Gs + i * kDh + d);
        qr[d] = qq.x; qr[d + 1] = qq.y; qr[d + 2] = qq.z; qr[d + 3] = qq.w;
        gr[d] = gg.x; gr[d + 1] = gg.y; gr[d + 2] = gg.z; gr[d + 3] = gg.w;
        s = fmaf(qq.x, k[d], s);
        s = fmaf(qq.y, k[d + 1], s);
        s = fmaf(qq.z, k[d + 2], s);
        s = fmaf(qq.w, k[d + 3], s);
        dp = fmaf(gg.x, v[d], dp);
        dp = fmaf(gg.y, v[d + 1], dp);
        dp = fmaf(gg.z, v[d + 2], dp);
        dp = fmaf(gg.w, v[d + 3], dp);
      }
      const float p = expf(s - lse[i]);
      const float ds = p * (dp - Dd[i]);
#pragma unroll
      for (int d = 0; d < kDh; ++d) {
        dv[d] = fmaf(p, gr[d], dv[d]);
        dk[d] = fmaf(ds, qr[d], dk[d]);
      }
    }
    float* dst = dbase + (size_t)j * (3 * kN);
#pragma unroll
    for (int d = 0; d < kDh; d += 4) {
      *reinterpret_cast<float4*>(dst + kN + d) =
          make_float4(dk[d] * scale, dk[d + 1] * scale, dk[d + 2] * scale, dk[d + 3] * scale);
      *reinterpret_cast<float4*>(dst + 2 * kN + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
    }
  }
}

int launch_attention_bwd(const float* qkv, const float* out, const float* d_out, int nseq, int n,
                         float* d_qkv, cudaStream_t st) {
  if (nseq <= 0 || n <= 0) return 0;
  if ((long long)nseq * kHeads > 2147483647LL) {
    set_error("attention_bwd: nseq=%d exceeds the grid", nseq);
    return 1;
  }
  const size_t smem = ((size_t)4 * n * kDh + 2 * (size_t)n) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("attention_bwd: sequence of %d tokens does not fit shared memory (limit 390)", n);
    return 1;
  }
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return 1;
    }
    once.mark_configured();
  }
  int threads = (n + 31) / 32 * 32;
  if (threads > kAttnBwdThreads) threads = kAttnBwdThreads;
  KernelScope prof(kClsAttention, st);
  attention_bwd_kernel<<<(unsigned)nseq * kHeads, threads, smem, st>>>(qkv, out, d_out, n, d_qkv);
  return check_launch("attention_bwd_kernel");
}

}  // namespace cse
