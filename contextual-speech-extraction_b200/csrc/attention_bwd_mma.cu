// attention_bwd_mma.cu — attention backward on the tensor cores (mma.sync m16n8k16 bf16), performance mode.
//
// Reference: autograd of nn.MultiheadAttention's core softmax(q k^T / sqrt(32)) v per head
// (CSE_transformer.py:535-557), as the reference's --bf16 / --fp16 training step differentiates it
// (train_ContSep.py:383-400).  The fp32 SIMT kernel it replaces under autocast (backward.cu:attention_bwd_kernel)
// was 34 % of the measured training step (420 us per launch at 68 x 251, profiles/r02_train_step_kernels_autocast.txt).
//
// One CTA (4 warps) per (sequence, head), n <= 256 tokens; Q, K, V, dO of the head staged once in shared memory as
// bf16 (64-byte rows, XOR-swizzled for ldmatrix), softmax statistics in fp32:
//   stage    dO fp32 -> bf16;  D_i = sum_d dO_id O_id
//   phase A  (a warp owns 16 QUERY rows)  S = Q K^T -> row max / sum -> L_i = log2-sum-exp, P in registers;
//            per 16 keys: dP = dO V^T -> dS = P (dP - D) / sqrt(32) -> dQ += dS K            -> dQ rows
//   phase B  (a warp owns 16 KEY rows, everything transposed so that P^T and dS^T come out of the MMAs in the
//            A-operand layout):  S^T = K Q^T -> P^T = 2^(S^T - L);  dP^T = V dO^T -> dS^T;
//            dV += P^T dO,  dK += dS^T Q                                                     -> dK, dV rows
// S and dP are computed twice (once per orientation): seven n x n x 32 contractions per head instead of five, but no
// shared-memory transposes, no atomics and no cross-warp reductions.  P and dS are rounded to bf16 for the second
// contractions, as autocast does; accumulation and the softmax algebra are fp32.
#include "common.cuh"
#include "mma_sync.cuh"
#include "tc_ptx.cuh"

namespace cse {

namespace {

constexpr int kBwdThreads = 128;
constexpr float kScale = 0.17677669529663687f;               // 1 / sqrt(32)
constexpr float kSl2 = 0.17677669529663687f * 1.4426950408889634f;  // log2(e) / sqrt(32)

// A-operand fragments of a 16-row tile, both k16 steps of the 32 head dimensions
__device__ __forceinline__ void load_a32(uint32_t (&a)[2][4], const bf16* M, int row0, int lane) {
  const int row = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int ch = lane >> 4;
  ldmatrix_x4(a[0], sw_ptr(M, row, ch));
  ldmatrix_x4(a[1], sw_ptr(M, row, ch + 2));
}

// C[16 x 16] (two n-tiles) += A[16 x 32] * M[rows r0..r0+15][0..31]^T  (B = the rows of M as stored: [n][k])
__device__ __forceinline__ void mma_abt(float (&c0)[4], float (&c1)[4], const uint32_t (&a)[2][4], const bf16* M,
                                        int r0, int lane) {
  const int row = r0 + (lane & 7) + (lane >> 4) * 8;
  const int ch = (lane >> 3) & 1;
  uint32_t b0[4], b1[4];
  ldmatrix_x4(b0, sw_ptr(M, row, ch));
  ldmatrix_x4(b1, sw_ptr(M, row, ch + 2));
  mma_bf16_16816(c0, a[0], b0[0], b0[1]);
  mma_bf16_16816(c1, a[0], b0[2], b0[3]);
  mma_bf16_16816(c0, a[1], b1[0], b1[1]);
  mma_bf16_16816(c1, a[1], b1[2], b1[3]);
}

// C[16 x 32] (four n-tiles over the head dimensions) += A[16 x 16] * M[rows r0..r0+15][0..31]  (B via ldmatrix.trans)
__device__ __forceinline__ void mma_ab(float (&c)[4][4], const uint32_t (&a)[4], const bf16* M, int r0, int lane) {
  const int row = r0 + ((lane >> 3) & 1) * 8 + (lane & 7);
  const int ch = lane >> 4;
  uint32_t v0[4], v1[4];
  ldmatrix_x4_trans(v0, sw_ptr(M, row, ch));
  ldmatrix_x4_trans(v1, sw_ptr(M, row, ch + 2));
  mma_bf16_16816(c[0], a, v0[0], v0[1]);
  mma_bf16_16816(c[1], a, v0[2], v0[3]);
  mma_bf16_16816(c[2], a, v1[0], v1[1]);
  mma_bf16_16816(c[3], a, v1[2], v1[3]);
}

// NP = maximal number of 16-token pairs (n_pad <= 16 * NP): sizes the register arrays
template <int NP>
__global__ void __launch_bounds__(kBwdThreads)
attention_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ O, const float* __restrict__ dO,
                         int n, int n_pad, float* __restrict__ dqkv) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + (size_t)n_pad * 32;
  bf16* Vs = Ks + (size_t)n_pad * 32;
  bf16* Gs = Vs + (size_t)n_pad * 32;                       // dO
  float* Ls = reinterpret_cast<float*>(Gs + (size_t)n_pad * 32);  // log2-sum-exp of the scaled scores, per query
  float* Ds = Ls + n_pad;                                   // D_i = dO_i . O_i
  const int h = blockIdx.x & 7;
  const size_t seq = blockIdx.x >> 3;
  const size_t row0 = seq * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int npairs = n_pad >> 4;

  // ---- stage Q, K, V (bf16, cp.async) and dO (fp32 -> bf16) of this head; D_i ----
  {
    const int ch = threadIdx.x & 3, j0 = threadIdx.x >> 2;  // 16-byte chunk of the 64-byte row, row
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      bf16* dstM = which == 0 ? Qs : which == 1 ? Ks : Vs;
      for (int j = j0; j < n_pad; j += kBwdThreads / 4) {
        bf16* dst = dstM + j * 32 + ((ch ^ ((j >> 1) & 3)) << 3);
        if (j < n) {
          const bf16* src = qkv + (row0 + j) * (3 * kN) + which * kN + h * kDh + ch * 8;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                       : "memory");
        } else {
          *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    for (int j = j0; j < n_pad; j += kBwdThreads / 4) {
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      float d = 0.f;
      if (j < n) {
        const float* gp = dO + (row0 + j) * kN + h * kDh + ch * 8;
        const float4 a = *reinterpret_cast<const float4*>(gp), b = *reinterpret_cast<const float4*>(gp + 4);
        const f8 o = ld8(O + (row0 + j) * kN + h * kDh + ch * 8);
        d = a.x * o.v[0] + a.y * o.v[1] + a.z * o.v[2] + a.w * o.v[3] + b.x * o.v[4] + b.y * o.v[5] + b.z * o.v[6] +
            b.w * o.v[7];
        u.x = pack_bf16(a.x, a.y);
        u.y = pack_bf16(a.z, a.w);
        u.z = pack_bf16(b.x, b.y);
        u.w = pack_bf16(b.z, b.w);
      }
      *reinterpret_cast<uint4*>(Gs + j * 32 + ((ch ^ ((j >> 1) & 3)) << 3)) = u;
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      if (ch == 0) Ds[j] = d;
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();

  // ================= phase A: query tiles -> L_i, dQ =================
  for (int mt = warp; mt < npairs; mt += kBwdThreads / 32) {
    uint32_t qa[2][4], ga[2][4];
    load_a32(qa, Qs, mt * 16, lane);
    load_a32(ga, Gs, mt * 16, lane);
    float s[2 * NP][4];
#pragma unroll
    for (int np = 0; np < NP; ++np) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[2 * np][e] = s[2 * np + 1][e] = 0.f;
      if (np < npairs) mma_abt(s[2 * np], s[2 * np + 1], qa, Ks, np * 16, lane);
    }
    // softmax statistics of rows g and g + 8 (keys >= n masked)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2 * NP; ++nt) {
      const int key = nt * 8 + t4 * 2;
      if (key >= n) s[nt][0] = s[nt][2] = -INFINITY;
      if (key + 1 >= n) s[nt][1] = s[nt][3] = -INFINITY;
      m0 = tc::max3_f32(m0, s[nt][0], s[nt][1]);
      m1 = tc::max3_f32(m1, s[nt][2], s[nt][3]);
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float ms0 = m0 * kSl2, ms1 = m1 * kSl2;  // finite: key 0 < n
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2 * NP; ++nt) {
      s[nt][0] = ex2_approx(fmaf(s[nt][0], kSl2, -ms0));  // masked: 2^-inf = 0
      s[nt][1] = ex2_approx(fmaf(s[nt][1], kSl2, -ms0));
      s[nt][2] = ex2_approx(fmaf(s[nt][2], kSl2, -ms1));
      s[nt][3] = ex2_approx(fmaf(s[nt][3], kSl2, -ms1));
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    if (t4 == 0) {  // padded query rows: P^T = 2^(s - inf) = 0 in phase B
      Ls[r0] = r0 < n ? ms0 + log2f(l0) : INFINITY;
      Ls[r1] = r1 < n ? ms1 + log2f(l1) : INFINITY;
    }
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const float d0 = Ds[r0], d1 = Ds[r1];
    float dq[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[i][e] = 0.f;
#pragma unroll
    for (int np = 0; np < NP; ++np) {
      if (np < npairs) {
        float dp0[4] = {0.f, 0.f, 0.f, 0.f}, dp1[4] = {0.f, 0.f, 0.f, 0.f};
        mma_abt(dp0, dp1, ga, Vs, np * 16, lane);  // dP = dO V^T for these 16 keys
        // dS = P (dP - D) / sqrt(32), packed as the A operand of dQ += dS K (k16 = these 16 keys)
        uint32_t dsa[4];
        dsa[0] = pack_bf16(s[2 * np][0] * i0 * (dp0[0] - d0) * kScale, s[2 * np][1] * i0 * (dp0[1] - d0) * kScale);
        dsa[1] = pack_bf16(s[2 * np][2] * i1 * (dp0[2] - d1) * kScale, s[2 * np][3] * i1 * (dp0[3] - d1) * kScale);
        dsa[2] = pack_bf16(s[2 * np + 1][0] * i0 * (dp1[0] - d0) * kScale, s[2 * np + 1][1] * i0 * (dp1[1] - d0) * kScale);
        dsa[3] = pack_bf16(s[2 * np + 1][2] * i1 * (dp1[2] - d1) * kScale, s[2 * np + 1][3] * i1 * (dp1[3] - d1) * kScale);
        mma_ab(dq, dsa, Ks, np * 16, lane);
      }
    }
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int col = h * kDh + dt * 8 + t4 * 2;
      if (r0 < n) *reinterpret_cast<float2*>(dqkv + (row0 + r0) * (3 * kN) + col) = make_float2(dq[dt][0], dq[dt][1]);
      if (r1 < n) *reinterpret_cast<float2*>(dqkv + (row0 + r1) * (3 * kN) + col) = make_float2(dq[dt][2], dq[dt][3]);
    }
  }
  __syncthreads();  // L_i of every query is in shared memory

  // ================= phase B: key tiles -> dK, dV (transposed orientation) =================
  for (int jt = warp; jt < npairs; jt += kBwdThreads / 32) {
    uint32_t ka[2][4], va[2][4];
    load_a32(ka, Ks, jt * 16, lane);
    load_a32(va, Vs, jt * 16, lane);
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;
    for (int ip = 0; ip < npairs; ++ip) {  // 16 queries per step
      float st0[4] = {0.f, 0.f, 0.f, 0.f}, st1[4] = {0.f, 0.f, 0.f, 0.f};
      float dp0[4] = {0.f, 0.f, 0.f, 0.f}, dp1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_abt(st0, st1, ka, Qs, ip * 16, lane);  // S^T = K Q^T   : rows = keys g, g + 8; columns = queries
      mma_abt(dp0, dp1, va, Gs, ip * 16, lane);  // dP^T = V dO^T
      const int c0 = ip * 16 + t4 * 2;           // this thread's query columns: c0, c0 + 1 (tile 0), + 8 (tile 1)
      const float2 La = *reinterpret_cast<const float2*>(Ls + c0), Lb = *reinterpret_cast<const float2*>(Ls + c0 + 8);
      const float2 Da = *reinterpret_cast<const float2*>(Ds + c0), Db = *reinterpret_cast<const float2*>(Ds + c0 + 8);
      float p[8];
      p[0] = ex2_approx(fmaf(st0[0], kSl2, -La.x));
      p[1] = ex2_approx(fmaf(st0[1], kSl2, -La.y));
      p[2] = ex2_approx(fmaf(st0[2], kSl2, -La.x));
      p[3] = ex2_approx(fmaf(st0[3], kSl2, -La.y));
      p[4] = ex2_approx(fmaf(st1[0], kSl2, -Lb.x));
      p[5] = ex2_approx(fmaf(st1[1], kSl2, -Lb.y));
      p[6] = ex2_approx(fmaf(st1[2], kSl2, -Lb.x));
      p[7] = ex2_approx(fmaf(st1[3], kSl2, -Lb.y));
      uint32_t pa[4], dsa[4];
      pa[0] = pack_bf16(p[0], p[1]);
      pa[1] = pack_bf16(p[2], p[3]);
      pa[2] = pack_bf16(p[4], p[5]);
      pa[3] = pack_bf16(p[6], p[7]);
      dsa[0] = pack_bf16(p[0] * (dp0[0] - Da.x) * kScale, p[1] * (dp0[1] - Da.y) * kScale);
      dsa[1] = pack_bf16(p[2] * (dp0[2] - Da.x) * kScale, p[3] * (dp0[3] - Da.y) * kScale);
      dsa[2] = pack_bf16(p[4] * (dp1[0] - Db.x) * kScale, p[5] * (dp1[1] - Db.y) * kScale);
      dsa[3] = pack_bf16(p[6] * (dp1[2] - Db.x) * kScale, p[7] * (dp1[3] - Db.y) * kScale);
      mma_ab(dv, pa, Gs, ip * 16, lane);   // dV += P^T dO
      mma_ab(dk, dsa, Qs, ip * 16, lane);  // dK += dS^T Q
    }
    const int r0 = jt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int col = h * kDh + dt * 8 + t4 * 2;
      if (r0 < n) {
        *reinterpret_cast<float2*>(dqkv + (row0 + r0) * (3 * kN) + kN + col) = make_float2(dk[dt][0], dk[dt][1]);
        *reinterpret_cast<float2*>(dqkv + (row0 + r0) * (3 * kN) + 2 * kN + col) = make_float2(dv[dt][0], dv[dt][1]);
      }
      if (r1 < n) {
        *reinterpret_cast<float2*>(dqkv + (row0 + r1) * (3 * kN) + kN + col) = make_float2(dk[dt][2], dk[dt][3]);
        *reinterpret_cast<float2*>(dqkv + (row0 + r1) * (3 * kN) + 2 * kN + col) = make_float2(dv[dt][2], dv[dt][3]);
      }
    }
  }
}

}  // namespace

// d_qkv [M,768] fp32 = gradient of the attention core wrt the packed bf16 qkv [M,768], given its bf16 output
// out [M,256] and the fp32 gradient d_out [M,256].  n <= 256 (the caller falls back to the fp32 kernel beyond).
int launch_attention_bwd_bf16(const bf16* qkv, const bf16* out, const float* d_out, int nseq, int n, float* d_qkv,
                              cudaStream_t st) {
  if (nseq <= 0 || n <= 0) return 0;
  if (n > 256) {
    set_error("attention_bwd_bf16: n=%d exceeds 256", n);
    return 1;
  }
  if ((long long)nseq * kHeads > 2147483647LL) {
    set_error("attention_bwd_bf16: nseq=%d exceeds the grid", nseq);
    return 1;
  }
  const int n_pad = (n + 15) / 16 * 16;
  const size_t smem = (size_t)4 * n_pad * 32 * sizeof(bf16) + 2 * (size_t)n_pad * sizeof(float);
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    if (e != cudaSuccess) {
      set_error("attention_bwd_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return 1;
    }
    once.mark_configured();
  }
  KernelScope prof(kClsAttention, st);
  if (n_pad <= 64)
    attention_bwd_mma_kernel<4><<<(unsigned)nseq * kHeads, kBwdThreads, smem, st>>>(qkv, out, d_out, n, n_pad, d_qkv);
  else
    attention_bwd_mma_kernel<16><<<(unsigned)nseq * kHeads, kBwdThreads, smem, st>>>(qkv, out, d_out, n, n_pad, d_qkv);
  return check_launch("attention_bwd_mma_kernel");
}

}  // namespace cse
