// attention.cu — fused flash-style multi-head attention for one chunk sequence.
//
// Reference: MultiheadAttention.forward -> nn.MultiheadAttention -> F.scaled_dot_product_attention
// (CSE_transformer.py:468-477,535-557; torch functional.py:6682): 8 heads x d=32, no mask, no
// dropout, need_weights=False.  Sequences are the 250(+c)-frame intra chunks and the S(+c)-chunk
// inter columns; the packed projection buffer is qkv [nseq*n, 768] (q | k | v column blocks,
// head h = columns h*32..h*32+31 of each block).
//
//  * CSE_FP32: SIMT kernel, K/V of one (sequence, head) staged in shared memory, one thread per
//    query row, exact expf, fp32 accumulate.
//  * CSE_BF16: tensor-core kernel (mma.sync.m16n8k16 bf16, fp32 accumulate), online softmax in
//    registers with ex2.approx, P kept in registers as the A operand of P*V (never written to
//    memory).  Q/K staged row-major (ldmatrix), V read through ldmatrix.trans.  At d=32 the kernel
//    is bound by the exp (MUFU) and issue rate, not by tensor throughput (QK^T + PV are ~10 % of
//    the path's FLOPs), so it is written to minimise instructions per score:
//    FMNMX + FFMA + MUFU.EX2 + FADD + 1/2 CVT.
//    Short sequences (inter stack at 2-8 s of audio, n <= 64) put all 8 heads of a sequence in
//    one CTA (one warp per head); long ones use one CTA per (sequence, head).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "mma_sync.cuh"  // sw_ptr, mma_bf16_16816, ldmatrix_x4(_trans), pack_bf16, ex2_approx

namespace cse {

// ------------------------------------------------------------------------------------------
// fp32 SIMT
// ------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 128;

__global__ void __launch_bounds__(kAttnThreads) attention_f32_kernel(const float* __restrict__ qkv,
                                                                     int n,
                                                                     float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_f[];
  float* Ks = smem_f;                  // [n][32]
  float* Vs = smem_f + (size_t)n * kDh;  // [n][32]
  const int h = blockIdx.x % kHeads;
  const size_t seq = blockIdx.x / kHeads;
  const float* base = qkv + seq * n * (3 * kN);
  for (int i = threadIdx.x; i < n * (kDh / 4); i += kAttnThreads) {
    const int j = i / (kDh / 4), d4 = i % (kDh / 4);
    const float4 kv = *reinterpret_cast<const float4*>(base + (size_t)j * (3 * kN) + kN + h * kDh + d4 * 4);
    const float4 vv = *reinterpret_cast<const float4*>(base + (size_t)j * (3 * kN) + 2 * kN + h * kDh + d4 * 4);
    *reinterpret_cast<float4*>(Ks + j * kDh + d4 * 4) = kv;
    *reinterpret_cast<float4*>(Vs + j * kDh + d4 * 4) = vv;
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  for (int r = threadIdx.x; r < n; r += kAttnThreads) {
    float q[kDh], acc[kDh];
    const float* qp = base + (size_t)r * (3 * kN) + h * kDh;
#pragma unroll
    for (int d = 0; d < kDh; d += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qp + d);
      q[d] = t.x * scale; q[d + 1] = t.y * scale; q[d + 2] = t.z * scale; q[d + 3] = t.w * scale;
    }
#pragma unroll
    for (int d = 0; d < kDh; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < n; j0 += 4) {
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < n) {
          float t = 0.f;
#pragma unroll
          for (int d = 0; d < kDh; d += 4) {
            const float4 kk = *reinterpret_cast<const float4*>(Ks + j * kDh + d);
            t = fmaf(q[d], kk.x, t);
            t = fmaf(q[d + 1], kk.y, t);
            t = fmaf(q[d + 2], kk.z, t);
            t = fmaf(q[d + 3], kk.w, t);
          }
          s[u] = t;
        } else {
          s[u] = -INFINITY;
        }
      }
      const float mb = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
      const float mn = fmaxf(m, mb);
      const float corr = expf(m - mn);  // m = -inf on the first block -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < kDh; ++d) acc[d] *= corr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < n) {
          const float p = expf(s[u] - mn);
          l += p;
#pragma unroll
          for (int d = 0; d < kDh; d += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(Vs + j * kDh + d);
            acc[d] = fmaf(p, vv.x, acc[d]);
            acc[d + 1] = fmaf(p, vv.y, acc[d + 1]);
            acc[d + 2] = fmaf(p, vv.z, acc[d + 2]);
            acc[d + 3] = fmaf(p, vv.w, acc[d + 3]);
          }
        }
      }
      m = mn;
    }
    const float inv = 1.0f / l;
    float* op = out + (seq * n + r) * kN + h * kDh;
#pragma unroll
    for (int d = 0; d < kDh; d += 4)
      *reinterpret_cast<float4*>(op + d) =
          make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
  }
}

// ------------------------------------------------------------------------------------------
// bf16 tensor-core (mma.sync m16n8k16)
// ------------------------------------------------------------------------------------------
struct TileState {
  float o[4][4];
  float m0, m1, l0, l1;  // running max (scaled log2 domain) and row sums
};

// One block of 16*NP keys (NP <= 4 pairs of 8-key tiles) for a 16-row query tile.  kTail: keys >= n
// are masked.  Everything is compile-time unrolled: 4*NP QK^T HMMAs, 16*NP scores per lane,
// 4*NP PV HMMAs.
template <int NP, bool kTail>
__device__ __forceinline__ void attn_block(TileState& st, const uint32_t (&qa)[2][4], const bf16* Ks,
                                           const bf16* Vs, int j0, int n, int lane) {
  const float sl2 = 0.17677669529663687f * 1.4426950408889634f;  // log2(e)/sqrt(32)
  const int t4 = lane & 3;
  float s[2 * NP][4];
#pragma unroll
  for (int np = 0; np < NP; ++np) {
#pragma unroll
    for (int e = 0; e < 4; ++e) s[2 * np][e] = s[2 * np + 1][e] = 0.f;
    // K fragments (B operand): matrices (keys 0-7,d 0-7) (keys 0-7,d 8-15) (keys 8-15,d 0-7)
    // (keys 8-15,d 8-15) for k-step 0; d + 16 for k-step 1
    const int krow = j0 + np * 16 + (lane & 7) + (lane >> 4) * 8;
    const int kch = (lane >> 3) & 1;
    uint32_t kb0[4], kb1[4];
    ldmatrix_x4(kb0, sw_ptr(Ks, krow, kch));
    ldmatrix_x4(kb1, sw_ptr(Ks, krow, kch + 2));
    mma_bf16_16816(s[2 * np], qa[0], kb0[0], kb0[1]);
    mma_bf16_16816(s[2 * np + 1], qa[0], kb0[2], kb0[3]);
    mma_bf16_16816(s[2 * np], qa[1], kb1[0], kb1[1]);
    mma_bf16_16816(s[2 * np + 1], qa[1], kb1[2], kb1[3]);
  }
  float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 2 * NP; ++nt) {
    if (kTail) {
      const int key = j0 + nt * 8 + t4 * 2;
      if (key >= n) s[nt][0] = s[nt][2] = -INFINITY;
      if (key + 1 >= n) s[nt][1] = s[nt][3] = -INFINITY;
    }
    bm0 = tc::max3_f32(bm0, s[nt][0], s[nt][1]);  // FMNMX3
    bm1 = tc::max3_f32(bm1, s[nt][2], s[nt][3]);
  }
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
  bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
  bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
  const float mn0 = fmaxf(st.m0, bm0 * sl2), mn1 = fmaxf(st.m1, bm1 * sl2);  // finite: key j0 < n
  const float c0 = ex2_approx(st.m0 - mn0), c1 = ex2_approx(st.m1 - mn1);
  st.l0 *= c0;
  st.l1 *= c1;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    st.o[dt][0] *= c0;
    st.o[dt][1] *= c0;
    st.o[dt][2] *= c1;
    st.o[dt][3] *= c1;
  }
  st.m0 = mn0;
  st.m1 = mn1;
#pragma unroll
  for (int kt = 0; kt < NP; ++kt) {  // 16 keys per step
    float p[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      tc::fma_f32x2(s[kt * 2 + u][0], s[kt * 2 + u][1], sl2, -mn0);  // FFMA2: two scores per instruction
      tc::fma_f32x2(s[kt * 2 + u][2], s[kt * 2 + u][3], sl2, -mn1);
      p[u][0] = ex2_approx(s[kt * 2 + u][0]);
      p[u][1] = ex2_approx(s[kt * 2 + u][1]);
      p[u][2] = ex2_approx(s[kt * 2 + u][2]);
      p[u][3] = ex2_approx(s[kt * 2 + u][3]);
      st.l0 += p[u][0] + p[u][1];
      st.l1 += p[u][2] + p[u][3];
    }
    uint32_t pa[4];
    pa[0] = pack_bf16(p[0][0], p[0][1]);
    pa[1] = pack_bf16(p[0][2], p[0][3]);
    pa[2] = pack_bf16(p[1][0], p[1][1]);
    pa[3] = pack_bf16(p[1][2], p[1][3]);
    // V fragments via ldmatrix.trans: (keys 0-7,d0) (keys 8-15,d0) (keys 0-7,d0+8) (keys 8-15,d0+8)
    const int vrow = j0 + kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
    const int vch = lane >> 4;
    uint32_t v0[4], v1[4];
    ldmatrix_x4_trans(v0, sw_ptr(Vs, vrow, vch));
    ldmatrix_x4_trans(v1, sw_ptr(Vs, vrow, vch + 2));
    mma_bf16_16816(st.o[0], pa, v0[0], v0[1]);
    mma_bf16_16816(st.o[1], pa, v0[2], v0[3]);
    mma_bf16_16816(st.o[2], pa, v1[0], v1[1]);
    mma_bf16_16816(st.o[3], pa, v1[2], v1[3]);
  }
}

// One 16-row query tile of one head against all n keys.  Qs/Ks/Vs: staged [n_pad][32] bf16
// (swizzled, pad rows zero).  Writes out rows r < n.
__device__ __forceinline__ void attn_tile(const bf16* Qs, const bf16* Ks, const bf16* Vs, int mt, int n,
                                          int n_pad, int lane, bf16* __restrict__ out_rows) {
  const int g = lane >> 2, t4 = lane & 3;
  // Q fragments (A operand): ldmatrix x4 = (rows 0-7,k0) (rows 8-15,k0) (rows 0-7,k0+8) (rows 8-15,k0+8)
  uint32_t qa[2][4];
  {
    const int qrow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int qch = lane >> 4;
    ldmatrix_x4(qa[0], sw_ptr(Qs, qrow, qch));
    ldmatrix_x4(qa[1], sw_ptr(Qs, qrow, qch + 2));
  }
  TileState st;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) st.o[i][j] = 0.f;
  st.m0 = st.m1 = -INFINITY;
  st.l0 = st.l1 = 0.f;

  const int n_full = n >> 6;  // unmasked 64-key blocks
  int j0 = 0;
  for (int b = 0; b < n_full; ++b, j0 += 64) attn_block<4, false>(st, qa, Ks, Vs, j0, n, lane);
  const int rem = (n_pad - j0) >> 4;  // 16-key pairs left (0..4), masked
  if (rem == 4) attn_block<4, true>(st, qa, Ks, Vs, j0, n, lane);
  else if (rem == 3) attn_block<3, true>(st, qa, Ks, Vs, j0, n, lane);
  else if (rem == 2) attn_block<2, true>(st, qa, Ks, Vs, j0, n, lane);
  else if (rem == 1) attn_block<1, true>(st, qa, Ks, Vs, j0, n, lane);

  float l0 = st.l0, l1 = st.l1;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    const int col = dt * 8 + t4 * 2;
    if (r0 < n) *reinterpret_cast<uint32_t*>(out_rows + (size_t)r0 * kN + col) = pack_bf16(st.o[dt][0] * i0, st.o[dt][1] * i0);
    if (r1 < n) *reinterpret_cast<uint32_t*>(out_rows + (size_t)r1 * kN + col) = pack_bf16(st.o[dt][2] * i1, st.o[dt][3] * i1);
  }
}

// HPC = heads per CTA (1: eight warps share one head's query tiles; 8: one warp per head).
template <int HPC>
__global__ void __launch_bounds__(256, HPC == 8 ? 3 : 2)
attention_bf16_kernel(const bf16* __restrict__ qkv, int n, int n_pad, bf16* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kThreads = 256;
  pdl_launch_dependents();
  pdl_wait();
  bf16* S = reinterpret_cast<bf16*>(smem_raw);  // [HPC][3][n_pad][32]
  const int h0 = (HPC == 1) ? (int)((gridDim.x - 1 - blockIdx.x) % kHeads) : 0;
  // CTAs are scheduled in blockIdx order: walk the sequences from the last one down (see launch_attention)
  const size_t rb = gridDim.x - 1 - blockIdx.x;
  const size_t seq = (HPC == 1) ? rb / kHeads : rb;
  const bf16* base = qkv + seq * n * (3 * kN);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t mat = (size_t)n_pad * 32;

  // stage Q, K, V rows: 16-byte chunks, coalesced along each row's 64 B; pad rows zero.  Chunk index =
  // (row, head, 16-byte chunk) with HPC*4 (a power of two) chunks per row, so the decomposition is shifts
  // and masks only (runtime divisions by n_pad used to be ~40 % of this kernel's instructions at n = 35).
  constexpr int kCpr = HPC * 4;               // chunks per (row, matrix)
  constexpr int kRowStep = kThreads / kCpr;   // rows advanced per loop trip: a multiple of 8, so the swizzle
  {                                           // term (j >> 1) & 3 and everything but the row are loop-invariant
    const int ch = threadIdx.x & 3;
    const int hh = (threadIdx.x >> 2) & (HPC - 1);
    const int j0 = threadIdx.x / kCpr;
    const bf16* src0 = base + (size_t)j0 * (3 * kN) + (h0 + hh) * kDh + ch * 8;
    bf16* dst0 = S + (size_t)hh * 3 * mat + j0 * 32 + ((ch ^ ((j0 >> 1) & 3)) << 3);
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      const bf16* src = src0 + which * kN;
      bf16* dst = dst0 + which * mat;
      for (int j = j0; j < n_pad; j += kRowStep, src += (size_t)kRowStep * 3 * kN, dst += kRowStep * 32) {
        if (j < n) {
          // cp.async (LDGSTS): every chunk of the CTA is in flight at once instead of one
          // load->store round trip per loop iteration
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                       "l"(src)
                       : "memory");
        } else {
          *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();

  const int n_mt = (n + 15) >> 4;
  if (HPC == 1) {
    const bf16* Qs = S;
    const bf16* Ks = S + mat;
    const bf16* Vs = S + 2 * mat;
    bf16* orow = out + seq * n * kN + h0 * kDh;
    for (int mt = wid; mt < n_mt; mt += kThreads / 32) attn_tile(Qs, Ks, Vs, mt, n, n_pad, lane, orow);
  } else {
    const bf16* Qs = S + (size_t)wid * 3 * mat;
    bf16* orow = out + seq * n * kN + wid * kDh;
    for (int mt = 0; mt < n_mt; ++mt) attn_tile(Qs, Qs + mat, Qs + 2 * mat, mt, n, n_pad, lane, orow);
  }
}

// Test hook (cse_debug_force_mma_attention): route n <= 256 through the long-sequence kernel too, so
// both attention kernels can be checked against the oracle at the same shapes.
int g_attention_mode = 0;

// Sweep direction: the GEMMs and the fused FFN walk the rows upwards; LayerNorm and attention walk them
// DOWNWARDS, so every kernel of a layer starts with the rows its producer touched last and that are still in
// the 126 MB L2 (an activation tensor of cfg2 is 70-280 MB).  Worth ~2 % of the forward.
int launch_attention(const void* qkv, int nseq, int n, int act, void* out, cudaStream_t st) {
  if (nseq <= 0 || n <= 0) return 0;
  if ((long long)nseq * kHeads > 2147483647LL) {
    set_error("attention: nseq=%d exceeds the grid", nseq);
    return 1;
  }
  KernelScope prof(kClsAttention, st);
  if (act == CSE_BF16) {
    // up to 256 tokens the whole score row fits TMEM: tcgen05 kernel (attention_tc.cu); longer
    // sequences (inter stack beyond ~30 s of audio) use the online-softmax mma.sync kernel below
    // Very short sequences (inter stack at <= 8 s of audio) are latency- not exp-bound and run
    // faster with all 8 heads of a sequence in one mma.sync CTA (measured 122 vs 163 us at n = 35).
    if (n <= 256 && g_attention_mode != 1 && (n > 64 || g_attention_mode >= 2))
      return launch_attention_tc((const bf16*)qkv, nseq, n, (bf16*)out, st);
    const int n_pad = (n + 15) / 16 * 16;
    static DeviceOnce once;
    if (!once.configured_on_this_device()) {
      cudaError_t e1 = cudaFuncSetAttribute(attention_bf16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaError_t e8 = cudaFuncSetAttribute(attention_bf16_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e1 != cudaSuccess || e8 != cudaSuccess) {
        set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e8));
        return 1;
      }
      once.mark_configured();
    }
    if (n <= 64) {  // all 8 heads of a sequence in one CTA, one warp per head
      const size_t smem = (size_t)8 * 3 * n_pad * 32 * sizeof(bf16);
      launch_pdl(attention_bf16_kernel<8>, dim3(nseq), dim3(256), smem, st, 1, (const bf16*)qkv, n, n_pad, (bf16*)out);
      return check_launch("attention_bf16_kernel<8>");
    }
    const size_t smem = (size_t)3 * n_pad * 32 * sizeof(bf16);
    if (smem > 200 * 1024) {
      set_error("attention: sequence of %d tokens does not fit shared memory", n);
      return 1;
    }
    launch_pdl(attention_bf16_kernel<1>, dim3((unsigned)nseq * kHeads), dim3(256), smem, st, 1, (const bf16*)qkv, n, n_pad, (bf16*)out);
    return check_launch("attention_bf16_kernel<1>");
  }
  const size_t smem = (size_t)n * kDh * sizeof(float) * 2;
  static DeviceOnce once32;
  if (!once32.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return 1;
    }
    once32.mark_configured();
  }
  if (smem > 200 * 1024) {
    set_error("attention: sequence of %d tokens does not fit shared memory (fp32)", n);
    return 1;
  }
  attention_f32_kernel<<<(unsigned)nseq * kHeads, kAttnThreads, smem, st>>>((const float*)qkv, n, (float*)out);
  return check_launch("attention_f32_kernel");
}

}  // namespace cse
