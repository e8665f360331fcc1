// attention_tc.cu — tcgen05 / TMEM attention for sequences of up to 256 tokens (every intra chunk,
// and the inter stack up to ~30 s of audio).
//
// Reference: nn.MultiheadAttention core = softmax(q k^T / sqrt(32)) v per head, no mask
// (CSE_transformer.py:535-557 -> torch functional.py:6682).
//
// One work item = one 128-row query tile of one head:
//   * n <= 128 ("packed"): floor(128/n) consecutive sequences share a tile; S is block-diagonal and
//     every row only exponentiates its own sequence's n keys (the inter stack at 2-16 s of audio).
//   * 128 < n <= 256 ("split"): two query tiles per sequence against all n keys (the intra stack).
// Pipeline per item (two items in flight per SM, one per 128-thread softmax group; warp 1 issues the S
// MMAs, warp 10 the PV MMAs, two producer lanes of warp 0 stream Q/K and V):
//   TMA (SWIZZLE_64B boxes of the packed qkv buffer: Q 128x32, K/V up to 256x32)
//   -> tcgen05.mma  S[128 x Ncols] = Q K^T   (both operands K-major, fp32 accumulators in TMEM)
//   -> 128 softmax threads, ONE ROW EACH, ONE pass over the row in 64-key chunks: online softmax — chunk c is exponentiated against
//      the running maximum m_c, P_c is packed to bf16 into a SWIZZLE_128B K-major smem k-block
//   -> tcgen05.mma  O_c[128 x 32] = P_c V_c per chunk, as soon as the four warps of the group have
//      delivered P_c (A = P from smem, B = V as an MN-major operand, exactly as TMA wrote it); O_c
//      lands in the first 32 TMEM columns of the chunk's own, already consumed, score columns
//   -> the same threads read the O_c, re-weight them by 2^(m_c - m_final) (<= 1), divide by the
//      re-weighted row sum and store bf16.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

using namespace tc;

constexpr int kAtThreads = 352;  // warp 0 TMA, warp 1 S-MMA, warps 2-5 softmax group 0, 6-9 group 1, warp 10 PV-MMA
constexpr int kAtQ = 128 * 64;        // 8 KB   Q tile  [128 rows x 32 bf16]
constexpr int kAtKV = 256 * 64;       // 16 KB  K or V  [256 rows x 32 bf16]
constexpr int kAtP = 4 * 128 * 128;   // 64 KB  P       4 k-blocks of [128 rows x 64 bf16]
constexpr int kAtBuf = kAtQ + 2 * kAtKV + kAtP;
constexpr size_t kAtSmem = 1024 + 2 * (size_t)kAtBuf + 384;

struct AtItem {
  int q_row0, q_rows, kv_row0, kv_rows, head;
};

__device__ __forceinline__ AtItem at_decode(int item, int n, int nseq, int g) {
  AtItem it;
  it.head = item & 7;
  const int unit = item >> 3;
  if (g > 0) {  // packed: g sequences per tile
    const int s0 = unit * g;
    const int cnt = min(g, nseq - s0);
    it.q_row0 = it.kv_row0 = s0 * n;
    it.q_rows = it.kv_rows = cnt * n;
  } else {      // split: two query tiles per sequence
    const int seq = unit >> 1, mt = unit & 1;
    it.q_row0 = seq * n + mt * 128;
    it.q_rows = min(128, n - mt * 128);
    it.kv_row0 = seq * n;
    it.kv_rows = n;
  }
  return it;
}

__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// max over 64 values: four independent chains of three-input FMNMX3 (a serial fmaxf chain costs 64 x 4 cycles)
__device__ __forceinline__ float at_max64(const float (&v)[64]) {
  float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
  for (int i = 4; i < 60; i += 8) {
    m0 = max3_f32(m0, v[i], v[i + 4]);
    m1 = max3_f32(m1, v[i + 1], v[i + 5]);
    m2 = max3_f32(m2, v[i + 2], v[i + 6]);
    m3 = max3_f32(m3, v[i + 3], v[i + 7]);
  }
  m0 = fmaxf(m0, v[60]);
  m1 = fmaxf(m1, v[61]);
  m2 = fmaxf(m2, v[62]);
  m3 = fmaxf(m3, v[63]);
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

__global__ void __launch_bounds__(kAtThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, int n, int nseq,
                    int g, int items) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t sBar = smem_base + 2 * kAtBuf;
  // per buffer b: qk_full, v_full, s_full, o_full, free  (8 bytes each)
  auto bar = [&](int which, int b) -> uint32_t { return sBar + (uint32_t)(which * 2 + b) * 8; };
  enum { QK_FULL = 0, V_FULL = 1, S_FULL = 2, O_FULL = 3, FREE = 4 };
  auto bar_p = [&](int b, int cp) -> uint32_t { return sBar + (uint32_t)(10 + b * 4 + cp) * 8; };  // P_c written
  const uint32_t tmem_slot = sBar + 18 * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));
  auto sQ = [&](int b) { return smem_base + (uint32_t)b * kAtBuf; };
  auto sK = [&](int b) { return smem_base + (uint32_t)b * kAtBuf + kAtQ; };
  auto sV = [&](int b) { return smem_base + (uint32_t)b * kAtBuf + kAtQ + kAtKV; };
  auto sP = [&](int b) { return smem_base + (uint32_t)b * kAtBuf + kAtQ + 2 * kAtKV; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(QK_FULL, b), 1);
      mbar_init(bar(V_FULL, b), 1);
      mbar_init(bar(S_FULL, b), 1);
      mbar_init(bar(O_FULL, b), 1);
      mbar_init(bar(FREE, b), 4);
      for (int cp = 0; cp < 4; ++cp) mbar_init(bar_p(b, cp), 4);  // one arrive per softmax warp of the group
    }
    mbar_fence_init();
  }
  pdl_launch_dependents();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // prologue done; from here on the previous kernel's output is read

  if (warp == 0) {
    // ================= TMA producers: lane 0 streams Q/K, lane 1 streams V =================
    // (separate lanes so a V load that waits for the previous PV-MMA never delays the next Q/K)
    if (lane < 2) {
      int k = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int b = k & 1;
        const uint32_t ph = (uint32_t)(k >> 1) & 1u;
        const AtItem it = at_decode(items - 1 - item, n, nseq, g);  // descending sweep, see launch_attention
        const int nk = it.kv_rows > 128 ? 2 : 1;
        if (lane == 0) {
          // Q/K of this buffer were last read by the S-MMA of item k-2
          if (k >= 2) mbar_wait(bar(S_FULL, b), ph ^ 1u, 10);
          mbar_expect_tx(bar(QK_FULL, b), (uint32_t)(kAtQ + nk * 8192));
          tma_load_2d(sQ(b), &tmQKV, bar(QK_FULL, b), it.head * kDh, it.q_row0);
          tma_load_2d(sK(b), &tmQKV, bar(QK_FULL, b), kN + it.head * kDh, it.kv_row0);
          if (nk == 2) tma_load_2d(sK(b) + 8192, &tmQKV, bar(QK_FULL, b), kN + it.head * kDh, it.kv_row0 + 128);
        } else {
          // V of this buffer was last read by the PV-MMA of item k-2
          if (k >= 2) mbar_wait(bar(O_FULL, b), ph ^ 1u, 11);
          mbar_expect_tx(bar(V_FULL, b), (uint32_t)(nk * 8192));
          tma_load_2d(sV(b), &tmQKV, bar(V_FULL, b), 2 * kN + it.head * kDh, it.kv_row0);
          if (nk == 2) tma_load_2d(sV(b) + 8192, &tmQKV, bar(V_FULL, b), 2 * kN + it.head * kDh, it.kv_row0 + 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= S = Q K^T issuer =================
    if (lane == 0) {
      int k = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int b = k & 1;
        const uint32_t ph = (uint32_t)(k >> 1) & 1u;
        const AtItem it = at_decode(items - 1 - item, n, nseq, g);  // descending sweep, see launch_attention
        const int ncols = (it.kv_rows + 15) & ~15;
        mbar_wait(bar(QK_FULL, b), ph, 22);
        if (k >= 2) mbar_wait(bar(FREE, b), ph ^ 1u, 23);  // O of item k-2 drained from TMEM
        fence_after();
        const uint32_t idesc_s = make_idesc_bf16(128, ncols, 0, 0);
        const uint64_t ad = make_desc(sQ(b), 512, kLayoutSw64);
        const uint64_t bd = make_desc(sK(b), 512, kLayoutSw64);
        const uint32_t d_s = tmem_base + (uint32_t)b * 256;
        umma_bf16(d_s, ad, bd, idesc_s, 0u);
        umma_bf16(d_s, ad + 2, bd + 2, idesc_s, 1u);  // second k16 step: +32 B inside the 64-B row
        umma_commit(bar(S_FULL, b));
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ================= O_c = P_c V_c issuer (items in order; the softmax warps never block on MMA issue) =====
    if (lane == 0) {
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 32, 0, 1);  // B = V is MN-major
      int k = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int b = k & 1;
        const uint32_t ph = (uint32_t)(k >> 1) & 1u;
        const AtItem it = at_decode(items - 1 - item, n, nseq, g);  // descending sweep, see launch_attention
        const int ncols = (it.kv_rows + 15) & ~15;
        const int npair = (ncols + 63) >> 6;
        const uint32_t d_o = tmem_base + (uint32_t)b * 256;
        mbar_wait(bar(V_FULL, b), ph, 21);
        for (int cp = 0; cp < 4; ++cp) {
          mbar_wait(bar_p(b, cp), ph, 24);
          fence_after();
          const int t1 = (cp < npair) ? min(4 * cp + 4, ncols / 16) : 0;
          for (int t = 4 * cp; t < t1; ++t) {
            const uint64_t ad = make_desc(sP(b) + (uint32_t)cp * 16384, 1024, kLayoutSw128) + (uint64_t)(2 * (t & 3));
            const uint64_t bd = make_desc(sV(b) + (uint32_t)t * 1024, 512, kLayoutSw64);
            umma_bf16(d_o + (uint32_t)cp * 64, ad, bd, idesc_pv, (t & 3) != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar(O_FULL, b));
      }
    }
    __syncwarp();
  } else {
    // ================= softmax + epilogue groups (128 threads each) =================
    const int grp = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const float sl2 = 0.17677669529663687f * 1.4426950408889634f;  // log2(e)/sqrt(32)
    int k = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
      if ((k & 1) != grp) continue;
      const int b = grp;
      const uint32_t ph = (uint32_t)(k >> 1) & 1u;
      const AtItem it = at_decode(items - 1 - item, n, nseq, g);  // descending sweep, see launch_attention
      const int ncols = (it.kv_rows + 15) & ~15;
      const int npair = (ncols + 63) >> 6;  // 64-key blocks
      int lo = 0, hi = it.kv_rows;
      if (g > 0) {  // packed: this row's own sequence
        const int sidx = min(r / n, it.q_rows / n - 1);
        lo = sidx * n;
        hi = lo + n;
      }
      // tcgen05.ld is warp-collective (.sync.aligned): chunk loops must be warp-uniform, so they run
      // over the union [wlo, whi) of the key ranges of the warp's 32 rows; per-row masks inside.
      const int wlo = __reduce_min_sync(0xffffffffu, lo);
      const int whi = __reduce_max_sync(0xffffffffu, hi);
      const bool uniform = (g == 0);  // split mode: every row of the tile has the same key range
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)b * 256;
      mbar_wait(bar(S_FULL, b), ph, 30);
      fence_after();
      // ---- single pass over S: online softmax per 64-key chunk ----
      float m_run = -INFINITY;
      float mc[4], lc[4];
      unsigned char* prow = smem_al + (sP(b) - smem_base) + r * 128;
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        mc[cp] = -INFINITY;
        lc[cp] = 0.f;
        if (cp < npair) {
          const int c0 = cp * 64;
          uint4* dst = reinterpret_cast<uint4*>(prow + cp * 16384);  // k-block cp: [128 rows x 64 keys]
          if (c0 + 64 <= wlo || c0 >= whi) {  // warp-uniform: no row of this warp attends these keys
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i ^ (r & 7)] = make_uint4(0u, 0u, 0u, 0u);
          } else {
            float v[64];
            tmem_ld64(taddr + c0, v);
            if (!(uniform && c0 + 64 <= hi)) {
#pragma unroll
              for (int i = 0; i < 64; ++i)
                if (c0 + i < lo || c0 + i >= hi) v[i] = -INFINITY;
            }
            const float m_new = fmaxf(m_run, at_max64(v) * sl2);
            const float msc = (m_new == -INFINITY) ? 0.f : m_new;  // row has no key yet (packed mode)
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {  // FFMA2 for the scale/shift, MUFU.EX2, FADD2 for the row sum
              fma_f32x2(v[i], v[i + 1], sl2, -msc);
              fma_f32x2(v[i + 2], v[i + 3], sl2, -msc);
              v[i] = at_ex2(v[i]);  // masked: 2^-inf = 0
              v[i + 1] = at_ex2(v[i + 1]);
              v[i + 2] = at_ex2(v[i + 2]);
              v[i + 3] = at_ex2(v[i + 3]);
              add_f32x2(s0, s1, v[i], v[i + 1]);
              add_f32x2(s2, s3, v[i + 2], v[i + 3]);
            }
            mc[cp] = m_new;
            lc[cp] = (s0 + s1) + (s2 + s3);
            m_run = m_new;
            // 16-byte chunk i XOR (row & 7): SWIZZLE_128B K-major
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 u;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
              h[0] = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
              h[1] = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
              h[2] = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
              h[3] = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
              dst[i ^ (r & 7)] = u;
            }
          }
          fence_before();        // this warp's reads of the chunk's score columns precede O_c overwriting them
          fence_proxy_async();   // P_c (generic-proxy stores) visible to the tensor core's async proxy
        }
        // (chunks past the item's keys are signalled too: the barrier phases must advance once per item)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p(b, cp));
      }
      // ---- epilogue: O = sum_c 2^(m_c - m) O_c, row sum likewise, O / rowsum -> bf16 ----
      float fc[4], sum = 0.f;
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        fc[cp] = (mc[cp] == -INFINITY) ? 0.f : at_ex2(mc[cp] - m_run);
        sum += fc[cp] * lc[cp];
      }
      mbar_wait(bar(O_FULL, b), ph, 31);
      fence_after();
      {
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
#pragma unroll
        for (int cp = 0; cp < 4; ++cp) {
          if (cp < npair) {
            float oc[32];
            tmem_ld32(taddr + cp * 64, oc);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = fmaf(fc[cp], oc[i], o[i]);
          }
        }
        const float inv = 1.0f / sum;
        if (r < it.q_rows) {
          uint4* dsto = reinterpret_cast<uint4*>(out + (size_t)(it.q_row0 + r) * kN + it.head * kDh);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
            h[0] = __floats2bfloat162_rn(o[8 * i] * inv, o[8 * i + 1] * inv);
            h[1] = __floats2bfloat162_rn(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
            h[2] = __floats2bfloat162_rn(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
            h[3] = __floats2bfloat162_rn(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
            dsto[i] = u;
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(FREE, b));
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// v4.  What bounded v1 (148 us at 544 x 251, ~4.8 k cycles per item) was neither MUFU nor the tensor pipe but the
// INSTRUCTION STREAM OF THE ISSUING THREADS: tools/micro/umma_pv_rate.cu issues one item's eighteen tcgen05.mma from a
// fully unrolled body in ~850 cycles (45-55 per instruction, whatever the accumulator pattern, P in shared or tensor
// memory, with the softmax warps' tcgen05.ld/st traffic beside it), while the kernels' issue loops — item decode,
// descriptor construction, barrier address arithmetic, R2UR moves, ~100 dependent instructions per 4-instruction
// chunk — took ~1000 cycles per chunk (clock64 traces, profiles/r02_attention_trace.txt).  v4 therefore has
//   * one issuing thread PER TMEM BUFFER (warps 1 and 10), each strictly sequential (S_k, P V_k chunks 0-3,
//     S_{k+2}, ...): no polling, descriptors are a precomputed low word + constant, the four instructions of a
//     chunk are unrolled;
//   * a softmax warp's time is dominated by TMEM round trips (~250-300 cycles per tcgen05.ld / tcgen05.st + wait;
//     a two-pass variant with the row maximum taken first measured 2.4 k cycles for that pass alone), so the
//     single-pass online softmax stays (per-chunk maxima m_c and accumulators O_c, re-weighted by 2^(m_c - m)
//     in the epilogue), with the score load of chunk c+1 in flight while chunk c is exponentiated, the
//     tcgen05.st of P_c completing under the maximum search of chunk c+1, and the four O_c loads of the
//     epilogue issued together;
//   * P never touches shared memory: bf16 pairs go back into the item's own consumed score columns and feed
//     P_c V_c as the TMEM A operand; the 128 KB of P staging of v1 are a 4-deep Q/K/V ring;
//   * POLY pairs out of every 8 pairs of exponentials can be evaluated on the FMA pipe (Cody-Waite split + degree-3
//     polynomial, relative error 7.5e-5, far below the bf16 rounding of P) instead of MUFU.EX2; measured no gain
//     (174 / 170 / 172 us at POLY = 0 / 2 / 3) while the kernel is not MUFU-bound, so only POLY = 0 is instantiated.
// TMEM columns of buffer b (256 per item): chunk c (64 keys) = [64c, 64c+64): scores -> P_c in [64c, 64c+32)
// (two bf16 per column), O_c (fp32 [128 x 32]) in [64c+32, 64c+64).
// ------------------------------------------------------------------------------------------
// warp 0 TMA, warps 1 / 2 issuers of TMEM buffer 0 / 1, warps 3-18 softmax: group (= buffer) x key half x lane quarter
constexpr int kA4Threads = 608;
constexpr long long kA4Stagger = 3500;  // cycles: about half of one group's item period
constexpr int kA4Regs = 96;   // five warps on one SM sub-partition (16 K registers each): 5 x 32 x 96 = 15360
constexpr int kA2Stages = 4;
constexpr int kA2StageBytes = kAtQ + 2 * kAtKV;  // 40 KB: Q tile + K + V
constexpr size_t kA2Smem = 1024 + (size_t)kA2Stages * kA2StageBytes + 512 + 2 * 2 * 2 * 128 * 16;
constexpr int kA2TraceSlots = 16, kA2TraceItems = 64;

__device__ __forceinline__ uint64_t desc_from(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;\n" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x for x <= 0 on the FMA pipe: x = j + f, j = round(x), |f| <= 0.5; 2^f by a degree-3 minimax polynomial
// (max relative error 7.5e-5), 2^j by adding j to the exponent field.  Inputs below -126 (masked keys: -inf)
// are clamped; they come out as ~1e-38 instead of 0.
__device__ __forceinline__ void ex2_poly2(float& x0, float& x1) {
  const float kMagic = 12582912.f;  // 1.5 * 2^23: adding it rounds to an integer in the low mantissa bits
  x0 = fmaxf(x0, -126.f);  // (-127 would wrap the exponent field)
  x1 = fmaxf(x1, -126.f);
  const unsigned long long x = pack_f32x2(x0, x1);
  const unsigned long long xf = fadd2(x, pack_f32x2(kMagic, kMagic));
  const unsigned long long j = fadd2(xf, pack_f32x2(-kMagic, -kMagic));
  const unsigned long long f = ffma2(j, pack_f32x2(-1.f, -1.f), x);
  unsigned long long pq = ffma2(pack_f32x2(0.055170804f, 0.055170804f), f, pack_f32x2(0.24260928f, 0.24260928f));
  pq = ffma2(pq, f, pack_f32x2(0.69326097f, 0.69326097f));
  pq = ffma2(pq, f, pack_f32x2(0.99992818f, 0.99992818f));
  float p0, p1, f0, f1;
  unpack_f32x2(pq, p0, p1);
  unpack_f32x2(xf, f0, f1);
  x0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(f0) << 23));
  x1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(f1) << 23));
}

__device__ __forceinline__ float at_max32(const uint32_t (&u)[32]) {
  float m0 = __uint_as_float(u[0]), m1 = __uint_as_float(u[1]);
#pragma unroll
  for (int i = 2; i < 30; i += 4) {
    m0 = max3_f32(m0, __uint_as_float(u[i]), __uint_as_float(u[i + 2]));
    m1 = max3_f32(m1, __uint_as_float(u[i + 1]), __uint_as_float(u[i + 3]));
  }
  return fmaxf(max3_f32(m0, m1, __uint_as_float(u[30])), __uint_as_float(u[31]));
}

template <int POLY>
__global__ void __maxnreg__(kA4Regs)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, int n, int nseq,
                     int g, int items, long long* __restrict__ trace) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t sBar = smem_base + kA2Stages * kA2StageBytes;
  // per stage: QK_FULL, V_FULL, QK_EMPTY, V_EMPTY; per TMEM buffer: S_FULL, O_FULL, FREE, P[4]
  auto bar_st = [&](int which, int st) -> uint32_t { return sBar + (uint32_t)(which * kA2Stages + st) * 8; };
  enum { QK_FULL = 0, V_FULL = 1, QK_EMPTY = 2, V_EMPTY = 3 };
  const uint32_t sBarB = sBar + 4 * kA2Stages * 8;
  auto bar_b = [&](int which, int b) -> uint32_t { return sBarB + (uint32_t)(which * 2 + b) * 8; };
  enum { S_FULL = 0, O_FULL = 1, FREE = 2 };
  auto bar_p = [&](int b, int cp) -> uint32_t { return sBarB + (uint32_t)(6 + b * 4 + cp) * 8; };
  const uint32_t tmem_slot = sBarB + 14 * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));
  auto sQ = [&](int st) { return smem_base + (uint32_t)st * kA2StageBytes; };
  auto sK = [&](int st) { return smem_base + (uint32_t)st * kA2StageBytes + kAtQ; };
  auto sV = [&](int st) { return smem_base + (uint32_t)st * kA2StageBytes + kAtQ + kAtKV; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = blockIdx.x < items ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const bool tracing = trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int k, int slot) {
    if (tracing && k < kA2TraceItems) trace[k * kA2TraceSlots + slot] = clock64();
  };
  // Every item has the same key count except, in packed mode, the items of the last (partial) unit.
  const int units = g > 0 ? (nseq + g - 1) / g : 2 * nseq;
  const int kv_full = g > 0 ? g * n : n, kv_last = g > 0 ? (nseq - (units - 1) * g) * n : n;
  auto item_index = [&](int k) { return items - 1 - ((int)blockIdx.x + k * (int)gridDim.x); };  // descending sweep
  auto item_ncols = [&](int k) { return (((item_index(k) >> 3) == units - 1 ? kv_last : kv_full) + 15) & ~15; };

  if (threadIdx.x == 0) {
    for (int st = 0; st < kA2Stages; ++st) {
      mbar_init(bar_st(QK_FULL, st), 1);
      mbar_init(bar_st(V_FULL, st), 1);
      mbar_init(bar_st(QK_EMPTY, st), 1);
      mbar_init(bar_st(V_EMPTY, st), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_b(S_FULL, b), 1);
      mbar_init(bar_b(O_FULL, b), 1);
      mbar_init(bar_b(FREE, b), 8);                                // the group's eight softmax warps
      for (int cp = 0; cp < 4; ++cp) mbar_init(bar_p(b, cp), 4);  // the four lane quarters of the chunk's key half
    }
    mbar_fence_init();
  }
  pdl_launch_dependents();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // prologue done; from here on the previous kernel's output is read

  if (warp == 0) {
    // ================= TMA producers: lane 0 streams Q/K, lane 1 streams V, up to kA2Stages items ahead ==========
    if (lane < 2) {
      for (int k = 0; k < nitems; ++k) {
        const int st = k & (kA2Stages - 1);
        const uint32_t ph = (uint32_t)(k / kA2Stages) & 1u;
        const AtItem it = at_decode(item_index(k), n, nseq, g);
        const int nk = it.kv_rows > 128 ? 2 : 1;
        if (lane == 0) {
          if (k >= kA2Stages) mbar_wait(bar_st(QK_EMPTY, st), ph ^ 1u, 10);  // S of item k-4 has consumed the stage
          mbar_expect_tx(bar_st(QK_FULL, st), (uint32_t)(kAtQ + nk * 8192));
          tma_load_2d(sQ(st), &tmQKV, bar_st(QK_FULL, st), it.head * kDh, it.q_row0);
          tma_load_2d(sK(st), &tmQKV, bar_st(QK_FULL, st), kN + it.head * kDh, it.kv_row0);
          if (nk == 2) tma_load_2d(sK(st) + 8192, &tmQKV, bar_st(QK_FULL, st), kN + it.head * kDh, it.kv_row0 + 128);
          stamp(k, 13);
        } else {
          if (k >= kA2Stages) mbar_wait(bar_st(V_EMPTY, st), ph ^ 1u, 11);   // PV of item k-4 has consumed the stage
          mbar_expect_tx(bar_st(V_FULL, st), (uint32_t)(nk * 8192));
          tma_load_2d(sV(st), &tmQKV, bar_st(V_FULL, st), 2 * kN + it.head * kDh, it.kv_row0);
          if (nk == 2) tma_load_2d(sV(st) + 8192, &tmQKV, bar_st(V_FULL, st), 2 * kN + it.head * kDh, it.kv_row0 + 128);
          stamp(k, 14);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ================= tensor-pipe issuer of TMEM buffer b: S_k, P V_k chunk 0..3, S_{k+2}, ... =================
    if (lane == 0) {
      const int b = warp - 1;
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 32, 0, 1);  // A = P from TMEM, B = V MN-major
      // descriptor words (tc::make_desc): low = start >> 4 | LBO field 1; high = SBO >> 4 | version | layout
      constexpr uint32_t kDescHi64 = (512u >> 4) | (1u << 14) | (kLayoutSw64 << 29);
      constexpr uint32_t kStep = (uint32_t)kA2StageBytes >> 4;
      const uint32_t q_lo0 = ((sQ(0) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t k_lo0 = ((sK(0) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t v_lo0 = ((sV(0) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t t_buf = tmem_base + (uint32_t)b * 256;
      const uint32_t bS = bar_b(S_FULL, b), bO = bar_b(O_FULL, b), bF = bar_b(FREE, b), bP = bar_p(b, 0);
      for (int k = b; k < nitems; k += 2) {
        const uint32_t st = (uint32_t)k & (kA2Stages - 1);
        const uint32_t ph_st = ((uint32_t)k / kA2Stages) & 1u, ph_b = ((uint32_t)k >> 1) & 1u;
        const int ncols = item_ncols(k);
        const int ksteps = ncols >> 4;
        // ---- S = Q K^T ----
        mbar_wait_spin(bar_st(QK_FULL, st), ph_st, 22);
        if (k >= 2) mbar_wait_spin(bF, ph_b ^ 1u, 23);  // O of item k-2 has been read out of this buffer
        fence_after();
        {
          const uint32_t idesc_s = make_idesc_bf16(128, ncols, 0, 0);
          const uint64_t qd = desc_from(q_lo0 + st * kStep, kDescHi64), kd = desc_from(k_lo0 + st * kStep, kDescHi64);
          umma_bf16(t_buf, qd, kd, idesc_s, 0u);
          umma_bf16(t_buf, qd + 2, kd + 2, idesc_s, 1u);  // second k16 step: +32 B inside the 64-B row
          umma_commit(bS);
          umma_commit(bar_st(QK_EMPTY, st));
        }
        stamp(k, 0);
        // ---- O_c = P_c V_c, c = 0..3 ----
        mbar_wait_spin(bar_st(V_FULL, st), ph_st, 21);
        const uint64_t vd = desc_from(v_lo0 + st * kStep, kDescHi64);
#pragma unroll
        for (int cp = 0; cp < 4; ++cp) {
          mbar_wait_spin(bP + 8 * cp, ph_b, 24);
          fence_after();
          const uint32_t a0 = t_buf + (uint32_t)cp * 64;
          const uint64_t vc = vd + (uint64_t)(cp * 4 * 64);   // 1024 B of V per 16 keys
          const int left = ksteps - 4 * cp;
          if (left >= 4) {
            umma_bf16_ts(a0 + 32, a0, vc, idesc_pv, 0u);
            umma_bf16_ts(a0 + 32, a0 + 8, vc + 64, idesc_pv, 1u);
            umma_bf16_ts(a0 + 32, a0 + 16, vc + 128, idesc_pv, 1u);
            umma_bf16_ts(a0 + 32, a0 + 24, vc + 192, idesc_pv, 1u);
          } else {
            for (int j = 0; j < left; ++j)
              umma_bf16_ts(a0 + 32, a0 + 8 * (uint32_t)j, vc + (uint64_t)(64 * j), idesc_pv, j != 0 ? 1u : 0u);
          }
          stamp(k, 1 + cp);
        }
        umma_commit(bO);
        umma_commit(bar_st(V_EMPTY, st));
      }
    }
    __syncwarp();
  } else {
    // ================= softmax + epilogue: 256 threads per item = (query row) x (key half) =================
    // Warp (quarter q, half h) of a group owns TMEM lanes [32q, 32q+32) and the score chunks 2h, 2h+1 (keys
    // [128h, 128h+128)): a single warp cannot keep MUFU busy (two warps per SM sub-partition ran the exponentials
    // at half the MUFU rate and idled it during their epilogues), four per sub-partition can.  Each warp runs the
    // online softmax over its own two chunks; the two halves of a row meet once per item, through shared memory,
    // to agree on the row maximum and the row sum, and each converts half of the 32 output columns.
    const int sw = warp - 3;
    const int grp = sw >> 3, half = (sw >> 2) & 1;
    const int quarter = warp & 3;       // TMEM lane quarter: fixed by hardware to warp id % 4
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const float sl2 = 0.17677669529663687f * 1.4426950408889634f;  // log2(e)/sqrt(32)
    const bool tr = tracing && quarter == 0 && lane == 0 && half == 0;
    const int b = grp;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)b * 256;
    const uint32_t bS = bar_b(S_FULL, b), bO = bar_b(O_FULL, b), bF = bar_b(FREE, b), bP = bar_p(b, 0);
    // exchange of (m_c, l_c) between the two halves of a row: [item parity][group][half][row] float4
    float4* xchg = reinterpret_cast<float4*>(smem_al + (sBarB + 16 * 8 - smem_base));
    const int r_seq = g > 0 ? r / n : 0;  // packed mode: which of the tile's sequences this row belongs to
    // The two groups would otherwise run in lock-step (same start, same period): both in their exponential phases
    // (MUFU shared four ways) and then both in their MUFU-free phases (exchange, O wait, epilogue).  Start group 1
    // about half an item late, so that one group's exponentials run under the other's epilogue.
    if (grp == 1 && nitems > 2) {
      const long long t0 = clock64();
      while (clock64() - t0 < kA4Stagger) {}
    }
    for (int k = grp; k < nitems; k += 2) {
      const uint32_t ph = (uint32_t)(k >> 1) & 1u;
      const int idx = item_index(k);
      const int ncols = item_ncols(k);
      const int npair = (ncols + 63) >> 6;  // 64-key chunks
      int lo = 0, hi = n;
      int q_row0, q_rows;
      if (g > 0) {  // packed: this row's own sequence
        const int unit = idx >> 3;
        const int cnt = unit == units - 1 ? nseq - unit * g : g;
        const int sidx = min(r_seq, cnt - 1);
        lo = sidx * n;
        hi = lo + n;
        q_row0 = unit * g * n;
        q_rows = cnt * n;
      } else {
        const int unit = idx >> 3;
        q_row0 = (unit >> 1) * n + (unit & 1) * 128;
        q_rows = min(128, n - (unit & 1) * 128);
      }
      // tcgen05.ld / .st are warp-collective: chunk loops must be warp-uniform, so they run over the union
      // [wlo, whi) of the key ranges of the warp's 32 rows; per-row masks inside.
      const int wlo = __reduce_min_sync(0xffffffffu, lo);
      const int whi = __reduce_max_sync(0xffffffffu, hi);
      const bool uniform = (g == 0);  // split mode: every row of the tile has the same key range
      auto live = [&](int cp) { return cp < npair && !(64 * cp + 64 <= wlo || 64 * cp >= whi); };
      auto mask = [&](uint32_t (&u)[32], int c0) {
        if (!(uniform && c0 + 32 <= hi)) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < lo || c0 + i >= hi) u[i] = 0xff800000u;  // -inf
        }
      };
      // 32 scores -> 16 packed bf16 columns of P, stored without waiting
      auto exps = [&](uint32_t (&u)[32], float msc, unsigned long long& sum2, uint32_t dst) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0 = __uint_as_float(u[i]), x1 = __uint_as_float(u[i + 1]);
          fma_f32x2(x0, x1, sl2, -msc);
          if (((i >> 1) & 7) < POLY) {
            ex2_poly2(x0, x1);
          } else {
            x0 = at_ex2(x0);  // masked: 2^-inf = 0
            x1 = at_ex2(x1);
          }
          sum2 = fadd2(sum2, pack_f32x2(x0, x1));
          pk[i >> 1] = cvt_bf16x2(x0, x1);  // column j of P_c = keys (2j, 2j+1)
        }
        tmem_st16_nowait(dst, pk);
      };
      if (tr) stamp(k, 5);
      mbar_wait(bS, ph, 30);
      fence_after();
      if (tr) stamp(k, 6);
      float m_run = -INFINITY;
      float mc[2] = {-INFINITY, -INFINITY}, lc[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int cp = 2 * half + j;
        if (cp < npair) {
          if (live(cp)) {
            // 96 registers per thread (five warps on a sub-partition) and no spills — local memory misses the
            // few KB of L1 left beside 160 KB of shared memory and costs an L2 round trip — so only ONE 32-score
            // piece is held at a time: second piece for its maximum, first piece for its maximum and its
            // exponentials, second piece again (TMEM loads are cheap with four warps per sub-partition to hide them)
            const uint32_t tA = taddr + 64 * cp, tB = tA + 32;
            uint32_t x[32];
            tmem_ld32_issue(tB, x);
            tmem_ld_wait32(x);
            mask(x, 64 * cp + 32);
            const float mB = at_max32(x);
            tmem_ld32_issue(tA, x);
            tmem_ld_wait32(x);
            mask(x, 64 * cp);
            const float m_new = fmaxf(m_run, fmaxf(at_max32(x), mB) * sl2);
            const float msc = (m_new == -INFINITY) ? 0.f : m_new;  // row has no key yet (packed mode)
            unsigned long long sum2 = pack_f32x2(0.f, 0.f);
            exps(x, msc, sum2, tA);           // P of keys [0,32) of the chunk -> columns [0,16): consumed scores
            tmem_ld32_issue(tB, x);
            tmem_ld_wait32(x);
            mask(x, 64 * cp + 32);
            exps(x, msc, sum2, tA + 16);      // keys [32,64) -> columns [16,32): scores of the first piece, consumed
            float s_lo, s_hi;
            unpack_f32x2(sum2, s_lo, s_hi);
            mc[j] = m_new;
            lc[j] = s_lo + s_hi;
            m_run = m_new;
          } else {  // no row of this warp attends these keys
            uint32_t zeros[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) zeros[i] = 0u;
            tmem_st32_nowait(taddr + 64 * cp, zeros);
          }
          tmem_st_wait();
          fence_before();
        }
        // (chunks past the item's keys are signalled too: the barrier phases must advance once per item)
        __syncwarp();
        if (lane == 0) mbar_arrive(bP + 8 * cp);
        if (tr) stamp(k, 7 + j);
      }
      // ---- the two halves of the row agree on the maximum and the sum ----
      float4* slot = xchg + (((k >> 1) & 1) * 4 + grp * 2) * 128;   // [parity][group] -> [half][row]
      slot[half * 128 + r] = make_float4(mc[0], lc[0], mc[1], lc[1]);
      named_bar_sync(1 + grp * 4 + quarter, 64);
      const float4 other = slot[(half ^ 1) * 128 + r];
      // chunk order 0..3: this warp's pair sits at [2 half, 2 half + 1] (selects, not indexed stores: registers)
      const bool h1 = half != 0;
      float m4[4], l4[4];
      m4[0] = h1 ? other.x : mc[0];
      l4[0] = h1 ? other.y : lc[0];
      m4[1] = h1 ? other.z : mc[1];
      l4[1] = h1 ? other.w : lc[1];
      m4[2] = h1 ? mc[0] : other.x;
      l4[2] = h1 ? lc[0] : other.y;
      m4[3] = h1 ? mc[1] : other.z;
      l4[3] = h1 ? lc[1] : other.w;
      const float m_all = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      float fc[4], sum = 0.f;
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        fc[cp] = (m4[cp] == -INFINITY) ? 0.f : at_ex2(m4[cp] - m_all);
        sum += fc[cp] * l4[cp];
      }
      const float inv = 1.0f / sum;
      if (tr) stamp(k, 9);
      // ---- epilogue: O = sum_c 2^(m_c - m) O_c / rowsum -> bf16; this warp converts output columns [16 half, +16) ----
      mbar_wait(bO, ph, 31);
      fence_after();
      if (tr) stamp(k, 11);
      {
        // The four O_c loads together (one round trip), UNCONDITIONALLY: loads guarded by `c < npair` make the
        // compiler merge "maybe loaded" register arrays through local memory.  Chunks past the item's keys hold
        // stale columns; their weight is forced to zero by a select, not by a multiplication (0 * NaN).
        uint32_t o0[16], o1[16], o2[16], o3[16];
        const uint32_t tO = taddr + 32 + 16 * half;
        tmem_ld16_issue(tO, o0);
        tmem_ld16_issue(tO + 64, o1);
        tmem_ld16_issue(tO + 128, o2);
        tmem_ld16_issue(tO + 192, o3);
        tmem_ld_wait16(o0);
        tmem_ld_wait16(o1);
        tmem_ld_wait16(o2);
        tmem_ld_wait16(o3);
        fence_before();  // the O_c reads precede the next S overwriting the buffer
        __syncwarp();
        if (lane == 0) mbar_arrive(bF);  // release TMEM before the arithmetic and the global stores
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float v1 = 1 < npair ? __uint_as_float(o1[i]) : 0.f;
          const float v2 = 2 < npair ? __uint_as_float(o2[i]) : 0.f;
          const float v3 = 3 < npair ? __uint_as_float(o3[i]) : 0.f;
          float acc = fc[0] * __uint_as_float(o0[i]);
          acc = fmaf(fc[1], v1, acc);
          acc = fmaf(fc[2], v2, acc);
          acc = fmaf(fc[3], v3, acc);
          o[i] = acc * inv;
        }
        if (r < q_rows) {
          uint4* dsto = reinterpret_cast<uint4*>(out + (size_t)(q_row0 + r) * kN + (idx & 7) * kDh + 16 * half);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint4 u;
            u.x = cvt_bf16x2(o[8 * i], o[8 * i + 1]);
            u.y = cvt_bf16x2(o[8 * i + 2], o[8 * i + 3]);
            u.z = cvt_bf16x2(o[8 * i + 4], o[8 * i + 5]);
            u.w = cvt_bf16x2(o[8 * i + 6], o[8 * i + 7]);
            dsto[i] = u;
          }
        }
      }
      if (tr) stamp(k, 12);
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int at_tensor_map(const void* qkv, uint64_t rows, CUtensorMap* out) {
  static AtEncodeFn fn = nullptr;
  static std::mutex mu;
  struct Key {
    const void* ptr;
    uint64_t rows;
    bool operator==(const Key& o) const { return ptr == o.ptr && rows == o.rows; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const { return std::hash<const void*>()(k.ptr) ^ (std::hash<uint64_t>()(k.rows) * 1000003ull); }
  };
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  std::lock_guard<std::mutex> g(mu);
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      set_error("attention_tc: cuTensorMapEncodeTiled is unavailable");
      return 1;
    }
    fn = reinterpret_cast<AtEncodeFn>(p);
  }
  const Key key{qkv, rows};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)(3 * kN), rows};
  cuuint64_t gstride[1] = {(cuuint64_t)(3 * kN) * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)kDh, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 1;
  }
  if (cache.size() > 1024) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// Default: the v1 kernel — still the fastest measured (148 us at 544 x 251 against 174 us for v4, whose groups
// fall into lock-step: both in their MUFU phases, then both in their MUFU-free phases; profiles/r02_experiments.md).
// CSE_ATTN_VER=4 selects v4 for A/B runs.
static int attention_version() {
  static const int ver = []() {
    const char* e = getenv("CSE_ATTN_VER");
    return (e != nullptr && atoi(e) == 4) ? 4 : 1;
  }();
  return ver;
}

long long* g_attention_trace = nullptr;  // cse_debug_attention_trace: device buffer [64 items][16 slots] of clock64 stamps

int launch_attention_tc(const bf16* qkv, int nseq, int n, bf16* out, cudaStream_t st) {
  if (n < 1 || n > 256) {
    set_error("attention_tc: n=%d outside [1,256]", n);
    return 1;
  }
  // cse_debug_force_mma_attention: 2 = v1, 4 = v4; otherwise CSE_ATTN_VER
  const int ver = g_attention_mode == 2 ? 1 : g_attention_mode == 4 ? 4 : attention_version();
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kAtSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_tc4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kA2Smem);

    if (e != cudaSuccess) {
      set_error("attention_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return 1;
    }
    once.mark_configured();
  }
  const int sms = sm_count();
  CUtensorMap tm;
  if (at_tensor_map(qkv, (uint64_t)nseq * n, &tm)) return 1;
  const int g = n <= 128 ? 128 / n : 0;
  const long long units = g > 0 ? ((long long)nseq + g - 1) / g : (long long)nseq * 2;
  const long long items = units * kHeads;
  if (items > 2147483647LL) {
    set_error("attention_tc: too many work items");
    return 1;
  }
  const int grid = (int)(items < sms ? items : sms);
  if (ver == 4) {
    launch_pdl(attention_tc4_kernel<0>, dim3(grid), dim3(kA4Threads), kA2Smem, st, 1, tm, out, n, nseq, g, (int)items,
               g_attention_trace);
    return check_launch("attention_tc4_kernel<0>");
  }
  launch_pdl(attention_tc_kernel, dim3(grid), dim3(kAtThreads), kAtSmem, st, 1, tm, out, n, nseq, g, (int)items);
  return check_launch("attention_tc_kernel");
}

}  // namespace cse
