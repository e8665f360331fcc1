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
// v3.  What bounded v1 (and a first re-plumbing, "v2", that kept its arithmetic) was not MUFU and not the issue
// order but the tensor pipe itself: a tcgen05.mma that OVERWRITES its accumulator (scale-d = 0) costs ~775 cycles
// at M = 128 whatever N is, against ~105 for one that accumulates (tools/micro/umma_pv_rate.cu, B200: groups of
// "1 overwrite + 3 accumulate" average 273 cycles per instruction, pure accumulate chains 105).  The online softmax
// gave every 64-key chunk its own O_c accumulator, i.e. 4 overwriting instructions + 1 for S per item: ~4.4 k of
// the ~4.8 k cycles an item took.  v3 therefore
//   * takes the row maximum over the WHOLE score row first (a second tcgen05.ld pass over S is ~150 cycles per
//     item), so every P_c is exponentiated against the final maximum and all sixteen P_c V_c instructions
//     ACCUMULATE into one O tile that the softmax warps zero-fill with tcgen05.st;
//   * (ZERO_S) lets the softmax warps zero-fill the score columns as well, so that Q K^T accumulates too;
//   * keeps v2's plumbing: ONE issuing warp polls (mbarrier.test_wait) every pending piece of tensor work of both
//     in-flight items and issues whichever is ready, S of the next item first; P never touches shared memory (bf16
//     pairs go back into the item's own consumed score columns and feed P_c V_c as the TMEM A operand), and the
//     128 KB of P staging v1 needed are a 4-deep Q/K/V ring.
// TMEM columns of buffer b (256 per item): chunk c (64 keys) = [64c, 64c+64): scores -> P_c in [64c, 64c+32)
// (two bf16 per column); O (fp32 [128 x 32]) in [32, 64), the dead upper half of chunk 0.
// ------------------------------------------------------------------------------------------
constexpr int kA2Threads = 320;  // warp 0 TMA, warp 1 S/PV issuer, warps 2-5 softmax group 0, warps 6-9 group 1
constexpr int kA2Stages = 4;
constexpr int kA2StageBytes = kAtQ + 2 * kAtKV;  // 40 KB: Q tile + K + V
constexpr size_t kA2Smem = 1024 + (size_t)kA2Stages * kA2StageBytes + 512;
constexpr int kA2TraceSlots = 16, kA2TraceItems = 64;

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

template <bool ZERO_S>
__global__ void __launch_bounds__(kA2Threads, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, int n, int nseq,
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
      mbar_init(bar_b(FREE, b), 4);
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
    // ================= TMA producers: lane 0 streams Q/K, lane 1 streams V, up to kA2Stages items ahead ==========
    if (lane < 2) {
      for (int k = 0; k < nitems; ++k) {
        const int st = k & (kA2Stages - 1);
        const uint32_t ph = (uint32_t)(k / kA2Stages) & 1u;
        const AtItem it = at_decode(items - 1 - ((int)blockIdx.x + k * (int)gridDim.x), n, nseq, g);
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
  } else if (warp == 1) {
    // ================= the one tensor-pipe issuer: S = Q K^T and O += P_c V_c of both in-flight items ==========
    if (lane == 0) {
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 32, 0, 1);  // A = P from TMEM, B = V MN-major
      int s_k[2] = {0, 1};   // next item (CTA-local index) whose S goes to TMEM buffer b
      int p_k[2] = {0, 1};   // item whose PV chunks are being issued on buffer b
      int p_cp[2] = {0, 0};  // its next chunk
      int done = 0;
      long long t_idle = clock64();
      while (done < nitems) {
        bool progress = false;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          // ---- S of the next item of this buffer (first: it unblocks a whole softmax group) ----
          int k = s_k[b];
          if (k < nitems) {
            const int st = k & (kA2Stages - 1);
            const uint32_t ph_st = (uint32_t)(k / kA2Stages) & 1u, ph_b = (uint32_t)(k >> 1) & 1u;
            // FREE completes once per item (its O has been read out) — and, with ZERO_S, once more at kernel start
            // (the group's first zero-fill), so use j of the buffer waits for completion j instead of j - 1
            const bool buf_free = ZERO_S ? mbar_test(bar_b(FREE, b), ph_b) : (k < 2 || mbar_test(bar_b(FREE, b), ph_b ^ 1u));
            if (buf_free && mbar_test(bar_st(QK_FULL, st), ph_st)) {
              fence_after();
              const AtItem it = at_decode(items - 1 - ((int)blockIdx.x + k * (int)gridDim.x), n, nseq, g);
              const int ncols = (it.kv_rows + 15) & ~15;
              const uint32_t idesc_s = make_idesc_bf16(128, ncols, 0, 0);
              const uint64_t ad = make_desc(sQ(st), 512, kLayoutSw64);
              const uint64_t bd = make_desc(sK(st), 512, kLayoutSw64);
              const uint32_t d_s = tmem_base + (uint32_t)b * 256;
              umma_bf16(d_s, ad, bd, idesc_s, ZERO_S ? 1u : 0u);
              umma_bf16(d_s, ad + 2, bd + 2, idesc_s, 1u);  // second k16 step: +32 B inside the 64-B row
              umma_commit(bar_b(S_FULL, b));
              umma_commit(bar_st(QK_EMPTY, st));
              stamp(k, 0);
              s_k[b] = k + 2;
              progress = true;
            }
          }
          // ---- next P_c V_c chunk of the item in flight on this buffer ----
          k = p_k[b];
          if (k < nitems && k < s_k[b]) {
            const int st = k & (kA2Stages - 1);
            const uint32_t ph_st = (uint32_t)(k / kA2Stages) & 1u, ph_b = (uint32_t)(k >> 1) & 1u;
            const int cp = p_cp[b];
            if ((cp > 0 || mbar_test(bar_st(V_FULL, st), ph_st)) && mbar_test(bar_p(b, cp), ph_b)) {
              fence_after();
              const AtItem it = at_decode(items - 1 - ((int)blockIdx.x + k * (int)gridDim.x), n, nseq, g);
              const int ncols = (it.kv_rows + 15) & ~15;
              const int npair = (ncols + 63) >> 6;
              const uint32_t t_buf = tmem_base + (uint32_t)b * 256;
              const int t1 = (cp < npair) ? min(4 * cp + 4, ncols / 16) : 0;
              for (int t = 4 * cp; t < t1; ++t) {
                const uint64_t bd = make_desc(sV(st) + (uint32_t)t * 1024, 512, kLayoutSw64);
                // every instruction accumulates: the softmax warps zero-filled O before signalling chunk 0
                umma_bf16_ts(t_buf + 32, t_buf + (uint32_t)(cp * 64 + 8 * (t & 3)), bd, idesc_pv, 1u);
              }
              stamp(k, 1 + cp);
              if (cp == 3) {
                umma_commit(bar_b(O_FULL, b));
                umma_commit(bar_st(V_EMPTY, st));
                p_k[b] = k + 2;
                p_cp[b] = 0;
                ++done;
              } else {
                p_cp[b] = cp + 1;
              }
              progress = true;
            }
          }
        }
        if (progress) {
          t_idle = clock64();
        } else if (clock64() - t_idle > 4000000000LL) {
          printf("attention_tc3: issuer stalled (block %d, S items %d/%d, PV items %d/%d chunks %d/%d, done %d of %d)\n",
                 (int)blockIdx.x, s_k[0], s_k[1], p_k[0], p_k[1], p_cp[0], p_cp[1], done, nitems);
          __trap();
        }
      }
    }
    __syncwarp();
  } else {
    // ================= softmax + epilogue groups (128 threads each, one query row per thread) =================
    const int grp = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const float sl2 = 0.17677669529663687f * 1.4426950408889634f;  // log2(e)/sqrt(32)
    const bool tr = tracing && quarter == 0 && lane == 0;
    const int b = grp;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)b * 256;
    uint32_t zeros[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) zeros[i] = 0u;
    if (ZERO_S && grp < nitems) {  // this warp's lanes of the whole buffer, before its first S accumulates into them
#pragma unroll
      for (int c = 0; c < 8; ++c) tmem_st32(taddr + 32 * c, zeros);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_b(FREE, b));
    }
    for (int k = grp; k < nitems; k += 2) {
      const uint32_t ph = (uint32_t)(k >> 1) & 1u;
      const AtItem it = at_decode(items - 1 - ((int)blockIdx.x + k * (int)gridDim.x), n, nseq, g);
      const int ncols = (it.kv_rows + 15) & ~15;
      const int npair = (ncols + 63) >> 6;  // 64-key chunks
      int lo = 0, hi = it.kv_rows;
      if (g > 0) {  // packed: this row's own sequence
        const int sidx = min(r / n, it.q_rows / n - 1);
        lo = sidx * n;
        hi = lo + n;
      }
      // tcgen05.ld / .st are warp-collective: chunk loops must be warp-uniform, so they run over the union
      // [wlo, whi) of the key ranges of the warp's 32 rows; per-row masks inside.
      const int wlo = __reduce_min_sync(0xffffffffu, lo);
      const int whi = __reduce_max_sync(0xffffffffu, hi);
      const bool uniform = (g == 0);  // split mode: every row of the tile has the same key range
      if (tr) stamp(k, 5);
      mbar_wait(bar_b(S_FULL, b), ph, 30);
      fence_after();
      if (tr) stamp(k, 6);
      // ---- pass 1: the row maximum over the whole score row ----
      float m_row = -INFINITY;
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        const int c0 = cp * 64;
        if (cp < npair && !(c0 + 64 <= wlo || c0 >= whi)) {
          float v[64];
          tmem_ld64(taddr + c0, v);
          if (!(uniform && c0 + 64 <= hi)) {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (c0 + i < lo || c0 + i >= hi) v[i] = -INFINITY;
          }
          m_row = fmaxf(m_row, at_max64(v));
        }
      }
      const float msc = m_row * sl2;  // every row has at least one key of its own
      if (tr) stamp(k, 7);
      // ---- pass 2: P_c = 2^(s * sl2 - msc) per 64-key chunk, bf16 pairs back into the chunk's own columns ----
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        if (cp < npair) {
          const int c0 = cp * 64;
          uint32_t pk[32];
          if (c0 + 64 <= wlo || c0 >= whi) {  // warp-uniform: no row of this warp attends these keys
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = 0u;
          } else {
            float v[64];
            tmem_ld64(taddr + c0, v);
            if (!(uniform && c0 + 64 <= hi)) {
#pragma unroll
              for (int i = 0; i < 64; ++i)
                if (c0 + i < lo || c0 + i >= hi) v[i] = -INFINITY;
            }
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
              pk[i >> 1] = cvt_bf16x2(v[i], v[i + 1]);          // column j of P_c = keys (2j, 2j+1)
              pk[(i >> 1) + 1] = cvt_bf16x2(v[i + 2], v[i + 3]);
            }
          }
          tmem_st32(taddr + c0, pk);                    // P_c over the first half of the chunk's own score columns
          if (cp == 0) tmem_st32(taddr + 32, zeros);    // O = 0 in the (consumed) second half of chunk 0
          fence_before();
        }
        // (chunks past the item's keys are signalled too: the barrier phases must advance once per item)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p(b, cp));
        if (tr) stamp(k, 8 + (cp < 3 ? cp : 2));
      }
      const float sum = (s0 + s1) + (s2 + s3);
      // ---- epilogue: O / rowsum -> bf16 ----
      mbar_wait(bar_b(O_FULL, b), ph, 31);
      fence_after();
      if (tr) stamp(k, 11);
      {
        float o[32];
        tmem_ld32(taddr + 32, o);
        if (ZERO_S) {  // score columns back to zero for the buffer's next Q K^T (this warp's lanes only)
          const int nz = (ncols + 31) >> 5;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < nz) tmem_st32(taddr + 32 * c, zeros);
        }
        fence_before();  // O read (and the zero-fill) precede the next S
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_b(FREE, b));  // release TMEM before the global stores
        const float inv = 1.0f / sum;
        if (r < it.q_rows) {
          uint4* dsto = reinterpret_cast<uint4*>(out + (size_t)(it.q_row0 + r) * kN + it.head * kDh);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = cvt_bf16x2(o[8 * i] * inv, o[8 * i + 1] * inv);
            u.y = cvt_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
            u.z = cvt_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
            u.w = cvt_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
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

// CSE_ATTN_VER = 1 keeps the v1 kernel (two issuer warps, P through shared memory, one accumulator per chunk),
// 3 = v3, 4 = v3 with zero-filled score columns (default; see the v3 header) — for A/B runs.
static int attention_version() {
  static const int ver = []() {
    const char* e = getenv("CSE_ATTN_VER");
    const int v = e != nullptr ? atoi(e) : 0;
    return (v == 1 || v == 3 || v == 4) ? v : 4;
  }();
  return ver;
}

long long* g_attention_trace = nullptr;  // cse_debug_attention_trace: device buffer [64 items][16 slots] of clock64 stamps

int launch_attention_tc(const bf16* qkv, int nseq, int n, bf16* out, cudaStream_t st) {
  if (n < 1 || n > 256) {
    set_error("attention_tc: n=%d outside [1,256]", n);
    return 1;
  }
  // cse_debug_force_mma_attention: 2 = v1, 3 = v3, 4 = v3 + zero-filled scores; otherwise CSE_ATTN_VER
  const int ver = g_attention_mode == 2 ? 1 : (g_attention_mode == 3 || g_attention_mode == 4) ? g_attention_mode
                                                                                                : attention_version();
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kAtSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kA2Smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kA2Smem);
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
    launch_pdl(attention_tc3_kernel<true>, dim3(grid), dim3(kA2Threads), kA2Smem, st, 1, tm, out, n, nseq, g,
               (int)items, g_attention_trace);
    return check_launch("attention_tc3_kernel<1>");
  }
  if (ver == 3) {
    launch_pdl(attention_tc3_kernel<false>, dim3(grid), dim3(kA2Threads), kA2Smem, st, 1, tm, out, n, nseq, g,
               (int)items, g_attention_trace);
    return check_launch("attention_tc3_kernel<0>");
  }
  launch_pdl(attention_tc_kernel, dim3(grid), dim3(kAtThreads), kAtSmem, st, 1, tm, out, n, nseq, g, (int)items);
  return check_launch("attention_tc_kernel");
}

}  // namespace cse
