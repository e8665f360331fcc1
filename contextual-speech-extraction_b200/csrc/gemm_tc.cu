// gemm_tc.cu — bf16 tensor-core GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// C[M,N] = A[M,K] * W[N,K]^T (+ bias_scale*bias)(relu)(+ residual): every nn.Linear / 1x1 conv of
// the path in CSE_BF16 mode (QKV, out-proj, FFN1/2: CSE_transformer.py:335-340,547;
// masknet.conv1d/conv2d/output/output_gate/end_conv1x1: ContSep.py:229,247,255,258).
// A is row-major activations, W is the nn.Linear weight [out,in]: BOTH are K-major, which is the
// native UMMA operand layout, so no transposes exist anywhere.
//
// Structure (persistent, warp-specialised, one CTA per SM, clusters of two CTAs on adjacent row tiles):
//   warp 0      : TMA producer — 128x64 A tile + HALF of the BNx64 W tile per stage (the other half is
//                 multicast in by the peer CTA), SWIZZLE_128B, mbarrier complete_tx
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32
//                 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the
//                 MMAs of tile i+1); tcgen05.commit releases smem stages / publishes accumulators
//   warps 2..9  : epilogue — tcgen05.ld (32 lanes x 64 columns) -> bias (FFMA2, prefetched one chunk
//                 ahead) / ReLU fused into the bf16 conversion -> 128B-swizzled shared-memory chunk ->
//                 TMA store.  The fp32 residual update
//                 R += A W^T + b is a TMA REDUCE-ADD (cp.reduce.async.bulk.tensor .add), so the
//                 residual stream is never loaded into the SM; the M tail is clipped by TMA.
// (Two round-1 variants — one cta_group::2 UMMA per CTA pair, and an A tile resident across n-tiles — measured slower,
// profiles/r01_experiments.md, and were removed in round 2.)  CSE_DBG_* macros strip parts of the kernel for
// tools/gemm_variants.py.
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.  The kernel is launched
// with programmatic dependent launch: its prologue overlaps the previous kernel's tail (pdl_wait()).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

constexpr int kTM = 128;  // UMMA M
constexpr int kTK = 64;   // K per stage: 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
#ifndef CSE_EPI_WARPS
#define CSE_EPI_WARPS 8
#endif
#ifndef CSE_STAGES_256
#define CSE_STAGES_256 4
#endif
#ifndef CSE_EPI_ROW_BYTES
#define CSE_EPI_ROW_BYTES 128  // bytes per row of one epilogue staging chunk: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
#endif

// ------------------------------------------------------------------------------------------
// PTX wrappers: tc_ptx.cuh (shared with ffn_tc.cu / attention_tc.cu / gemm_ln_tc.cu); only what is
// specific to this kernel is defined here.
// ------------------------------------------------------------------------------------------
using namespace tc;

__device__ __forceinline__ void tc_fence_before() { fence_before(); }
__device__ __forceinline__ void tc_fence_after() { fence_after(); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// K-major SWIZZLE_128B operand: 1024 B between 8-row groups (tc::make_desc)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  return make_desc(smem_addr, 1024, kLayoutSw128);
}
// MN-major SWIZZLE_128B operand: atoms of 8 k-rows x 64 elements (128 B); SBO = 1024 B between 8-k groups, LBO =
// 8192 B between 64-element blocks along M / N (one [64 k x 64 mn] TMA box each)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)kLayoutSw128 << 61;
  return d;
}
// kind::f16, A/B = bf16 K-major, D = fp32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) { return tc::make_idesc_bf16(M, N, 0, 0); }

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
template <int BN, bool LNE = false>
struct TcCfg {
  // LNE (K = 256: four k-blocks per tile, an epilogue-bound kernel): three ring stages are enough, and the 48 KB
  // they free hold the bias / gamma / beta vectors (read from global memory per piece they cost an L2 round trip
  // each: the L1 left beside 225 KB of shared memory does not keep them)
  static constexpr int kStages = LNE ? 3 : ((BN == 256) ? CSE_STAGES_256 : 6);
  static constexpr int kABytes = kTM * kTK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kTK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;   // bytes one k-block brings in (A + W)
  static constexpr int kTmemCols = 2 * BN;       // double-buffered accumulator
  static constexpr int kEpiWarps = (BN == 256) ? CSE_EPI_WARPS : 8;  // 4 TMEM lane quarters x column groups
  static constexpr int kThreads = 64 + 32 * kEpiWarps;              // producer + MMA + epilogue warps
  // 4 KB of staging per epilogue warp, cut into chunks of 32 rows x kRowBytes.  With 64-byte rows a warp
  // has TWO chunks, so the TMA store of one chunk drains while the next is being converted (a 4 KB
  // store takes ~0.8 us to release its source when the write path is busy, as long as the conversion).
  static constexpr int kRowBytes = CSE_EPI_ROW_BYTES;
  static constexpr int kEpiBufs = 128 / kRowBytes;
  static constexpr int kEpiBufBytes = 32 * kRowBytes;
  static constexpr int kEpiBytes = kEpiWarps * kEpiBufs * kEpiBufBytes;
  static constexpr int kRingBytes = kStages * (kABytes + kBBytes);
  static constexpr int kVecBytes = LNE ? 3 * 256 * 4 : 0;   // bias, gamma, beta
  static constexpr size_t kSmem = 1024 /*align slack*/ + (size_t)kRingBytes + kEpiBytes + 384 + kVecBytes;
};

// LNE (BN = 256 = N, fp32 residual stream): the epilogue is `R += A W^T + b; H = LayerNorm(R)` — the attention
// output projection, its residual add and the pre-FFN LayerNorm of a transformer layer in one kernel
// (CSE_transformer.py:399-408).  Unfused, LayerNorm re-reads the 140 MB residual stream the projection has just
// written; here the row is normalised while it is on chip.
struct LnEpilogue {
  float* R;            // [M,256] fp32 residual stream, updated in place
  const float* gamma;  // LayerNorm weight / bias
  const float* beta;
  bf16* H;             // [M,256] bf16 LayerNorm(R_new)
  float eps;
};

// MNM = 2 (dgrad): only the B operand is MN-major — C[M,N] = A[M,K] W[K,N] for the row-major weight W as stored
// ([out, in] = [K, N]: dA = dC W needs no transposed weight copy); A stays K-major.
// MNM = 1 (wgrad): BOTH operands are "MN-major" — C[M,N] += X^T Y for row-major X [T,M], Y [T,N] contracted over their
// slow dimension T (tokens).  TMA brings [64 tokens x 64 features] SWIZZLE_128B boxes (64 lines of 128 B: exactly
// the canonical MN-major atom stack, 8 token-rows per atom), two per 128 output rows / four per 256 output columns;
// descriptors: SBO = 1024 B between 8-token groups, LBO = 8192 B between 64-feature blocks; a K = 16 step advances
// two atoms (2048 B).  No transposed copies of the operands exist (they were 6 ms of a 28.6 ms training step).
template <int BN, bool OUT_F32, bool LNE = false, int MNM = 0>
__global__ void __launch_bounds__((TcCfg<BN, LNE>::kThreads), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias,
               float bias_scale, int accumulate_into_c, void* Cout, int ldc, int M, int N, int K,
               int relu, int ksplit, LnEpilogue ln) {
  using Cfg = TcCfg<BN, LNE>;
  static_assert(!LNE || (BN == 256 && OUT_F32 && Cfg::kEpiWarps == 8),
                "the LayerNorm epilogue is built on the 256-wide fp32 per-CTA variant");
  static_assert(MNM == 0 || !LNE, "MN-major operands and the LayerNorm epilogue are separate variants");
  pdl_launch_dependents();
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SWIZZLE_128B: 1024-B aligned
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + Cfg::kStages * Cfg::kABytes;
  const uint32_t sEpi = smem_base + Cfg::kRingBytes;  // multiple of 1024
  const uint32_t sBar = sEpi + Cfg::kEpiBytes;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2]; then the TMEM base slot
  const uint32_t bar_full = sBar;
  const uint32_t bar_empty = sBar + 8 * Cfg::kStages;
  const uint32_t bar_tfull = sBar + 16 * Cfg::kStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t sVec = sBar + 384;  // LNE: bias[256], gamma[256], beta[256]
  unsigned char* smem_aligned = smem_dyn + (smem_base - smem_u32(smem_dyn));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_aligned + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kTM - 1) / kTM, n_tiles = N / BN;
  const int num_kb = K / kTK;
  // Split-K (ksplit > 1; wgrad: few output tiles, K = the row count): a tile job covers k-blocks
  // [ks * kb_per, min(num_kb, (ks + 1) * kb_per)) and the jobs of one tile meet in C through the reduce-add
  // epilogue.  The host picks ksplit so that no job is empty and only with accumulate_into_c and no bias.
  const int kb_per = (num_kb + ksplit - 1) / ksplit;
  // CTA pairs (clusters of 2): both CTAs of a pair walk the same n-tile sequence on adjacent
  // m-tiles, so every weight tile is fetched from L2 ONCE per pair — each CTA loads half of it and
  // TMA-multicasts that half into both CTAs' shared memory.  (These GEMMs were bound by L2->SM
  // bandwidth, ~9 TB/s aggregate, because the weights are re-streamed for every 128-row tile.)
  const int cta_rank = (int)cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_pt = ((m_tiles + 1) >> 1) * n_tiles * ksplit;
  // Tile walk: pair p takes tiles p, p + npairs, ... (n fastest, so the pairs that share an A tile run at the same time)
  const int t_begin = pair_id, t_end = total_pt, t_step = npairs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 2);  // released by the MMA issuers of BOTH CTAs
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, Cfg::kEpiWarps);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  if constexpr (LNE) {  // parameters, not produced by the previous kernel: may be read before pdl_wait()
    float* vec = reinterpret_cast<float*>(smem_aligned + (sVec - smem_base));
    for (int i = threadIdx.x; i < 256; i += Cfg::kThreads) {
      vec[i] = bias[i];
      vec[256 + i] = ln.gamma[i];
      vec[512 + i] = ln.beta[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // prologue done; from here on the previous kernel's output is read / this kernel's output written

  if (warp == 0) {
    // ================= TMA producer =================
    // The whole warp walks the loop (keeps it convergent for the teardown barrier); lane 0 acts.
    uint32_t stage = 0, phase = 0;
    for (int pt = t_begin; pt < t_end; pt += t_step) {
      const int tl = pt / ksplit, ks = pt - tl * ksplit;
      const int m0 = (2 * (tl / n_tiles) + cta_rank) * kTM, n0 = (tl % n_tiles) * BN;
      const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      if constexpr (LNE) {
        // the residual rows this tile's epilogue will read, one tile of lead: HBM -> L2 now, so that the epilogue's
        // loads are L2 hits (its eight warps cannot keep enough HBM requests in flight themselves)
        const int rows_here = min(kTM, M - m0);
        const char* rbase = reinterpret_cast<const char*>(ln.R + (size_t)m0 * kN);
        for (int i = lane; i < rows_here * (kN * 4 / 128); i += 32)
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(rbase + (size_t)i * 128));
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        if constexpr (MNM != 0) {
          if (lane == 0) {
            mbar_wait_spin(bar_empty + 8 * stage, phase ^ 1, 1);   // both CTAs have consumed this stage
            mbar_expect_tx(bar_full + 8 * stage, Cfg::kStageBytes);
            // MN-major boxes [64 k x 64 m/n]: coordinates (m/n, k)
            if constexpr (MNM == 1) {
#pragma unroll
              for (int i = 0; i < kTM / 64; ++i)
                tma_load_2d(sA + stage * Cfg::kABytes + i * 8192, &tmA, bar_full + 8 * stage, m0 + 64 * i, kb * kTK);
            } else {
              tma_load_2d(sA + stage * Cfg::kABytes, &tmA, bar_full + 8 * stage, kb * kTK, m0);
            }
#pragma unroll
            for (int i = 0; i < BN / 128; ++i) {  // this CTA's half of the 64-feature blocks -> both CTAs
              const int blk = cta_rank * (BN / 128) + i;
              tma_load_2d_mcast(sB + stage * Cfg::kBBytes + blk * 8192, &tmB, bar_full + 8 * stage, n0 + 64 * blk,
                                kb * kTK, (uint16_t)3);
            }
          }
        } else if (lane == 0) {
          mbar_wait_spin(bar_empty + 8 * stage, phase ^ 1, 1);   // both CTAs have consumed this stage
          mbar_expect_tx(bar_full + 8 * stage, Cfg::kStageBytes);
          tma_load_2d(sA + stage * Cfg::kABytes, &tmA, bar_full + 8 * stage, kb * kTK, m0);
          // this CTA's half of the weight tile -> both CTAs (rows past the M tail read as zeros)
          tma_load_2d_mcast(sB + stage * Cfg::kBBytes + cta_rank * (Cfg::kBBytes / 2), &tmB,
                            bar_full + 8 * stage, kb * kTK, n0 + cta_rank * (BN / 2), (uint16_t)3);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (lane 0 issues; warp stays convergent) =================
    constexpr uint32_t idesc = MNM != 0 ? tc::make_idesc_bf16(kTM, BN, MNM == 1 ? 1 : 0, 1) : make_idesc_bf16(kTM, BN);
    uint32_t stage = 0, phase = 0, astage = 0, aphase = 0;
    for (int pt = t_begin; pt < t_end; pt += t_step) {
      if (lane == 0) {
        mbar_wait_spin(bar_tempty + 8 * astage, aphase ^ 1, 2);
        tc_fence_after();
      }
      __syncwarp();
      const uint32_t d_tmem = tmem_base + astage * BN;
      const int ks = pt % ksplit;
      const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        if (lane == 0) {
          mbar_wait_spin(bar_full + 8 * stage, phase, 3);
          tc_fence_after();
          const uint64_t adesc = MNM == 1 ? make_mnmajor_sw128_desc(sA + stage * Cfg::kABytes)
                                          : make_kmajor_sw128_desc(sA + stage * Cfg::kABytes);
          const uint64_t bdesc = MNM != 0 ? make_mnmajor_sw128_desc(sB + stage * Cfg::kBBytes)
                                          : make_kmajor_sw128_desc(sB + stage * Cfg::kBBytes);
          if constexpr (MNM != 0) {
            // a K = 16 step: MN-major = two 8-k atoms = 2048 B (+128 in the addr>>4 field); K-major = 32 B (+2)
            constexpr int kAStep = MNM == 1 ? 128 : 2;
#pragma unroll
            for (int k = 0; k < kTK / kUmmaK; ++k)
              umma_bf16(d_tmem, adesc + kAStep * k, bdesc + 128 * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          } else
#pragma unroll
          for (int k = 0; k < kTK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in the (addr>>4) field
#ifndef CSE_DBG_NOMMA
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
#endif
          }
          // release the stage in both CTAs
          umma_commit_mcast(bar_empty + 8 * stage, (uint16_t)3);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) umma_commit(bar_tfull + 8 * astage);  // accumulator complete -> epilogue
      __syncwarp();
      if (++astage == 2) { astage = 0; aphase ^= 1; }
    }
  } else {
    // ================= epilogue warps 2..9 =================
    // warp -> TMEM lane quarter (warp & 3, fixed by hardware) x column half ((warp-2) >> 2).
    // Per chunk: TMEM -> registers (one output row per thread) -> bias/ReLU/pack -> warp-private
    // SWIZZLE_128B staging tile -> one elected lane issues a TMA store (or reduce-add).  Two
    // staging buffers per warp: the store of chunk i overlaps the TMEM load of chunk i+1.
    constexpr int CW = Cfg::kRowBytes / (OUT_F32 ? 4 : 2);  // columns per chunk
    constexpr int GW = BN / (Cfg::kEpiWarps / 4);  // columns per warp
    constexpr int NCH = GW / CW;
    const int quarter = warp & 3;
    const int cgroup = (warp - 2) >> 2;
    const uint32_t stg0 = sEpi + (warp - 2) * Cfg::kEpiBufs * Cfg::kEpiBufBytes;
    unsigned char* stg0_ptr = smem_aligned + (stg0 - smem_base);
    uint32_t astage = 0, aphase = 0, chunk_ctr = 0;
    float bv[CW];  // bias of the upcoming chunk (lane-uniform), prefetched one chunk ahead
    auto load_bias = [&](int col) {
#pragma unroll
      for (int i = 0; i < CW; i += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col + i));
        bv[i] = b4.x;  // bias_scale is applied in the packed FMA below
        bv[i + 1] = b4.y;
        bv[i + 2] = b4.z;
        bv[i + 3] = b4.w;
      }
    };
    if (bias != nullptr && t_begin < t_end) load_bias((t_begin % n_tiles) * BN + cgroup * GW);  // (bias: ksplit == 1)
    for (int pt = t_begin; pt < t_end; pt += t_step) {
      const int tl = pt / ksplit;
      const int m0 = (2 * (tl / n_tiles) + cta_rank) * kTM, n0 = (tl % n_tiles) * BN;
      const int row_base = m0 + quarter * 32;
      float4 rq[LNE ? 8 : 1];  // LNE: the next 32 x 32 piece of the residual stream, in the coalesced walk's layout
      if constexpr (LNE) {     // (does not depend on the accumulator: requested before waiting for it)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row_base + 4 * i + (lane >> 3);
          rq[i] = r < M ? *reinterpret_cast<const float4*>(ln.R + (size_t)r * kN + cgroup * 128 + 4 * (lane & 7))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(bar_tfull + 8 * astage, aphase, 4);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + astage * BN;
      if constexpr (LNE) {
        // One output row per thread pair: this thread holds columns [128 cgroup, +128) of row `row` (TMEM lane =
        // row).  Global memory is touched row-contiguously instead — a 32-row x 32-column piece goes through the
        // warp's 4 KB staging tile both ways (a first version with each lane walking its own row issued 32
        // separate lines per load instruction and was 40 us slower than the two kernels it replaces):
        //   pass 1: R piece -> staging -> x = acc + b + R -> staging -> R;  x parked in the accumulator's own TMEM
        //           columns; half-row sums of (x - shift) and (x - shift)^2 -> half-row mean and M2; the two halves
        //           of a row meet through shared memory and combine (Chan)
        //   pass 2: x from TMEM -> (x - mean) rstd gamma + beta -> bf16 -> staging -> H
        const float* s_vec = reinterpret_cast<const float*>(smem_aligned + (sVec - smem_base));
        float4* stg4 = reinterpret_cast<float4*>(stg0_ptr);          // [32 rows][8 x 16 B], slot ^ (row & 7)
        const int cr = lane >> 3, cc = lane & 7;                       // coalesced walk: 4 rows x 128 B per step
        float s1 = 0.f, s2 = 0.f, shift = 0.f;  // sums of (x - shift), (x - shift)^2; shift = the half-row's first x
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          const int col = cgroup * 128 + ch * 32;
          float v[32];
          tmem_ld32(t_row + col, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) stg4[(4 * i + cr) * 8 + (cc ^ ((4 * i + cr) & 7))] = rq[i];
          if (ch < 3) {  // the next piece's loads fly while this one is processed
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = row_base + 4 * i + cr;
              rq[i] = r < M ? *reinterpret_cast<const float4*>(ln.R + (size_t)r * kN + col + 32 + 4 * cc)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          __syncwarp();
          uint32_t xb[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r4 = stg4[lane * 8 + (j ^ (lane & 7))];
            const float4 b4 = *reinterpret_cast<const float4*>(s_vec + col + 4 * j);
            float4 x;
            x.x = (v[4 * j] + b4.x) + r4.x;
            x.y = (v[4 * j + 1] + b4.y) + r4.y;
            x.z = (v[4 * j + 2] + b4.z) + r4.z;
            x.w = (v[4 * j + 3] + b4.w) + r4.w;
            if (ch == 0 && j == 0) shift = x.x;
            const float d0 = x.x - shift, d1 = x.y - shift, d2 = x.z - shift, d3 = x.w - shift;
            s1 += (d0 + d1) + (d2 + d3);
            s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
            stg4[lane * 8 + (j ^ (lane & 7))] = x;
            xb[4 * j] = __float_as_uint(x.x);
            xb[4 * j + 1] = __float_as_uint(x.y);
            xb[4 * j + 2] = __float_as_uint(x.z);
            xb[4 * j + 3] = __float_as_uint(x.w);
          }
          tmem_st32(t_row + col, xb);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = row_base + 4 * i + cr;
            const float4 x = stg4[(4 * i + cr) * 8 + (cc ^ ((4 * i + cr) & 7))];
            if (r < M) *reinterpret_cast<float4*>(ln.R + (size_t)r * kN + col + 4 * cc) = x;
          }
          __syncwarp();  // the staging tile is rewritten by the next piece
        }
        // half-row mean and M2 from the shifted sums (the shift is a sample of the row: no cancellation)
        const float mean_h = shift + s1 * (1.0f / 128.f);
        const float m2 = fmaxf(s2 - s1 * s1 * (1.0f / 128.f), 0.f);
        // (mean, M2) of this half-row -> the upper 2 KB of this warp's staging tile (pass 3 uses the lower 2 KB
        // only; the next tile's pass 1 is behind the barrier at the end of the tile); the partner warp — same lane
        // quarter, other column half, four warps away — reads it from there
        float2* xch_mine = reinterpret_cast<float2*>(stg0_ptr + 2048);
        const int partner = (warp - 2) ^ 4;
        const float2* xch_other = reinterpret_cast<const float2*>(
            smem_aligned + (sEpi - smem_base) + partner * Cfg::kEpiBufs * Cfg::kEpiBufBytes + 2048);
        xch_mine[lane] = make_float2(mean_h, m2);
        named_bar_sync(1 + quarter, 64);  // the two warps of this lane quarter
        const float2 o = xch_other[lane];
        const float mean = 0.5f * (mean_h + o.x);
        const float dm = mean_h - o.x;
        const float var = (m2 + o.y + dm * dm * 64.f) * (1.0f / 256.f);
        const float rstd = rsqrtf(var + ln.eps);
        uint4* stgh = reinterpret_cast<uint4*>(stg0_ptr);               // [32 rows][4 x 16 B], slot ^ ((row >> 1) & 3)
        const int hr = lane >> 2, hc = lane & 3;                       // coalesced walk: 8 rows x 64 B per step
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          const int col = cgroup * 128 + ch * 32;
          float v[32];
          tmem_ld32(t_row + col, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 g0 = *reinterpret_cast<const float4*>(s_vec + 256 + col + 8 * j);
            const float4 g1 = *reinterpret_cast<const float4*>(s_vec + 256 + col + 8 * j + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(s_vec + 512 + col + 8 * j);
            const float4 b1 = *reinterpret_cast<const float4*>(s_vec + 512 + col + 8 * j + 4);
            const float* x = v + 8 * j;
            uint4 u;
            u.x = cvt_bf16x2((x[0] - mean) * rstd * g0.x + b0.x, (x[1] - mean) * rstd * g0.y + b0.y);
            u.y = cvt_bf16x2((x[2] - mean) * rstd * g0.z + b0.z, (x[3] - mean) * rstd * g0.w + b0.w);
            u.z = cvt_bf16x2((x[4] - mean) * rstd * g1.x + b1.x, (x[5] - mean) * rstd * g1.y + b1.y);
            u.w = cvt_bf16x2((x[6] - mean) * rstd * g1.z + b1.z, (x[7] - mean) * rstd * g1.w + b1.w);
            stgh[lane * 4 + (j ^ ((lane >> 1) & 3))] = u;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + hr, r = row_base + rr;
            const uint4 u = stgh[rr * 4 + (hc ^ ((rr >> 1) & 3))];
            if (r < M) *reinterpret_cast<uint4*>(ln.H + (size_t)r * kN + col + 8 * hc) = u;
          }
          __syncwarp();
        }
        named_bar_sync(1 + quarter, 64);  // the exchange slots are rewritten by the next tile
      }
#ifdef CSE_DBG_NOEPI
      constexpr int kNch = 0;
#else
      constexpr int kNch = LNE ? 0 : NCH;
#endif
#pragma unroll 1
      for (int ch = 0; ch < kNch; ++ch, ++chunk_ctr) {
        const int col_local = cgroup * GW + ch * CW;
        const int col0 = n0 + col_local;
        const uint32_t buf = chunk_ctr % Cfg::kEpiBufs;
        if (chunk_ctr >= (uint32_t)Cfg::kEpiBufs) {  // the TMA store that last used this staging tile has drained it
          if (lane == 0) bulk_wait_read<Cfg::kEpiBufs - 1>();
          __syncwarp();
        }
        uint4* stg = reinterpret_cast<uint4*>(stg0_ptr + buf * Cfg::kEpiBufBytes);
        float v[CW];
        if constexpr (CW == 64) tmem_ld64(t_row + col_local, v);
        else if constexpr (CW == 32) tmem_ld32(t_row + col_local, v);
        else tmem_ld16(t_row + col_local, v);
        if (bias != nullptr) {
#pragma unroll
          for (int i = 0; i < CW; i += 2) axpy_f32x2(v[i], v[i + 1], bv[i], bv[i + 1], bias_scale);  // FFMA2
          // prefetch the bias of the chunk this warp handles next (this tile's next chunk, or the first
          // chunk of its next tile): the load latency hides behind the pack/store and the next wait
          const int npt = pt + t_step;
          load_bias(ch + 1 < kNch ? col0 + CW : (npt < t_end ? (npt % n_tiles) * BN : n0) + cgroup * GW);
        }
        if (relu && OUT_F32) {  // (bf16 outputs fuse the ReLU into the conversion below)
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        // row = lane; 16-byte slot index XOR-swizzled exactly as the TMA box expects
        // (SWIZZLE_128B: slot ^ (row & 7); SWIZZLE_64B: slot ^ ((row >> 1) & 3)) -> conflict-free STS.128
        constexpr int kSlots = Cfg::kRowBytes / 16;
        const int sw = (kSlots == 8) ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
          uint4 u;
          if constexpr (OUT_F32) {
            u.x = __float_as_uint(v[4 * i]);
            u.y = __float_as_uint(v[4 * i + 1]);
            u.z = __float_as_uint(v[4 * i + 2]);
            u.w = __float_as_uint(v[4 * i + 3]);
          } else if (relu) {
            u.x = cvt_bf16x2_relu(v[8 * i], v[8 * i + 1]);
            u.y = cvt_bf16x2_relu(v[8 * i + 2], v[8 * i + 3]);
            u.z = cvt_bf16x2_relu(v[8 * i + 4], v[8 * i + 5]);
            u.w = cvt_bf16x2_relu(v[8 * i + 6], v[8 * i + 7]);
          } else {
            u.x = cvt_bf16x2(v[8 * i], v[8 * i + 1]);
            u.y = cvt_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            u.z = cvt_bf16x2(v[8 * i + 4], v[8 * i + 5]);
            u.w = cvt_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          }
          stg[lane * kSlots + (i ^ sw)] = u;
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          const uint32_t src = stg0 + buf * Cfg::kEpiBufBytes;
          // in-place residual update R += tile: TMA reduce-add (the stream is never loaded by the SM);
          // plain outputs: TMA store. (Direct coalesced st.global from the staging tile was measured
          // slower: QKV 75 vs 63 us.)
#ifndef CSE_DBG_NOSTORE
          if (accumulate_into_c) tma_reduce_add_2d(&tmC, src, col0, row_base);
          else tma_store_2d(&tmC, src, col0, row_base);
          bulk_commit();
#else
          (void)src;
#endif
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_tempty + 8 * astage);
      }
      if (++astage == 2) { astage = 0; aphase ^= 1; }
    }
    if (lane == 0) bulk_wait_all();  // all output writes complete before the CTA retires
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA retires while the peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point fetched through cudart; no libcuda link dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_rows, box_cols, esize;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           box_cols == o.box_cols && esize == o.esize;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<uint64_t>()(k.rows * 1315423911ull + k.cols * 2654435761ull + k.ld * 97ull +
                               k.box_rows * 131ull + k.box_cols * 7ull + k.esize);
    return h;
  }
};

// Row-major [rows, cols] matrix (bf16 or fp32) with leading dimension ld (elements);
// box = [box_rows x box_cols] with box_cols * esize == 128 B (SWIZZLE_128B) or 64 B (SWIZZLE_64B).
int get_tensor_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, uint32_t esize, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, rows, cols, ld, box_rows, box_cols, esize};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("gemm_tc: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols * esize == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%llu cols=%llu ld=%llu box=%ux%u esize=%u",
              (int)r, ptr, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld,
              box_rows, box_cols, esize);
    return 1;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

int sm_count() {  // of the CURRENT device (cached per ordinal)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

template <int BN, bool OUT_F32, bool LNE = false, int MNM = 0>
static int launch_tc_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                          const float* bias, float bias_scale, int accumulate, void* C, int ldc, int M,
                          int N, int K, int relu, cudaStream_t st, int ksplit = 1, LnEpilogue ln = LnEpilogue{}) {
  using Cfg = TcCfg<BN, LNE>;
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, OUT_F32, LNE, MNM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute(%zu B smem) failed: %s", Cfg::kSmem, cudaGetErrorString(e));
      return 1;
    }
    once.mark_configured();
  }
  const int pair_tiles = ceil_div(ceil_div(M, kTM), 2) * (N / BN) * ksplit;
  const int max_pairs = sm_count() / 2;
  const int grid = 2 * (pair_tiles < max_pairs ? pair_tiles : max_pairs);
  KernelScope prof(kClsGemmTc, st);
  cudaError_t le = launch_pdl(gemm_tc_kernel<BN, OUT_F32, LNE, MNM>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmem, st, 2, tmA, tmB, tmC, bias, bias_scale,
                                      accumulate, C, ldc, M, N, K, relu, ksplit, ln);
  if (le != cudaSuccess) {
    set_error("gemm_tc_kernel cluster launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("gemm_tc_kernel");
}

int launch_gemm_tc(const bf16* A, int lda, const bf16* W, const float* bias, float bias_scale,
                   const float* residual, void* C, int ldc, int M, int N, int K, int relu,
                   int out_fp32, cudaStream_t st) {
  if (M <= 0) return 0;
  if (K % kTK != 0 || N % 128 != 0 || lda % 8 != 0 || (ldc % 8) != 0) {
    set_error("gemm_tc: need K %% 64 == 0, N %% 128 == 0, lda/ldc %% 8 == 0 (M=%d N=%d K=%d lda=%d ldc=%d)",
              M, N, K, lda, ldc);
    return 1;
  }
  if (residual != nullptr && (!out_fp32 || (const void*)residual != (const void*)C)) {
    set_error("gemm_tc: the residual must be the fp32 output itself (in-place stream update via TMA reduce-add)");
    return 1;
  }
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) & 15) {
    set_error("gemm_tc: operands must be 16-byte aligned");
    return 1;
  }
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmB, tmC;
  if (get_tensor_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kTM, kTK, 2, &tmA)) return 1;
  if (get_tensor_map(W, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)(BN / 2), kTK, 2, &tmB)) return 1;  // half tile per CTA
  if (get_tensor_map(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, CSE_EPI_ROW_BYTES / (out_fp32 ? 4 : 2),
                     out_fp32 ? 4 : 2, &tmC))
    return 1;
  const int acc = residual != nullptr ? 1 : 0;
  if (BN == 256) {
    return out_fp32 ? launch_tc_impl<256, true>(tmA, tmB, tmC, bias, bias_scale, acc, C, ldc, M, N, K, relu, st)
                    : launch_tc_impl<256, false>(tmA, tmB, tmC, bias, bias_scale, acc, C, ldc, M, N, K, relu, st);
  }
  return out_fp32 ? launch_tc_impl<128, true>(tmA, tmB, tmC, bias, bias_scale, acc, C, ldc, M, N, K, relu, st)
                  : launch_tc_impl<128, false>(tmA, tmB, tmC, bias, bias_scale, acc, C, ldc, M, N, K, relu, st);
}

// R[M,256] += A[M,K] W[256,K]^T + b in place, H[M,256] = LayerNorm(R) (bf16): out-proj + residual + norm2 of a
// transformer layer in one kernel (CSE_transformer.py:399-408).
int launch_gemm_tc_residual_ln(const bf16* A, int lda, const bf16* W, const float* bias, float* R, const float* gamma,
                               const float* beta, float eps, bf16* H, int M, int K, cudaStream_t st) {
  if (M <= 0) return 0;
  if (K % kTK != 0 || lda % 8 != 0) {
    set_error("gemm_tc_residual_ln: need K %% 64 == 0 and lda %% 8 == 0 (K=%d lda=%d)", K, lda);
    return 1;
  }
  if (A == nullptr || W == nullptr || bias == nullptr || R == nullptr || gamma == nullptr || beta == nullptr || H == nullptr) {
    set_error("gemm_tc_residual_ln: NULL argument");
    return 1;
  }
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)R | (uintptr_t)H | (uintptr_t)bias | (uintptr_t)gamma | (uintptr_t)beta) & 15) {
    set_error("gemm_tc_residual_ln: operands must be 16-byte aligned");
    return 1;
  }
  CUtensorMap tmA, tmB;
  if (get_tensor_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kTM, kTK, 2, &tmA)) return 1;
  if (get_tensor_map(W, (uint64_t)kN, (uint64_t)K, (uint64_t)K, 128, kTK, 2, &tmB)) return 1;  // half tile per CTA
  LnEpilogue ln{R, gamma, beta, H, eps};
  return launch_tc_impl<256, true, true>(tmA, tmB, tmA /*unused*/, bias, 1.f, 0, R, kN, M, kN, K, 0, st, 1, ln);
}

// C[M,N] = A[M,K] W[K,N] for the row-major bf16 W [K, N] as nn.Linear stores it ([out, in]): the input gradient
// dA = dC W without a transposed weight copy (B operand MN-major).  C fp32 or bf16, overwritten.
int launch_gemm_tc_dgrad(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int out_fp32, int M, int N,
                         int K, cudaStream_t st) {
  if (M <= 0) return 0;
  if (K % kTK != 0 || N % 128 != 0 || lda % 8 != 0 || ldw % 8 != 0 || ldc % 8 != 0) {
    set_error("gemm_tc_dgrad: need K %% 64 == 0, N %% 128 == 0, lda/ldw/ldc %% 8 == 0 (M=%d N=%d K=%d)", M, N, K);
    return 1;
  }
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) & 15) {
    set_error("gemm_tc_dgrad: operands must be 16-byte aligned");
    return 1;
  }
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmB, tmC;
  if (get_tensor_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kTM, kTK, 2, &tmA)) return 1;
  if (get_tensor_map(W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, 64, 64, 2, &tmB)) return 1;  // [64 k x 64 n] boxes
  if (get_tensor_map(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, CSE_EPI_ROW_BYTES / (out_fp32 ? 4 : 2),
                     out_fp32 ? 4 : 2, &tmC))
    return 1;
  if (BN == 256)
    return out_fp32 ? launch_tc_impl<256, true, false, 2>(tmA, tmB, tmC, nullptr, 0.f, 0, C, ldc, M, N, K, 0, st)
                    : launch_tc_impl<256, false, false, 2>(tmA, tmB, tmC, nullptr, 0.f, 0, C, ldc, M, N, K, 0, st);
  return out_fp32 ? launch_tc_impl<128, true, false, 2>(tmA, tmB, tmC, nullptr, 0.f, 0, C, ldc, M, N, K, 0, st)
                  : launch_tc_impl<128, false, false, 2>(tmA, tmB, tmC, nullptr, 0.f, 0, C, ldc, M, N, K, 0, st);
}

// C[M,N] (fp32) += X^T Y for row-major bf16 X [T, M] (ldx), Y [T, N] (ldy), contracted over the T tokens and split
// over the CTA pairs: the weight gradient dW = dC^T A straight from the row-major dC and A (MN-major operands).
int launch_gemm_tc_wgrad(const bf16* X, int ldx, const bf16* Y, int ldy, float* C, int ldc, int M, int N, int T,
                         cudaStream_t st) {
  if (M <= 0 || T <= 0) return 0;
  if (M % 128 != 0 || N % 128 != 0 || ldx % 8 != 0 || ldy % 8 != 0 || ldc % 8 != 0) {
    set_error("gemm_tc_wgrad: need M %% 128 == 0, N %% 128 == 0, ldx/ldy/ldc %% 8 == 0 (M=%d N=%d)", M, N);
    return 1;
  }
  if (((uintptr_t)X | (uintptr_t)Y | (uintptr_t)C) & 15) {
    set_error("gemm_tc_wgrad: operands must be 16-byte aligned");
    return 1;
  }
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmX, tmY, tmC;
  if (get_tensor_map(X, (uint64_t)T, (uint64_t)M, (uint64_t)ldx, 64, 64, 2, &tmX)) return 1;  // [64 tokens x 64 features]
  if (get_tensor_map(Y, (uint64_t)T, (uint64_t)N, (uint64_t)ldy, 64, 64, 2, &tmY)) return 1;
  if (get_tensor_map(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, CSE_EPI_ROW_BYTES / 4, 4, &tmC)) return 1;
  const int Tk = (int)align_up((size_t)T, kTK);     // the token tail of the last k-block reads as zeros (TMA)
  const int num_kb = Tk / kTK;
  const int tiles = ceil_div(ceil_div(M, kTM), 2) * (N / BN);
  int want = (sm_count() / 2) / tiles;
  if (want > num_kb / 4) want = num_kb / 4;
  if (want < 1) want = 1;
  const int kb_per = ceil_div(num_kb, want);
  const int ksplit = ceil_div(num_kb, kb_per);
  return BN == 256
             ? launch_tc_impl<256, true, false, 1>(tmX, tmY, tmC, nullptr, 0.f, 1, C, ldc, M, N, Tk, 0, st, ksplit)
             : launch_tc_impl<128, true, false, 1>(tmX, tmY, tmC, nullptr, 0.f, 1, C, ldc, M, N, Tk, 0, st, ksplit);
}

// C[M,N] (fp32) += A[M,K] W[N,K]^T with the K dimension split over CTAs: the weight-gradient shape (M, N = a layer's
// out / in features, K = the token count), whose 2-8 output tiles would otherwise occupy 2-8 of the 148 SMs.
int launch_gemm_tc_splitk(const bf16* A, int lda, const bf16* W, int ldw, float* C, int ldc, int M, int N, int K,
                          cudaStream_t st) {
  if (M <= 0) return 0;
  if (K % kTK != 0 || N % 128 != 0 || lda % 8 != 0 || ldw % 8 != 0 || ldc % 8 != 0) {
    set_error("gemm_tc_splitk: need K %% 64 == 0, N %% 128 == 0, lda/ldw/ldc %% 8 == 0 (M=%d N=%d K=%d)", M, N, K);
    return 1;
  }
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) & 15) {
    set_error("gemm_tc_splitk: operands must be 16-byte aligned");
    return 1;
  }
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmB, tmC;
  if (get_tensor_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kTM, kTK, 2, &tmA)) return 1;
  if (get_tensor_map(W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(BN / 2), kTK, 2, &tmB)) return 1;
  if (get_tensor_map(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, CSE_EPI_ROW_BYTES / 4, 4, &tmC)) return 1;
  const int num_kb = K / kTK;
  const int tiles = ceil_div(ceil_div(M, kTM), 2) * (N / BN);
  int want = (sm_count() / 2) / tiles;              // one job per CTA pair
  if (want > num_kb / 4) want = num_kb / 4;         // at least four k-blocks (one ring) per job
  if (want < 1) want = 1;
  const int kb_per = ceil_div(num_kb, want);
  const int ksplit = ceil_div(num_kb, kb_per);      // no empty job
  return BN == 256 ? launch_tc_impl<256, true>(tmA, tmB, tmC, nullptr, 0.f, 1, C, ldc, M, N, K, 0, st, ksplit)
                   : launch_tc_impl<128, true>(tmA, tmB, tmC, nullptr, 0.f, 1, C, ldc, M, N, K, 0, st, ksplit);
}

}  // namespace cse
