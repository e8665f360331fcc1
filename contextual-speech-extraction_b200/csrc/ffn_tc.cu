// ffn_tc.cu — the whole position-wise feed-forward sub-block in ONE tcgen05 kernel (CSE_BF16):
//
//   R[M,256] += relu( A[M,256] · W1[1024,256]^T + b1 ) · W2[256,1024]^T + b2
//
// Reference: `src = src + pos_ffn(norm2(src))` with pos_ffn = Linear(256,1024) -> ReLU -> Linear(1024,256)
// (CSE_transformer.py:407-411, PositionalwiseFeedForward :547-566); A is the bf16 output of norm2.
// Unfused, the [M,1024] hidden activation is written by one GEMM (280 MB for cfg2) and read back by the
// next; both GEMMs are bound by L2 / HBM traffic, not by the tensor core.  Here the hidden activation
// never leaves the SM — it does not even leave TENSOR MEMORY: the epilogue warps turn the fp32
// accumulator into bf16 pairs in place (tcgen05.ld -> bias/ReLU -> tcgen05.st) and the second GEMM
// reads it from TMEM as its A operand.
//
// Measured on B200 (tools/micro/umma_rate.cu): back-to-back 128-row UMMAs into one accumulator cost at least
// ~96 cycles each however small N is (107 at N = 128), so only N = 256 instructions run near the full rate.
// Both GEMMs are therefore issued as 128x256x16 UMMAs and the hidden dimension is walked in 4 chunks of 256:
//   G1_c : Hacc (TMEM, 256 cols) = A(128x256) · W1_c(256x256)^T             16 UMMAs (SS)
//   E1_c : Hacc -> +b1 -> ReLU -> bf16 pairs -> back into the SAME TMEM columns, in four 64-unit pieces;
//          all eight E1 warps work on the same piece (32 columns each): a piece costs a warp ~500 cycles
//          (mostly the convert/pack arithmetic, the TMEM load itself is <100), and what the second GEMM
//          waits for is the FIRST piece
//   G2_c : Y (TMEM, 256 cols) += H_c(128x256, A operand from TMEM) · W2[:, chunk c]^T   16 UMMAs (TS)
// TMEM is full (Y 256 + Hacc 256 columns), so Hacc is single-buffered; to keep the tensor pipe busy the
// G2 k-blocks are issued piece by piece as E1 finishes them, and G1_{c+1} follows G2_c in pipe order,
// which is all the protection Hacc needs.
// After G2_3 the output warps add b2 and fold Y into the fp32 residual stream with TMA reduce-add (the
// stream is never loaded into the SM).
// Shared memory: A 64 KB (4 k-blocks, resident for the tile) + a 4-stage x 32 KB weight ring (one
// stage = 256 weight rows x 64 k = four UMMAs) + 32 KB output staging.  CTAs run in clusters of two on
// adjacent row tiles: every weight box is fetched from L2 once per pair (each CTA loads 128 of the 256
// rows and multicasts them).
//
// LNF variant (launch_ffn_tc_ln, 512 threads): the two LayerNorms either side of the sub-block ride along in this
// compute-bound kernel instead of being two HBM-bound launches of their own (64 launches, 14.6 % of the cfg 2 forward):
//   norm2 (CSE_transformer.py:406): warps 14-15 normalise the fp32 residual rows of the tiles AHEAD of the tensor pipe
//     into a bf16 scratch matrix in global memory (it stays in L2) and the TMA producer loads the A tile from there once
//     a shared-memory counter says the tile's rows are written (generic-proxy writes -> fence.proxy.async -> release /
//     acquire -> TMA).  All fourteen non-pipe warps share the first tile so the pipe starts after two batches of rows.
//   norm1 of the NEXT layer (CSE_transformer.py:387): each output warp owns 32 complete rows of the tile it has just
//     folded into R; once its reduce-adds have completed it reads those rows back (L2 hits) and writes their LayerNorm
//     for the next layer's in_proj GEMM.
// A transformer layer is then four launches: in_proj GEMM, attention, out_proj GEMM (+R), this kernel.  The LNF tile
// walk runs from the last row tile down: out_proj walked the rows upwards, so the walk starts in L2.
//
// Roles (448 threads): warp 0 TMA producer | warp 1 single-thread tcgen05.mma issuer | warps 2-9
// hidden-chunk epilogue E1 (TMEM lane quarter = warp & 3, column half = (warp-2) >> 2) | warps 10-13
// output epilogue Y -> R (one per lane quarter, two staging chunks each).  All mbarrier waits are bounded.
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

using namespace tc;

int get_tensor_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, uint32_t esize, CUtensorMap* out);  // gemm_tc.cu
int sm_count();

#ifdef FFN_TRACE
// development aid (tools/gemm_variants.py): per-stage timestamps of block 0, never compiled into the product
__device__ unsigned long long g_ffn_trace[4][64];
#define FFN_TRACE_PUT(role, idx) \
  if (blockIdx.x == 0 && (idx) < 64) g_ffn_trace[role][idx] = clock64()
extern "C" __attribute__((visibility("default"))) int cse_debug_ffn_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ffn_trace, sizeof(g_ffn_trace));
}
#else
#define FFN_TRACE_PUT(role, idx)
#endif

namespace {

#ifdef FFN_SOLO
#define FFN_COMMIT_WEMPTY(bar) umma_commit(bar)
#else
#define FFN_COMMIT_WEMPTY(bar) umma_commit_mcast(bar, (uint16_t)3)
#endif

constexpr int kFfnThreads = 448;
// + warps 14-15: LayerNorm (norm2) producers.  Sixteen warps = four per SM sub-partition: a fifth warp on a
// sub-partition would cap every thread at 96 registers and spill the E1 warps (the critical path)
constexpr int kFfnLnThreads = 512;
constexpr int kD = 256;        // d_model
constexpr int kH = 1024;       // d_ffn
constexpr int kHC = 256;       // hidden units per chunk
constexpr int kChunks = kH / kHC;
constexpr int kKb = 128 * 128;          // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kStageBytes = 2 * kKb;    // 32 KB: 256 weight rows x 64 k
#ifndef FFN_STAGES
#define FFN_STAGES 4
#endif
constexpr int kStages = FFN_STAGES;
constexpr int kABytes = 4 * kKb;        // resident A tile: 4 k-blocks
constexpr int kStgBytes = 32 * 128;     // output staging chunk: 32 rows x 32 fp32 (two per output warp)
constexpr int kBiasBytes = kD * 4;      // b2 (b1 is read through L1)
constexpr size_t kFfnSmem = 1024 + kABytes + kStages * kStageBytes + 8 * kStgBytes + kBiasBytes + 256;

struct FfnLn {            // LNF only
  const float* R;         // [M,256] fp32 residual stream (also updated through tmR)
  const float* g2;        // this layer's norm2 weight / bias
  const float* b2;
  const float* g1n;       // next layer's norm1 weight / bias (H1 == nullptr: not produced)
  const float* b1n;
  bf16* A;                // [M,256] scratch: norm2(R), written by the LayerNorm warps, read back through tmA
  bf16* H1;               // [M,256] norm1_next(R + FFN(norm2(R)))
  float eps;
  int dbg;                // timing experiments only (cse_debug_ffn_ln): 1 = norm2 warps skip their rows, 2 = no norm1
};
int g_ffn_ln_dbg = 0;
// dbg & 4: block 0 stamps clock64 per tile — [4 it + k]: output warp 10 (k = 0 Y ready, 1 Y drained, 2 reduce-adds
// complete, 3 norm1 rows written); [128 + it]: norm2 warp 14 delivered tile it; [160 + it]: the producer got tile it's rows
__device__ unsigned long long g_ffn_ln_trace[192];
#define LN_STAMP(cond, idx) \
  if ((ln.dbg & 4) && blockIdx.x == 0 && lane == 0 && (cond) && (idx) < 192) g_ffn_ln_trace[idx] = clock64()

// LayerNorm of the eight rows [row0, row0 + 8) of R (clipped at M) -> bf16 rows of H.  One warp, layernorm_kernel's
// arithmetic and channel ownership (norm.cu: same results); the eight rows' loads and reductions go together.  Loads go
// to L2 only: the rows are rewritten by TMA reduce-adds.
__device__ __forceinline__ void ln_batch_bf16(const float* __restrict__ R, int M, int row0, const f8& gg, const f8& bb,
                                              float eps, bf16* __restrict__ H, int lane) {
  f8 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ld_row8_cg(R + (size_t)min(row0 + i, M - 1) * kD, lane);
  ln_rows<8>(v, gg, bb, eps);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (row0 + i < M) st_row8(H + (size_t)(row0 + i) * kD, lane, v[i]);
}
__device__ __forceinline__ void ln_rows_bf16(const float* __restrict__ R, int M, int row0, int nrows,
                                             const float* __restrict__ g, const float* __restrict__ b, float eps,
                                             bf16* __restrict__ H, int lane) {
  const f8 gg = ld_row8(g, lane), bb = ld_row8(b, lane);
  for (int r = 0; r < nrows; r += 8) ln_batch_bf16(R, M, row0 + r, gg, bb, eps, H, lane);
}
// eight rows (8 KB, contiguous) -> L2, one instruction
__device__ __forceinline__ void prefetch_rows8(const float* R, int M, int row0) {
  if (row0 + 8 <= M)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(R + (size_t)row0 * kD), "r"(8 * kD * 4) : "memory");
}

// rows written (generic proxy) -> visible to the TMA producer's loads (async proxy): proxy fence, then a release
// increment of the shared-memory row counter; the producer acquires it
__device__ __forceinline__ void ln_signal(uint32_t ctr, int lane) {
  asm volatile("fence.proxy.async;\n" ::: "memory");
  __syncwarp();
  if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;\n" ::"r"(ctr) : "memory");
}
__device__ __forceinline__ void ln_wait(uint32_t ctr, uint32_t target) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(ctr) : "memory");
    if (v >= target) break;
    if (clock64() - t0 > 4000000000LL) {
      printf("ffn_tc: LayerNorm rows timeout (block %d, have %u, need %u)\n", (int)blockIdx.x, v, target);
      __trap();
    }
  }
}

// G2 consumes the four 64-unit pieces of a hidden chunk in the order E1 finishes them
__device__ __forceinline__ int g2_piece(int s) { return s; }

template <bool LNF>
__global__ void __launch_bounds__(LNF ? kFfnLnThreads : kFfnThreads, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR,
              const float* __restrict__ b1, const float* __restrict__ b2, int M, FfnLn ln) {
  constexpr int kThreads = LNF ? kFfnLnThreads : kFfnThreads;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t sA = smem_base;
  const uint32_t sW = sA + kABytes;
  const uint32_t sStg = sW + kStages * kStageBytes;
  const uint32_t sBias = sStg + 8 * kStgBytes;
  const uint32_t sBar = sBias + kBiasBytes;
  float* s_b2 = reinterpret_cast<float*>(smem_al + (sBias - smem_base));
  const uint32_t bar_wfull = sBar;              // [<=6]
  const uint32_t bar_wempty = sBar + 48;        // [<=6]
  const uint32_t bar_afull = sBar + 96;         // [4] k-blocks of the resident A tile
  const uint32_t bar_aempty = sBar + 128;       // [4]
  const uint32_t bar_hfull = sBar + 160;        // G1_c complete: Hacc holds fp32 pre-activations
  const uint32_t bar_pfull = sBar + 168;        // [4] piece p of the chunk is bf16 in TMEM
  const uint32_t bar_yfull = sBar + 200;
  const uint32_t bar_yempty = sBar + 208;
  const uint32_t tmem_slot = sBar + 216;
  const uint32_t ln_ctr = sBar + 224;           // LNF: warps that have delivered their norm2 rows (14 for the first tile, 2 per later tile)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_pairs = ((M + 127) / 128 + 1) >> 1;
  // tile pair of walk step i: LNF walks the row tiles downwards (see the header)
  auto pair_at = [&](int i) { return LNF ? m_pairs - 1 - i : i; };

  if (threadIdx.x == 0) {
    if constexpr (LNF) *reinterpret_cast<volatile uint32_t*>(smem_al + (ln_ctr - smem_base)) = 0u;
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
#ifdef FFN_SOLO
      mbar_init(bar_wempty + 8 * i, 1);
#else
      mbar_init(bar_wempty + 8 * i, 2);  // released by the MMA issuers of both CTAs of the pair
#endif
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_afull + 8 * i, 1);
      mbar_init(bar_aempty + 8 * i, 1);
      mbar_init(bar_pfull + 8 * i, 8);   // all eight E1 warps contribute to every piece
    }
    mbar_init(bar_hfull, 1);
    mbar_init(bar_yfull, 1);
    mbar_init(bar_yempty, 4);   // one arrive per output-epilogue warp
    mbar_fence_init();
  }
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < kD; i += kThreads) s_b2[i] = b2[i];  // parameter: not produced by the previous kernel
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // prologue done; from here on the previous kernel's output is read and R is updated
  const uint32_t tmem_y = tmem_base;            // columns [0, 256)
  const uint32_t tmem_h = tmem_base + 256;      // columns [256, 512)

  if (warp == 0) {
    // ================= TMA producer (lane 0 acts; the warp stays convergent) =================
    uint32_t stage = 0, wphase = 0;
    int it = 0;
    for (int mpi = pair_id; mpi < m_pairs; mpi += npairs, ++it) {
      const int mp = pair_at(mpi);
      const int m0 = (2 * mp + rank) * 128;
      if (lane == 0) {
        if (!LNF && mp + npairs < m_pairs) {  // next tile's A -> L2 while this tile computes
          for (int kb = 0; kb < 4; ++kb) tma_prefetch_l2_2d(&tmA, kb * 64, (2 * (mp + npairs) + rank) * 128);
        }
        for (int c = 0; c < kChunks; ++c) {
          for (int s = 0; s < 8; ++s) {  // stages 0..3: W1 k-blocks of G1_c; 4..7: W2 pieces of G2_c
            if (c == 0 && s < 4) {  // the previous tile's last G1 has finished with this k-block of A
              if (LNF && s == 0) {
                ln_wait(ln_ctr, 14u + 2u * (uint32_t)it);  // this tile's norm2 rows are written
                LN_STAMP(it < 32, 160 + it);
              }
              mbar_wait_spin(bar_aempty + 8 * s, ((uint32_t)it & 1u) ^ 1u, 1);
              mbar_expect_tx(bar_afull + 8 * s, kKb);
              tma_load_2d(sA + s * kKb, &tmA, bar_afull + 8 * s, s * 64, m0);
            }
            mbar_wait_spin(bar_wempty + 8 * stage, wphase ^ 1u, 2);
            mbar_expect_tx(bar_wfull + 8 * stage, kStageBytes);
#ifdef FFN_SOLO
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t dst = sW + stage * kStageBytes + hh * kKb;
              if (s < 4) tma_load_2d(dst, &tmW1, bar_wfull + 8 * stage, s * 64, c * kHC + hh * 128);
              else tma_load_2d(dst, &tmW2, bar_wfull + 8 * stage, c * kHC + g2_piece(s - 4) * 64, hh * 128);
            }
#else
            const uint32_t dst = sW + stage * kStageBytes + rank * kKb;  // this CTA fetches 128 of the 256 rows
            if (s < 4)   // W1 rows [256c, +256) x k-block s
              tma_load_2d_mcast(dst, &tmW1, bar_wfull + 8 * stage, s * 64, c * kHC + rank * 128, (uint16_t)3);
            else         // W2 rows [0, 256) x hidden units [256c + 64p, +64)
              tma_load_2d_mcast(dst, &tmW2, bar_wfull + 8 * stage, c * kHC + g2_piece(s - 4) * 64, rank * 128,
                                (uint16_t)3);
#endif
            if (++stage == kStages) { stage = 0; wphase ^= 1u; }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer (lane 0 issues; the warp stays convergent) =================
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
    uint32_t stage = 0, wphase = 0;
    int it = 0;
    for (int mpi = pair_id; mpi < m_pairs; mpi += npairs, ++it) {
      if (lane == 0) {
        for (int c = 0; c < kChunks; ++c) {
          const uint32_t use = (uint32_t)(it * kChunks + c);
          // ---- G1_c: Hacc is free — G2_{c-1}, its last reader, precedes this in pipe order ----
          for (int kb = 0; kb < 4; ++kb) {
            FFN_TRACE_PUT(0, (it * kChunks + c) * 8 + kb);
            if (c == 0) mbar_wait_spin(bar_afull + 8 * kb, (uint32_t)it & 1u, 5);
            mbar_wait_spin(bar_wfull + 8 * stage, wphase, 6);
            fence_after();
            FFN_TRACE_PUT(1, (it * kChunks + c) * 8 + kb);
            const uint64_t adesc = make_desc(sA + kb * kKb, 1024, kLayoutSw128);
            const uint64_t bdesc = make_desc(sW + stage * kStageBytes, 1024, kLayoutSw128);
#if !defined(FFN_DBG_NOMMA)
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_h, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
#endif
            FFN_COMMIT_WEMPTY(bar_wempty + 8 * stage);
            if (c == kChunks - 1) umma_commit(bar_aempty + 8 * kb);  // A k-block free for the next tile
            if (++stage == kStages) { stage = 0; wphase ^= 1u; }
          }
          umma_commit(bar_hfull);
          // ---- G2_c, piece by piece as E1 delivers them ----
          for (int s = 0; s < 4; ++s) {
            const int p = g2_piece(s);
            FFN_TRACE_PUT(0, (it * kChunks + c) * 8 + 4 + s);
            if (c == 0 && s == 0) mbar_wait_spin(bar_yempty, ((uint32_t)it & 1u) ^ 1u, 7);  // previous tile's Y drained
            mbar_wait_spin(bar_pfull + 8 * p, use & 1u, 8);
            FFN_TRACE_PUT(1, (it * kChunks + c) * 8 + 4 + s);
            mbar_wait_spin(bar_wfull + 8 * stage, wphase, 9);
            fence_after();
            FFN_TRACE_PUT(2, (it * kChunks + c) * 8 + 4 + s);
            // hidden units [64p, 64p+64) of the chunk = packed bf16 pairs in TMEM columns [64p, 64p+32) of Hacc
            const uint32_t a_tmem = tmem_h + p * 64;
            const uint64_t bdesc = make_desc(sW + stage * kStageBytes, 1024, kLayoutSw128);
#if !defined(FFN_DBG_NOMMA)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tmem_y, a_tmem + 8 * k, bdesc + 2 * k, idesc, (c | s | k) != 0 ? 1u : 0u);
#endif
            FFN_COMMIT_WEMPTY(bar_wempty + 8 * stage);
            if (++stage == kStages) { stage = 0; wphase ^= 1u; }
          }
          if (c == kChunks - 1) umma_commit(bar_yfull);
        }
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ================= hidden-chunk epilogue E1, warps 2..9 =================
    const int q = warp & 3;            // TMEM lane quarter (fixed by hardware: warp id % 4)
    const int h = (warp - 2) >> 2;     // column half: hidden units [128h, 128h+128) of the chunk
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    if constexpr (LNF) {  // first tile: every non-pipe warp normalises eight of its rows
      ln_rows_bf16(ln.R, M, (2 * pair_at(pair_id) + rank) * 128 + (warp - 2) * 8, 8, ln.g2, ln.b2, ln.eps, ln.A, lane);
      ln_signal(ln_ctr, lane);
    }
    int it = 0;
    for (int mpi = pair_id; mpi < m_pairs; mpi += npairs, ++it) {
      for (int c = 0; c < kChunks; ++c) {
        const uint32_t use = (uint32_t)(it * kChunks + c);
        mbar_wait(bar_hfull, use & 1u, 11);
        fence_after();
#ifdef FFN_TRACE
        const bool tr = (warp == 2 && lane == 0 && it == 1 && c == 1);
        if (tr) { FFN_TRACE_PUT(3, 39); }
#define E1_STAMP(j) if (tr) { FFN_TRACE_PUT(3, 40 + p * 5 + (j)); }
#else
#define E1_STAMP(j)
#endif
#pragma unroll 1
        for (int p = 0; p < 4; ++p) {  // piece: hidden units [64p, 64p+64) of the chunk; this warp: 32 of them
#ifndef FFN_DBG_NOE1
          // (issuing the load of piece p+1 before converting piece p was measured slower)
          // the piece's 32 bias values are requested BEFORE the TMEM load (an asm the compiler will not move
          // loads across), so their L1 round trips overlap it
          float4 bq[8];
          const float4* bsrc = reinterpret_cast<const float4*>(b1 + c * kHC + p * 64 + h * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) bq[i] = __ldg(bsrc + i);  // lane-uniform, L1-resident
          if constexpr (LNF) {
            if (ln.dbg & 8) {
#pragma unroll
              for (int i = 0; i < 8; ++i) bq[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // timing experiment: no bias loads
            }
          }
          float v[32];
          tmem_ld32(tmem_h + lane_off + p * 64 + h * 32, v);
          E1_STAMP(0)
          // bias + ReLU, pack to bf16 pairs (low half = even hidden unit)
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // two packed adds + two ReLU-fused conversions per four hidden units
            const float4 t = bq[i];
            add_f32x2(v[4 * i], v[4 * i + 1], t.x, t.y);
            add_f32x2(v[4 * i + 2], v[4 * i + 3], t.z, t.w);
            pk[2 * i] = cvt_bf16x2_relu(v[4 * i], v[4 * i + 1]);
            pk[2 * i + 1] = cvt_bf16x2_relu(v[4 * i + 2], v[4 * i + 3]);
          }
          // The packed piece occupies columns [64p, 64p+32): the h = 1 warp writes where the h = 0 warp of
          // the same lane quarter has just READ, so the two warps meet before either stores.
          E1_STAMP(1)
          named_bar_sync(1 + q, 64);
          E1_STAMP(2)
          tmem_st16(tmem_h + lane_off + p * 64 + h * 16, pk);
          E1_STAMP(3)
#endif
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pfull + 8 * p);
          E1_STAMP(4)
        }
      }
    }
  } else if (LNF && warp >= 14) {
    // ================= norm2 producers, warps 14..15 (LNF): rows of the tiles ahead -> bf16 scratch A =================
    // first tile: rows [96, 128) (the E1 and output warps take eight rows each of [0, 96)); later tiles: 64 rows each.
    // The warp's work is a stream of 8-row batches; the batch two places ahead is pulled into L2 while this one is
    // normalised (a warp holds 8 KB of loads in registers — at HBM latency that is a third of the rate the tensor pipe
    // consumes rows at; a whole tile of lead was too long: the lines were evicted again before use).
    const f8 gg = ld_row8(ln.g2, lane), bb = ld_row8(ln.b2, lane);
    const int w = warp - 14;
    auto batch_row = [&](int mpi, int it, int b) {   // first row of batch b of walk step mpi
      return (2 * pair_at(mpi) + rank) * 128 + (it == 0 ? 96 + w * 16 : w * 64) + b * 8;
    };
    int it = 0;
    for (int mpi = pair_id; mpi < m_pairs; mpi += npairs, ++it) {
      const int nb = it == 0 ? 2 : 8;
      const bool more = mpi + npairs < m_pairs;
      if (it == 0 && lane == 0) prefetch_rows8(ln.R, M, batch_row(mpi, 0, 1));
      for (int b = 0; b < nb; ++b) {
        if (lane == 0) {
          if (b + 2 < nb) prefetch_rows8(ln.R, M, batch_row(mpi, it, b + 2));
          else if (more) prefetch_rows8(ln.R, M, batch_row(mpi + npairs, it + 1, b + 2 - nb));
        }
        if (!(ln.dbg & 1)) ln_batch_bf16(ln.R, M, batch_row(mpi, it, b), gg, bb, ln.eps, ln.A, lane);
      }
      ln_signal(ln_ctr, lane);
      LN_STAMP(warp == 14 && it < 32, 128 + it);
    }
  } else {
    // ================= output epilogue Y + b2 -> R, warps 10..13 =================
    // fp32 residual stream updated in place by TMA reduce-add; rows past M are clipped by the tensor map
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t stg0 = sStg + (warp - 10) * 2 * kStgBytes;  // two staging chunks per warp
    if constexpr (LNF) {
      ln_rows_bf16(ln.R, M, (2 * pair_at(pair_id) + rank) * 128 + (warp - 2) * 8, 8, ln.g2, ln.b2, ln.eps, ln.A, lane);
      ln_signal(ln_ctr, lane);
    }
    int it = 0;
    for (int mpi = pair_id; mpi < m_pairs; mpi += npairs, ++it) {
      const int row_base = (2 * pair_at(mpi) + rank) * 128 + q * 32;
      mbar_wait(bar_yfull, (uint32_t)it & 1u, 13);
      fence_after();
      if constexpr (LNF) { LN_STAMP(warp == 10 && it < 32, 4 * it); }
      if (warp == 10 && lane == 0) { FFN_TRACE_PUT(2, it * 16); }
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int col = c * 32;
        float y[32];
        tmem_ld32(tmem_y + lane_off + col, y);
        if (warp == 10 && lane == 0) { FFN_TRACE_PUT(2, it * 16 + 1 + c); }
        if (c == 7) {  // Y fully read by this warp: the next tile's G2_0 may overwrite it
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_yempty);
        }
#ifndef FFN_DBG_NOFINAL
        // (coalesced red.global.add.v4.f32 from the staging tile was measured slower: 156 vs 148 us)
        const uint32_t stg = stg0 + (c & 1) * kStgBytes;
        uint4* stg_ptr = reinterpret_cast<uint4*>(smem_al + (stg - smem_base));
        if (lane == 0) bulk_wait_read<1>();  // the reduce-add issued two chunks ago has drained this staging tile
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = *reinterpret_cast<const float4*>(s_b2 + col + 4 * i);
          add_f32x2(y[4 * i], y[4 * i + 1], t.x, t.y);
          add_f32x2(y[4 * i + 2], y[4 * i + 3], t.z, t.w);
          uint4 u;
          u.x = __float_as_uint(y[4 * i]);
          u.y = __float_as_uint(y[4 * i + 1]);
          u.z = __float_as_uint(y[4 * i + 2]);
          u.w = __float_as_uint(y[4 * i + 3]);
          stg_ptr[lane * 8 + (i ^ (lane & 7))] = u;  // SWIZZLE_128B, conflict-free
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmR, stg, col, row_base);
          bulk_commit();
        }
#endif
      }
      if constexpr (LNF) {
        LN_STAMP(warp == 10 && it < 32, 4 * it + 1);
        if (ln.H1 != nullptr && !(ln.dbg & 2) && mpi + npairs < m_pairs) {   // (the last tile: all warps, below)
          // norm1 of the next layer on the 32 rows this warp has just updated: its own reduce-adds have completed
          // (wait_group without .read), the rows come back from L2
          if (lane == 0) {
            if (ln.dbg & 16) bulk_wait_read<0>();   // timing experiment: no L1 invalidate (rows may be stale)
            else bulk_wait_all();   // completion of the writes; includes the L1 invalidate generic loads need
          }
          __syncwarp();
          LN_STAMP(warp == 10 && it < 32, 4 * it + 2);
          ln_rows_bf16(ln.R, M, row_base, 32, ln.g1n, ln.b1n, ln.eps, ln.H1, lane);
          LN_STAMP(warp == 10 && it < 32, 4 * it + 3);
        }
      }
    }
    if (lane == 0) bulk_wait_all();  // all residual updates complete before the CTA retires
    __syncwarp();
  }

  if constexpr (LNF) {
    // The CTA's LAST tile: nothing is left to overlap its norm1 rows with, so all fourteen non-pipe warps share them
    // (one 8- or 16-row trip each instead of four trips by the output warps: the kernel's tail).  The output warps
    // arrive after their reduce-adds have completed (bulk_wait_all above, which also invalidates L1).
    if (ln.H1 != nullptr && !(ln.dbg & 2) && warp >= 2) {
      named_bar_sync(8, 14 * 32);
      const int n_mine = (m_pairs - pair_id + npairs - 1) / npairs;
      const int m0 = (2 * pair_at(pair_id + (n_mine - 1) * npairs) + rank) * 128;
      const int w = warp - 2;
      if (w < 12) ln_rows_bf16(ln.R, M, m0 + w * 8, 8, ln.g1n, ln.b1n, ln.eps, ln.H1, lane);
      else ln_rows_bf16(ln.R, M, m0 + 96 + (w - 12) * 16, 16, ln.g1n, ln.b1n, ln.eps, ln.H1, lane);
    }
  }

  fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA retires while the peer may still multicast into it
  if (warp == 1) {
    fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

static int launch_ffn_common(bool lnf, const bf16* A, const bf16* W1, const float* b1, const bf16* W2, const float* b2,
                             float* R, int M, const FfnLn& ln, cudaStream_t st) {
  if (M <= 0) return 0;
  if (((uintptr_t)A | (uintptr_t)W1 | (uintptr_t)W2 | (uintptr_t)R | (uintptr_t)b1 | (uintptr_t)b2) & 15) {
    set_error("ffn_tc: operands must be 16-byte aligned");
    return 1;
  }
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFfnSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(ffn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFfnSmem);
    if (e != cudaSuccess) {
      set_error("ffn_tc: cudaFuncSetAttribute(%zu B smem) failed: %s", kFfnSmem, cudaGetErrorString(e));
      return 1;
    }
    once.mark_configured();
  }
  CUtensorMap tmA, tmW1, tmW2, tmR;
  if (get_tensor_map(A, (uint64_t)M, kD, kD, 128, 64, 2, &tmA)) return 1;
  if (get_tensor_map(W1, kH, kD, kD, 128, 64, 2, &tmW1)) return 1;   // half of a 256-row stage per CTA
  if (get_tensor_map(W2, kD, kH, kH, 128, 64, 2, &tmW2)) return 1;
  if (get_tensor_map(R, (uint64_t)M, kD, kD, 32, 32, 4, &tmR)) return 1;
  const int m_pairs = (ceil_div(M, 128) + 1) / 2;
  const int max_pairs = sm_count() / 2;
  const int grid = 2 * (m_pairs < max_pairs ? m_pairs : max_pairs);
  KernelScope prof(kClsFfn, st);
  cudaError_t le = lnf ? launch_pdl(ffn_tc_kernel<true>, dim3(grid), dim3(kFfnLnThreads), kFfnSmem, st, 2, tmA, tmW1,
                                    tmW2, tmR, b1, b2, M, ln)
                       : launch_pdl(ffn_tc_kernel<false>, dim3(grid), dim3(kFfnThreads), kFfnSmem, st, 2, tmA, tmW1,
                                    tmW2, tmR, b1, b2, M, ln);
  if (le != cudaSuccess) {
    set_error("ffn_tc_kernel cluster launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("ffn_tc_kernel");
}

extern "C" __attribute__((visibility("default"))) void cse_debug_ffn_ln(int flags) { g_ffn_ln_dbg = flags; }
extern "C" __attribute__((visibility("default"))) int cse_debug_ffn_ln_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ffn_ln_trace, sizeof(g_ffn_ln_trace));
}

int launch_ffn_tc(const bf16* A, const bf16* W1, const float* b1, const bf16* W2, const float* b2, float* R,
                  int M, cudaStream_t st) {
  return launch_ffn_common(false, A, W1, b1, W2, b2, R, M, FfnLn{}, st);
}

// R += FFN(norm2(R)); H1 = norm1_next(R) (H1 == nullptr: not produced).  `scratch` [M,256] bf16 receives norm2(R).
int launch_ffn_tc_ln(float* R, const float* ln2_g, const float* ln2_b, float eps, bf16* scratch, const bf16* W1,
                     const float* b1, const bf16* W2, const float* b2, const float* ln1n_g, const float* ln1n_b,
                     bf16* H1, int M, cudaStream_t st) {
  if (((uintptr_t)ln2_g | (uintptr_t)ln2_b | (uintptr_t)ln1n_g | (uintptr_t)ln1n_b | (uintptr_t)H1) & 15) {
    set_error("ffn_tc_ln: LayerNorm parameters / outputs must be 16-byte aligned");
    return 1;
  }
  if (H1 != nullptr && (ln1n_g == nullptr || ln1n_b == nullptr)) {
    set_error("ffn_tc_ln: H1 requested without the next layer's norm1 parameters");
    return 1;
  }
  FfnLn ln;
  ln.dbg = g_ffn_ln_dbg;
  ln.R = R; ln.g2 = ln2_g; ln.b2 = ln2_b; ln.g1n = ln1n_g; ln.b1n = ln1n_b; ln.A = scratch; ln.H1 = H1; ln.eps = eps;
  return launch_ffn_common(true, scratch, W1, b1, W2, b2, R, M, ln, st);
}

}  // namespace cse
