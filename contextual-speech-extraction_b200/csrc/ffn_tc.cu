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

// G2 consumes the four 64-unit pieces of a hidden chunk in the order E1 finishes them
__device__ __forceinline__ int g2_piece(int s) { return s; }

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR,
              const float* __restrict__ b1, const float* __restrict__ b2, int M) {
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
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_pairs = ((M + 127) / 128 + 1) >> 1;

  if (threadIdx.x == 0) {
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
  for (int i = threadIdx.x; i < kD; i += kFfnThreads) s_b2[i] = b2[i];  // parameter: not produced by the previous kernel
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
    for (int mp = pair_id; mp < m_pairs; mp += npairs, ++it) {
      const int m0 = (2 * mp + rank) * 128;
      if (lane == 0) {
        if (mp + npairs < m_pairs) {  // next tile's A -> L2 while this tile computes
          for (int kb = 0; kb < 4; ++kb) tma_prefetch_l2_2d(&tmA, kb * 64, (2 * (mp + npairs) + rank) * 128);
        }
        for (int c = 0; c < kChunks; ++c) {
          for (int s = 0; s < 8; ++s) {  // stages 0..3: W1 k-blocks of G1_c; 4..7: W2 pieces of G2_c
            if (c == 0 && s < 4) {  // the previous tile's last G1 has finished with this k-block of A
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
    for (int mp = pair_id; mp < m_pairs; mp += npairs, ++it) {
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
    int it = 0;
    for (int mp = pair_id; mp < m_pairs; mp += npairs, ++it) {
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
  } else {
    // ================= output epilogue Y + b2 -> R, warps 10..13 =================
    // fp32 residual stream updated in place by TMA reduce-add; rows past M are clipped by the tensor map
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t stg0 = sStg + (warp - 10) * 2 * kStgBytes;  // two staging chunks per warp
    int it = 0;
    for (int mp = pair_id; mp < m_pairs; mp += npairs, ++it) {
      const int row_base = (2 * mp + rank) * 128 + q * 32;
      mbar_wait(bar_yfull, (uint32_t)it & 1u, 13);
      fence_after();
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
    }
    if (lane == 0) bulk_wait_all();  // all residual updates complete before the CTA retires
    __syncwarp();
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

int launch_ffn_tc(const bf16* A, const bf16* W1, const float* b1, const bf16* W2, const float* b2, float* R,
                  int M, cudaStream_t st) {
  if (M <= 0) return 0;
  if (((uintptr_t)A | (uintptr_t)W1 | (uintptr_t)W2 | (uintptr_t)R | (uintptr_t)b1 | (uintptr_t)b2) & 15) {
    set_error("ffn_tc: operands must be 16-byte aligned");
    return 1;
  }
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFfnSmem);
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
  KernelScope prof(kClsGemmTc, st);
  cudaError_t le = launch_pdl(ffn_tc_kernel, dim3(grid), dim3(kFfnThreads), kFfnSmem, st, 2, tmA, tmW1, tmW2, tmR, b1, b2, M);
  if (le != cudaSuccess) {
    set_error("ffn_tc_kernel cluster launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("ffn_tc_kernel");
}

}  // namespace cse
