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
// One CTA owns a 128-row tile; the hidden dimension is walked in 8 chunks of 128 units:
//   G1_j : Hacc[j&1] (TMEM, 128 cols) = A(128x256) · W1_j(128x256)^T      16 UMMAs 128x128x16
//   E1_j : Hacc[j&1] -> +b1 -> ReLU -> bf16 pairs -> back into the SAME TMEM columns
//   G2_j : Y (TMEM, 256 cols) += H_j(128x128, A operand from TMEM) · W2[:, chunk j]^T
//                                                                            8 UMMAs 128x256x16
// issued as G1_0 G1_1 G2_0 G1_2 G2_1 ... G1_7 G2_6 G2_7 so that E1_j overlaps G1_{j+1} (and G2_{j-1}).
// The tensor pipe executes in issue order, so G1_{j+2} cannot overwrite Hacc[j&1] before G2_j has read
// it: the hidden buffers need no "empty" barriers.  After G2_7 the epilogue adds b2 and folds Y into
// the fp32 residual stream with TMA reduce-add (the stream is never loaded into the SM).
// TMEM: Y 256 + Hacc 2 x 128 = 512 columns.  Shared memory: A 64 KB (4 k-blocks, resident for the
// tile) + a 4-stage x 32 KB weight ring + 32 KB epilogue staging.  Every op is two ring stages: G1
// stages hold two k-blocks of the W1 chunk, G2 stages one k-block of all 256 W2 rows.  CTAs run in
// clusters of two on adjacent row tiles: every weight box is fetched from L2 once per pair (each CTA
// loads half of it and multicasts).
//
// Roles (448 threads): warp 0 TMA producer | warp 1 single-thread tcgen05.mma issuer | warps 2-9
// hidden-chunk epilogue E1 (TMEM lane quarter = warp & 3, column half = (warp-2) >> 2) | warps 10-13
// output epilogue Y -> R (one per lane quarter, two staging chunks each), so E1 of the next tile never
// waits behind the residual update of the previous one.  All mbarrier waits are bounded.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

using namespace tc;

int get_tensor_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, uint32_t esize, CUtensorMap* out);  // gemm_tc.cu
int sm_count();

namespace {

constexpr int kFfnThreads = 448;
constexpr int kD = 256;        // d_model
constexpr int kH = 1024;       // d_ffn
constexpr int kHC = 128;       // hidden units per chunk
constexpr int kChunks = kH / kHC;
constexpr int kKb = 128 * 128;          // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kStageBytes = 2 * kKb;    // 32 KB
#ifndef FFN_STAGES
#define FFN_STAGES 4
#endif
constexpr int kStages = FFN_STAGES;
constexpr int kABytes = 4 * kKb;        // resident A tile: 4 k-blocks
constexpr int kStgBytes = 32 * 128;     // output staging chunk: 32 rows x 32 fp32 (two per output warp)
constexpr int kBiasBytes = kD * 4;      // b2 (b1 is read through L1)
constexpr int kOps = 2 * kChunks;
constexpr size_t kFfnSmem = 1024 + kABytes + kStages * kStageBytes + 8 * kStgBytes + kBiasBytes + 256;

// op i of the per-tile schedule G1_0 G1_1 G2_0 G1_2 G2_1 ... G1_7 G2_6 G2_7
__device__ __forceinline__ void decode_op(int i, bool& is_g1, int& j) {
  if (i == 0) { is_g1 = true; j = 0; }
  else if (i == kOps - 1) { is_g1 = false; j = kChunks - 1; }
  else if (i & 1) { is_g1 = true; j = (i + 1) >> 1; }
  else { is_g1 = false; j = (i >> 1) - 1; }
}

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
  const uint32_t bar_wfull = sBar;              // [4]
  const uint32_t bar_wempty = sBar + 32;        // [4]
  const uint32_t bar_afull = sBar + 64;         // [2] k-blocks {0,1} / {2,3} of the resident A tile
  const uint32_t bar_aempty = sBar + 80;        // [2]
  const uint32_t bar_hfull = sBar + 96;         // [2] G1_j complete: Hacc[b] holds fp32 pre-activations
  const uint32_t bar_pfull = sBar + 112;        // [2] E1_j complete: Hacc[b] holds the bf16 hidden chunk
  const uint32_t bar_yfull = sBar + 128;
  const uint32_t bar_yempty = sBar + 136;
  const uint32_t tmem_slot = sBar + 144;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_pairs = ((M + 127) / 128 + 1) >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, 2);  // released by the MMA issuers of both CTAs of the pair
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_afull + 8 * i, 1);
      mbar_init(bar_aempty + 8 * i, 1);
      mbar_init(bar_hfull + 8 * i, 1);
      mbar_init(bar_pfull + 8 * i, 8);   // one arrive per epilogue warp
    }
    mbar_init(bar_yfull, 1);
    mbar_init(bar_yempty, 4);   // one arrive per output-epilogue warp
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < kD; i += kFfnThreads) s_b2[i] = b2[i];
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_y = tmem_base;            // columns [0, 256)
  const uint32_t tmem_h = tmem_base + 256;      // columns [256, 384) and [384, 512)

  if (warp == 0) {
    // ================= TMA producer (lane 0 acts; the warp stays convergent) =================
    uint32_t stage = 0, wphase = 0;
    int it = 0;
    for (int mp = pair_id; mp < m_pairs; mp += npairs, ++it) {
      const int m0 = (2 * mp + rank) * 128;
      if (lane == 0 && mp + npairs < m_pairs) {  // next tile's A -> L2 while this tile computes
        for (int kb = 0; kb < 4; ++kb) tma_prefetch_l2_2d(&tmA, kb * 64, (2 * (mp + npairs) + rank) * 128);
      }
      for (int i = 0; i < kOps; ++i) {
        bool is_g1;
        int j;
        decode_op(i, is_g1, j);
        if (lane == 0) {
          for (int t = 0; t < 2; ++t) {
            if (is_g1 && j == 0) {  // the previous tile's last G1 has finished with these two k-blocks of A
              mbar_wait_spin(bar_aempty + 8 * t, ((uint32_t)it & 1u) ^ 1u, 1);
              mbar_expect_tx(bar_afull + 8 * t, 2 * kKb);
              tma_load_2d(sA + (2 * t) * kKb, &tmA, bar_afull + 8 * t, (2 * t) * 64, m0);
              tma_load_2d(sA + (2 * t + 1) * kKb, &tmA, bar_afull + 8 * t, (2 * t + 1) * 64, m0);
            }
            mbar_wait_spin(bar_wempty + 8 * stage, wphase ^ 1u, 2);
            mbar_expect_tx(bar_wfull + 8 * stage, kStageBytes);
            const uint32_t dst = sW + stage * kStageBytes;
            if (is_g1) {
              // W1 rows [j*128, +128) x k-blocks 2t, 2t+1; this CTA fetches 64 of the rows for both CTAs
              tma_load_2d_mcast(dst + rank * (kKb / 2), &tmW1, bar_wfull + 8 * stage, (2 * t) * 64,
                                j * kHC + rank * 64, (uint16_t)3);
              tma_load_2d_mcast(dst + kKb + rank * (kKb / 2), &tmW1, bar_wfull + 8 * stage, (2 * t + 1) * 64,
                                j * kHC + rank * 64, (uint16_t)3);
            } else {
              // W2 rows [0, 256) x hidden k-block j*128 + t*64; this CTA fetches 128 of the rows
              tma_load_2d_mcast(dst + rank * kKb, &tmW2, bar_wfull + 8 * stage, j * kHC + t * 64, rank * 128,
                                (uint16_t)3);
            }
            if (++stage == kStages) { stage = 0; wphase ^= 1u; }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (lane 0 issues; the warp stays convergent) =================
    constexpr uint32_t idesc_g1 = make_idesc_bf16(128, kHC, 0, 0);
    constexpr uint32_t idesc_g2 = make_idesc_bf16(128, kD, 0, 0);
    uint32_t stage = 0, wphase = 0;
    int it = 0;
    for (int mp = pair_id; mp < m_pairs; mp += npairs, ++it) {
      for (int i = 0; i < kOps; ++i) {
        bool is_g1;
        int j;
        decode_op(i, is_g1, j);
        if (lane == 0) {
          const int b = j & 1;
          const uint32_t use = (uint32_t)(it * (kChunks / 2) + (j >> 1));  // prior uses of buffer b
          if (is_g1) {
            // Hacc[b] is free: G2_{j-2}, its last reader, was issued earlier and the pipe runs in order
            const uint32_t d_tmem = tmem_h + b * kHC;
            for (int t = 0; t < 2; ++t) {
              if (j == 0) mbar_wait_spin(bar_afull + 8 * t, (uint32_t)it & 1u, 5);
              mbar_wait_spin(bar_wfull + 8 * stage, wphase, 6);
              fence_after();
#if !defined(FFN_DBG_NOMMA) && !defined(FFN_DBG_NOG1)
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t adesc = make_desc(sA + (2 * t + (kk >> 2)) * kKb, 1024, kLayoutSw128);
                const uint64_t bdesc = make_desc(sW + stage * kStageBytes + (kk >> 2) * kKb, 1024, kLayoutSw128);
                umma_bf16(d_tmem, adesc + 2 * (kk & 3), bdesc + 2 * (kk & 3), idesc_g1, (t | kk) != 0 ? 1u : 0u);
              }
#endif
              umma_commit_mcast(bar_wempty + 8 * stage, (uint16_t)3);
              if (j == kChunks - 1) umma_commit(bar_aempty + 8 * t);  // A k-blocks free for the next tile
              if (++stage == kStages) { stage = 0; wphase ^= 1u; }
            }
            umma_commit(bar_hfull + 8 * b);
          } else {
            if (j == 0) mbar_wait_spin(bar_yempty, ((uint32_t)it & 1u) ^ 1u, 7);  // previous tile's Y drained
            mbar_wait_spin(bar_pfull + 8 * b, use & 1u, 8);                      // E1_j has written H_j
            fence_after();
            for (int t = 0; t < 2; ++t) {
              mbar_wait_spin(bar_wfull + 8 * stage, wphase, 9);
              fence_after();
              // hidden units [64t, 64t+64) of the chunk = packed bf16 pairs in TMEM columns [64t, 64t+32)
              // of Hacc[b] (each epilogue column-half packs into its own columns)
              const uint32_t a_tmem = tmem_h + b * kHC + t * 64;
              const uint64_t bdesc = make_desc(sW + stage * kStageBytes, 1024, kLayoutSw128);
#if !defined(FFN_DBG_NOMMA) && !defined(FFN_DBG_NOG2)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ts(tmem_y, a_tmem + 8 * k, bdesc + 2 * k, idesc_g2, (j | t | k) != 0 ? 1u : 0u);
#endif
              umma_commit_mcast(bar_wempty + 8 * stage, (uint16_t)3);
              if (++stage == kStages) { stage = 0; wphase ^= 1u; }
            }
            if (j == kChunks - 1) umma_commit(bar_yfull);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // ================= hidden-chunk epilogue E1, warps 2..9 =================
    const int q = warp & 3;            // TMEM lane quarter (fixed by hardware: warp id % 4)
    const int h = (warp - 2) >> 2;     // column half
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int mp = pair_id; mp < m_pairs; mp += npairs, ++it) {
      for (int j = 0; j < kChunks; ++j) {
        const int b = j & 1;
        const uint32_t use = (uint32_t)(it * (kChunks / 2) + (j >> 1));
        mbar_wait(bar_hfull + 8 * b, use & 1u, 11);
        fence_after();
#ifndef FFN_DBG_NOE1
        float v[64];
        tmem_ld64(tmem_h + lane_off + b * kHC + h * 64, v);
        // bias + ReLU, pack to bf16 pairs (low half = even hidden unit) and write them over this warp's
        // own first 32 fp32 columns: the A operand of G2_j, read by the tensor core straight from TMEM
        uint32_t pk[32];
        const float4* bsrc = reinterpret_cast<const float4*>(b1 + j * kHC + h * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 t = __ldg(bsrc + i);  // lane-uniform, L1-resident
          const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(v[4 * i] + t.x, 0.f), fmaxf(v[4 * i + 1] + t.y, 0.f));
          const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(v[4 * i + 2] + t.z, 0.f), fmaxf(v[4 * i + 3] + t.w, 0.f));
          pk[2 * i] = *reinterpret_cast<const uint32_t*>(&lo);
          pk[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&hi);
        }
        tmem_st32(tmem_h + lane_off + b * kHC + h * 64, pk);
#endif
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pfull + 8 * b);
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
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int col = c * 32;
        float y[32];
        tmem_ld32(tmem_y + lane_off + col, y);
        if (c == 7) {  // Y fully read by this warp: the next tile's G2_0 may overwrite it
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_yempty);
        }
#ifndef FFN_DBG_NOFINAL
        const uint32_t stg = stg0 + (c & 1) * kStgBytes;
        uint4* stg_ptr = reinterpret_cast<uint4*>(smem_al + (stg - smem_base));
        if (lane == 0) bulk_wait_read<1>();  // the reduce-add issued two chunks ago has drained this staging tile
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = *reinterpret_cast<const float4*>(s_b2 + col + 4 * i);
          uint4 u;
          u.x = __float_as_uint(y[4 * i] + t.x);
          u.y = __float_as_uint(y[4 * i + 1] + t.y);
          u.z = __float_as_uint(y[4 * i + 2] + t.z);
          u.w = __float_as_uint(y[4 * i + 3] + t.w);
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
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFfnSmem);
    if (e != cudaSuccess) {
      set_error("ffn_tc: cudaFuncSetAttribute(%zu B smem) failed: %s", kFfnSmem, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  CUtensorMap tmA, tmW1, tmW2, tmR;
  if (get_tensor_map(A, (uint64_t)M, kD, kD, 128, 64, 2, &tmA)) return 1;
  if (get_tensor_map(W1, kH, kD, kD, 64, 64, 2, &tmW1)) return 1;    // half of a 128-row k-block per CTA
  if (get_tensor_map(W2, kD, kH, kH, 128, 64, 2, &tmW2)) return 1;   // half of the 256 rows per CTA
  if (get_tensor_map(R, (uint64_t)M, kD, kD, 32, 32, 4, &tmR)) return 1;
  const int m_pairs = (ceil_div(M, 128) + 1) / 2;
  const int max_pairs = sm_count() / 2;
  const int grid = 2 * (m_pairs < max_pairs ? m_pairs : max_pairs);
  KernelScope prof(kClsGemmTc, st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kFfnThreads);
  cfg.dynamicSmemBytes = kFfnSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, ffn_tc_kernel, tmA, tmW1, tmW2, tmR, b1, b2, M);
  if (le != cudaSuccess) {
    set_error("ffn_tc_kernel cluster launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("ffn_tc_kernel");
}

}  // namespace cse
