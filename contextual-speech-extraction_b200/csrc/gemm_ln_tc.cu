// gemm_ln_tc.cu — LayerNorm fused into the A-operand producer of a tcgen05 GEMM (CSE_BF16):
//
//   C[M,N] (bf16) = act( LN(R[M,256]; gamma, beta, eps) · W[N,256]^T + bias ),  N % 256 == 0
//
// Reference: the pre-norm sub-block head `src1 = norm1(src); self_att(src1)` (CSE_transformer.py:385-390):
// norm1 -> in_proj (QKV).  Unfused this is a LayerNorm kernel (1 KB read + 512 B write per row) plus a
// GEMM that re-reads the 512 B once per 256-column tile; here the fp32 residual row is read ONCE by eight
// LayerNorm warps, normalised in registers (warp-shuffle statistics, two-pass variance as nn.LayerNorm),
// rounded to bf16 and written straight into the SWIZZLE_128B K-major shared-memory tile the tensor core
// consumes.  K = 256 is the whole row, so the normalised [128 x 256] A tile (64 KB) stays resident while
// the CTA walks the N/256 weight tiles of its rows.
//
// The A tile is single-buffered (shared memory also holds a 4 x 32 KB weight ring and the epilogue
// staging), so the LayerNorm warps run ONE ROW TILE AHEAD IN REGISTERS: while the tensor core works on
// row tile i they load and normalise row tile i+1 (16 rows per warp, 64 packed registers per lane) and
// dump it into shared memory k-block by k-block the moment the last UMMA reading that k-block retires.
// All UMMAs are 128x256x16 (a 128-row UMMA costs >= ~96 cycles however small N is, so 128-wide tiles run
// at ~60 % of the rate: one reason the first version of this kernel lost to the unfused path).  CTAs run in clusters of two on adjacent row tiles and multicast the weight tiles.
// Tiles are dealt to CTA pairs as contiguous runs in (row tile, n-tile) order, so the work is balanced to
// one tile while consecutive tiles share their row tile.
//
// Roles (576 threads; 18 warps are allocated as 20, which leaves 96 registers per thread): warp 0 weight-tile TMA producer | warp 1 single-thread
// tcgen05.mma issuer | warps 2-9 epilogue (tcgen05.ld -> bias/ReLU/bf16 -> swizzled staging -> TMA store) |
// warps 10-17 LayerNorm / A producers (64 of their registers hold the packed row tile).  All mbarrier
// waits are bounded.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

using namespace tc;

int get_tensor_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, uint32_t esize, CUtensorMap* out);  // gemm_tc.cu
int sm_count();

namespace {

constexpr int kLnThreads = 576;
constexpr int kLnBN = 256;
constexpr int kLnKb = 128 * 128;              // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kLnStageBytes = 2 * kLnKb;      // 32 KB: 256 weight rows x 64 k
constexpr int kLnStages = 4;
constexpr int kLnEpiBuf = 32 * 128;           // 4 KB: 32 rows x 64 bf16
constexpr size_t kLnSmem = 1024 + 4 * kLnKb + kLnStages * kLnStageBytes + 8 * kLnEpiBuf + 256;

__global__ void __launch_bounds__(kLnThreads, 1)
gemm_ln_tc_kernel(const float* __restrict__ R, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, int M, int N,
                  int relu) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t sA = smem_base;
  const uint32_t sW = sA + 4 * kLnKb;
  const uint32_t sEpi = sW + kLnStages * kLnStageBytes;
  const uint32_t sBar = sEpi + 8 * kLnEpiBuf;
  const uint32_t bar_wfull = sBar;            // [4]
  const uint32_t bar_wempty = sBar + 32;      // [4]
  const uint32_t bar_afull = sBar + 64;       // [4] k-block of the normalised A tile written
  const uint32_t bar_aempty = sBar + 96;      // [4] k-block no longer read by any UMMA
  const uint32_t bar_tfull = sBar + 128;      // [2]
  const uint32_t bar_tempty = sBar + 144;     // [2]
  const uint32_t tmem_slot = sBar + 160;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_tiles = N / kLnBN;
  const int total_pt = (((M + 127) / 128 + 1) >> 1) * n_tiles;
  // pair p owns a contiguous run of tiles in (row-tile pair, n-tile) order
  const int t_begin = (int)((long long)pair_id * total_pt / npairs);
  const int t_end = (int)((long long)(pair_id + 1) * total_pt / npairs);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kLnStages; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, 2);  // released by the MMA issuers of both CTAs of the pair
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_afull + 8 * i, 8);   // one arrive per LayerNorm warp
      mbar_init(bar_aempty + 8 * i, 1);  // tcgen05.commit after the row tile's last UMMA on this k-block
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 8);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kLnBN);
  fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= weight-tile TMA producer (lane 0 acts; the warp stays convergent) =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int pt = t_begin; pt < t_end; ++pt) {
        const int n0 = (pt % n_tiles) * kLnBN;
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait_spin(bar_wempty + 8 * stage, phase ^ 1u, 1);
          mbar_expect_tx(bar_wfull + 8 * stage, kLnStageBytes);
          // this CTA fetches 128 of the tile's 256 weight rows for both CTAs
          tma_load_2d_mcast(sW + stage * kLnStageBytes + rank * kLnKb, &tmW, bar_wfull + 8 * stage, kb * 64,
                            n0 + rank * 128, (uint16_t)3);
          if (++stage == kLnStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (lane 0 issues; the warp stays convergent) =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kLnBN, 0, 0);
      uint32_t stage = 0, phase = 0, astage = 0, aphase = 0, seg = 0;
      for (int pt = t_begin; pt < t_end; ++pt) {
        const bool new_rows = (pt == t_begin) || (pt % n_tiles == 0);
        const bool last_use = (pt + 1 == t_end) || ((pt + 1) % n_tiles == 0);
        if (new_rows && pt != t_begin) ++seg;
        mbar_wait_spin(bar_tempty + 8 * astage, aphase ^ 1u, 2);
        fence_after();
        const uint32_t d_tmem = tmem_base + astage * kLnBN;
        for (int kb = 0; kb < 4; ++kb) {
          if (new_rows) mbar_wait_spin(bar_afull + 8 * kb, seg & 1u, 3);
          mbar_wait_spin(bar_wfull + 8 * stage, phase, 4);
          fence_after();
          const uint64_t adesc = make_desc(sA + kb * kLnKb, 1024, kLayoutSw128);
          const uint64_t bdesc = make_desc(sW + stage * kLnStageBytes, 1024, kLayoutSw128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_mcast(bar_wempty + 8 * stage, (uint16_t)3);
          if (last_use) umma_commit(bar_aempty + 8 * kb);  // A k-block free for the next row tile
          if (++stage == kLnStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tfull + 8 * astage);
        if (++astage == 2) { astage = 0; aphase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // ================= epilogue warps 2..9 =================
    // warp -> TMEM lane quarter (warp & 3, fixed by hardware) x column half ((warp-2) >> 2)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t stg = sEpi + (warp - 2) * kLnEpiBuf;
    uint4* stg_ptr = reinterpret_cast<uint4*>(smem_al + (stg - smem_base));
    uint32_t astage = 0, aphase = 0, chunk_ctr = 0;
    for (int pt = t_begin; pt < t_end; ++pt) {
      const int m0 = (2 * (pt / n_tiles) + rank) * 128, n0 = (pt % n_tiles) * kLnBN;
      mbar_wait(bar_tfull + 8 * astage, aphase, 5);
      fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch, ++chunk_ctr) {
        const int col_local = half * 128 + ch * 64;
        if (chunk_ctr >= 1) {  // the TMA store of the previous chunk must have drained the staging tile
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int part = 0; part < 2; ++part) {  // 32 columns at a time: these warps run on 96 registers
          float v[32];
          tmem_ld32(tmem_base + lane_off + astage * kLnBN + col_local + part * 32, v);
          if (ch == 1 && part == 1) {  // accumulator fully read by this warp: the TMEM stage may be reused
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * astage);
          }
          if (bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + col_local + part * 32 + i));
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
            hh[0] = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
            hh[1] = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
            hh[2] = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
            hh[3] = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
            stg_ptr[lane * 8 + ((part * 4 + i) ^ (lane & 7))] = u;  // row = lane, SWIZZLE_128B slots: conflict-free
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, stg, n0 + col_local, m0 + q * 32);  // rows past M are clipped by the tensor map
          bulk_commit();
        }
      }
      if (++astage == 2) { astage = 0; aphase ^= 1u; }
    }
    if (lane == 0) bulk_wait_all();  // all output writes complete before the CTA retires
    __syncwarp();
  } else {
    // ================= LayerNorm -> bf16 A tile producers, warps 10..17 (16 rows each) =================
    const int w = warp - 10;
    const int kb = lane >> 3, chunk = lane & 7;  // this lane's 8 channels: k-block, 16-byte slot
    uint32_t seg = 0;
    for (int pt = t_begin; pt < t_end; ++pt) {
      const bool new_rows = (pt == t_begin) || (pt % n_tiles == 0);
      if (!new_rows) continue;
      const int row0 = (2 * (pt / n_tiles) + rank) * 128 + w * 16;
      // pull the row tile this warp will normalise NEXT into L2 while it works on this one
      {
        int npt = (pt / n_tiles + 1) * n_tiles;  // first tile of the next row-tile pair of this run
        if (npt < t_end) {
          const int nrow = (2 * (npt / n_tiles) + rank) * 128 + w * 16 + (lane >> 1);
          if (nrow < M) {
            const char* p = reinterpret_cast<const char*>(R + (size_t)nrow * kN) + (lane & 1) * 512;
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p + i * 128));
          }
        }
      }
      uint4 pk[16];  // 16 rows x this lane's 8 channels, normalised, bf16
#pragma unroll
      for (int b = 0; b < 8; ++b) {  // two rows per step: 64 registers already hold finished rows
        f8 x[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int row = row0 + b * 2 + u;
          if (row < M) {
            x[u] = ld8(R + (size_t)row * kN + lane * 8);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[u].v[i] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const f8 gg = ld8(gamma + lane * 8), bb = ld8(beta + lane * 8);  // L1-resident; deliberately not kept
          ln_row(x[u], gg, bb, eps);
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&pk[b * 2 + u]);
#pragma unroll
          for (int i = 0; i < 4; ++i) hh[i] = __floats2bfloat162_rn(x[u].v[2 * i], x[u].v[2 * i + 1]);
        }
      }
      // dump: each lane group (8 lanes = one k-block) waits only for ITS k-block of the previous row tile
      mbar_wait(bar_aempty + 8 * kb, (seg & 1u) ^ 1u, 6);
      unsigned char* abase = smem_al + (sA - smem_base) + kb * kLnKb;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int trow = w * 16 + i;  // row inside the 128-row tile
        *reinterpret_cast<uint4*>(abase + trow * 128 + ((chunk ^ (trow & 7)) << 4)) = pk[i];
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (chunk == 0) mbar_arrive(bar_afull + 8 * kb);
      ++seg;
    }
  }

  fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA retires while the peer may still multicast into it
  if (warp == 1) {
    fence_after();
    tmem_dealloc(tmem_base, 2 * kLnBN);
  }
}

}  // namespace

int launch_gemm_ln_tc(const float* R, const float* gamma, const float* beta, float eps, const bf16* W,
                      const float* bias, bf16* C, int ldc, int M, int N, int relu, cudaStream_t st) {
  if (M <= 0) return 0;
  if (N % kLnBN != 0 || ldc % 8 != 0) {
    set_error("gemm_ln_tc: need N %% 256 == 0 and ldc %% 8 == 0 (N=%d ldc=%d)", N, ldc);
    return 1;
  }
  if (((uintptr_t)R | (uintptr_t)W | (uintptr_t)C | (uintptr_t)gamma | (uintptr_t)beta) & 15) {
    set_error("gemm_ln_tc: operands must be 16-byte aligned");
    return 1;
  }
  static DeviceOnce once;
  if (!once.configured_on_this_device()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kLnSmem);
    if (e != cudaSuccess) {
      set_error("gemm_ln_tc: cudaFuncSetAttribute(%zu B smem) failed: %s", kLnSmem, cudaGetErrorString(e));
      return 1;
    }
    once.mark_configured();
  }
  CUtensorMap tmW, tmC;
  if (get_tensor_map(W, (uint64_t)N, (uint64_t)kN, (uint64_t)kN, 128, 64, 2, &tmW)) return 1;  // half tile per CTA
  if (get_tensor_map(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, 64, 2, &tmC)) return 1;
  const int pair_tiles = ((ceil_div(M, 128) + 1) / 2) * (N / kLnBN);
  const int max_pairs = sm_count() / 2;
  const int grid = 2 * (pair_tiles < max_pairs ? pair_tiles : max_pairs);
  KernelScope prof(kClsGemmTc, st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kLnThreads);
  cfg.dynamicSmemBytes = kLnSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_ln_tc_kernel, R, gamma, beta, eps, tmW, tmC, bias, M, N, relu);
  if (le != cudaSuccess) {
    set_error("gemm_ln_tc_kernel cluster launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("gemm_ln_tc_kernel");
}

}  // namespace cse
