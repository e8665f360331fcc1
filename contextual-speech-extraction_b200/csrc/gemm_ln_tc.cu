// gemm_ln_tc.cu — LayerNorm fused into the A-operand producer of a tcgen05 GEMM.
//
//   C[M,N] (bf16) = act( LN(R[M,256]; gamma, beta, eps) · W[N,256]^T + bias )
//
// Reference: the pre-norm sub-blocks `src1 = norm1(src); self_att(src1)` and
// `src1 = norm2(src); pos_ffn(src1)` (CSE_transformer.py:385-411): norm1 -> in_proj (QKV) and
// norm2 -> ffn.0 (+ReLU).  Unfused, each pair costs a LayerNorm kernel (1 KB read + 512 B write per
// row) plus a GEMM that re-reads the 512 B; here the fp32 residual row is read ONCE, normalised in
// registers (warp-shuffle statistics, two-pass variance as nn.LayerNorm), rounded to bf16 and written
// straight into the SWIZZLE_128B K-major shared-memory tile the tensor core consumes.  K = 256 is the
// whole row, so the normalised [128 x 256] A tile (64 KB) stays resident while all N/128 weight
// tiles stream past it; A is double-buffered so the LayerNorm of the next 128 rows overlaps the MMAs
// of the current ones.
//
// Roles (448 threads): warp 0 W-tile TMA producer | warp 1 single-thread tcgen05.mma issuer |
// warps 2-5 LayerNorm/A producers (32 rows each, 8 rows = 8 KB in flight per warp) | warps 6-13
// epilogue (TMEM -> smem-staged bias/ReLU/bf16 -> swizzled staging -> coalesced global stores).
// With K = 256 every output element costs only 256 MACs, so the kernel is epilogue-heavy: eight
// epilogue warps (two per TMEM lane quarter, one 64-column chunk each per tile) keep up with the MMAs.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cse {

using namespace tc;

int get_tensor_map(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, uint32_t esize, CUtensorMap* out);  // gemm_tc.cu
int sm_count();

constexpr int kLnThreads = 448;
constexpr int kLnWarps = 4;   // LayerNorm producer warps
constexpr int kLnEpiWarps = 8;  // epilogue warps
constexpr int kLnBN = 128;                  // weight tile rows (output columns) per MMA tile
constexpr int kLnKb = 4;                    // K = 256 = 4 k-blocks of 64
constexpr int kLnABuf = kLnKb * 128 * 128;  // 64 KB: [4 k-blocks][128 rows][128 B]
constexpr int kLnBStage = kLnBN * 128;      // 16 KB
constexpr int kLnBStages = 3;
constexpr int kLnEpiBuf = 32 * 128;         // 4 KB: 32 rows x 64 bf16
constexpr int kLnBiasMax = 1024;            // bias[N] staged in shared memory (N <= 1024)
constexpr int kLnEpi = kLnEpiWarps * kLnEpiBuf + kLnBiasMax * 4;  // one staging chunk per epilogue warp + bias
constexpr size_t kLnSmem = 1024 + 2 * (size_t)kLnABuf + (size_t)kLnBStages * kLnBStage + kLnEpi + 256;

__global__ void __launch_bounds__(kLnThreads, 1)
gemm_ln_tc_kernel(const float* __restrict__ R, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, const __grid_constant__ CUtensorMap tmW,
                  bf16* __restrict__ Cout, int ldc, const float* __restrict__ bias, int M, int N,
                  int relu) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + 2 * kLnABuf;
  const uint32_t sEpi = sB + kLnBStages * kLnBStage;
  const uint32_t sBar = sEpi + kLnEpi;
  const uint32_t bar_afull = sBar;            // [2]
  const uint32_t bar_aempty = sBar + 16;      // [2]
  const uint32_t bar_bfull = sBar + 32;       // [stages <= 4]
  const uint32_t bar_bempty = sBar + 64;      // [stages <= 4]
  const uint32_t bar_tfull = sBar + 96;       // [2]
  const uint32_t bar_tempty = sBar + 112;     // [2]
  const uint32_t tmem_slot = sBar + 128;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + 127) / 128, n_tiles = N / kLnBN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_afull + 8 * i, kLnWarps);   // one arrive per LayerNorm warp
      mbar_init(bar_aempty + 8 * i, 1);  // tcgen05.commit after the tile's last MMA
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, kLnEpiWarps);  // one arrive per epilogue warp
    }
    for (int i = 0; i < kLnBStages; ++i) {
      mbar_init(bar_bfull + 8 * i, 1);
      mbar_init(bar_bempty + 8 * i, 1);
    }
    mbar_fence_init();
  }
  // bias[N] -> shared memory once per CTA: the epilogue's per-chunk global bias loads were its
  // largest stall (ncu source view, profiles/)
  float* s_bias = reinterpret_cast<float*>(smem_al + (sEpi + kLnEpiWarps * kLnEpiBuf - smem_base));
  if (bias != nullptr)
    for (int i = threadIdx.x; i < N; i += kLnThreads) s_bias[i] = bias[i];
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kLnBN);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= weight-tile TMA producer =================
    uint32_t stage = 0, phase = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      for (int nt = 0; nt < n_tiles; ++nt) {
        for (int kb = 0; kb < kLnKb; ++kb) {
          if (lane == 0) {
            mbar_wait(bar_bempty + 8 * stage, phase ^ 1u, 1);
            mbar_expect_tx(bar_bfull + 8 * stage, kLnBStage);
            tma_load_2d(sB + stage * kLnBStage, &tmW, bar_bfull + 8 * stage, kb * 64, nt * kLnBN);
          }
          __syncwarp();
          if (++stage == kLnBStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_bf16(128, kLnBN, 0, 0);
    uint32_t stage = 0, phase = 0, astage = 0, aphase = 0;
    int j = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++j) {
      const int abuf = j & 1;
      if (lane == 0) {
        mbar_wait(bar_afull + 8 * abuf, (uint32_t)(j >> 1) & 1u, 2);
        fence_after();
      }
      __syncwarp();
      for (int nt = 0; nt < n_tiles; ++nt) {
        if (lane == 0) {
          mbar_wait(bar_tempty + 8 * astage, aphase ^ 1u, 3);
          fence_after();
          const uint32_t d_tmem = tmem_base + astage * kLnBN;
          for (int kb = 0; kb < kLnKb; ++kb) {
            mbar_wait(bar_bfull + 8 * stage, phase, 4);
            fence_after();
            const uint64_t adesc = make_desc(sA + abuf * kLnABuf + kb * (128 * 128), 1024, kLayoutSw128);
            const uint64_t bdesc = make_desc(sB + stage * kLnBStage, 1024, kLayoutSw128);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(bar_bempty + 8 * stage);
            if (++stage == kLnBStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(bar_tfull + 8 * astage);
          if (nt == n_tiles - 1) umma_commit(bar_aempty + 8 * abuf);  // A tile fully consumed
        } else {
          for (int kb = 0; kb < kLnKb; ++kb)
            if (++stage == kLnBStages) { stage = 0; phase ^= 1u; }
        }
        __syncwarp();
        if (++astage == 2) { astage = 0; aphase ^= 1u; }
      }
    }
  } else if (warp < 2 + kLnWarps) {
    // ================= LayerNorm -> bf16 A tile producers (warps 2-5, 32 rows each) =================
    const int w = warp - 2;
    const f8 gg = ld8(gamma + lane * 8), bb = ld8(beta + lane * 8);
    const int kb = lane >> 3, chunk = lane & 7;  // this lane's 8 channels: k-block, 16-byte chunk
    int j = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++j) {
      const int abuf = j & 1;
      mbar_wait(bar_aempty + 8 * abuf, ((uint32_t)(j >> 1) & 1u) ^ 1u, 5);
      unsigned char* abase = smem_al + (sA - smem_base) + abuf * kLnABuf + kb * (128 * 128);
      const int row0 = mt * 128 + w * 32;
#pragma unroll 1
      for (int rr = 0; rr < 32; rr += 8) {
        f8 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // eight rows (8 KB) in flight per warp
          const int row = row0 + rr + u;
          if (row < M) {
            x[u] = ld8(R + (size_t)row * kN + lane * 8);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[u].v[i] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          ln_row(x[u], gg, bb, eps);
          const int trow = w * 32 + rr + u;  // row inside the 128-row tile
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x[u].v[2 * i], x[u].v[2 * i + 1]);
          *reinterpret_cast<uint4*>(abase + trow * 128 + ((chunk ^ (trow & 7)) << 4)) = pk;
        }
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull + 8 * abuf);
    }
  } else {
    // ================= epilogue warps 6-13 =================
    const int quarter = warp & 3;
    const int half = (warp - 2 - kLnWarps) >> 2;
    const uint32_t stg0 = sEpi + (warp - 2 - kLnWarps) * kLnEpiBuf;
    unsigned char* stg0_ptr = smem_al + (stg0 - smem_base);
    uint32_t astage = 0, aphase = 0, chunk_ctr = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      const int row_base = mt * 128 + quarter * 32;
      for (int nt = 0; nt < n_tiles; ++nt) {
        mbar_wait(bar_tfull + 8 * astage, aphase, 6);
        fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + astage * kLnBN;
#pragma unroll 1
        for (int ch = half; ch <= half; ++ch, ++chunk_ctr) {  // this warp's 64-column half of the tile
          const int col0 = nt * kLnBN + ch * 64;
          const uint32_t buf = 0;
          float v[64];
          tmem_ld64(t_row + ch * 64, v);
          if (bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + col0 + i);  // smem broadcast
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          uint4* stg = reinterpret_cast<uint4*>(stg0_ptr + buf * kLnEpiBuf);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
            h[0] = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
            h[1] = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
            h[2] = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
            h[3] = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
            stg[lane * 8 + (i ^ (lane & 7))] = u;
          }
          // Read the chunk back with lanes running along the row (8 lanes x 16 B = one 128-byte line,
          // 4 rows per instruction) and store straight to global memory.  Measured: TMA bulk stores
          // of 4 KB chunks take ~3 us to release their shared-memory source, which capped this
          // epilogue at ~11 GB/s per SM; plain coalesced stores have no such wait.
          __syncwarp();
          {
            const int c = lane & 7;
            bf16* cbase = Cout + (size_t)col0 + c * 8;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int rr = jj * 4 + (lane >> 3);
              const uint4 x = stg[rr * 8 + (c ^ (rr & 7))];
              const int grow = row_base + rr;
              if (grow < M) *reinterpret_cast<uint4*>(cbase + (size_t)grow * ldc) = x;
            }
          }
          __syncwarp();  // staging buffer `buf` is reused two chunks later
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * astage);
        if (++astage == 2) { astage = 0; aphase ^= 1u; }
      }
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc(tmem_base, 2 * kLnBN);
  }
}

int launch_gemm_ln_tc(const float* R, const float* gamma, const float* beta, float eps, const bf16* W,
                      const float* bias, bf16* C, int ldc, int M, int N, int relu, cudaStream_t st) {
  if (M <= 0) return 0;
  if (N % kLnBN != 0 || ldc % 8 != 0 || N > kLnBiasMax) {
    set_error("gemm_ln_tc: need N %% 128 == 0, N <= 1024 and ldc %% 8 == 0 (N=%d ldc=%d)", N, ldc);
    return 1;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kLnSmem);
    if (e != cudaSuccess) {
      set_error("gemm_ln_tc: cudaFuncSetAttribute(%zu B smem) failed: %s", kLnSmem, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  CUtensorMap tmW;
  if (get_tensor_map(W, (uint64_t)N, (uint64_t)kN, (uint64_t)kN, kLnBN, 64, 2, &tmW)) return 1;
  const int m_tiles = ceil_div(M, 128);
  const int grid = m_tiles < sm_count() ? m_tiles : sm_count();
  KernelScope prof(kClsGemmTc, st);
  gemm_ln_tc_kernel<<<grid, kLnThreads, kLnSmem, st>>>(R, gamma, beta, eps, tmW, C, ldc, bias, M, N, relu);
  return check_launch("gemm_ln_tc_kernel");
}

}  // namespace cse
