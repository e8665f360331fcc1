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
//    registers with exp2, P kept in registers as the A operand of P*V (never written to memory).
//    K staged row-major, V read through ldmatrix.trans.  Attention is ~10 % of the path's FLOPs
//    and exp-bound at d=32, so this kernel is written for the SFU/LSU balance, not UMMA tiles.
#include "common.cuh"

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
constexpr int kRowStride = 40;  // bf16 per staged K/V row (32 + 8 pad): conflict-free LDS/ldmatrix

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(kAttnThreads) attention_bf16_kernel(const bf16* __restrict__ qkv,
                                                                      int n, int n_pad,
                                                                      bf16* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);        // [n_pad][40]
  bf16* Vs = Ks + (size_t)n_pad * kRowStride;           // [n_pad][40]
  const int h = blockIdx.x % kHeads;
  const size_t seq = blockIdx.x / kHeads;
  const bf16* base = qkv + seq * n * (3 * kN);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  // stage K and V rows of this head: 4 x 16-byte chunks per row each; pad rows are zero
  for (int i = threadIdx.x; i < n_pad * 4; i += kAttnThreads) {
    const int j = i >> 2, ch = i & 3;
    uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = make_uint4(0u, 0u, 0u, 0u);
    if (j < n) {
      kv = *reinterpret_cast<const uint4*>(base + (size_t)j * (3 * kN) + kN + h * kDh + ch * 8);
      vv = *reinterpret_cast<const uint4*>(base + (size_t)j * (3 * kN) + 2 * kN + h * kDh + ch * 8);
    }
    *reinterpret_cast<uint4*>(Ks + j * kRowStride + ch * 8) = kv;
    *reinterpret_cast<uint4*>(Vs + j * kRowStride + ch * 8) = vv;
  }
  __syncthreads();

  const float sl2 = 0.17677669529663687f * 1.4426950408889634f;  // log2(e)/sqrt(32)
  const int g = lane >> 2, t4 = lane & 3;
  const int n_mt = (n + 15) >> 4;
  for (int mt = wid; mt < n_mt; mt += kAttnThreads / 32) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    // Q fragments: 2 k-steps of 16 over d=32
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int col = h * kDh + ks * 16 + t4 * 2;
      qa[ks][0] = (r0 < n) ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * (3 * kN) + col) : 0u;
      qa[ks][1] = (r1 < n) ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * (3 * kN) + col) : 0u;
      qa[ks][2] = (r0 < n) ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * (3 * kN) + col + 8) : 0u;
      qa[ks][3] = (r1 < n) ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * (3 * kN) + col + 8) : 0u;
    }
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int j0 = 0; j0 < n_pad; j0 += 64) {
      const int nt_cnt = min(8, (n_pad - j0) >> 3);  // 8-key tiles in this block (even number)
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        if (nt < nt_cnt) {
          const bf16* kr = Ks + (j0 + nt * 8 + g) * kRowStride + t4 * 2;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
            mma_bf16_16816(s[nt], qa[ks], b0, b1);
          }
        }
      }
      // scale to log2 domain, mask keys >= n, block row max
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = j0 + nt * 8 + t4 * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool valid = (nt < nt_cnt) && (key + (e & 1) < n);
          s[nt][e] = valid ? s[nt][e] * sl2 : -INFINITY;
        }
        bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
        bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);  // finite: every block has key j0 < n
      const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
      l0 *= c0;
      l1 *= c1;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        o[dt][0] *= c0;
        o[dt][1] *= c0;
        o[dt][2] *= c1;
        o[dt][3] *= c1;
      }
      m0 = mn0;
      m1 = mn1;
      // P = exp2(s - m), row sums in fp32, P packed to bf16 A fragments; O += P V
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {  // 16 keys per step
        if (kt * 2 < nt_cnt) {
          uint32_t pa[4];
          float p[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            p[u][0] = exp2f(s[kt * 2 + u][0] - mn0);
            p[u][1] = exp2f(s[kt * 2 + u][1] - mn0);
            p[u][2] = exp2f(s[kt * 2 + u][2] - mn1);
            p[u][3] = exp2f(s[kt * 2 + u][3] - mn1);
            l0 += p[u][0] + p[u][1];
            l1 += p[u][2] + p[u][3];
          }
          pa[0] = pack_bf16(p[0][0], p[0][1]);
          pa[1] = pack_bf16(p[0][2], p[0][3]);
          pa[2] = pack_bf16(p[1][0], p[1][1]);
          pa[3] = pack_bf16(p[1][2], p[1][3]);
          // V fragments via ldmatrix.trans: matrices (keys 0-7,d0) (keys 8-15,d0) (keys 0-7,d0+8)
          // (keys 8-15,d0+8); lane -> row address of matrix lane/8, row lane%8
          const int key = j0 + kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
          for (int dp = 0; dp < 2; ++dp) {  // d tiles (0,1) then (2,3)
            const bf16* vp = Vs + key * kRowStride + dp * 16 + (lane >> 4) * 8;
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(vp);
            uint32_t v0, v1, v2, v3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                         : "r"(addr));
            mma_bf16_16816(o[dp * 2], pa, v0, v1);
            mma_bf16_16816(o[dp * 2 + 1], pa, v2, v3);
          }
        }
      }
    }
    // finalise: quad-reduce the row sums, normalise, store bf16 pairs
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int col = h * kDh + dt * 8 + t4 * 2;
      if (r0 < n)
        *reinterpret_cast<uint32_t*>(out + (seq * n + r0) * kN + col) = pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
      if (r1 < n)
        *reinterpret_cast<uint32_t*>(out + (seq * n + r1) * kN + col) = pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
    }
  }
}

int launch_attention(const void* qkv, int nseq, int n, int act, void* out, cudaStream_t st) {
  if (nseq <= 0 || n <= 0) return 0;
  if ((long long)nseq * kHeads > 2147483647LL) {
    set_error("attention: nseq=%d exceeds the grid", nseq);
    return 1;
  }
  dim3 grid((unsigned)nseq * kHeads);
  KernelScope prof(kClsAttention, st);  // CTA = (sequence, head); heads of a sequence are adjacent
  if (act == CSE_BF16) {
    const int n_pad = (n + 15) / 16 * 16;
    const size_t smem = (size_t)n_pad * kRowStride * sizeof(bf16) * 2;
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(attention_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      configured = true;
    }
    if (smem > 200 * 1024) {
      set_error("attention: sequence of %d tokens does not fit shared memory", n);
      return 1;
    }
    attention_bf16_kernel<<<grid, kAttnThreads, smem, st>>>((const bf16*)qkv, n, n_pad, (bf16*)out);
    return check_launch("attention_bf16_kernel");
  }
  const size_t smem = (size_t)n * kDh * sizeof(float) * 2;
  static bool configured32 = false;
  if (!configured32) {
    cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured32 = true;
  }
  if (smem > 200 * 1024) {
    set_error("attention: sequence of %d tokens does not fit shared memory (fp32)", n);
    return 1;
  }
  attention_f32_kernel<<<grid, kAttnThreads, smem, st>>>((const float*)qkv, n, (float*)out);
  return check_launch("attention_f32_kernel");
}

}  // namespace cse
