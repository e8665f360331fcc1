// head.cu — mask-estimation tail: PReLU + overlap-add, tanh/sigmoid gate, mask*mix_w and the
// ConvTranspose1d overlap-add decoder (HBM-bound kernels).
//
// Reference: Dual_Path_Model_CSE.forward tail (ContSep.py:244-266), _over_add (ContSep.py:337-370),
// Sepformer.forward mask/decode/length fix (ContSep.py:79-95; ContExt.py:116-129),
// speechbrain Decoder = nn.ConvTranspose1d(256,1,16,stride=8,bias=False).
#include <cstdlib>

#include "common.cuh"
#include "mma_sync.cuh"

namespace cse {

// U[b,l,:] = prelu(X[b,s1,k1,:]) + prelu(X[b,s1-1,k1+P,:]),  p = l+P, s1 = p/P, k1 = p - s1*P.
// (Every kept frame is covered by exactly two chunks because gap >= 1, ContSep.py:287.)
// The 1x1 conv2d that the reference applies BEFORE the overlap-add is linear and position-wise,
// so it commutes with it (bias counted twice): this halves the conv2d GEMM rows.
template <typename T>
__global__ void __launch_bounds__(256) prelu_ola_kernel(const float* __restrict__ X,
                                                        const float* __restrict__ prelu, int S,
                                                        int L, size_t rows, T* __restrict__ U) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  const float a = prelu[0];
  for (size_t r = warp; r < rows; r += nwarps) {
    const int l = (int)(r % L);
    const int b = (int)(r / L);
    const int p = l + kP;
    const int s1 = p / kP, k1 = p - s1 * kP;
    const f8 x1 = ld8(X + (((size_t)b * S + s1) * kK + k1) * kN + lane * 8);
    const f8 x2 = ld8(X + (((size_t)b * S + s1 - 1) * kK + k1 + kP) * kN + lane * 8);
    f8 u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y1 = x1.v[i] >= 0.f ? x1.v[i] : a * x1.v[i];
      const float y2 = x2.v[i] >= 0.f ? x2.v[i] : a * x2.v[i];
      u.v[i] = y1 + y2;
    }
    st8(U + r * kN + lane * 8, u);
  }
}

int launch_prelu_ola(const float* X, const float* prelu, int B, int S, int L, int act, void* U,
                     cudaStream_t st) {
  const size_t rows = (size_t)B * L;
  const int grid = (int)min((size_t)148 * 8, (rows + 7) / 8);
  if (act == CSE_BF16)
    prelu_ola_kernel<bf16><<<grid, 256, 0, st>>>(X, prelu, S, L, rows, (bf16*)U);
  else
    prelu_ola_kernel<float><<<grid, 256, 0, st>>>(X, prelu, S, L, rows, (float*)U);
  return check_launch("prelu_ola_kernel");
}

// out = tanh(o) * sigmoid(g), 8 elements per thread.  bf16 mode: one MUFU.TANH per function (sigmoid(x) =
// 0.5 + 0.5 tanh(x / 2); tanh.approx is good to 2^-11, the bf16 result to 2^-9) — libm tanhf + expf + a division
// made the kernel XU-bound (59 % XU, 0.44 of the HBM roofline).  fp32 mode keeps the libm functions.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T>
__global__ void __launch_bounds__(256) gate_kernel(const T* o, const T* g,
                                                   size_t n8, T* out) {  // out may alias o
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
    const f8 a = ld8(o + i * 8), b = ld8(g + i * 8);
    f8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if constexpr (sizeof(T) == 2) r.v[k] = tanh_approx(a.v[k]) * fmaf(0.5f, tanh_approx(0.5f * b.v[k]), 0.5f);
      else r.v[k] = tanhf(a.v[k]) * (1.0f / (1.0f + expf(-b.v[k])));
    }
    st8(out + i * 8, r);
  }
}

int launch_gate(const void* o, const void* g, size_t n, int act, void* out, cudaStream_t st) {
  if (n % 8 != 0) {
    set_error("gate: element count %zu not a multiple of 8", n);
    return 1;
  }
  const size_t n8 = n / 8;
  const int grid = (int)min((size_t)148 * 8, (n8 + 255) / 256);
  if (act == CSE_BF16)
    gate_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)o, (const bf16*)g, n8, (bf16*)out);
  else
    gate_kernel<float><<<grid, 256, 0, st>>>((const float*)o, (const float*)g, n8, (float*)out);
  return check_launch("gate_kernel");
}

// frames[r, k] = sum_n relu(mask_pre[r,n]) * E[b,l,n] * dec_w[n,k],  r = (b*L + l)*n_masks + s.
// One warp per row; the 16 tap sums are reduced with a butterfly that halves the live values at
// each shuffle step (16 -> 8 -> 4 -> 2 -> 1 per lane), 31 shuffles per row instead of 80.
template <typename T>
__global__ void __launch_bounds__(256) decode_frames_kernel(const T* __restrict__ mask_pre,
                                                            const T* __restrict__ E,
                                                            const float* __restrict__ dec_w,
                                                            int n_masks, size_t rows,
                                                            float* __restrict__ frames) {
  // taps of channel n = lane*8+i live at s_w[(i*32+lane)*20 + k]: the 20-float lane stride makes
  // the per-lane float4 reads below bank-conflict-free (a plain [n][16] layout is 32-way conflicted)
  __shared__ __align__(16) float s_w[kN * 20];
  for (int idx = threadIdx.x; idx < kN * kEncK; idx += 256) {
    const int n = idx / kEncK, k = idx % kEncK;
    s_w[((n & 7) * 32 + (n >> 3)) * 20 + k] = dec_w[idx];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * 8;
  for (size_t r = warp; r < rows; r += nwarps) {
    const size_t bl = r / n_masks;
    const f8 m = ld8(mask_pre + r * kN + lane * 8);
    f8 e;
    if (E != nullptr) e = ld8(E + bl * kN + lane * 8);
    float t[kEncK];
#pragma unroll
    for (int k = 0; k < kEncK; ++k) t[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // E == NULL: plain Decoder.forward on an already-masked input (no ReLU, no multiply)
      const float v = (E != nullptr) ? fmaxf(m.v[i], 0.f) * e.v[i] : m.v[i];
      const float* wr = s_w + (i * 32 + lane) * 20;
#pragma unroll
      for (int k = 0; k < kEncK; k += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
        t[k] = fmaf(v, w4.x, t[k]);
        t[k + 1] = fmaf(v, w4.y, t[k + 1]);
        t[k + 2] = fmaf(v, w4.z, t[k + 2]);
        t[k + 3] = fmaf(v, w4.w, t[k + 3]);
      }
    }
    // butterfly: after the step with offset o, lane keeps the half of the taps selected by its
    // bit (lane & o); after 4 steps each lane holds one tap summed over its 16-lane half... then
    // a final xor-16 step adds the two halves.
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      const int o = 1 << step;           // lane bit used at this step
      const int half = kEncK >> (step + 1);  // values kept
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int k = 0; k < half; ++k) {
        const float keep = up ? t[k + half] : t[k];
        const float send = up ? t[k] : t[k + half];
        t[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    float v = t[0] + __shfl_xor_sync(0xffffffffu, t[0], 16);
    // lane's tap index: bit(step) of lane selects the upper half at that step
    // tap = b0*8 + b1*4 + b2*2 + b3 where b_i = bit i of lane
    const int tap = ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3);
    if (lane < 16) frames[r * kEncK + tap] = v;
  }
}

// The same contraction on the tensor cores (bf16 mode): [rows, 256] x [256, 16] with mma.sync m16n8k16.  Under autocast
// the reference's ConvTranspose1d runs in bf16 as well (its input mask * mix_w is cast), so v = relu(mask) * mix_w is
// rounded to bf16 here too.  A warp takes 16 rows at a time: coalesced 16-byte loads of mask and mix_w, v staged as
// bf16 in its own padded shared-memory tile, ldmatrix A fragments, the decoder filter as the B operand
// (bf16, [16 taps][256 channels]) staged once per CTA and read by ldmatrix.  ~8 instructions per row instead of ~250: the SIMT
// version's butterfly made it issue-bound at 0.12 of the HBM roofline.
__device__ __forceinline__ uint4 ldg128_ordered(const void* p) {
  uint4 u;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}
constexpr int kDfRowBytes = 2 * kN + 16;  // 528: consecutive rows land 4 banks apart -> conflict-free ldmatrix
constexpr int kDfTile = 16 * kDfRowBytes;
__global__ void __launch_bounds__(256) decode_frames_mma_kernel(const bf16* __restrict__ mask_pre,
                                                                const bf16* __restrict__ E,
                                                                const float* __restrict__ dec_w, int n_masks,
                                                                size_t rows, float* __restrict__ frames) {
  extern __shared__ __align__(16) unsigned char df_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  // the decoder filter as the B operand: [16 taps][256 channels] bf16 (K-major rows, padded like the A tiles),
  // staged once per CTA; fragments come from ldmatrix each k-step
  unsigned char* wtile = df_smem;
  unsigned char* tile = df_smem + kDfTile + wid * kDfTile;
  for (int idx = threadIdx.x; idx < kN * kEncK; idx += 256) {
    const int c = idx / kEncK, tap = idx % kEncK;   // dec_w [256][16]: coalesced read
    *reinterpret_cast<bf16*>(wtile + tap * kDfRowBytes + c * 2) = __float2bfloat16_rn(dec_w[idx]);
  }
  __syncthreads();
  // ldmatrix x4 of B = (taps 0-7, k0) (taps 0-7, k0 + 8) (taps 8-15, k0) (taps 8-15, k0 + 8)
  const uint32_t b_base = (uint32_t)__cvta_generic_to_shared(wtile) + ((lane & 7) + (lane >> 4) * 8) * kDfRowBytes +
                          ((lane >> 3) & 1) * 16;
  const size_t ntiles = (rows + 15) / 16;
  const size_t nwarps = (size_t)gridDim.x * 8;
  for (size_t tl = (size_t)blockIdx.x * 8 + wid; tl < ntiles; tl += nwarps) {
    const size_t r0 = tl * 16;
    // stage v = relu(mask) * mix_w: lane -> 8 channels of a row, 16 rows.  The 16 loads of an 8-row batch are
    // issued before any of them is used (row index clamped, not branched on: a predicated load is not hoisted over
    // its guard, and the row-at-a-time version spent one DRAM round trip per row — 0.22 of the HBM roofline)
    // (volatile asm loads keep their program order, so the compiler cannot sink them next to their uses; the
    // mix_w row of row r0 + i is q0 + (rem0 + i) / n_masks: one division per tile, a multiply-shift per row)
    const uint32_t last = (uint32_t)min((size_t)15, rows - 1 - r0), nm = (uint32_t)n_masks;
    const uint32_t q0 = (uint32_t)r0 / nm, rem0 = (uint32_t)r0 - q0 * nm, inv = (256u + nm - 1u) / nm;
    const bf16* mrow = mask_pre + r0 * kN + lane * 8;
    const bf16* erow = E + (size_t)q0 * kN + lane * 8;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      uint4 mq[8], eq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t ri = min((uint32_t)(hb * 8 + i), last);
        mq[i] = ldg128_ordered(mrow + (size_t)ri * kN);
        eq[i] = ldg128_ordered(erow + (size_t)(((rem0 + ri) * inv) >> 8) * kN);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mq[i]);
        const __nv_bfloat162* eh = reinterpret_cast<const __nv_bfloat162*>(&eq[i]);
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 m = __bfloat1622float2(mh[j]), e = __bfloat1622float2(eh[j]);
          w[j] = pack_bf16(fmaxf(m.x, 0.f) * e.x, fmaxf(m.y, 0.f) * e.y);
        }
        const bool live = r0 + hb * 8 + i < rows;   // rows past the end stage zeros
        *reinterpret_cast<uint4*>(tile + (hb * 8 + i) * kDfRowBytes + lane * 16) =
            live ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    __syncwarp();
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
    // A fragments: ldmatrix x4 = (rows 0-7, k0) (rows 8-15, k0) (rows 0-7, k0 + 8) (rows 8-15, k0 + 8)
    const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(tile) + ((lane & 7) + ((lane >> 3) & 1) * 8) * kDfRowBytes +
                            (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      uint32_t a[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                   : "r"(a_base + ks * 32));
      uint32_t w4[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                   : "=r"(w4[0]), "=r"(w4[1]), "=r"(w4[2]), "=r"(w4[3])
                   : "r"(b_base + ks * 32));
      mma_bf16_16816(c0, a, w4[0], w4[1]);
      mma_bf16_16816(c1, a, w4[2], w4[3]);
    }
    __syncwarp();  // the tile is rewritten by the next iteration
    // C fragments: (row g, taps 2 t4, 2 t4 + 1) and (row g + 8, ...), tap tile 0 / 1
    const size_t ra = r0 + g, rb = ra + 8;
    if (ra < rows) {
      *reinterpret_cast<float2*>(frames + ra * kEncK + 2 * t4) = make_float2(c0[0], c0[1]);
      *reinterpret_cast<float2*>(frames + ra * kEncK + 8 + 2 * t4) = make_float2(c1[0], c1[1]);
    }
    if (rb < rows) {
      *reinterpret_cast<float2*>(frames + rb * kEncK + 2 * t4) = make_float2(c0[2], c0[3]);
      *reinterpret_cast<float2*>(frames + rb * kEncK + 8 + 2 * t4) = make_float2(c1[2], c1[3]);
    }
  }
}

// est[b,t,s] = frames[(b,l,s), t-8l] + frames[(b,l-1,s), t-8(l-1)], l = t/8; zero beyond T_est
// (F.pad branch, ContSep.py:92-93); only t < T is produced (trim branch, :95).
__global__ void __launch_bounds__(256) decode_ola_kernel(const float* __restrict__ frames, int L,
                                                         int T, int n_masks, size_t total,
                                                         float* __restrict__ est) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int s = (int)(i % n_masks);
    const size_t bt = i / n_masks;
    const int t = (int)(bt % T);
    const size_t b = bt / T;
    const int l = t >> 3, k = t & 7;
    float v = 0.f;
    if (l < L) v += frames[((b * L + l) * n_masks + s) * kEncK + k];
    if (l >= 1 && l - 1 < L) v += frames[((b * L + l - 1) * n_masks + s) * kEncK + k + 8];
    est[i] = v;
  }
}

// mask = relu(mask_pre) as fp32 (Dual_Path_Model.forward's return value, ContSep.py:263).
template <typename T>
__global__ void __launch_bounds__(256) relu_f32_kernel(const T* __restrict__ x, size_t n8,
                                                       float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
    f8 v = ld8(x + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = fmaxf(v.v[k], 0.f);
    st8(out + i * 8, v);
  }
}

int launch_relu_f32(const void* x, size_t n, int act, float* out, cudaStream_t st) {
  const size_t n8 = n / 8;
  const int grid = (int)min((size_t)148 * 8, (n8 + 255) / 256);
  if (act == CSE_BF16)
    relu_f32_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, n8, out);
  else
    relu_f32_kernel<float><<<grid, 256, 0, st>>>((const float*)x, n8, out);
  return check_launch("relu_f32_kernel");
}

int launch_mask_decode(const void* mask_pre, const void* E, const float* dec_w, int B, int L, int T,
                       int n_masks, int act, float* frames, float* est, cudaStream_t st) {
  const size_t rows = (size_t)B * L * n_masks;
  const int grid = (int)min((size_t)148 * 4, (rows + 7) / 8);
  static const bool simt_only = []() {  // CSE_DECODE_SIMT=1 keeps the SIMT kernel in bf16 mode (A/B aid)
    const char* e = getenv("CSE_DECODE_SIMT");
    return e != nullptr && e[0] == '1';
  }();
  if (act == CSE_BF16 && E != nullptr && !simt_only) {
    static DeviceOnce once;
    if (!once.configured_on_this_device()) {
      cudaError_t e = cudaFuncSetAttribute(decode_frames_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * kDfTile);
      if (e != cudaSuccess) {
        set_error("decode_frames: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return 1;
      }
      once.mark_configured();
    }
    const size_t tiles = (rows + 15) / 16;
    const int gridm = (int)min((size_t)148 * 2, (tiles + 7) / 8);
    decode_frames_mma_kernel<<<gridm, 256, 9 * kDfTile, st>>>((const bf16*)mask_pre, (const bf16*)E, dec_w, n_masks,
                                                              rows, frames);
  } else if (act == CSE_BF16)
    decode_frames_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)mask_pre, (const bf16*)E, dec_w,
                                                     n_masks, rows, frames);
  else
    decode_frames_kernel<float><<<grid, 256, 0, st>>>((const float*)mask_pre, (const float*)E,
                                                      dec_w, n_masks, rows, frames);
  if (check_launch("decode_frames_kernel")) return 1;
  const size_t total = (size_t)B * T * n_masks;
  const int grid2 = (int)min((size_t)148 * 8, (total + 255) / 256);
  decode_ola_kernel<<<grid2, 256, 0, st>>>(frames, L, T, n_masks, total, est);
  return check_launch("decode_ola_kernel");
}

}  // namespace cse
