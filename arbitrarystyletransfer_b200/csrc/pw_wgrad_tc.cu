// K4w: weight gradient of the pointwise (1x1) convolutions as a tcgen05 GEMM over the PIXEL axis.
//
//   D[i][j] = sum_p A[p][i] * B[p][j]          A = [P][Ca] bf16 (row stride ld_a), B = [P][Cb] bf16
//
// i.e. dW[co][ci] = sum_p dZ[p][co] * X[p][ci] for nn.Conv2d(.., 1, 1, 0) (mobilenetv2.py:103-150) with NHWC
// activations.  The contraction index (pixels) is the OUTER dimension of both operands, so both are
// MN-major UMMA operands: a TMA box {64 channels, 64 pixels} lands as 64 rows (pixels = K) of 128 B
// (64 channels = M or N) with the 128-byte swizzle, which is exactly the canonical MN-major SWIZZLE_128B
// layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: SBO = 1024 B (8 pixels), LBO = one box
// (64 channels further along M/N).  One tcgen05.mma consumes 16 pixels = 2048 B of each box.
// HBM-bound (every operand byte is read once per M block): split-K over CTAs, fp32 atomics into dW.
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int WG_THREADS = 192;             // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int WG_KP = 64;                   // pixels per stage
constexpr int WG_BOX = WG_KP * 128;         // 8 KB: one {64 channels x 64 pixels} box
constexpr int WG_A_BOXES = 2;               // M = 128 channels
constexpr int WG_B_BOXES = 4;               // N <= 256 channels
constexpr int WG_STAGE = (WG_A_BOXES + WG_B_BOXES) * WG_BOX;   // 48 KB
constexpr int WG_STAGES = 4;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + (2 * WG_STAGES + 1) * 8 + 16 + 1024;

struct WgParams {
  int Ca, Cb, BN, b_boxes;
  int64_t P;
  int64_t chunks;          // number of 64-row K chunks
  float* out;              // out[b * sb + i * si + j * sj] (+)= D[i][j]
  int64_t si, sj, sb;
  int split;               // K splits per batch item (gridDim.x = split * batch)
  int accumulate;          // 1: atomicAdd into a zero-filled buffer (split-K); 0: plain stores (split == 1)
};

// MN-major, SWIZZLE_128B shared-memory descriptor (see header comment).
__device__ __forceinline__ uint64_t make_sdesc_mn128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor with both operands MN-major ("transpose" bits 15 and 16).
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
pw_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + WG_STAGES * WG_STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (WG_STAGES + s); };
  const uint32_t done_bar = bars + 8u * (2 * WG_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * WG_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + WG_STAGES * WG_STAGE + 8 * (2 * WG_STAGES + 1));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int mb = blockIdx.y, nb = blockIdx.z;
  const int bi = blockIdx.x / p.split;               // batch item
  const int sidx = blockIdx.x - bi * p.split;
  // contiguous K-chunk range of this split
  const int64_t per = (p.chunks + p.split - 1) / p.split;
  const int64_t c0 = (int64_t)sidx * per;
  const int64_t c1 = c0 + per < p.chunks ? c0 + per : p.chunks;
  const int64_t nk = c1 > c0 ? c1 - c0 : 0;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)(WG_A_BOXES + p.b_boxes) * WG_BOX;
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t c = c0; c < c1; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), bytes);
          const uint32_t dst = base + stage * WG_STAGE;
          const int pix = (int)(c * WG_KP);
          for (int b = 0; b < WG_A_BOXES; ++b)
            tma_load_3d(dst + b * WG_BOX, &tmA, full_bar(stage), mb * 128 + b * 64, pix, bi);
          for (int b = 0; b < p.b_boxes; ++b)
            tma_load_3d(dst + (WG_A_BOXES + b) * WG_BOX, &tmB, full_bar(stage), nb * p.BN + b * 64, pix, bi);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_bf16_mn(128, p.BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int64_t c = c0; c < c1; ++c) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t ad = make_sdesc_mn128(base + stage * WG_STAGE, WG_BOX);
        const uint64_t bd = make_sdesc_mn128(base + stage * WG_STAGE + WG_A_BOXES * WG_BOX, WG_BOX);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < WG_KP / 16; ++k)   // 16 pixels = 2048 B = 128 descriptor units per MMA
            umma_bf16(tmem_base, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), idesc, k ? 1u : accum);
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        accum = 1u;
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(done_bar);
      __syncwarp();
    } else {
      const int e = warp & 3;
      mbar_wait(done_bar, 0u);
      tc_fence_after();
      const int i = mb * 128 + e * 32 + lane;
      for (int col = 0; col < p.BN; col += 16) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)col, v);
        tmem_ld_wait();
        if (i >= p.Ca) continue;
        float* orow = p.out + (int64_t)bi * p.sb + (int64_t)i * p.si;
        const int j0 = nb * p.BN + col;
        if (!p.accumulate && p.sj == 1 && j0 + 16 <= p.Cb && ((reinterpret_cast<uintptr_t>(orow + j0) & 15) == 0)) {
#pragma unroll
          for (int t = 0; t < 16; t += 4)
            *reinterpret_cast<float4*>(orow + j0 + t) = make_float4(__uint_as_float(v[t]), __uint_as_float(v[t + 1]),
                                                                    __uint_as_float(v[t + 2]), __uint_as_float(v[t + 3]));
          continue;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const int j = j0 + t;
          if (j < p.Cb) {
            if (p.accumulate) atomicAdd(orow + (int64_t)j * p.sj, __uint_as_float(v[t]));
            else orow[(int64_t)j * p.sj] = __uint_as_float(v[t]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace tc
}  // namespace ast

using namespace ast;
using namespace ast::tc;

// out[b*sb + i*si + j*sj] (+)= sum_k a[b][k][i] * b[b][k][j]   (both operands [K rows][columns contiguous], bf16)
static int gemm_tn_bf16(const void* a, int ld_a, int Ca, int64_t a_bs, const void* b, int ld_b, int Cb, int64_t b_bs,
                        int64_t K, int batch, float* out, int64_t si, int64_t sj, int64_t sb, int accumulate,
                        cudaStream_t stream) {
  if (!a || !b || !out || Ca <= 0 || Cb <= 0 || K <= 0 || batch <= 0) return AST_E_BADARG;
  if (Ca % 8 != 0 || Cb % 8 != 0 || ld_a % 8 != 0 || ld_b % 8 != 0 || ld_a < Ca || ld_b < Cb || K >= 0x7fffffffLL ||
      (batch > 1 && (a_bs % 8 != 0 || b_bs % 8 != 0)))
    return AST_E_SHAPE;
  if (!aligned16(a) || !aligned16(b)) return AST_E_ALIGN;
  int n_blocks = 1, BN = (Cb + 15) / 16 * 16;
  while (BN > 256) {
    ++n_blocks;
    BN = ((Cb + n_blocks - 1) / n_blocks + 15) / 16 * 16;
  }
  const int m_blocks = (Ca + 127) / 128;
  if (n_blocks > 65535 || m_blocks > 65535) return AST_E_SHAPE;
  WgParams p = {};
  p.Ca = Ca; p.Cb = Cb; p.BN = BN; p.b_boxes = (BN + 63) / 64;
  p.P = K; p.chunks = (K + WG_KP - 1) / WG_KP;
  p.out = out; p.si = si; p.sj = sj; p.sb = sb; p.accumulate = accumulate;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)Ca, (uint64_t)K, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ld_a * 2, (uint64_t)(batch > 1 ? a_bs : (int64_t)K * ld_a) * 2};
    const uint32_t box[3] = {64, (uint32_t)WG_KP, 1};
    int r = encode_bf16_map(&tmA, a, 3, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cb, (uint64_t)K, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ld_b * 2, (uint64_t)(batch > 1 ? b_bs : (int64_t)K * ld_b) * 2};
    const uint32_t box[3] = {64, (uint32_t)WG_KP, 1};
    int r = encode_bf16_map(&tmB, b, 3, dims, str, box);
    if (r) return r;
  }
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    attr_done = true;
  }
  int64_t split = 1;
  if (accumulate) {
    split = 148 / ((int64_t)m_blocks * n_blocks * batch);
    if (split < 1) split = 1;
    if (split > p.chunks / 4) split = p.chunks / 4;     // >= 4 K chunks per CTA: the 128 x BN atomic epilogue must amortise
    if (split < 1) split = 1;
  }
  p.split = (int)split;
  if (split * batch >= 0x7fffffffLL) return AST_E_SHAPE;
  pw_wgrad_tc_kernel<<<dim3((unsigned)(split * batch), m_blocks, n_blocks), WG_THREADS, WG_SMEM, stream>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_pw_wgrad(const void* a, int ld_a, int Ca, const void* b, int ld_b, int Cb, int64_t P,
                            float* out, int64_t si, int64_t sj, void* stream) {
  return gemm_tn_bf16(a, ld_a, Ca, 0, b, ld_b, Cb, 0, P, 1, out, si, sj, 0, 1, (cudaStream_t)stream);
}

// ---- Gram-matrix backward on the tensor cores (losses.py:105-109 autograd) -------------------------------
// gx[b][c][p] = sum_c' S[b][c'][c] * X[b][c'][p],  S = (gg + gg^T) / (C*HW)  (symmetric), X = the fp32 NCHW tap.
// X's contraction index c' is its OUTER dimension, so it is an MN-major operand: only 16-bit types take the plain
// 128-byte swizzle there, hence bf16 copies of X and S (the gradient is rounded to bf16 right afterwards anyway
// when it enters the VGG backward pass).
namespace ast {
__global__ void gram_bwd_prep_s_kernel(const float* __restrict__ gg, __nv_bfloat16* __restrict__ s, int C, float scale,
                                       int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int cp = (int)(r % C);
    const int64_t b = r / C;
    s[i] = __float2bfloat16_rn((gg[i] + gg[(b * C + c) * C + cp]) * scale);
  }
}
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}
}  // namespace ast

extern "C" size_t ast_gram_bwd_tc_ws_bytes(int B, int C, int64_t HW) {
  return ((size_t)B * C * HW + (size_t)B * C * C) * 2 + 256;
}

extern "C" int ast_gram_bwd_tc(const float* x, const float* gg, float* gx, int B, int C, int64_t HW, void* ws,
                               size_t ws_bytes, void* stream) {
  if (!x || !gg || !gx || !ws || B <= 0 || C <= 0 || HW <= 0) return AST_E_BADARG;
  if (C % 8 != 0 || HW % 8 != 0) return AST_E_SHAPE;
  if (ws_bytes < ast_gram_bwd_tc_ws_bytes(B, C, HW)) return AST_E_WORKSPACE;
  if (!aligned16(x) || !aligned16(ws)) return AST_E_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(ws);
  const size_t xbytes = ((size_t)B * C * HW * 2 + 255) / 256 * 256;
  __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(ws) + xbytes);
  const int64_t n4 = (int64_t)B * C * HW / 4;
  int64_t nb = (n4 + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  cast_bf16_kernel<<<(unsigned)nb, 256, 0, s>>>(x, xb, n4);
  AST_CHECK_LAUNCH();
  const int64_t ts = (int64_t)B * C * C;
  int64_t nbs = (ts + 255) / 256;
  if (nbs > 148 * 4) nbs = 148 * 4;
  gram_bwd_prep_s_kernel<<<(unsigned)nbs, 256, 0, s>>>(gg, sb, C, 1.f / ((float)C * (float)HW), ts);
  AST_CHECK_LAUNCH();
  return gemm_tn_bf16(sb, C, C, (int64_t)C * C, xb, (int)HW, (int)HW, (int64_t)C * HW, C, B, gx, HW, 1, (int64_t)C * HW, 0, s);
}
