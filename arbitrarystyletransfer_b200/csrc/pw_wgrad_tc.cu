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
  int64_t chunks;          // number of 64-pixel K chunks
  float* out;              // out[i * si + j * sj] += D[i][j]
  int64_t si, sj;
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
  // contiguous K-chunk range of this split
  const int64_t per = (p.chunks + gridDim.x - 1) / gridDim.x;
  const int64_t c0 = (int64_t)blockIdx.x * per;
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
            tma_load_2d(dst + b * WG_BOX, &tmA, full_bar(stage), mb * 128 + b * 64, pix);
          for (int b = 0; b < p.b_boxes; ++b)
            tma_load_2d(dst + (WG_A_BOXES + b) * WG_BOX, &tmB, full_bar(stage), nb * p.BN + b * 64, pix);
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
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const int j = nb * p.BN + col + t;
          if (j < p.Cb) atomicAdd(p.out + (int64_t)i * p.si + (int64_t)j * p.sj, __uint_as_float(v[t]));
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

extern "C" int ast_pw_wgrad(const void* a, int ld_a, int Ca, const void* b, int ld_b, int Cb, int64_t P,
                            float* out, int64_t si, int64_t sj, void* stream) {
  if (!a || !b || !out || Ca <= 0 || Cb <= 0 || P <= 0) return AST_E_BADARG;
  if (Ca % 8 != 0 || Cb % 8 != 0 || ld_a % 8 != 0 || ld_b % 8 != 0 || ld_a < Ca || ld_b < Cb ||
      P >= 0x7fffffffLL)
    return AST_E_SHAPE;
  if (!aligned16(a) || !aligned16(b)) return AST_E_ALIGN;
  int n_blocks = 1, BN = (Cb + 15) / 16 * 16;
  while (BN > 256) {
    ++n_blocks;
    BN = ((Cb + n_blocks - 1) / n_blocks + 15) / 16 * 16;
  }
  const int m_blocks = (Ca + 127) / 128;
  WgParams p = {};
  p.Ca = Ca; p.Cb = Cb; p.BN = BN; p.b_boxes = (BN + 63) / 64;
  p.P = P; p.chunks = (P + WG_KP - 1) / WG_KP;
  p.out = out; p.si = si; p.sj = sj;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)Ca, (uint64_t)P};
    const uint64_t str[1] = {(uint64_t)ld_a * 2};
    const uint32_t box[2] = {64, (uint32_t)WG_KP};
    int r = encode_bf16_map(&tmA, a, 2, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Cb, (uint64_t)P};
    const uint64_t str[1] = {(uint64_t)ld_b * 2};
    const uint32_t box[2] = {64, (uint32_t)WG_KP};
    int r = encode_bf16_map(&tmB, b, 2, dims, str, box);
    if (r) return r;
  }
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    attr_done = true;
  }
  int64_t split = 148 / ((int64_t)m_blocks * n_blocks);
  if (split < 1) split = 1;
  if (split > p.chunks) split = p.chunks;
  pw_wgrad_tc_kernel<<<dim3((unsigned)split, m_blocks, n_blocks), WG_THREADS, WG_SMEM, (cudaStream_t)stream>>>(
      tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}
