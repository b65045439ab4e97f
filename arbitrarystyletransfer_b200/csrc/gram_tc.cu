// K3g: Gram matrix G[n] = X[n] X[n]^T / (C*H*W) (losses.py:105-109) on tcgen05 in TF32.
//
// X is the reference's own fp32 NCHW tap: per image a [C][HW] matrix whose rows are contiguous along
// HW, i.e. already the K-major operand of X X^T.  Both operands are therefore loaded straight from
// the fp32 tensor by TMA (rows of 32 floats = 128 B, SWIZZLE_128B) and multiplied with
// tcgen05.mma.kind::tf32 (K = 8 per instruction, fp32 accumulation in TMEM); no cast, no transpose.
// Work item = (image, 128-row block, N block, split-K chunk); partial tiles are reduced with fp32
// red.global.add into G (zeroed first).  TF32 rounds the inputs to 10 mantissa bits, products are
// exact in fp32: G agrees with the fp32 reference to ~1e-4 relative (sum of 10^3..10^5 terms);
// ast_gram_fwd (CUDA-core fp32) remains for 1e-6 agreement.
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int GR_THREADS = 256;
constexpr int GR_KBLK = 32;                        // floats per stage row = 128 B
constexpr int GR_A_BYTES = 128 * GR_KBLK * 4;      // 16 KB

template <int BN>
struct GrCfg {
  static constexpr int B_BYTES = BN * GR_KBLK * 4;
  static constexpr int STAGE_BYTES = GR_A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
};

struct GrParams {
  int C, m_blocks, n_blocks, k_chunks, ksteps_total, ksteps_per_chunk;
  float scale;
  float* g;   // [N][C][C]
};

// D fp32, A/B tf32 (format code 2), K-major both
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(GR_THREADS, 1)
gram_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GrParams p) {
  using C = GrCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  const uint32_t done_bar = bars + 8u * (2 * C::STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 1));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  int t = blockIdx.x;
  const int kc = t % p.k_chunks; t /= p.k_chunks;
  const int nb = t % p.n_blocks; t /= p.n_blocks;
  const int mb = t % p.m_blocks;
  const int n = t / p.m_blocks;
  const int ks0 = kc * p.ksteps_per_chunk;
  int ks1 = ks0 + p.ksteps_per_chunk;
  if (ks1 > p.ksteps_total) ks1 = p.ksteps_total;
  const int nsteps = ks1 - ks0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = ks0; ks < ks1; ++ks) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
        const uint32_t a_dst = base + stage * C::STAGE_BYTES;
        tma_load_3d(a_dst, &tmA, full_bar(stage), ks * GR_KBLK, mb * 128, n);
        tma_load_3d(a_dst + GR_A_BYTES, &tmB, full_bar(stage), ks * GR_KBLK, nb * BN, n);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(128, BN);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum = 0;
    for (int ks = 0; ks < nsteps; ++ks) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint64_t ad = make_sdesc_k128(base + stage * C::STAGE_BYTES);
      const uint64_t bd = make_sdesc_k128(base + stage * C::STAGE_BYTES + GR_A_BYTES);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < GR_KBLK / 8; ++k)   // 8 tf32 = 32 B per K slice
          umma_tf32(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k ? 1u : accum);
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      accum = 1u;
      if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
    }
    if (elect_one_sync()) umma_commit(done_bar);
    __syncwarp();
  } else if (warp >= 4) {
    const int e = warp - 4;
    mbar_wait(done_bar, 0u);
    tc_fence_after();
    const int r = mb * 128 + e * 32 + lane;
    constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
    for (int chunk = 0; chunk < BN / CH; ++chunk) {
      uint32_t v[CH];
      tmem_ld_cols(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(chunk * CH), v);
      tmem_ld_wait();
      if (nsteps > 0 && r < p.C) {
        float* dst = p.g + ((int64_t)n * p.C + r) * p.C + nb * BN + chunk * CH;
#pragma unroll
        for (int i = 0; i < CH; ++i)
          if (nb * BN + chunk * CH + i < p.C) atomicAdd(dst + i, __uint_as_float(v[i]) * p.scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
  }
}

template <int BN>
static int launch_gram(const CUtensorMap& tmA, const CUtensorMap& tmB, const GrParams& p, int grid,
                       cudaStream_t s) {
  using C = GrCfg<BN>;
  auto kern = gram_tf32_kernel<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  kern<<<grid, GR_THREADS, C::SMEM_BYTES, s>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}

}  // namespace tc
}  // namespace ast

using namespace ast;
using namespace ast::tc;

extern "C" int ast_gram_fwd_tf32(const float* x, float* g, int B, int C, int64_t HW, void* stream) {
  if (!x || !g || B <= 0 || C <= 0 || HW <= 0) return AST_E_BADARG;
  if (HW % 4 != 0 || HW >= 0x7fffffffLL || !aligned16(x)) return AST_E_SHAPE;   // TMA: 16 B row stride
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return AST_E_NODRIVER;
  cudaStream_t s = (cudaStream_t)stream;
  int BN = 16;
  if (C % 256 == 0) BN = 256;
  else if (C % 128 == 0) BN = 128;
  else if (C % 64 == 0) BN = 64;
  else if (C % 32 == 0) BN = 32;
  else if (C % 16 != 0) return AST_E_SHAPE;
  GrParams p = {};
  p.C = C;
  p.m_blocks = (C + 127) / 128;
  p.n_blocks = C / BN;
  p.ksteps_total = (int)((HW + GR_KBLK - 1) / GR_KBLK);
  p.scale = 1.f / ((float)C * (float)HW);
  p.g = g;
  const int items = B * p.m_blocks * p.n_blocks;
  int chunks = (2 * 148 + items - 1) / items;
  int max_chunks = p.ksteps_total / 16;
  if (max_chunks < 1) max_chunks = 1;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  p.ksteps_per_chunk = (p.ksteps_total + chunks - 1) / chunks;
  p.k_chunks = (p.ksteps_total + p.ksteps_per_chunk - 1) / p.ksteps_per_chunk;
  const int grid = items * p.k_chunks;
  AST_CUDA(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)B * C * C, s));
  // both operands are views of x: {HW, C, B} fp32, boxes {32, 128, 1} and {32, BN, 1}
  CUtensorMap tmA, tmB;
  for (int which = 0; which < 2; ++which) {
    cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * 4 * C};
    cuuint32_t bx[3] = {GR_KBLK, which == 0 ? 128u : (cuuint32_t)BN, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(which == 0 ? &tmA : &tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x),
                     gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return AST_E_SHAPE;
  }
  switch (BN) {
    case 256: return launch_gram<256>(tmA, tmB, p, grid, s);
    case 128: return launch_gram<128>(tmA, tmB, p, grid, s);
    case 64: return launch_gram<64>(tmA, tmB, p, grid, s);
    case 32: return launch_gram<32>(tmA, tmB, p, grid, s);
    default: return launch_gram<16>(tmA, tmB, p, grid, s);
  }
}
