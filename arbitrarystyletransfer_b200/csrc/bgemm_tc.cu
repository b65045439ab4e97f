// K6g: batched bf16 GEMM on tcgen05 with either operand K-major or MN-major -- the contraction engine of the
// AdaAttN layer (models.py:70-115) and of its backward pass.
//
//   D[b][i][j] = sum_k A(b, i, k) * B(b, j, k)           fp32 accumulation in TMEM, D fp32 or bf16, row-major
//
//   a_mn = 0:  A stored [M rows][K contiguous]  (K-major:  q, k, P = softmax(QK^T) read by rows, dS, [dMean dM2])
//   a_mn = 1:  A stored [K rows][M contiguous]  (MN-major: P^T, dS^T -- the same buffers, contraction over their rows)
//   b_mn = 0:  B stored [N rows][K contiguous]  (K-major:  k in Q K^T, [v v^2] in dA)
//   b_mn = 1:  B stored [K rows][N contiguous]  (MN-major: [v v^2] in P [v v^2], k in dS k, q in dS^T q)
//
// so that every product the layer needs -- Q K^T (models.py:97), P V and P V^2 (:101-103), and the five products
// of the backward pass -- reads the NHWC activations exactly as they lie in HBM: no transposes are materialised.
// K-major tiles are one TMA box {64 k, rows} (rows of 128 B, SWIZZLE_128B, descriptor advance 32 B per 16-wide
// K step); MN-major tiles are boxes {64 m/n, 64 k} (canonical MN-major SWIZZLE_128B, SBO = 1024 B, LBO = one box,
// advance 2048 B per K step) -- see pw_wgrad_tc.cu.  Out-of-range rows / columns / K are zero-filled by TMA, so
// ragged sizes need no special case; the epilogue masks.  One 128 x BN tile per CTA, 4-stage TMA ring, warp 0 TMA,
// warp 1 MMA (elect.sync), warps 2-5 epilogue.  These products are small (<= 1 GFLOP per image at 256^2 input):
// the kernel is latency/HBM-bound, the point of tcgen05 here is that the contraction costs nothing next to the
// streaming passes around it.
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int BG_THREADS = 192;
constexpr int BG_KS = 64;                        // K elements per stage
constexpr int BG_A_BYTES = 128 * BG_KS * 2;      // 16 KB
constexpr int BG_B_BYTES = 256 * BG_KS * 2;      // 32 KB (BN <= 256)
constexpr int BG_STAGE = BG_A_BYTES + BG_B_BYTES;
constexpr int BG_STAGES = 4;
constexpr int BG_SMEM = BG_STAGES * BG_STAGE + (2 * BG_STAGES + 1) * 8 + 16 + 1024;
constexpr int BG_BOX = 64 * 64 * 2;              // one MN-major box, 8 KB

struct BgParams {
  int M, N, K, BN, a_mn, b_mn, b_boxes, out_bf16;
  int kchunks;
  void* out;
  int64_t ld_d, sb;
};

__device__ __forceinline__ uint64_t bg_sdesc_mn128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(BG_BOX >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(BG_THREADS, 1)
bgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const BgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + BG_STAGES * BG_STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (BG_STAGES + s); };
  const uint32_t done_bar = bars + 8u * (2 * BG_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * BG_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + BG_STAGES * BG_STAGE + 8 * (2 * BG_STAGES + 1));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int bi = blockIdx.x, mb = blockIdx.y, nb = blockIdx.z;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < BG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)BG_A_BYTES + (uint32_t)(p.b_mn ? p.b_boxes * BG_BOX : p.BN * BG_KS * 2);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < p.kchunks; ++c) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), bytes);
        const uint32_t dstA = base + stage * BG_STAGE;
        const uint32_t dstB = dstA + BG_A_BYTES;
        const int k0 = c * BG_KS;
        if (p.a_mn) {
          tma_load_3d(dstA, &tmA, full_bar(stage), mb * 128, k0, bi);
          tma_load_3d(dstA + BG_BOX, &tmA, full_bar(stage), mb * 128 + 64, k0, bi);
        } else {
          tma_load_3d(dstA, &tmA, full_bar(stage), k0, mb * 128, bi);
        }
        if (p.b_mn) {
          for (int b = 0; b < p.b_boxes; ++b)
            tma_load_3d(dstB + b * BG_BOX, &tmB, full_bar(stage), nb * p.BN + b * 64, k0, bi);
        } else {
          tma_load_3d(dstB, &tmB, full_bar(stage), k0, nb * p.BN, bi);
        }
        if (++stage == BG_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, p.BN) | ((uint32_t)(p.a_mn != 0) << 15) | ((uint32_t)(p.b_mn != 0) << 16);
    const uint64_t a_step = p.a_mn ? 128u : 2u;     // descriptor units (16 B) per 16-wide K step
    const uint64_t b_step = p.b_mn ? 128u : 2u;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum = 0;
    for (int c = 0; c < p.kchunks; ++c) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint32_t sa = base + stage * BG_STAGE;
      const uint64_t ad = p.a_mn ? bg_sdesc_mn128(sa) : make_sdesc_k128(sa);
      const uint64_t bd = p.b_mn ? bg_sdesc_mn128(sa + BG_A_BYTES) : make_sdesc_k128(sa + BG_A_BYTES);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < BG_KS / 16; ++k)
          if (c * BG_KS + k * 16 < p.K)        // K tail: steps that are pure TMA zero fill are skipped
            umma_bf16(tmem_base, ad + a_step * k, bd + b_step * k, idesc, k ? 1u : accum);
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      accum = 1u;
      if (++stage == BG_STAGES) { stage = 0; phase ^= 1u; }
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
      const int j0 = nb * p.BN + col;
      if (i >= p.M || j0 >= p.N) continue;
      if (p.out_bf16) {
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)bi * p.sb + (int64_t)i * p.ld_d;
        if (j0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(orow + j0) & 15) == 0)) {
          uint4 w0 = make_uint4(pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])),
                                pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                                pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])),
                                pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7])));
          uint4 w1 = make_uint4(pack_bf16(__uint_as_float(v[8]), __uint_as_float(v[9])),
                                pack_bf16(__uint_as_float(v[10]), __uint_as_float(v[11])),
                                pack_bf16(__uint_as_float(v[12]), __uint_as_float(v[13])),
                                pack_bf16(__uint_as_float(v[14]), __uint_as_float(v[15])));
          reinterpret_cast<uint4*>(orow + j0)[0] = w0;
          reinterpret_cast<uint4*>(orow + j0)[1] = w1;
        } else {
#pragma unroll
          for (int t = 0; t < 16; ++t)
            if (j0 + t < p.N) orow[j0 + t] = __float2bfloat16_rn(__uint_as_float(v[t]));
        }
      } else {
        float* orow = reinterpret_cast<float*>(p.out) + (int64_t)bi * p.sb + (int64_t)i * p.ld_d;
        if (j0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(orow + j0) & 15) == 0)) {
#pragma unroll
          for (int t = 0; t < 16; t += 4)
            *reinterpret_cast<float4*>(orow + j0 + t) = make_float4(__uint_as_float(v[t]), __uint_as_float(v[t + 1]),
                                                                    __uint_as_float(v[t + 2]), __uint_as_float(v[t + 3]));
        } else {
#pragma unroll
          for (int t = 0; t < 16; ++t)
            if (j0 + t < p.N) orow[j0 + t] = __uint_as_float(v[t]);
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

extern "C" int ast_bgemm(const void* a, int a_mn, int ld_a, int64_t a_bs, const void* b, int b_mn, int ld_b,
                         int64_t b_bs, void* d, int d_bf16, int64_t ld_d, int64_t d_bs, int M, int N, int K,
                         int batch, void* stream) {
  if (!a || !b || !d || M <= 0 || N <= 0 || K <= 0 || batch <= 0) return AST_E_BADARG;
  const int a_inner = a_mn ? M : K, b_inner = b_mn ? N : K;
  if (ld_a % 8 != 0 || ld_b % 8 != 0 || ld_a < a_inner || ld_b < b_inner || ld_d < N ||
      (batch > 1 && (a_bs % 8 != 0 || b_bs % 8 != 0)) || batch > 65535)
    return AST_E_SHAPE;
  if (!aligned16(a) || !aligned16(b)) return AST_E_ALIGN;
  int n_blocks = 1, BN = (N + 15) / 16 * 16;
  while (BN > 256) {
    ++n_blocks;
    BN = ((N + n_blocks - 1) / n_blocks + 15) / 16 * 16;
  }
  const int m_blocks = (M + 127) / 128;
  if (n_blocks > 65535 || m_blocks > 65535) return AST_E_SHAPE;
  BgParams p = {};
  p.M = M; p.N = N; p.K = K; p.BN = BN; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.b_boxes = (BN + 63) / 64; p.out_bf16 = d_bf16 ? 1 : 0;
  p.kchunks = (K + BG_KS - 1) / BG_KS;
  p.out = d; p.ld_d = ld_d; p.sb = d_bs;
  CUtensorMap tmA, tmB;
  {
    const uint64_t rows = a_mn ? K : M;
    const uint64_t dims[3] = {(uint64_t)a_inner, rows, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ld_a * 2, (uint64_t)(batch > 1 ? a_bs : (int64_t)rows * ld_a) * 2};
    const uint32_t box[3] = {64, (uint32_t)(a_mn ? 64 : 128), 1};
    int r = encode_bf16_map(&tmA, a, 3, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t rows = b_mn ? K : N;
    const uint64_t dims[3] = {(uint64_t)b_inner, rows, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ld_b * 2, (uint64_t)(batch > 1 ? b_bs : (int64_t)rows * ld_b) * 2};
    const uint32_t box[3] = {64, (uint32_t)(b_mn ? 64 : BN), 1};
    int r = encode_bf16_map(&tmB, b, 3, dims, str, box);
    if (r) return r;
  }
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(bgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM));
    attr_done = true;
  }
  bgemm_tc_kernel<<<dim3((unsigned)batch, m_blocks, n_blocks), BG_THREADS, BG_SMEM, (cudaStream_t)stream>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}
