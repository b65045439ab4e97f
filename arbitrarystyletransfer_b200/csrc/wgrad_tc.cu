// K2w: weight gradient of a 3x3 stride-1 convolution on tcgen05 / TMEM.
//
// Replaces what autograd computes for nn.Conv2d.weight in the reference's decoder training step
// (train.py:287-300 over the convs of models.py:598-628).
//
//   dW[co][ci][kh][kw] = sum_{n,h,w} dZ[n][co][h][w] * Xpad[n][ci][h+kh][w+kw]
//
// Both operands are taken from CHANNEL-PLANAR bf16 copies whose pixel axis is the padded linear
// index q = (n*(H+2) + h+1)*wp + (w+1), row pitch wp = a multiple of 8 pixels >= W+2:
//   dzT [Cout][ldq]  (zero halo)            xT [3][Cin][ldq]  (halo = the padding the conv saw)
// With that index the tap (kh,kw) is a pure shift along q, delta = (kh-1)*wp + (kw-1), and because
// dZ's halo is zero no product ever crosses an image or row boundary.  So the whole gradient is
// nine K-major GEMMs  D_tap[co][ci] = dzT[co][:] . xT[ci][: + delta]  with K = ldq: the same SW128
// K-major operand form as the forward kernel (rows of 64 pixels = 128 B).
// TMA tiled loads need a 16-byte aligned innermost coordinate (measured on B200: an odd element
// offset never completes), so the +-1 column shift is materialised as three copies of xT
// (xT[s][ci][q] = x[ci][q + s - 1]) and only the row shift (kh-1)*wp -- a multiple of 8 -- goes into
// the coordinate.
// Work item = (tap, 128-row block of Cout, N block of Cin, split-K chunk); partial sums are
// reduced with fp32 red.global.add into dwpk [9][Cout][Cin] (zeroed by the caller's memset).
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int WG_THREADS = 256;
constexpr int WG_KBLK = 64;                       // pixels per stage
constexpr int WG_A_BYTES = 128 * WG_KBLK * 2;     // 16 KB

template <int BN>
struct WgCfg {
  static constexpr int B_BYTES = BN * WG_KBLK * 2;
  static constexpr int STAGE_BYTES = WG_A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
};

struct WgParams {
  int Cin, Cout, wp;         // wp = planar row pitch
  int m_blocks, n_blocks, k_chunks;
  int ksteps_total, ksteps_per_chunk;
  float* dwpk;               // [9][Cout][Cin] fp32
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const WgParams p) {
  using C = WgCfg<BN>;
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

  // work item decode
  int t = blockIdx.x;
  const int kc = t % p.k_chunks; t /= p.k_chunks;
  const int nb = t % p.n_blocks; t /= p.n_blocks;
  const int mb = t % p.m_blocks;
  const int tap = t / p.m_blocks;
  const int kh = tap / 3, kw = tap - 3 * kh;
  const int delta = (kh - 1) * p.wp;  // multiple of 8 elements; the kw shift selects the xT copy
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
        tma_load_2d(a_dst, &tmA, full_bar(stage), ks * WG_KBLK, mb * 128);
        tma_load_3d(a_dst + WG_A_BYTES, &tmB, full_bar(stage), ks * WG_KBLK + delta, nb * BN, kw);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int ks = 0; ks < nsteps; ++ks) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t ad = make_sdesc_k128(base + stage * C::STAGE_BYTES);
        const uint64_t bd = make_sdesc_k128(base + stage * C::STAGE_BYTES + WG_A_BYTES);
#pragma unroll
        for (int k = 0; k < WG_KBLK / 16; ++k) {
          umma_bf16(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, accum);
          accum = 1u;
        }
        umma_commit(empty_bar(stage));
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    // epilogue: TMEM lane = co row, columns = ci; fp32 reduction into dwpk
    const int e = warp - 4;
    mbar_wait(done_bar, 0u);
    tc_fence_after();
    const int co = mb * 128 + e * 32 + lane;
    constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
    for (int chunk = 0; chunk < BN / CH; ++chunk) {
      uint32_t v[CH];
      tmem_ld_cols(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(chunk * CH), v);
      tmem_ld_wait();
      if (nsteps > 0 && co < p.Cout) {
        float* dst = p.dwpk + ((int64_t)tap * p.Cout + co) * p.Cin + nb * BN + chunk * CH;
#pragma unroll
        for (int i = 0; i < CH; ++i)
          if (nb * BN + chunk * CH + i < p.Cin) atomicAdd(dst + i, __uint_as_float(v[i]));
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
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgParams& p, int grid,
                        cudaStream_t s) {
  using C = WgCfg<BN>;
  auto kern = wgrad_tc_kernel<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  kern<<<grid, WG_THREADS, C::SMEM_BYTES, s>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}

}  // namespace tc
}  // namespace ast

using namespace ast;
using namespace ast::tc;

extern "C" int ast_conv3x3_wgrad(const void* dz_planar, const void* x_planar3, float* dwpk, int N,
                                 int H, int W, int Cin, int Cout, int wp, void* stream) {
  const void* x_planar = x_planar3;
  if (!dz_planar || !x_planar || !dwpk || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0)
    return AST_E_BADARG;
  if (wp < W + 2 || wp % 8 != 0) return AST_E_SHAPE;
  const int64_t ldq = (int64_t)N * (H + 2) * wp;
  const int64_t Q = ldq;
  if (ldq >= 0x7fffffffLL) return AST_E_SHAPE;
  if (Cin % 16 != 0) return AST_E_SHAPE;
  if (!aligned16(dz_planar) || !aligned16(x_planar)) return AST_E_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  int BN = 16;
  if (Cin % 256 == 0) BN = 256;
  else if (Cin % 128 == 0) BN = 128;
  else if (Cin % 64 == 0) BN = 64;
  else if (Cin % 32 == 0) BN = 32;
  WgParams p = {};
  p.Cin = Cin; p.Cout = Cout; p.wp = wp;
  p.m_blocks = (Cout + 127) / 128;
  p.n_blocks = Cin / BN;
  p.ksteps_total = (int)((Q + WG_KBLK - 1) / WG_KBLK);
  p.dwpk = dwpk;
  // split K so that the grid is a few waves of the 148 SMs, with chunks of at least 16 stages
  const int items = 9 * p.m_blocks * p.n_blocks;
  int chunks = (3 * 148 + items - 1) / items;
  int max_chunks = p.ksteps_total / 16;
  if (max_chunks < 1) max_chunks = 1;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  p.ksteps_per_chunk = (p.ksteps_total + chunks - 1) / chunks;
  p.k_chunks = (p.ksteps_total + p.ksteps_per_chunk - 1) / p.ksteps_per_chunk;
  const int grid = items * p.k_chunks;

  AST_CUDA(cudaMemsetAsync(dwpk, 0, sizeof(float) * 9 * (size_t)Cout * Cin, s));
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)ldq, (uint64_t)Cout};
    const uint64_t str[1] = {(uint64_t)ldq * 2};
    const uint32_t box[2] = {WG_KBLK, 128};
    int r = encode_bf16_map(&tmA, dz_planar, 2, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {(uint64_t)ldq, (uint64_t)Cin, 3};
    const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)ldq * 2 * Cin};
    const uint32_t box[3] = {WG_KBLK, (uint32_t)BN, 1};
    int r = encode_bf16_map(&tmB, x_planar, 3, dims, str, box);
    if (r) return r;
  }
  switch (BN) {
    case 256: return launch_wgrad<256>(tmA, tmB, p, grid, s);
    case 128: return launch_wgrad<128>(tmA, tmB, p, grid, s);
    case 64: return launch_wgrad<64>(tmA, tmB, p, grid, s);
    case 32: return launch_wgrad<32>(tmA, tmB, p, grid, s);
    default: return launch_wgrad<16>(tmA, tmB, p, grid, s);
  }
}
