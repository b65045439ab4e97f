// K2wn: weight (and bias) gradient of a 3x3 stride-1 convolution straight from the NATIVE NHWC tensors, on tcgen05.
//
// Replaces what autograd computes for nn.Conv2d.weight / .bias in the reference's decoder training step
// (train.py:287-300 over the convs of models.py:598-628), and replaces K2w + its four channel-planar copies
// (csrc/wgrad_tc.cu, ast_native_to_planar): no transpose of either operand is made.
//
//   dW[co][ci][kh][kw] = sum_{n,y,x} dZ[n][y][x][co] * Xpad[n][y+kh][x+kw][ci]
//
// The contraction runs over pixels, the OUTER dimension of both NHWC operands, so both are MN-major UMMA operands
// (as in K4w, csrc/pw_wgrad_tc.cu): a TMA box {64 channels, bw, bh} lands as bw*bh pixel rows of 128 B with the
// 128-byte swizzle = the canonical MN-major SWIZZLE_128B layout (8 pixels per 1024-byte group, SBO = 1024 B,
// LBO = the next 64-channel box).  One K chunk = a bw x bh = 64-pixel patch of one image (bw = 64 / 32 / 16).
//
// What makes it cheap.  K2w ran one tap per work item: both operands were re-read from L2 nine times and the
// 64-channel layers moved 6.6 TB/s for 200 TFLOP/s.  Here a CTA owns (kh, 128 rows of Cout, <= 128 columns of Cin)
// and computes the THREE kw taps of its kh from ONE dZ box and ONE X box {64 ch, bw + 2, bh} that carries the two
// extra columns: the kw shift is a 128-byte offset of the B descriptor's start address (the tensor core's 128-byte
// swizzle is a function of the absolute shared-memory address -- DESIGN.md K2p -- so a start inside a 1024-byte
// group reads what TMA wrote), each 16-pixel MMA slice lies within one patch row, and the three taps accumulate in
// three TMEM accumulators.  Operand bytes per MAC: 3-4.4x fewer than K2w.  AST_WGRAD_SEP=1 loads three aligned
// {64, bw, bh} boxes instead (A/B check of the shifted-start reads).
// dZ's halo must be ZERO (it is: train_ops._zeros_native): patches that stick out of the image multiply X values
// with zeros (TMA zero-fills out-of-tensor elements), so any H, W works.
// Split-K over CTAs, fp32 red.global.add.v4 into dwpk [9][Cout][Cin] (zeroed here); the bias gradient (column sums of
// dZ) comes from the same dZ boxes times a constant box of ones.
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int WN_THREADS = 192;            // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: epilogue
constexpr int WN_ABOX = 64 * 128;          // {64 channels x 64 pixels} = 8 KB
constexpr int WN_BSLOT = 9 * 1024;         // {64 channels x (bw + 2) * bh <= 72 pixels}
constexpr int WN_MAX_STAGES = 6;
constexpr int WN_SMEM_OPERANDS = 198 * 1024;
constexpr int WN_ONES = 2048;              // 16 pixel rows x 128 B of bf16 1.0: the B operand of the bias-gradient MMAs
constexpr int WN_SMEM = WN_SMEM_OPERANDS + WN_ONES + (2 * WN_MAX_STAGES + 1) * 8 + 16 + 1024;

struct WnParams {
  int Cin, Cout, BN, a_boxes_last, b_boxes;   // a_boxes_last: 64-channel A boxes of the last M block (1 or 2)
  int m_blocks;
  int N, H, W, bw, bh, tiles_x, tiles_y, dz_halo;
  int chunks, split, sep;
  int stage_bytes, stages, b_off, b_slot;     // B region offset within a stage, bytes per 64-channel B slot
  float* dwpk;
  float* db;                                  // optional [Cout]: column sums of dz (zeroed by the host)
};

__device__ __forceinline__ uint64_t wn_sdesc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void red_add_v4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(dst), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)),
                 "f"(__uint_as_float(d))
               : "memory");
}

__global__ void __launch_bounds__(WN_THREADS, 1)
wgrad3x3_mn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const WnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ones = base + WN_SMEM_OPERANDS;
  const uint32_t bars = ones + WN_ONES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (WN_MAX_STAGES + s); };
  const uint32_t done_bar = bars + 8u * (2 * WN_MAX_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * WN_MAX_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + WN_SMEM_OPERANDS + WN_ONES + 8 * (2 * WN_MAX_STAGES + 1));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int mb = blockIdx.y, nb = blockIdx.z;
  const int kh = blockIdx.x / p.split;
  const int sidx = blockIdx.x - kh * p.split;
  const int per = (p.chunks + p.split - 1) / p.split;
  const int c0 = sidx * per;
  const int c1 = c0 + per < p.chunks ? c0 + per : p.chunks;
  const int nk = c1 > c0 ? c1 - c0 : 0;
  const int a_boxes = (mb == p.m_blocks - 1) ? p.a_boxes_last : 2;
  // Bias gradient db[co] = sum_p dz[p][co]: the kh = 0 CTAs of the first N block multiply their dZ boxes with a constant
  // box of ones as well (four N = 16 MMAs per patch into a fourth accumulator): no separate pass over dZ.
  const bool do_db = p.db != nullptr && kh == 0 && nb == 0;
  if (threadIdx.x >= 64) {
    reinterpret_cast<uint4*>(smem + WN_SMEM_OPERANDS)[threadIdx.x - 64] =
        make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < WN_MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t bbytes = p.sep ? (uint32_t)(3 * WN_ABOX) : (uint32_t)((p.bw + 2) * p.bh * 128);
        const uint32_t bytes = (uint32_t)a_boxes * WN_ABOX + (uint32_t)p.b_boxes * bbytes;
        const int tiles_img = p.tiles_x * p.tiles_y;
        int n = c0 / tiles_img;
        int t = c0 - n * tiles_img;
        int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        int stage = 0;
        uint32_t phase = 0;
        for (int c = c0; c < c1; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), bytes);
          const uint32_t dst = base + stage * p.stage_bytes;
          const int x0 = tx * p.bw, y0 = ty * p.bh;
          for (int b = 0; b < a_boxes; ++b)
            tma_load_4d(dst + b * WN_ABOX, &tmA, full_bar(stage), mb * 128 + b * 64, x0 + p.dz_halo, y0 + p.dz_halo, n);
          for (int b = 0; b < p.b_boxes; ++b) {
            const uint32_t bd = dst + p.b_off + b * p.b_slot;
            if (p.sep) {
              for (int kw = 0; kw < 3; ++kw)
                tma_load_4d(bd + kw * WN_ABOX, &tmB, full_bar(stage), nb * p.BN + b * 64, x0 + kw, y0 + kh, n);
            } else {
              tma_load_4d(bd, &tmB, full_bar(stage), nb * p.BN + b * 64, x0, y0 + kh, n);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++n; } }
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_bf16(128, p.BN) | (1u << 15) | (1u << 16);   // both operands MN-major
      // Descriptors are start-address adds on two constants: the offsets of the 4 x 3 (slice, kw) B windows are worked
      // out once (computed per MMA -- a division and ~40 dependent scalar instructions in the one issuing thread --
      // they made every MMA cost ~160 cycles: 113 -> 86 us on dec_conv8 with the MMAs removed).
      const int slices_per_row = p.bw >> 4;
      const uint32_t row_bytes = (uint32_t)(p.bw + 2) * 128u;
      uint32_t boff[4][3];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int ry = s / slices_per_row, xk = (s - ry * slices_per_row) << 4;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
          boff[s][kw] = (p.sep ? (uint32_t)(kw * WN_ABOX + s * 2048) : ry * row_bytes + (uint32_t)(xk + kw) * 128u) >> 4;
      }
      const uint64_t a_desc0 = wn_sdesc(base, WN_ABOX);
      const uint64_t b_desc0 = wn_sdesc(base + p.b_off, (uint32_t)p.b_slot);
      const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
      const uint32_t d1 = tmem_base + (uint32_t)p.BN, d2 = tmem_base + (uint32_t)(2 * p.BN);
      const uint32_t d3 = tmem_base + (uint32_t)(3 * p.BN);
      const uint32_t idesc_db = make_idesc_bf16(128, 16) | (1u << 15) | (1u << 16);
      const uint64_t ones_desc = wn_sdesc(ones, 0);
      int stage = 0;
      uint32_t soff = 0;               // stage * stage_bytes in 16-byte units (the address field never carries out)
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int c = c0; c < c1; ++c) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t ad = a_desc0 + soff, bd = b_desc0 + soff;
#pragma unroll
          for (int s = 0; s < 4; ++s) {            // 16 pixels of the patch per MMA, three kw taps each
            const uint32_t acc = s ? 1u : accum;
            umma_bf16(tmem_base, ad + (uint64_t)(s * 128), bd + boff[s][0], idesc, acc);
            umma_bf16(d1, ad + (uint64_t)(s * 128), bd + boff[s][1], idesc, acc);
            umma_bf16(d2, ad + (uint64_t)(s * 128), bd + boff[s][2], idesc, acc);
            if (do_db) umma_bf16(d3, ad + (uint64_t)(s * 128), ones_desc, idesc_db, acc);
          }
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        accum = 1u;
        soff += stage16;
        if (++stage == p.stages) { stage = 0; soff = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(done_bar);
      __syncwarp();
    } else {
      const int e = warp & 3;                      // TMEM lane quarter this warp may read
      mbar_wait(done_bar, 0u);
      tc_fence_after();
      const int co = mb * 128 + e * 32 + lane;
      const bool vec = (p.Cin & 3) == 0;
      if (do_db) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(3 * p.BN), v);
        tmem_ld_wait();
        if (co < p.Cout) atomicAdd(p.db + co, __uint_as_float(v[0]));
      }
      for (int kw = 0; kw < 3; ++kw) {
        float* orow = p.dwpk + ((int64_t)(kh * 3 + kw) * p.Cout + co) * p.Cin;
        for (int col = 0; col < p.BN; col += 16) {
          uint32_t v[16];
          tmem_ld_32x16(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(kw * p.BN + col), v);
          tmem_ld_wait();
          if (co >= p.Cout) continue;
          const int j0 = nb * p.BN + col;
          if (vec && j0 + 16 <= p.Cin) {
#pragma unroll
            for (int t = 0; t < 16; t += 4) red_add_v4(orow + j0 + t, v[t], v[t + 1], v[t + 2], v[t + 3]);
          } else {
#pragma unroll
            for (int t = 0; t < 16; ++t)
              if (j0 + t < p.Cin) atomicAdd(orow + j0 + t, __uint_as_float(v[t]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace tc
}  // namespace ast

using namespace ast;
using namespace ast::tc;

extern "C" int ast_conv3x3_wgrad_native(const void* dz, int cz, int dz_halo, const void* x, float* dwpk, float* db,
                                        int N, int H, int W, int Cin, int Cout, void* stream) {
  if (!dz || !x || !dwpk || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || cz <= 0) return AST_E_BADARG;
  if (dz_halo < 0 || dz_halo > 2 || Cout > cz || cz % 8 != 0 || Cin % 8 != 0 || cz > 2048) return AST_E_SHAPE;
  if (!aligned16(dz) || !aligned16(x)) return AST_E_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  static const int sep_env = getenv("AST_WGRAD_SEP") ? atoi(getenv("AST_WGRAD_SEP")) : 0;
  WnParams p = {};
  p.Cin = Cin; p.Cout = Cout; p.N = N; p.H = H; p.W = W; p.dz_halo = dz_halo; p.sep = sep_env ? 1 : 0;
  p.bw = W > 32 ? 64 : (W > 16 ? 32 : 16);
  p.bh = 64 / p.bw;
  p.tiles_x = (W + p.bw - 1) / p.bw;
  p.tiles_y = (H + p.bh - 1) / p.bh;
  const int64_t chunks = (int64_t)N * p.tiles_x * p.tiles_y;
  if (chunks >= 0x7fffffffLL) return AST_E_SHAPE;
  p.chunks = (int)chunks;
  int n_blocks = 1, BN = (Cin + 15) / 16 * 16;
  while (BN > 128) {
    ++n_blocks;
    BN = ((Cin + n_blocks - 1) / n_blocks + 15) / 16 * 16;
  }
  p.BN = BN; p.b_boxes = (BN + 63) / 64;
  p.m_blocks = (Cout + 127) / 128;
  p.a_boxes_last = (Cout - (p.m_blocks - 1) * 128 > 64) ? 2 : 1;
  if (n_blocks > 65535 || p.m_blocks > 65535) return AST_E_SHAPE;
  p.b_off = 2 * WN_ABOX;
  p.b_slot = p.sep ? 3 * WN_ABOX : WN_BSLOT;
  p.stage_bytes = p.b_off + p.b_boxes * p.b_slot;
  p.stages = WN_SMEM_OPERANDS / p.stage_bytes;
  if (p.stages > WN_MAX_STAGES) p.stages = WN_MAX_STAGES;
  p.dwpk = dwpk;
  int split = 148 / (3 * p.m_blocks * n_blocks);
  if (split > p.chunks / 4) split = p.chunks / 4;     // >= 4 K chunks per CTA: the 128 x 3 BN atomic epilogue must amortise
  if (split < 1) split = 1;
  p.split = split;

  AST_CUDA(cudaMemsetAsync(dwpk, 0, sizeof(float) * 9 * (size_t)Cout * Cin, s));
  if (db) AST_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, s));
  p.db = db;
  CUtensorMap tmA, tmB;
  {
    const int Hp = H + 2 * dz_halo, Wp = W + 2 * dz_halo;
    const uint64_t dims[4] = {(uint64_t)cz, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)N};
    const uint64_t str[3] = {(uint64_t)cz * 2, (uint64_t)Wp * cz * 2, (uint64_t)Hp * Wp * cz * 2};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int r = encode_bf16_map(&tmA, dz, 4, dims, str, box);
    if (r) return r;
  }
  {
    const int Hp = H + 2, Wp = W + 2;
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)N};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Wp * Cin * 2, (uint64_t)Hp * Wp * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)(p.sep ? p.bw : p.bw + 2), (uint32_t)p.bh, 1};
    int r = encode_bf16_map(&tmB, x, 4, dims, str, box);
    if (r) return r;
  }
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(wgrad3x3_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WN_SMEM));
    attr_done = true;
  }
  wgrad3x3_mn_kernel<<<dim3((unsigned)(3 * split), p.m_blocks, n_blocks), WN_THREADS, WN_SMEM, s>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}
