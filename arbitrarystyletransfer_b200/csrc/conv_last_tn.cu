// K2l: the last decoder convolution (models.py:626-627: ReflectionPad2d(1) + Conv2d(64 -> 3, 3x3), no ReLU; Hardtanh
// when exporting) with the nine taps moved from K into N.
//
// With 3 output channels the implicit GEMM of conv_tc.cu has N = 16 (padded) and K = 9 x 64: 36 tcgen05.mma per
// 128-pixel tile, each of which costs the fixed ~45 cycles any N <= 32 instruction costs (profiles/r1_ubench_mma_rate.txt),
// so that kernel is MMA-ISSUE bound at 3.4x the time of its HBM traffic (0.58 ms per 32 images at 512^2).  Here the
// GEMM is transposed in the tap index:
//     D[p][(tap, co)] = sum_ci X[p][ci] * W[tap][co][ci]          M = 128 padded-input pixels, N = 9 x 3 (-> 32), K = 64
//     out[h][w][co]   = bias[co] + sum_{kh,kw} D[(h + kh, w + kw)][(kh*3 + kw, co)]
// i.e. ONE K = 64 product per input pixel against all nine taps at once (4 instructions of N = 32 per tile instead
// of 36 of N = 16), followed by a 3x3 gather of the per-tap partial sums in shared memory.  The B operand is cut
// out of the packed weights [9][16][64] that ast_pack_conv_weight(cout_pad = 16) already produces by ONE 3-D TMA box
// {64 ci, 3 co, 9 taps}: it lands as 27 consecutive 128-byte rows (row = tap*3 + co), i.e. the K-major [27][64]
// matrix; rows 27..31 are zeroed.  (A first version used all 16 padded channels per tap, N = 144: 85 cycles per
// instruction and nine 4-column TMEM loads per thread; N = 32 costs 45 cycles and one 32-column load.)  A D tile of 8 x 16 padded-input pixels yields 6 x 14 outputs (66 % of the MMA rows are useful;
// neighbouring tiles re-read the overlap from L2), so the kernel does ~6 instruction-equivalents per 128 outputs
// and becomes HBM-bound: algorithmic bytes per launch = N*(H+2)*(W+2)*64*2 read + N*3*H*W*4 written.
// Roles: warp 0 TMA (A ring of 6 tiles = 96 KB in flight per SM: 4 left the loads latency-bound, 10 gained nothing, B once), warp 1 MMA + TMEM (2 accumulator stages of 32 columns), then FOUR
// epilogue groups of 4 warps (tcgen05.ld -> shared fp32 [128][29] -> 3x3 gather -> fp32 NCHW stores), group g owning
// accumulator stage g, its own gather buffer and every fourth tile (1 group: 330 us, 2: 280 us, 4: 245 us, 6: 247 us at 32 x 512^2).  The epilogue is ~250 dependent instructions per
// warp per tile (27 shared stores, 81 shared loads + adds, two barriers): with one or two warps per scheduler that
// chain, not HBM or the tensor pipe, set the pace (ncu: DRAM 48 %, tensor pipe 41 %, issue 34 %).
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int LT_GROUPS = 4;                         // epilogue groups = TMEM accumulator stages
constexpr int LT_TMEM = LT_GROUPS <= 1 ? 32 : LT_GROUPS <= 2 ? 64 : LT_GROUPS <= 4 ? 128 : 256;   // 32 columns per stage
constexpr int LT_THREADS = 64 + LT_GROUPS * 128;     // TMA, MMA, LT_GROUPS x 4 epilogue warps
constexpr int LT_DH = 8, LT_DW = 16;                 // D tile (padded-input pixels)
constexpr int LT_OH = LT_DH - 2, LT_OW = LT_DW - 2;  // outputs per tile
constexpr int LT_N = 32;                             // 9 taps x 3 output channels = 27 columns, padded to 32
constexpr int LT_A_BYTES = 128 * 64 * 2;             // 16 KB
constexpr int LT_B_BYTES = LT_N * 64 * 2;            // 4 KB
constexpr int LT_B_TX = 27 * 64 * 2;                 // bytes the weight box delivers (27 rows)
constexpr int LT_STAGES = 6;
constexpr int LT_DS = 29;                            // fp32 row stride of the gather buffer (odd: conflict-free)
constexpr int LT_GATHER_BYTES = 128 * LT_DS * 4;
constexpr int LT_BAR_OFF = LT_STAGES * LT_A_BYTES + LT_B_BYTES + 1024 /* B starts 1024-aligned */;
constexpr int LT_NBARS = 2 * LT_STAGES + 1 + 2 * LT_GROUPS;   // full/empty ring, B, accumulator full/empty per group
constexpr int LT_GBUF_OFF = LT_BAR_OFF + (LT_NBARS + 1) * 8 + 8;   // + TMEM slot; 16-byte aligned
constexpr int LT_SMEM_USED = LT_GBUF_OFF + LT_GROUPS * LT_GATHER_BYTES + 1024;
constexpr int LT_SMEM = LT_SMEM_USED;

struct LtParams {
  int N, H, W, Cout, tiles_h, tiles_w, num_tiles, clamp01;
  const float* bias;
  float* out;
};

__global__ void __launch_bounds__(LT_THREADS, 1)
conv3x3_last_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const LtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t b_addr = base + LT_STAGES * LT_A_BYTES;                  // 64 KB: 1024-aligned
  const uint32_t bars = base + LT_BAR_OFF;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (LT_STAGES + s); };
  const uint32_t b_bar = bars + 8u * (2 * LT_STAGES);
  auto acc_full = [&](int s) { return bars + 8u * (2 * LT_STAGES + 1 + s); };
  auto acc_empty = [&](int s) { return bars + 8u * (2 * LT_STAGES + 1 + LT_GROUPS + s); };
  const uint32_t tmem_slot = bars + 8u * LT_NBARS;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + LT_BAR_OFF + 8 * LT_NBARS);
  float* gbuf = reinterpret_cast<float*>(smem + LT_GBUF_OFF);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  for (int i = threadIdx.x; i < (LT_B_BYTES - LT_B_TX) / 4; i += LT_THREADS)      // B rows 27..31 (never delivered)
    reinterpret_cast<uint32_t*>(smem + LT_STAGES * LT_A_BYTES + LT_B_TX)[i] = 0u;
  fence_proxy_async_smem();
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < LT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(b_bar, 1);
      for (int s = 0; s < LT_GROUPS; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 128); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<LT_TMEM>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_bar, LT_B_TX);
      tma_load_3d(b_addr, &tmB, b_bar, 0, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int n = t / per_img;
        const int r = t - n * per_img;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), LT_A_BYTES);
        tma_load_4d(base + stage * LT_A_BYTES, &tmA, full_bar(stage), 0, tx * LT_OW, ty * LT_OH, n);
        if (++stage == LT_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, LT_N);
    mbar_wait(b_bar, 0u);
    tc_fence_after();
    const uint64_t bd = make_sdesc_k128(b_addr);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      mbar_wait(acc_empty(as), aphase ^ 1u);
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint64_t ad = make_sdesc_k128(base + stage * LT_A_BYTES);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + (uint32_t)(as * 32), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc,
                    k ? 1u : 0u);
        umma_commit(empty_bar(stage));
        umma_commit(acc_full(as));
      }
      __syncwarp();
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1u; }
      if (++as == LT_GROUPS) { as = 0; aphase ^= 1u; }
    }
  } else {
    const int e = warp & 3;                       // TMEM lane quarter of this warp
    const int grp = (warp - 2) >> 2;              // epilogue group = accumulator stage
    const int px = e * 32 + lane;                 // D row = padded-input pixel of the tile (r * 16 + c)
    const int et = ((warp - 2) & 3) * 32 + lane;  // 0..127: output slot of this thread
    const int orow = et / LT_OW, ocol = et - orow * LT_OW;
    float* gb = gbuf + grp * (128 * LT_DS);
    float bias[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bias[c] = (p.bias && c < p.Cout) ? p.bias[c] : 0.f;
    uint32_t aphase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      if ((it % LT_GROUPS) != grp) continue;
      const int n = t / per_img;
      const int r = t - n * per_img;
      const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
      mbar_wait(acc_full(grp), aphase);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(grp * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(acc_empty(grp));                // accumulator stage free for the MMA warp
      aphase ^= 1u;
      asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");   // previous gather finished reading gb
#pragma unroll
      for (int j = 0; j < 27; ++j) gb[px * LT_DS + j] = __uint_as_float(v[j]);
      asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
      if (et < LT_OH * LT_OW) {
        const int h = ty * LT_OH + orow, w = tx * LT_OW + ocol;
        if (h < p.H && w < p.W) {
          float a[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float* g = gb + ((orow + kh) * LT_DW + ocol + kw) * LT_DS + (kh * 3 + kw) * 3;
              a[0] += g[0]; a[1] += g[1]; a[2] += g[2];
            }
          const int64_t plane = (int64_t)p.H * p.W;
          float* o = p.out + (int64_t)n * p.Cout * plane + (int64_t)h * p.W + w;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < p.Cout) {
              float y = a[c];
              if (p.clamp01) y = fminf(fmaxf(y, 0.f), 1.f);
              o[c * plane] = y;
            }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<LT_TMEM>(tmem_base);
  }
}

// in: bf16 [N][H+2][W+2][64] with its halo filled; wpk16: bf16 [9][16][64]; out: fp32 [N][Cout][H][W], Cout <= 3.
int conv3x3_last_tn(const void* in, const void* wpk16, const float* bias, float* out, int N, int H, int W, int Cout,
                    int clamp01, int sm_count, cudaStream_t s) {
  if (Cout < 1 || Cout > 3 || H < 1 || W < 1) return AST_E_SHAPE;
  if (!aligned16(in) || !aligned16(wpk16)) return AST_E_ALIGN;
  LtParams p = {};
  p.N = N; p.H = H; p.W = W; p.Cout = Cout; p.clamp01 = clamp01; p.bias = bias; p.out = out;
  p.tiles_h = (H + LT_OH - 1) / LT_OH;
  p.tiles_w = (W + LT_OW - 1) / LT_OW;
  const int64_t nt = (int64_t)N * p.tiles_h * p.tiles_w;
  if (nt >= 0x7fffffffLL) return AST_E_SHAPE;
  p.num_tiles = (int)nt;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {64, (uint64_t)(W + 2), (uint64_t)(H + 2), (uint64_t)N};
    const uint64_t str[3] = {64 * 2, (uint64_t)(W + 2) * 64 * 2, (uint64_t)(H + 2) * (W + 2) * 64 * 2};
    const uint32_t box[4] = {64, LT_DW, LT_DH, 1};
    int r = encode_bf16_map(&tmA, in, 4, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {64, 16, 9};                    // [tap][co (padded to 16)][ci]
    const uint64_t str[2] = {64 * 2, 16 * 64 * 2};
    const uint32_t box[3] = {64, 3, 9};                      // 27 rows: tap*3 + co
    int r = encode_bf16_map(&tmB, wpk16, 3, dims, str, box);
    if (r) return r;
  }
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(conv3x3_last_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    attr_done = true;
  }
  const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  conv3x3_last_tn_kernel<<<grid, LT_THREADS, LT_SMEM, s>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}

}  // namespace tc
}  // namespace ast
