// K2l: the last decoder convolution (models.py:626-627: ReflectionPad2d(1) + Conv2d(64 -> 3, 3x3), no ReLU; Hardtanh
// when exporting) with the nine taps moved from K into N.
//
// With 3 output channels the implicit GEMM of conv_tc.cu has N = 16 (padded) and K = 9 x 64: 36 tcgen05.mma per
// 128-pixel tile, each of which costs the fixed ~45 cycles any N <= 32 instruction costs (profiles/r1_ubench_mma_rate.txt),
// so that kernel is MMA-ISSUE bound at 3.4x the time of its HBM traffic (0.58 ms per 32 images at 512^2).  Here the
// GEMM is transposed in the tap index:
//     D[p][(tap, co)] = sum_ci X[p][ci] * W[tap][co][ci]          M = 128 padded-input pixels, N = 9 x 16, K = 64
//     out[h][w][co]   = bias[co] + sum_{kh,kw} D[(h + kh, w + kw)][(kh*3 + kw, co)]
// i.e. ONE K = 64 product per input pixel against all nine taps at once (4 instructions of N = 144 per tile instead
// of 36 of N = 16), followed by a 3x3 gather of the per-tap partial sums in shared memory.  The packed weights
// [9][16][64] that ast_pack_conv_weight(cout_pad = 16) already produces are, read as a [144][64] matrix, exactly the
// K-major B operand.  A D tile of 8 x 16 padded-input pixels yields 6 x 14 outputs (66 % of the MMA rows are useful;
// neighbouring tiles re-read the overlap from L2), so the kernel does ~6 instruction-equivalents per 128 outputs
// and becomes HBM-bound: algorithmic bytes per launch = N*(H+2)*(W+2)*64*2 read + N*3*H*W*4 written.
// Roles: warp 0 TMA (A ring of 6 tiles = 96 KB in flight per SM: 4 left the loads latency-bound, 10 gained nothing, B once), warp 1 MMA + TMEM (2 accumulator stages of 256 columns), then TWO
// epilogue groups of 4 warps (tcgen05.ld -> shared fp32 [128][29] -> 3x3 gather -> fp32 NCHW stores), group g owning
// accumulator stage g, its own gather buffer and every second tile: one group's chain of latencies (TMEM load, two
// barriers, shared round trip, ~900 cycles per tile measured with a single group) overlaps the other's.
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int LT_THREADS = 320;                      // TMA, MMA, 2 x 4 epilogue warps
constexpr int LT_DH = 8, LT_DW = 16;                 // D tile (padded-input pixels)
constexpr int LT_OH = LT_DH - 2, LT_OW = LT_DW - 2;  // outputs per tile
constexpr int LT_N = 144;                            // 9 taps x 16 (padded) output channels
constexpr int LT_A_BYTES = 128 * 64 * 2;             // 16 KB
constexpr int LT_B_BYTES = LT_N * 64 * 2;            // 18 KB
constexpr int LT_STAGES = 6;
constexpr int LT_DS = 29;                            // fp32 row stride of the gather buffer (odd: conflict-free)
constexpr int LT_GATHER_BYTES = 128 * LT_DS * 4;
constexpr int LT_BAR_OFF = LT_STAGES * LT_A_BYTES + LT_B_BYTES + 1024 /* B starts 1024-aligned */;
constexpr int LT_NBARS = 2 * LT_STAGES + 5;            // full/empty ring, B, 2 x accumulator full/empty
constexpr int LT_GBUF_OFF = LT_BAR_OFF + (LT_NBARS + 1) * 8 + 8;   // + TMEM slot; 16-byte aligned
constexpr int LT_SMEM_USED = LT_GBUF_OFF + 2 * LT_GATHER_BYTES + 1024;
// request > half of the SM's shared memory: the kernel allocates all 512 TMEM columns, so two CTAs must never share an SM
constexpr int LT_SMEM = LT_SMEM_USED > 120 * 1024 ? LT_SMEM_USED : 120 * 1024;

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
  auto acc_empty = [&](int s) { return bars + 8u * (2 * LT_STAGES + 3 + s); };
  const uint32_t tmem_slot = bars + 8u * LT_NBARS;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + LT_BAR_OFF + 8 * LT_NBARS);
  float* gbuf = reinterpret_cast<float*>(smem + LT_GBUF_OFF);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < LT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(b_bar, 1);
      for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 128); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_bar, LT_B_BYTES);
      tma_load_2d(b_addr, &tmB, b_bar, 0, 0);
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
          umma_bf16(tmem_base + (uint32_t)(as * 256), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc,
                    k ? 1u : 0u);
        umma_commit(empty_bar(stage));
        umma_commit(acc_full(as));
      }
      __syncwarp();
      if (++stage == LT_STAGES) { stage = 0; phase ^= 1u; }
      if (++as == 2) { as = 0; aphase ^= 1u; }
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
      if ((it & 1) != grp) continue;
      const int n = t / per_img;
      const int r = t - n * per_img;
      const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
      mbar_wait(acc_full(grp), aphase);
      tc_fence_after();
      uint32_t v[9][4];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[tap][0]), "=r"(v[tap][1]), "=r"(v[tap][2]), "=r"(v[tap][3])
                     : "r"(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(grp * 256 + tap * 16))
                     : "memory");
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(acc_empty(grp));                // accumulator stage free for the MMA warp
      aphase ^= 1u;
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // previous gather finished reading gb
      else          asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int c = 0; c < 3; ++c) gb[px * LT_DS + tap * 3 + c] = __uint_as_float(v[tap][c]);
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else          asm volatile("bar.sync 2, 128;" ::: "memory");
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
    tmem_dealloc<512>(tmem_base);
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
    const uint64_t dims[2] = {64, LT_N};
    const uint64_t str[1] = {64 * 2};
    const uint32_t box[2] = {64, LT_N};
    int r = encode_bf16_map(&tmB, wpk16, 2, dims, str, box);
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
