// K4p: pointwise (1x1) convolution of the MobileNet-style blocks as a tcgen05 GEMM.
//
// Reference work replaced (paths relative to /root/reference): the nn.Conv2d(.., 1, 1, 0) layers of
// DepthWiseConv (mobilenetv2.py:95-165) with their eval-mode BatchNorm2d folded into (weight, bias),
// the following nn.Hardswish, the SELayer channel scaling of the INPUT (mobilenetv2.py:63-81,
// folded into per-sample weights) and the block's residual add (mobilenetv2.py:161-162).
//
// Layout: plain NHWC bf16, i.e. a [pixels][C] matrix per image -- already the K-major A operand.
//   D[pixel][co] = sum_ci X[pixel][ci] * W[n?][co][ci]      M = 128 pixels, N = Cout block, K = Cin
// A: 3-D TMA box {64 ci, 128 pixels, 1 image} (row stride ld_in, so a channel slice of a wider buffer
//    works: that is how torch.cat of the two encoder taps, models.py:332, costs nothing);
// B: 3-D TMA box {64 ci, BN co, 1 (image or 0)}.  Channel counts that are not multiples of 64 / 16
// rely on TMA zero fill for the K tail and on epilogue masking for the N tail.
// These layers are HBM-bound (K <= 768).  PERSISTENT CTAs (one or two per SM) walk the 128-pixel tiles: a TMA warp
// runs a ring of operand stages AHEAD across tile boundaries, the MMA warp alternates two TMEM accumulators, and eight
// epilogue warps (two per TMEM lane quarter, splitting the columns) drain one accumulator while the next tile's MMAs
// run.  (Round 1 launched one CTA per tile: TMEM allocation, barrier set-up, the first TMA round trip and the
// epilogue's latency chain were serial inside every CTA and only 2-3 CTAs fitted an SM: 19-25 % of the DRAM
// bandwidth, profiles/r1_ncu_pointwise_summary.csv.)
#include <cstdlib>
#include "tc.cuh"
#ifndef AST_KERNEL_DEBUG
#define AST_KERNEL_DEBUG 0
#endif

namespace ast {
namespace tc {

constexpr int PW_EPI_WARPS = 16;           // four per TMEM lane quarter: one 16-column chunk each per 64-channel block
constexpr int PW_THREADS = 64 + 32 * PW_EPI_WARPS;   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-17: epilogue
constexpr int PW_A_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int PW_MAX_STAGES = 8;
__host__ __device__ constexpr int pw_stage_bytes(int BN) { return PW_A_BYTES + ((BN * 128 + 1023) / 1024) * 1024; }
constexpr int PW_STG_BYTES = 128 * 128;   // one staged output block: 128 pixel rows x 64 channels, 128-byte swizzle
__host__ __device__ constexpr int pw_b_slot_bytes(int BN) { return ((BN * 128 + 1023) / 1024) * 1024; }
// resident > 0: the whole weight matrix (resident = K steps of one slot each) sits behind an A-only ring
__host__ __device__ constexpr int pw_smem_bytes(int BN, int stages, int staging = 0, int resident = 0) {
  return stages * (resident ? PW_A_BYTES : pw_stage_bytes(BN)) + resident * pw_b_slot_bytes(BN) + staging +
         (2 * PW_MAX_STAGES + 5) * 8 + 16 + 1024;
}

struct PwParams {
  int N, Cin, Cout, BN, n_blocks, tiles_per_img, per_sample_w, act;
  int stages, total_tiles;         // operand ring depth; N * tiles_per_img * n_blocks
  int staging;                     // bytes of TMA-store staging after the ring: 2 buffers (x 2 with out_act)
  int resident;                    // > 0: weights loaded ONCE per CTA into `resident` (= K steps) slots; else streamed
                                   // with every A stage.  (Measured: the TMA unit handles one box ROW per ~18 cycles;
                                   // re-streaming BN weight rows per 128 pixel rows made the expand layers
                                   // TMA-row-bound: 368 rows per tile for 40 -> 240.)
  int dbg_flags;                   // AST_PW_DBGFLAGS (bottleneck elimination, results are WRONG): 1 = no TMA stores,
                                   // 2 = no staging writes, 4 = no TMEM loads
  int64_t HW;
  int ld_out, ld_res;              // row strides (elements) of out / residual
  const float* bias;               // [Cout] or null
  int f16;                         // all 16-bit tensors of the call are fp16 (forward activations / weights) instead of
                                   // bf16 (gradients, attention rows): instruction descriptor + epilogue conversions
  const uint16_t* residual;        // [N*HW][ld_res] or null
  uint16_t* out;                   // [N*HW][ld_out]
  uint16_t* out_act;               // optional [N*HW][ld_act]: Hardswish(out) while out keeps the raw value
  int ld_act;
  int res_w;                       // > 0: residual is read through a nearest x2 upsample; res_w = output width
};

__device__ __forceinline__ float hardswish(float x) {
  return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);   // nn.Hardswish; this rounding is the one the tests pin bit for bit
}

// F16: every 16-bit tensor of the call is fp16 (forward activations / weights), else bf16 (gradients, attention rows);
// compile-time, so the epilogue's conversions carry no selects
template <int TMEM_COLS, bool F16, bool DBG>
__global__ void __launch_bounds__(PW_THREADS)
pw_conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAct, const PwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int dflags = DBG ? p.dbg_flags : 0;      // elimination flags live in a separate instantiation
  const int STAGE = p.resident ? PW_A_BYTES : pw_stage_bytes(p.BN);
  const int NS = p.stages;
  const int B_SLOT = pw_b_slot_bytes(p.BN);
  const uint32_t bres_base = base + NS * STAGE;    // resident weights (p.resident slots); 1024-byte aligned
  const uint32_t stg_base = bres_base + p.resident * B_SLOT;
  const uint32_t bars = stg_base + p.staging;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (PW_MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * PW_MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * PW_MAX_STAGES + 2 + s); };
  const uint32_t bres_bar = bars + 8u * (2 * PW_MAX_STAGES + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * PW_MAX_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + NS * STAGE + p.resident * B_SLOT + p.staging + 8 * (2 * PW_MAX_STAGES + 5));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  constexpr uint32_t ACC_COLS = TMEM_COLS / 2;     // two accumulators

  const int ksteps = (p.Cin + 63) / 64;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;
  // tile t -> (n-block, 128-pixel tile of the image, image); the n-blocks of one pixel tile are neighbours in the
  // walk, so its A tile is re-read from L2
  auto decode = [&](int t, int& nb, int& ti, int& n) {
    nb = t % p.n_blocks; t /= p.n_blocks;
    ti = t % p.tiles_per_img;
    n = t / p.tiles_per_img;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmOut);
    if (p.out_act) tma_prefetch_desc(&tmAct);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), PW_EPI_WARPS); }
      mbar_init(bres_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (p.resident) {      // one n-block, shared weights: the whole matrix, once
        mbar_expect_tx(bres_bar, (uint32_t)ksteps * b_bytes);
        for (int ks = 0; ks < ksteps; ++ks) tma_load_3d(bres_base + ks * B_SLOT, &tmB, bres_bar, ks * 64, 0, 0);
      }
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int nb, ti, n;
        decode(t, nb, ti, n);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), PW_A_BYTES + (p.resident ? 0u : b_bytes));
          const uint32_t a_dst = base + stage * STAGE;
          tma_load_3d(a_dst, &tmA, full_bar(stage), ks * 64, ti * 128, n);
          if (!p.resident)
            tma_load_3d(a_dst + PW_A_BYTES, &tmB, full_bar(stage), ks * 64, nb * p.BN, p.per_sample_w ? n : 0);
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = F16 ? make_idesc_f16(128, p.BN) : make_idesc_bf16(128, p.BN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, aphase = 0;
    if (p.resident && (int)blockIdx.x < p.total_tiles) mbar_wait(bres_bar, 0u);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(tempty_bar(acc), aphase ^ 1u);       // the epilogue has drained this accumulator (two tiles ago)
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
      uint32_t accum = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t ad = make_sdesc_k128(base + stage * STAGE);
        const uint64_t bd = make_sdesc_k128(p.resident ? bres_base + ks * B_SLOT : base + stage * STAGE + PW_A_BYTES);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (ks * 64 + k * 16 < p.Cin)   // K tail: skip 16-channel steps that are pure TMA zero fill
              umma_bf16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k ? 1u : accum);
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        accum = 1u;
        if (++stage == NS) { stage = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(tfull_bar(acc));
      __syncwarp();
      if (++acc == 2) { acc = 0; aphase ^= 1u; }
    }
  } else {
    // epilogue: warps 2..17 -> TMEM lane quarter (warp % 4); the four warps of a quarter take one 16-column chunk
    // each of a 64-channel block.  The output leaves through shared memory: per block the sixteen warps pack their
    // 128 pixel rows (128 B each) into a staging buffer in the 128-byte-swizzle layout (chunk c of row r at
    // c ^ (r & 7): conflict-free 16-byte st.shared), fence.proxy.async, a named barrier, and ONE lane issues a
    // cp.async.bulk.tensor store of the {64 ch, 128 px} box (two with the Hardswish copy): whole 128-byte lines
    // instead of one 32-byte sector per lane and store.  TMA clips ragged pixel tiles and channel tails against the
    // tensor's bounds.  Two buffers alternate; cp.async.bulk.wait_group.read guards their reuse.
    // (ncu: the round-1 epilogue was instruction-bound -- 25 000 warp-instructions per 128 x 240 tile on ten warps per
    // SM -- hence sixteen warps, the dtype as a template parameter and vector bias loads.)
    const int e = warp & 3;
    const int ew = warp - 2;
    const int sub = ew >> 2;                         // 16-column chunk of the block
    const int r_t = e * 32 + lane;                   // pixel row of the tile
    const uint32_t sw = (uint32_t)(r_t & 7);
    const uint32_t act_off = 2 * PW_STG_BYTES;       // the Hardswish copy's two buffers follow the raw ones
    const bool bias_vec = p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
    int acc = 0, buf = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int nb, ti, n;
      decode(t, nb, ti, n);
      mbar_wait(tfull_bar(acc), aphase);
      tc_fence_after();
      const uint32_t acc_tmem = tmem_base + (uint32_t)acc * ACC_COLS + ((uint32_t)(e * 32) << 16);
      const int64_t pix = (int64_t)ti * 128 + r_t;     // pixel inside the image
      const bool ok = pix < p.HW;
      const int64_t row = (int64_t)n * p.HW + pix;
      int64_t rrow = row;
      if (p.res_w > 0 && ok) {   // residual = the block input BEFORE the nearest x2 upsample (models.py:265-267)
        const int h = (int)(pix / p.res_w), w = (int)(pix % p.res_w);
        rrow = (int64_t)n * (p.HW >> 2) + (int64_t)(h >> 1) * (p.res_w >> 1) + (w >> 1);
      }
      for (int b0 = 0; b0 < p.BN; b0 += 64) {
        if (nb * p.BN + b0 >= p.Cout) break;           // uniform: nothing of this block exists
        if (ew == 0 && lane == 0) bulk_wait_group_read1();   // the stores issued two blocks ago have read `buf`
        named_bar_sync(1, 32 * PW_EPI_WARPS);
        const uint32_t srow = stg_base + (uint32_t)buf * PW_STG_BYTES + (uint32_t)r_t * 128u;
        const int c0 = b0 + sub * 16;
        if (c0 < p.BN) {
          uint32_t v[16];
          if (!(dflags & 4)) {
            tmem_ld_32x16(acc_tmem + (uint32_t)c0, v);
            tmem_ld_wait();
          }
          const int co0 = nb * p.BN + c0;
          const int valid = ok ? p.Cout - co0 : 0;     // < 16: channel tail (multiple of 8); <= 0: TMA clips the row / chunk
          float f[16];
          if (valid >= 16 && (bias_vec || !p.bias)) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + co0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
              f[i + 0] = __uint_as_float(v[i + 0]) + b.x;
              f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
              f[i + 2] = __uint_as_float(v[i + 2]) + b.z;
              f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              f[i] = __uint_as_float(v[i]);
              if (p.bias && i < valid) f[i] += __ldg(p.bias + co0 + i);
            }
          }
          if (p.act && !p.out_act) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = hardswish(f[i]);
          }
          if (p.residual) {
            const uint16_t* rp = p.residual + rrow * p.ld_res + co0;
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              if (i < valid) {
                float r[8];
                unpack8_dt(__ldg(reinterpret_cast<const uint4*>(rp + i)), r, F16);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[i + j] += r[j];
              }
            }
          }
          uint32_t pk[8], pa[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pk[j] = pk2_dt(f[2 * j], f[2 * j + 1], F16);
            if (p.out_act) {   // training: raw pre-activation in `out`, Hardswish of the ROUNDED value beside it
              const float2 rr = un2_dt(pk[j], F16);
              pa[j] = pk2_dt(hardswish(rr.x), hardswish(rr.y), F16);
            }
          }
          const uint32_t c16 = (uint32_t)(sub * 2);
          if (!(dflags & 2)) {
          st_shared_v4(srow + ((c16 ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(srow + (((c16 + 1) ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
          if (p.out_act) {
            st_shared_v4(srow + act_off + ((c16 ^ sw) << 4), pa[0], pa[1], pa[2], pa[3]);
            st_shared_v4(srow + act_off + (((c16 + 1) ^ sw) << 4), pa[4], pa[5], pa[6], pa[7]);
          }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 32 * PW_EPI_WARPS);
        if (ew == 0 && lane == 0 && !(dflags & 1)) {
          const uint32_t src = stg_base + (uint32_t)buf * PW_STG_BYTES;
          tma_store_3d(&tmOut, src, nb * p.BN + b0, ti * 128, n);
          if (p.out_act) tma_store_3d(&tmAct, src + act_off, nb * p.BN + b0, ti * 128, n);
          bulk_commit_group();
        }
        buf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));     // accumulator free for the tile after next
      if (++acc == 2) { acc = 0; aphase ^= 1u; }
    }
    if (ew == 0 && lane == 0) bulk_wait_group0();      // the staging buffers outlive their stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace tc
}  // namespace ast

using namespace ast;
using namespace ast::tc;

extern "C" int ast_pw_conv(const void* x, int ld_in, const void* w, int per_sample_w, const float* bias,
                           int act, const void* residual, int ld_res, void* out, int ld_out, int N,
                           int64_t HW, int Cin, int Cout, void* out_act, int ld_act, int res_up2_w, int dtype,
                           void* stream) {
  if (!x || !w || !out || N <= 0 || HW <= 0 || Cin <= 0 || Cout <= 0) return AST_E_BADARG;
  if (dtype != AST_DT_BF16 && dtype != AST_DT_F16) return AST_E_BADARG;
  if (out_act && (ld_act % 8 != 0 || ld_act < Cout || !aligned16(out_act) || residual)) return AST_E_SHAPE;
  if (Cin % 8 != 0 || Cout % 8 != 0 || ld_in % 8 != 0 || ld_out % 8 != 0 || (residual && ld_res % 8 != 0))
    return AST_E_SHAPE;
  if (ld_in < Cin || ld_out < Cout || HW >= 0x7fffffffLL) return AST_E_SHAPE;
  if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (residual && !aligned16(residual))) return AST_E_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  // N block: whole Cout (rounded up to 16) when <= 256, else the fewest equal blocks of a multiple of 16
  // (with more than one block the N block is a multiple of 64: the epilogue stores {64 ch, 128 px} boxes, and a box
  // that crossed into the next block's channels would overwrite them; only the tensor's own edge is clipped by TMA)
  int n_blocks = 1, BN = (Cout + 15) / 16 * 16;
  while (BN > 256) {
    ++n_blocks;
    BN = ((Cout + n_blocks - 1) / n_blocks + 63) / 64 * 64;
  }
  PwParams p = {};
  p.N = N; p.Cin = Cin; p.Cout = Cout; p.BN = BN; p.n_blocks = n_blocks; p.HW = HW;
  p.tiles_per_img = (int)((HW + 127) / 128);
  p.per_sample_w = per_sample_w; p.act = act;
  p.ld_out = ld_out; p.ld_res = ld_res;
  p.f16 = dtype == AST_DT_F16;
  p.bias = bias; p.residual = reinterpret_cast<const uint16_t*>(residual);
  p.out = reinterpret_cast<uint16_t*>(out);
  p.out_act = reinterpret_cast<uint16_t*>(out_act);
  p.ld_act = ld_act;
  if (res_up2_w < 0 || (res_up2_w > 0 && (!residual || res_up2_w % 2 != 0 || HW % res_up2_w != 0 ||
                                          (HW / res_up2_w) % 2 != 0)))
    return AST_E_SHAPE;
  p.res_w = res_up2_w;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)HW, (uint64_t)N};
    const uint64_t str[2] = {(uint64_t)ld_in * 2, (uint64_t)HW * ld_in * 2};
    const uint32_t box[3] = {64, 128, 1};
    int r = encode_bf16_map(&tmA, x, 3, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(per_sample_w ? N : 1)};
    const uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    int r = encode_bf16_map(&tmB, w, 3, dims, str, box);
    if (r) return r;
  }
  const int64_t total = (int64_t)N * p.tiles_per_img * n_blocks;
  if (total >= 0x7fffffffLL) return AST_E_SHAPE;
  p.total_tiles = (int)total;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    AST_CUDA(cudaGetDevice(&dev));
    AST_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  // two CTAs per SM while two accumulators of BN columns each fit twice into the 512 TMEM columns (BN <= 128),
  // else one; the operand ring takes what is left of the shared memory (2 .. 8 stages)
  p.staging = (out_act ? 4 : 2) * PW_STG_BYTES;
  static const int dbg_flags = getenv("AST_PW_DBGFLAGS") ? atoi(getenv("AST_PW_DBGFLAGS")) : 0;
  p.dbg_flags = dbg_flags;
  // weights resident when they are shared by all images, one n-block covers Cout and they take <= 96 KB
  const int ksteps = (Cin + 63) / 64;
  static const int res_env = getenv("AST_PW_RESIDENT") ? atoi(getenv("AST_PW_RESIDENT")) : 1;
  p.resident = (res_env && !per_sample_w && n_blocks == 1 && ksteps * pw_b_slot_bytes(BN) <= 96 * 1024) ? ksteps : 0;
  const int stage_b = p.resident ? PW_A_BYTES : pw_stage_bytes(BN);
  int per_sm = BN <= 128 ? 2 : 1;
  int stages = (110 * 1024 - pw_smem_bytes(BN, 0, p.staging, p.resident)) / stage_b;
  if (per_sm == 2 && stages < 2) per_sm = 1;       // staging + a two-stage ring do not fit twice: one CTA per SM
  if (per_sm == 1) stages = (220 * 1024 - pw_smem_bytes(BN, 0, p.staging, p.resident)) / stage_b;
  stages = stages < 2 ? 2 : (stages > PW_MAX_STAGES ? PW_MAX_STAGES : stages);
  p.stages = stages;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const PwParams);
  // first index: TMEM columns = two accumulators of the N block rounded up to a power of two
  static const KernelFn kerns[4][2] = {{pw_conv_tc_kernel<64, false, false>, pw_conv_tc_kernel<64, true, false>},
                                       {pw_conv_tc_kernel<128, false, false>, pw_conv_tc_kernel<128, true, false>},
                                       {pw_conv_tc_kernel<256, false, false>, pw_conv_tc_kernel<256, true, false>},
                                       {pw_conv_tc_kernel<512, false, false>, pw_conv_tc_kernel<512, true, false>}};
#if AST_KERNEL_DEBUG
  static const KernelFn kerns_dbg[4][2] = {{pw_conv_tc_kernel<64, false, true>, pw_conv_tc_kernel<64, true, true>},
                                           {pw_conv_tc_kernel<128, false, true>, pw_conv_tc_kernel<128, true, true>},
                                           {pw_conv_tc_kernel<256, false, true>, pw_conv_tc_kernel<256, true, true>},
                                           {pw_conv_tc_kernel<512, false, true>, pw_conv_tc_kernel<512, true, true>}};
#else
  const KernelFn (*kerns_dbg)[2] = kerns;      // elimination flags need a build with AST_KERNEL_DEBUG=1
#endif
  static bool attr_done = false;
  if (!attr_done) {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 2; ++j) {
        AST_CUDA(cudaFuncSetAttribute(kerns[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        AST_CUDA(cudaFuncSetAttribute(kerns_dbg[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      }
    attr_done = true;
  }
  const int64_t max_ctas = (int64_t)sm_count * per_sm;
  const unsigned grid = (unsigned)(total < max_ctas ? total : max_ctas);
  const int smem = pw_smem_bytes(BN, stages, p.staging, p.resident);
  CUtensorMap tmOut, tmAct;
  {
    const uint64_t dims[3] = {(uint64_t)Cout, (uint64_t)HW, (uint64_t)N};
    const uint32_t box[3] = {64, 128, 1};
    const uint64_t str[2] = {(uint64_t)ld_out * 2, (uint64_t)HW * ld_out * 2};
    int r = encode_bf16_map(&tmOut, out, 3, dims, str, box);
    if (r) return r;
    tmAct = tmOut;
    if (out_act) {
      const uint64_t stra[2] = {(uint64_t)ld_act * 2, (uint64_t)HW * ld_act * 2};
      r = encode_bf16_map(&tmAct, out_act, 3, dims, stra, box);
      if (r) return r;
    }
  }
  const int ki = BN <= 32 ? 0 : (BN <= 64 ? 1 : (BN <= 128 ? 2 : 3));
  ((AST_KERNEL_DEBUG && dbg_flags) ? kerns_dbg : kerns)[ki][p.f16 ? 1 : 0]<<<grid, PW_THREADS, smem, s>>>(tmA, tmB, tmOut, tmAct, p);
  AST_CHECK_LAUNCH();
  return 0;
}
