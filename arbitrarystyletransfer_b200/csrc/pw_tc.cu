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
// These layers are HBM-bound (K <= 768): one 128-pixel tile per CTA, shared memory and TMEM sized to the
// layer's N block so that 2-6 CTAs share an SM, and eight epilogue warps per CTA (two per TMEM lane quarter,
// splitting the columns) to keep enough stores in flight.
#include "tc.cuh"

namespace ast {
namespace tc {

constexpr int PW_THREADS = 320;            // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue
constexpr int PW_A_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int PW_STAGES = 2;
__host__ __device__ constexpr int pw_stage_bytes(int BN) { return PW_A_BYTES + ((BN * 128 + 1023) / 1024) * 1024; }
__host__ __device__ constexpr int pw_smem_bytes(int BN) {
  return PW_STAGES * pw_stage_bytes(BN) + (2 * PW_STAGES + 1) * 8 + 16 + 1024;
}

struct PwParams {
  int N, Cin, Cout, BN, n_blocks, tiles_per_img, per_sample_w, act;
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
  return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);   // nn.Hardswish
}

template <int TMEM_COLS>
__global__ void __launch_bounds__(PW_THREADS)
pw_conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const PwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int STAGE = pw_stage_bytes(p.BN);
  const uint32_t bars = base + PW_STAGES * STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (PW_STAGES + s); };
  const uint32_t done_bar = bars + 8u * (2 * PW_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * PW_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + PW_STAGES * STAGE + 8 * (2 * PW_STAGES + 1));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  int t = blockIdx.x;
  const int nb = t % p.n_blocks; t /= p.n_blocks;
  const int ti = t % p.tiles_per_img;
  const int n = t / p.tiles_per_img;
  const int ksteps = (p.Cin + 63) / 64;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < PW_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(done_bar, 1);
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
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), PW_A_BYTES + b_bytes);
        const uint32_t a_dst = base + stage * STAGE;
        tma_load_3d(a_dst, &tmA, full_bar(stage), ks * 64, ti * 128, n);
        tma_load_3d(a_dst + PW_A_BYTES, &tmB, full_bar(stage), ks * 64, nb * p.BN, p.per_sample_w ? n : 0);
        if (++stage == PW_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = p.f16 ? make_idesc_f16(128, p.BN) : make_idesc_bf16(128, p.BN);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint64_t ad = make_sdesc_k128(base + stage * STAGE);
      const uint64_t bd = make_sdesc_k128(base + stage * STAGE + PW_A_BYTES);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (ks * 64 + k * 16 < p.Cin)   // K tail: skip 16-channel steps that are pure TMA zero fill
            umma_bf16(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k ? 1u : accum);
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      accum = 1u;
      if (++stage == PW_STAGES) { stage = 0; phase ^= 1u; }
    }
    if (elect_one_sync()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // epilogue: warps 2..9 -> TMEM lane quarter (warp % 4); the two warps of a quarter alternate 16-column chunks
    const int e = warp & 3;
    const int half = (warp - 2) >> 2;
    mbar_wait(done_bar, 0u);
    tc_fence_after();
    const int64_t pix = (int64_t)ti * 128 + e * 32 + lane;   // pixel inside the image
    const bool ok = pix < p.HW;
    const int64_t row = (int64_t)n * p.HW + pix;
    int64_t rrow = row;
    if (p.res_w > 0 && ok) {   // residual = the block input BEFORE the nearest x2 upsample (models.py:265-267)
      const int h = (int)(pix / p.res_w), w = (int)(pix % p.res_w);
      rrow = (int64_t)n * (p.HW >> 2) + (int64_t)(h >> 1) * (p.res_w >> 1) + (w >> 1);
    }
    const bool wide_ok = ((p.ld_out & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0) &&
                         (!p.out_act || ((p.ld_act & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out_act) & 31) == 0));
    for (int c0 = half * 16; c0 < p.BN; c0 += 32) {
      uint32_t v[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      const int co0 = nb * p.BN + c0;
      if (!ok || co0 >= p.Cout) continue;
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        f[i] = __uint_as_float(v[i]);
        if (p.bias && co0 + i < p.Cout) f[i] += __ldg(p.bias + co0 + i);
        if (p.act && !p.out_act) f[i] = hardswish(f[i]);
      }
      const int valid = min(16, p.Cout - co0);   // multiple of 8 (host enforces Cout % 8 == 0)
      if (p.residual) {
        const uint16_t* rp = p.residual + rrow * p.ld_res + co0;
        for (int i = 0; i < valid; i += 8) {
          float r[8];
          unpack8_dt(__ldg(reinterpret_cast<const uint4*>(rp + i)), r, p.f16);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[i + j] += r[j];
        }
      }
      uint16_t* op = p.out + row * p.ld_out + co0;
      if (wide_ok && valid == 16) {   // one 32-byte sector per lane per store
        uint32_t pk[8], pa[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          pk[j] = pk2_dt(f[2 * j], f[2 * j + 1], p.f16);
          if (p.out_act) {
            const float2 rr = un2_dt(pk[j], p.f16);
            pa[j] = pk2_dt(hardswish(rr.x), hardswish(rr.y), p.f16);
          }
        }
        st_global_v8(op, pk);
        if (p.out_act) st_global_v8(p.out_act + row * p.ld_act + co0, pa);
        continue;
      }
      for (int i = 0; i < valid; i += 8) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = f[i + j];
        const uint4 ov = pack8_dt(o, p.f16);
        *reinterpret_cast<uint4*>(op + i) = ov;
        if (p.out_act) {   // training: raw pre-activation in `out`, Hardswish of the ROUNDED value here
          float r[8];
          unpack8_dt(ov, r, p.f16);
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = hardswish(r[j]);
          *reinterpret_cast<uint4*>(p.out_act + row * p.ld_act + co0 + i) = pack8_dt(r, p.f16);
        }
      }
    }
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
  int n_blocks = 1, BN = (Cout + 15) / 16 * 16;
  while (BN > 256) {
    ++n_blocks;
    BN = ((Cout + n_blocks - 1) / n_blocks + 15) / 16 * 16;
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
  static bool attr_done = false;
  if (!attr_done) {
    const int mx = pw_smem_bytes(256);
    AST_CUDA(cudaFuncSetAttribute(pw_conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    AST_CUDA(cudaFuncSetAttribute(pw_conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    AST_CUDA(cudaFuncSetAttribute(pw_conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    AST_CUDA(cudaFuncSetAttribute(pw_conv_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    attr_done = true;
  }
  const int64_t grid = (int64_t)N * p.tiles_per_img * n_blocks;
  if (grid >= 0x7fffffffLL) return AST_E_SHAPE;
  const int smem = pw_smem_bytes(BN);
  if (BN <= 32) pw_conv_tc_kernel<32><<<(unsigned)grid, PW_THREADS, smem, s>>>(tmA, tmB, p);
  else if (BN <= 64) pw_conv_tc_kernel<64><<<(unsigned)grid, PW_THREADS, smem, s>>>(tmA, tmB, p);
  else if (BN <= 128) pw_conv_tc_kernel<128><<<(unsigned)grid, PW_THREADS, smem, s>>>(tmA, tmB, p);
  else pw_conv_tc_kernel<256><<<(unsigned)grid, PW_THREADS, smem, s>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}
