// K2 (CUDA-core side): the layers that are not GEMM-shaped, the weight packer, a direct 3x3 kernel
// for odd shapes / on-device cross-checks, and the ast_conv3x3_fwd dispatcher.
//
// Reference work replaced (paths relative to /root/reference):
//   Normalization + conv_1 + relu_1   models.py:129-131, 198-224   (3 -> 64, zero pad)
//   last decoder conv (64 -> 3)       models.py:626-627            (reflect pad, no ReLU)
#include "common.cuh"

namespace ast {
namespace tc {
bool tc_supported(const ast_conv_desc* d);
int conv3x3_tc(const ast_conv_desc* d, const void* in, const void* wpk, const float* bias, void* out,
               float* tap, cudaStream_t s);
int conv3x3_first_tc(const float* img, const float* w, const float* bias, const float* mean,
                     const float* std_, void* out, float* tap, int tap_prerelu, int N, int H, int W,
                     cudaStream_t s);
int conv3x3_last_tc(const void* in, const void* wpk16, const float* bias, float* out, int N, int H,
                    int W, int Cin, int Cout, int clamp01, int kwbox, cudaStream_t s);
int conv12_fused(const float* img, const float* w1, const float* b1, const float* mean, const float* std_,
                 const void* wpk2, const float* b2, void* out, int N, int H, int W, cudaStream_t s);
}  // namespace tc

// ---- weight packing: OIHW fp32 -> bf16 [9][Cout][Cin] -------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o,
                                   int Cout, int Cin, int flip, int rows) {
  // flip: out[t][ci][co] = w[co][ci][8 - t]  (a conv from Cout channels to Cin: the data gradient)
  // rows >= Cout (no flip): zero-padded output-channel rows
  const int64_t total = (int64_t)9 * rows * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i / ((int64_t)rows * Cin));
    int64_t r = i - (int64_t)t * rows * Cin;
    float v;
    if (!flip) {
      int co = (int)(r / Cin), ci = (int)(r % Cin);
      v = co < Cout ? w[((int64_t)co * Cin + ci) * 9 + t] : 0.f;
    } else {
      int ci = (int)(r / Cout), co = (int)(r % Cout);
      v = w[((int64_t)co * Cin + ci) * 9 + (8 - t)];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}

__global__ void pack_weight_ex_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o,
                                      int Cout, int Cin, int flip, int R, int Cc,
                                      const float* __restrict__ row_scale) {
  const int64_t total = (int64_t)9 * R * Cc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cc);
    const int r = (int)((i / Cc) % R);
    const int t = (int)(i / ((int64_t)Cc * R));
    float v = 0.f;
    if (!flip) {
      if (r < Cout && c < Cin) v = w[((int64_t)r * Cin + c) * 9 + t];
    } else {
      if (r < Cin && c < Cout) v = w[((int64_t)c * Cin + r) * 9 + (8 - t)];
    }
    if (row_scale) v *= row_scale[r < (flip ? Cin : Cout) ? r : 0];
    o[i] = __float2bfloat16_rn(v);
  }
}

// ---- generic direct kernel: one thread per element of the PADDED output grid --------------------
__device__ __forceinline__ int reflect_idx(int p, int X) {  // ReflectionPad2d(1) source index
  return p < 0 ? -p : (p >= X ? 2 * X - 2 - p : p);
}

__device__ __forceinline__ float conv_at(const __nv_bfloat16* __restrict__ in,
                                         const __nv_bfloat16* __restrict__ wpk, int n, int h, int w,
                                         int co, int H, int W, int Cin, int Cout) {
  float acc = 0.f;
  for (int t = 0; t < 9; ++t) {
    const int kh = t / 3, kw = t % 3;
    const __nv_bfloat16* ip = in + (((int64_t)n * (H + 2) + h + kh) * (W + 2) + w + kw) * Cin;
    const __nv_bfloat16* wp = wpk + ((int64_t)t * Cout + co) * Cin;
    for (int ci = 0; ci < Cin; ++ci)
      acc = fmaf(__bfloat162float(ip[ci]), __bfloat162float(wp[ci]), acc);
  }
  return acc;
}

__global__ void conv3x3_direct_kernel(const __nv_bfloat16* __restrict__ in,
                                      const __nv_bfloat16* __restrict__ wpk,
                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                      float* __restrict__ tap, ast_conv_desc d, int Ho, int Wo) {
  const int64_t total = (int64_t)d.N * (Ho + 2) * (Wo + 2) * d.Cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % d.Cout);
    int64_t r = i / d.Cout;
    const int pw = (int)(r % (Wo + 2)) - 1; r /= (Wo + 2);
    const int ph = (int)(r % (Ho + 2)) - 1;
    const int n = (int)(r / (Ho + 2));
    const bool halo = ph < 0 || pw < 0 || ph >= Ho || pw >= Wo;
    if (halo && d.halo != AST_HALO_REFLECT) continue;
    const int oh = reflect_idx(ph, Ho), ow = reflect_idx(pw, Wo);
    const float b = bias ? bias[co] : 0.f;
    float v;
    if (d.epilogue == AST_EPI_POOL2) {
      v = -INFINITY;
      for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 2; ++c) {
          float x = conv_at(in, wpk, n, 2 * oh + a, 2 * ow + c, co, d.H, d.W, d.Cin, d.Cout) + b;
          if (tap && !halo) {
            float tv = d.tap_prerelu ? x : (d.relu ? fmaxf(x, 0.f) : x);
            tap[(((int64_t)n * d.Cout + co) * d.H + 2 * oh + a) * d.W + 2 * ow + c] = tv;
          }
          if (d.relu) x = fmaxf(x, 0.f);
          // pool the bf16-rounded values (rounding is monotonic: same result as rounding the max)
          v = fmaxf(v, __bfloat162float(__float2bfloat16_rn(x)));
        }
    } else {
      const int h = d.epilogue == AST_EPI_UP2 ? oh >> 1 : oh;
      const int w = d.epilogue == AST_EPI_UP2 ? ow >> 1 : ow;
      float x = conv_at(in, wpk, n, h, w, co, d.H, d.W, d.Cin, d.Cout) + b;
      const bool tap_owner = !halo && (d.epilogue != AST_EPI_UP2 || (((oh | ow) & 1) == 0));
      if (tap && tap_owner && d.tap_prerelu)
        tap[(((int64_t)n * d.Cout + co) * d.H + h) * d.W + w] = x;
      if (d.relu) x = fmaxf(x, 0.f);
      if (tap && tap_owner && !d.tap_prerelu)
        tap[(((int64_t)n * d.Cout + co) * d.H + h) * d.W + w] = x;
      v = x;
    }
    if (out) out[i] = __float2bfloat16_rn(v);
  }
  // odd H/W with POOL2: the tap rows/cols dropped by the floor are written by nobody (callers
  // that need them use AST_EPI_PLAIN).
}

// ---- first layer: NCHW fp32 image -> normalise -> conv 3->Cout -> ReLU -> native bf16 ------------
// One thread per pixel, all Cout (<= 64 per pass) accumulators in registers, weights in shared
// memory as [27][Cout] (warp-uniform float4 broadcasts), output staged through shared memory so
// that the warp writes whole 128-byte pixel rows.
constexpr int kFirstThreads = 128;
constexpr int kFirstCout = 64;

__global__ void __launch_bounds__(kFirstThreads)
conv3x3_first_kernel(const float* __restrict__ img, const float* __restrict__ w,
                     const float* __restrict__ bias, float3 mean, float3 rstd, int normalise,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ tap, int tap_prerelu,
                     int N, int H, int W, int Cout, int co0) {
  __shared__ __align__(16) float s_w[27][kFirstCout];
  __shared__ __align__(16) float s_b[kFirstCout];
  __shared__ __align__(16) uint32_t s_o[kFirstThreads][kFirstCout / 2 + 4];  // +16 B pad per pixel
  for (int i = threadIdx.x; i < 27 * kFirstCout; i += kFirstThreads) {
    const int k = i / kFirstCout, c = i % kFirstCout;  // k = ci*9 + tap
    s_w[k][c] = (co0 + c < Cout) ? w[(int64_t)(co0 + c) * 27 + k] : 0.f;
  }
  for (int i = threadIdx.x; i < kFirstCout; i += kFirstThreads)
    s_b[i] = (bias && co0 + i < Cout) ? bias[co0 + i] : 0.f;
  __syncthreads();

  const int64_t npix = (int64_t)N * H * W;
  const int64_t pix0 = (int64_t)blockIdx.x * kFirstThreads;
  const int64_t pix = pix0 + threadIdx.x;
  const bool valid = pix < npix;
  int n = 0, h = 0, x = 0;
  float acc[kFirstCout];
#pragma unroll
  for (int c = 0; c < kFirstCout; ++c) acc[c] = s_b[c];
  if (valid) {
    x = (int)(pix % W);
    h = (int)((pix / W) % H);
    n = (int)(pix / ((int64_t)W * H));
    const float m[3] = {mean.x, mean.y, mean.z};
    const float rs[3] = {rstd.x, rstd.y, rstd.z};
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* ip = img + ((int64_t)n * 3 + ci) * H * W;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int ih = h + kh - 1, iw = x + kw - 1;
          float v = 0.f;  // zero padding applies to the NORMALISED image (models.py:131 then pad)
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            v = __ldg(ip + (int64_t)ih * W + iw);
            if (normalise) v = (v - m[ci]) * rs[ci];
          }
          const float4* wr = reinterpret_cast<const float4*>(&s_w[ci * 9 + kh * 3 + kw][0]);
#pragma unroll
          for (int c4 = 0; c4 < kFirstCout / 4; ++c4) {
            const float4 ww = wr[c4];
            acc[4 * c4 + 0] = fmaf(v, ww.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(v, ww.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(v, ww.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(v, ww.w, acc[4 * c4 + 3]);
          }
        }
      }
    }
  }
  if (tap && valid && tap_prerelu) {
#pragma unroll
    for (int c = 0; c < kFirstCout; ++c)
      if (co0 + c < Cout) tap[(((int64_t)n * Cout + co0 + c) * H + h) * W + x] = acc[c];
  }
#pragma unroll
  for (int c = 0; c < kFirstCout; ++c) acc[c] = fmaxf(acc[c], 0.f);
  if (tap && valid && !tap_prerelu) {
#pragma unroll
    for (int c = 0; c < kFirstCout; ++c)
      if (co0 + c < Cout) tap[(((int64_t)n * Cout + co0 + c) * H + h) * W + x] = acc[c];
  }
  if (!out) return;
#pragma unroll
  for (int c = 0; c < kFirstCout / 2; ++c) s_o[threadIdx.x][c] = pack_bf16(acc[2 * c], acc[2 * c + 1]);
  __syncthreads();
  // cooperative store: 8 consecutive threads write the 128 B (64 ch) of one pixel
  const int nch = min(kFirstCout, Cout - co0);  // multiple of 8 enforced by the host
  const int vec_per_pix = nch / 8;
  for (int i = threadIdx.x; i < kFirstThreads * vec_per_pix; i += kFirstThreads) {
    const int pl = i / vec_per_pix, vj = i % vec_per_pix;
    const int64_t pp = pix0 + pl;
    if (pp >= npix) break;
    const int px = (int)(pp % W), phh = (int)((pp / W) % H), pn = (int)(pp / ((int64_t)W * H));
    const uint4 val = *reinterpret_cast<const uint4*>(&s_o[pl][vj * 4]);
    __nv_bfloat16* o = out + (((int64_t)pn * (H + 2) + phh + 1) * (W + 2) + px + 1) * Cout + co0 + vj * 8;
    *reinterpret_cast<uint4*>(o) = val;
  }
}

// ---- last layer: native bf16 (halo = padding) -> conv Cin->Cout(<=4) -> NCHW fp32 ----------------
constexpr int kLastThreads = 128;
constexpr int kLastMaxCout = 4;

template <int CIN>
__global__ void __launch_bounds__(kLastThreads)
conv3x3_last_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W,
                    int Cout, int clamp01) {
  // weights as [tap][ci][co4] fp32 in shared memory
  __shared__ __align__(16) float s_w[9][CIN][kLastMaxCout];
  for (int i = threadIdx.x; i < 9 * CIN * kLastMaxCout; i += kLastThreads) {
    const int co = i % kLastMaxCout, ci = (i / kLastMaxCout) % CIN, t = i / (kLastMaxCout * CIN);
    s_w[t][ci][co] = co < Cout ? w[((int64_t)co * CIN + ci) * 9 + t] : 0.f;
  }
  __syncthreads();
  const int64_t npix = (int64_t)N * H * W;
  const int64_t pix = (int64_t)blockIdx.x * kLastThreads + threadIdx.x;
  if (pix >= npix) return;
  const int x = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  float acc[kLastMaxCout];
#pragma unroll
  for (int c = 0; c < kLastMaxCout; ++c) acc[c] = (bias && c < Cout) ? bias[c] : 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int kh = t / 3, kw = t % 3;
    const uint4* ip = reinterpret_cast<const uint4*>(
        in + (((int64_t)n * (H + 2) + h + kh) * (W + 2) + x + kw) * CIN);
#pragma unroll
    for (int v = 0; v < CIN / 8; ++v) {
      const uint4 u = __ldg(ip + v);
      float f[8];
      Vec16<true>::unpack(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 ww = *reinterpret_cast<const float4*>(&s_w[t][v * 8 + j][0]);
        acc[0] = fmaf(f[j], ww.x, acc[0]);
        acc[1] = fmaf(f[j], ww.y, acc[1]);
        acc[2] = fmaf(f[j], ww.z, acc[2]);
        acc[3] = fmaf(f[j], ww.w, acc[3]);
      }
    }
  }
  for (int c = 0; c < Cout; ++c) {
    float v = acc[c];
    if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    out[(((int64_t)n * Cout + c) * H + h) * W + x] = v;
  }
}

}  // namespace ast

using namespace ast;

extern "C" int ast_pack_conv_weight(const float* w_oihw, void* wpk, int Cout, int Cin, int flip,
                                    int cout_pad, void* stream) {
  if (!w_oihw || !wpk || Cout <= 0 || Cin <= 0) return AST_E_BADARG;
  if (cout_pad != 0 && (flip || cout_pad < Cout)) return AST_E_BADARG;
  const int rows = cout_pad ? cout_pad : Cout;
  const int64_t total = (int64_t)9 * rows * Cin;
  int64_t nb = (total + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  pack_weight_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, reinterpret_cast<__nv_bfloat16*>(wpk), Cout, Cin, flip, rows);
  AST_CHECK_LAUNCH();
  return 0;
}

// ---- fold packing: OIHW fp32 -> bf16 [16][Cout][Cin] (see ast_b200.h) ---------------------------------
__global__ void pack_weight_fold_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int Cout, int Cin) {
  const int64_t total = (int64_t)16 * Cout * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int t = (int)(i / ((int64_t)Cin * Cout));
    const int b = t & 1, a = (t >> 1) & 1, px = (t >> 2) & 1, py = t >> 3;
    // taps of the 3x3 kernel that read low-res row offset a (column offset b) for output parity py (px)
    const int kh0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), kh1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
    const int kw0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), kw1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
    const float* wp = w + ((int64_t)co * Cin + ci) * 9;
    float acc = 0.f;
    for (int kh = kh0; kh <= kh1; ++kh)
      for (int kw = kw0; kw <= kw1; ++kw) acc += wp[kh * 3 + kw];
    o[i] = __float2bfloat16_rn(acc);
  }
}

extern "C" int ast_pack_conv_weight_fold(const float* w_oihw, void* wpk, int Cout, int Cin, void* stream) {
  if (!w_oihw || !wpk || Cout <= 0 || Cin <= 0) return AST_E_BADARG;
  const int64_t total = (int64_t)16 * Cout * Cin;
  int64_t nb = (total + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  pack_weight_fold_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, reinterpret_cast<__nv_bfloat16*>(wpk), Cout, Cin);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_pack_conv_weight_ex(const float* w_oihw, void* wpk, int Cout, int Cin, int flip,
                                       int rows_pad, int cols_pad, const float* row_scale,
                                       void* stream) {
  if (!w_oihw || !wpk || Cout <= 0 || Cin <= 0) return AST_E_BADARG;
  const int rows = flip ? Cin : Cout, cols = flip ? Cout : Cin;
  const int R = rows_pad ? rows_pad : rows, Cc = cols_pad ? cols_pad : cols;
  if (R < rows || Cc < cols) return AST_E_BADARG;
  const int64_t total = (int64_t)9 * R * Cc;
  int64_t nb = (total + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  pack_weight_ex_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, reinterpret_cast<__nv_bfloat16*>(wpk), Cout, Cin, flip, R, Cc, row_scale);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_conv3x3_fwd(const ast_conv_desc* d, const void* in, const void* wpk,
                               const float* bias, void* out, float* tap, void* stream) {
  if (!d || !in || !wpk || (!out && !tap)) return AST_E_BADARG;
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) return AST_E_BADARG;
  const bool fold = d->epilogue == AST_EPI_UPFOLD;
  if ((d->epilogue < AST_EPI_PLAIN || d->epilogue > AST_EPI_UP2) && !fold) return AST_E_BADARG;
  if (d->halo < AST_HALO_KEEP || d->halo > AST_HALO_CLAMP) return AST_E_BADARG;
  if (d->halo == AST_HALO_REFLECT) {
    const bool up = d->epilogue == AST_EPI_UP2 || fold;
    const int Ho = d->epilogue == AST_EPI_POOL2 ? d->H / 2 : (up ? 2 * d->H : d->H);
    const int Wo = d->epilogue == AST_EPI_POOL2 ? d->W / 2 : (up ? 2 * d->W : d->W);
    if (Ho < 2 || Wo < 2) return AST_E_SHAPE;  // ReflectionPad2d(1) needs at least 2 pixels
  }
  cudaStream_t s = (cudaStream_t)stream;
  const bool want_tc = d->impl == AST_CONV_TC || d->impl == AST_CONV_TC_TAPBOX || d->impl >= 64 ||
                       (d->impl == AST_CONV_AUTO && tc::tc_supported(d));
  if (want_tc) return tc::conv3x3_tc(d, in, wpk, bias, out, tap, s);
  if (fold || d->halo == AST_HALO_CLAMP) return AST_E_SHAPE;   // folded upsample conv: tcgen05 path only
  if (d->impl != AST_CONV_DIRECT && d->impl != AST_CONV_AUTO) return AST_E_BADARG;
  const int Ho = d->epilogue == AST_EPI_POOL2 ? d->H / 2 : (d->epilogue == AST_EPI_UP2 ? 2 * d->H : d->H);
  const int Wo = d->epilogue == AST_EPI_POOL2 ? d->W / 2 : (d->epilogue == AST_EPI_UP2 ? 2 * d->W : d->W);
  const int64_t total = (int64_t)d->N * (Ho + 2) * (Wo + 2) * d->Cout;
  int64_t nb = (total + 255) / 256;
  if (nb > 148 * 32) nb = 148 * 32;
  conv3x3_direct_kernel<<<(unsigned)nb, 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<const __nv_bfloat16*>(wpk), bias,
      reinterpret_cast<__nv_bfloat16*>(out), tap, *d, Ho, Wo);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_conv12_fused(const float* img, const float* w1, const float* b1, const float* mean,
                                const float* std_, const void* wpk2, const float* b2, void* out, int N, int H, int W,
                                void* stream) {
  return tc::conv12_fused(img, w1, b1, mean, std_, wpk2, b2, out, N, H, W, (cudaStream_t)stream);
}

extern "C" int ast_conv3x3_first(const float* img, const float* w, const float* bias,
                                 const float* mean, const float* std_, void* out, float* tap,
                                 int tap_prerelu, int N, int H, int W, int Cout, int impl,
                                 void* stream) {
  if (!img || !w || (!out && !tap) || N <= 0 || H <= 0 || W <= 0 || Cout <= 0) return AST_E_BADARG;
  if (Cout % 8 != 0) return AST_E_SHAPE;
  if (out && !aligned16(out)) return AST_E_ALIGN;
  if (impl == AST_CONV_TC && Cout != 64) return AST_E_SHAPE;
  if (impl != AST_CONV_DIRECT && Cout == 64 && (!bias || aligned16(bias)))
    return tc::conv3x3_first_tc(img, w, bias, mean, std_, out, tap, tap_prerelu, N, H, W,
                                (cudaStream_t)stream);
  float3 m = make_float3(0.f, 0.f, 0.f), rs = make_float3(1.f, 1.f, 1.f);
  const int normalise = (mean && std_) ? 1 : 0;
  if (normalise) {
    m = make_float3(mean[0], mean[1], mean[2]);
    rs = make_float3(1.f / std_[0], 1.f / std_[1], 1.f / std_[2]);
  }
  const int64_t npix = (int64_t)N * H * W;
  const int64_t nb = (npix + kFirstThreads - 1) / kFirstThreads;
  if (nb >= 0x7fffffffLL) return AST_E_SHAPE;
  for (int co0 = 0; co0 < Cout; co0 += kFirstCout) {
    conv3x3_first_kernel<<<(unsigned)nb, kFirstThreads, 0, (cudaStream_t)stream>>>(
        img, w, bias, m, rs, normalise, reinterpret_cast<__nv_bfloat16*>(out), tap, tap_prerelu, N,
        H, W, Cout, co0);
    AST_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int ast_conv3x3_last(const void* in, const float* w, const void* wpk16, const float* bias,
                                float* out, int N, int H, int W, int Cin, int Cout, int clamp01,
                                int impl, void* stream) {
  if (!in || (!w && !wpk16) || !out || N <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if ((impl == AST_CONV_TC || impl == AST_CONV_TC_TAPBOX) && (!wpk16 || Cin % 64 != 0)) return AST_E_SHAPE;
  if (impl != AST_CONV_DIRECT && wpk16 && Cin % 64 == 0 && Cout <= 16)
    return tc::conv3x3_last_tc(in, wpk16, bias, out, N, H, W, Cin, Cout, clamp01,
                               impl == AST_CONV_TC_TAPBOX ? 0 : 1, (cudaStream_t)stream);
  if (!w) return AST_E_BADARG;
  if (Cout < 1 || Cout > kLastMaxCout) return AST_E_SHAPE;
  if (!aligned16(in)) return AST_E_ALIGN;
  const int64_t npix = (int64_t)N * H * W;
  const int64_t nb = (npix + kLastThreads - 1) / kLastThreads;
  if (nb >= 0x7fffffffLL) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  const __nv_bfloat16* ip = reinterpret_cast<const __nv_bfloat16*>(in);
  if (Cin == 64)
    conv3x3_last_kernel<64><<<(unsigned)nb, kLastThreads, 0, s>>>(ip, w, bias, out, N, H, W, Cout, clamp01);
  else if (Cin == 16)
    conv3x3_last_kernel<16><<<(unsigned)nb, kLastThreads, 0, s>>>(ip, w, bias, out, N, H, W, Cout, clamp01);
  else if (Cin == 32)
    conv3x3_last_kernel<32><<<(unsigned)nb, kLastThreads, 0, s>>>(ip, w, bias, out, N, H, W, Cout, clamp01);
  else
    return AST_E_SHAPE;
  AST_CHECK_LAUNCH();
  return 0;
}
