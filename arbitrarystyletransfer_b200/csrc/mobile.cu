// K4: the non-GEMM pieces of the MobileNet-style Encoder / Decoder / AutoEncoder blocks
// (reference: mobilenetv2.py:38-43 conv_3x3_bn, :63-81 SELayer, :95-165 DepthWiseConv;
// models.py:140-184 Encoder, :242-320 DecoderBlock / Decoder, :322-338 AutoEncoder), eval mode.
// Layout: plain NHWC fp16 [N][H][W][C] (act_t, common.cuh) (no halo: these blocks use reflect padding of 1 or 2 and
// stride 1 or 2, resolved by index arithmetic in the stencil).  All HBM-bound.
#include "common.cuh"
#include <stdlib.h>

namespace ast {

constexpr int kM = 256;

__device__ __forceinline__ float hswish(float x) {
  return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
}
__device__ __forceinline__ int reflect(int p, int X) {  // padding_mode="reflect" / ReflectionPad2d
  p = p < 0 ? -p : p;
  return p >= X ? 2 * X - 2 - p : p;
}

// ---- depthwise k x k (k = 3 or 5), stride 1 or 2, reflect padding (k-1)/2, + bias (folded BN) +
// Hardswish, with the SELayer's global average pool fused: per-(n,c) sums of the OUTPUT.
// up2 != 0: the input is read through a virtual nearest x2 upsample (DecoderBlock._upsample_3,
// models.py:254, 265-267), reflect padding applied on the upsampled grid.
// grid = (pixel chunks, N); thread -> 8 channels of one pixel group (as native_stats_kernel).
template <typename AT>
__global__ void __launch_bounds__(kM)
dw_conv_kernel(const AT* __restrict__ x, const float* __restrict__ w /*[k*k][C]*/,
               const float* __restrict__ bias, AT* __restrict__ out, float* __restrict__ pool,
               int C, int H, int W, int Ho, int Wo, int k, int stride, int up2, int act, int chunks) {
  extern __shared__ float s_pool[];  // [groups][C]
  const int cv = C / 8;
  const int groups = kM / cv;
  const int g = threadIdx.x / cv, v = threadIdx.x % cv;
  const int n = blockIdx.y;
  const int pad = (k - 1) / 2;
  const int Hin = up2 ? 2 * H : H, Win = up2 ? 2 * W : W;   // size of the (virtual) conv input
  const int64_t npix = (int64_t)Ho * Wo;
  const int64_t per = (npix + chunks - 1) / chunks;
  const int64_t p0 = blockIdx.x * per, p1 = (p0 + per < npix) ? p0 + per : npix;
  float psum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) psum[j] = 0.f;
  if (g < groups) {
    float b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = bias ? __ldg(bias + v * 8 + j) : 0.f;
    const AT* xin = x + (int64_t)n * H * W * C + v * 8;
    for (int64_t p = p0 + g; p < p1; p += groups) {
      const int ho = (int)(p / Wo), wo = (int)(p % Wo);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = b[j];
      for (int kh = 0; kh < k; ++kh) {
        int ih = reflect(ho * stride + kh - pad, Hin);
        if (up2) ih >>= 1;
        for (int kw = 0; kw < k; ++kw) {
          int iw = reflect(wo * stride + kw - pad, Win);
          if (up2) iw >>= 1;
          float xv[8];
          ld8(xin + ((int64_t)ih * W + iw) * C, xv);
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(kh * k + kw) * C + v * 8));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(kh * k + kw) * C + v * 8 + 4));
          acc[0] = fmaf(xv[0], w0.x, acc[0]); acc[1] = fmaf(xv[1], w0.y, acc[1]);
          acc[2] = fmaf(xv[2], w0.z, acc[2]); acc[3] = fmaf(xv[3], w0.w, acc[3]);
          acc[4] = fmaf(xv[4], w1.x, acc[4]); acc[5] = fmaf(xv[5], w1.y, acc[5]);
          acc[6] = fmaf(xv[6], w1.z, acc[6]); acc[7] = fmaf(xv[7], w1.w, acc[7]);
        }
      }
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = hswish(acc[j]);
      }
      const uint4 o = H16<AT>::pack(acc);
      *reinterpret_cast<uint4*>(out + (((int64_t)n * Ho + ho) * Wo + wo) * C + v * 8) = o;
      // pool what the next layer will actually read (the rounded value); act == 2 (training):
      // the RAW value is stored for the backward pass and the pool sees Hardswish of it
      float r[8];
      H16<AT>::unpack(o, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) psum[j] += (act == 2) ? hswish(r[j]) : r[j];
    }
  }
  if (!pool) return;
  if (g < groups) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_pool[g * C + v * 8 + j] = psum[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kM) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += s_pool[gg * C + c];
    atomicAdd(pool + (int64_t)n * C + c, s);
  }
}

// ---- SELayer excitation (mobilenetv2.py:66-80): y = Hardtanh(0,1)(W2 relu(W1 mean + b1) + b2) ------
// one block per image; pool holds the per-channel SUMS, inv_hw turns them into means.
constexpr int kSe = 1024;
__global__ void __launch_bounds__(kSe)
se_fc_kernel(const float* __restrict__ pool, float inv_hw, const float* __restrict__ w1,
             const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
             float* __restrict__ scale, float* __restrict__ hid_out, float* __restrict__ pre_out, int C,
             int S) {
  // Both mat-vecs walk their weight rows with one WARP per output (lanes along the contiguous input index, shuffle
  // reduce): coalesced, and 32 warps deep instead of one serial dot product per thread -- at one CTA per image the
  // kernel is pure latency (measured 29 us per launch before, independent of the batch).
  extern __shared__ float sm[];  // mean[C] + hid[S]
  float* mean = sm;
  float* hid = sm + C;
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += kSe) mean[c] = pool[(int64_t)n * C + c] * inv_hw;
  __syncthreads();
  for (int j = warp; j < S; j += kSe / 32) {
    const float* wr = w1 + (int64_t)j * C;
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(wr[c], mean[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      a = fmaxf(a + b1[j], 0.f);
      hid[j] = a;
      if (hid_out) hid_out[(int64_t)n * S + j] = a;
    }
  }
  __syncthreads();
  for (int c = warp; c < C; c += kSe / 32) {
    const float* wr = w2 + (int64_t)c * S;
    float a = 0.f;
    for (int j = lane; j < S; j += 32) a = fmaf(wr[j], hid[j], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      a += b2[c];
      scale[(int64_t)n * C + c] = fminf(fmaxf(a, 0.f), 1.f);
      if (pre_out) pre_out[(int64_t)n * C + c] = a;
    }
  }
}

// ---- per-sample weights of the pw-linear conv: W'[n][co][ci] = W[co][ci] * se[n][ci] (bf16) --------
// (x * y in SELayer.forward, mobilenetv2.py:81, moved from the activation into the weights: exact in
// real arithmetic, and it saves one full pass over the widest tensor of the block.)
template <typename AT>
__global__ void scale_weights_kernel(const float* __restrict__ w, const float* __restrict__ se,
                                     AT* __restrict__ out, int N, int Cout, int Cin) {
  const int64_t total = (int64_t)N * Cout * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int64_t r = i / Cin;
    const int co = (int)(r % Cout);
    const int n = (int)(r / Cout);
    const float s = se ? se[(int64_t)n * Cin + ci] : 1.f;
    H16<AT>::store1(out, i, w[(int64_t)co * Cin + ci] * s);
  }
}

// ---- stem: NCHW fp32 image -> conv 3x3 (reflect pad 1, stride 1, no bias) -> Hardswish -> NHWC bf16
// (conv_3x3_bn, mobilenetv2.py:38-43; Cout <= 32).  One thread per pixel.
constexpr int kStemMaxCout = 32;
template <typename AT>
__global__ void __launch_bounds__(128)
stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w /*OIHW [Cout][3][3][3]*/,
                 AT* __restrict__ out, AT* __restrict__ out_raw, int N, int H, int W,
                 int Cout) {
  __shared__ float s_w[27][kStemMaxCout];
  for (int i = threadIdx.x; i < 27 * kStemMaxCout; i += 128) {
    const int kk = i / kStemMaxCout, c = i % kStemMaxCout;
    s_w[kk][c] = c < Cout ? w[c * 27 + kk] : 0.f;
  }
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (pix >= (int64_t)N * H * W) return;
  const int x = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  float acc[kStemMaxCout];
#pragma unroll
  for (int c = 0; c < kStemMaxCout; ++c) acc[c] = 0.f;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float v = __ldg(img + (((int64_t)n * 3 + ci) * H + reflect(h + kh - 1, H)) * W + reflect(x + kw - 1, W));
#pragma unroll
        for (int c = 0; c < kStemMaxCout; ++c) acc[c] = fmaf(v, s_w[ci * 9 + kh * 3 + kw][c], acc[c]);
      }
  AT* o = out + pix * Cout;
#pragma unroll
  for (int c = 0; c < kStemMaxCout; c += 8) {
    if (c >= Cout) break;
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = acc[c + j];
    if (out_raw) {   // training: keep the rounded pre-activation, activate the rounded value
      const uint4 rv = H16<AT>::pack(t);
      *reinterpret_cast<uint4*>(out_raw + pix * Cout + c) = rv;
      H16<AT>::unpack(rv, t);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = hswish(t[j]);
    st8(o + c, t);
  }
}

// ---- image head: NHWC bf16 -> ReflectionPad2d(1) -> conv 3x3 (bias) -> NCHW fp32 (+ Hardtanh(0,1))
// (Decoder._ref_out + _img_out + last_act, models.py:300-316; Cin <= 32, Cout <= 4)
template <typename AT>
__global__ void __launch_bounds__(128)
head_conv_kernel(const AT* __restrict__ x, const float* __restrict__ w /*OIHW*/,
                 const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W, int Cin,
                 int Cout, int clamp01) {
  __shared__ float s_w[9][32][4];
  for (int i = threadIdx.x; i < 9 * 32 * 4; i += 128) {
    const int co = i % 4, ci = (i / 4) % 32, t = i / 128;
    s_w[t][ci][co] = (co < Cout && ci < Cin) ? w[((int64_t)co * Cin + ci) * 9 + t] : 0.f;
  }
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (pix >= (int64_t)N * H * W) return;
  const int xw = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  float acc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = (bias && c < Cout) ? bias[c] : 0.f;
  for (int t = 0; t < 9; ++t) {
    const int ih = reflect(h + t / 3 - 1, H), iw = reflect(xw + t % 3 - 1, W);
    const AT* ip = x + (((int64_t)n * H + ih) * W + iw) * Cin;
    for (int v = 0; v < Cin / 8; ++v) {
      float f[8];
      ld8(ip + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 ww = *reinterpret_cast<const float4*>(&s_w[t][v * 8 + j][0]);
        acc[0] = fmaf(f[j], ww.x, acc[0]); acc[1] = fmaf(f[j], ww.y, acc[1]);
        acc[2] = fmaf(f[j], ww.z, acc[2]); acc[3] = fmaf(f[j], ww.w, acc[3]);
      }
    }
  }
  for (int c = 0; c < Cout; ++c) {
    float v = acc[c];
    if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    out[(((int64_t)n * Cout + c) * H + h) * W + xw] = v;
  }
}

// ---- NHWC bf16 (row stride ld) -> NCHW fp32 (feature taps at the reference boundary) ---------------
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const uint16_t* __restrict__ src, int ld,
                                                           float* __restrict__ dst, int C, int64_t HW, int f16) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int pl = ty; pl < 32; pl += 8) {
    const int64_t p = p0 + pl;
    const uint32_t raw = (p < HW && c0 + tx < C) ? src[((int64_t)n * HW + p) * ld + c0 + tx] : 0u;
    tile[pl][tx] = un2_dt(raw, f16).x;
  }
  __syncthreads();
  for (int cl = ty; cl < 32; cl += 8) {
    const int64_t p = p0 + tx;
    if (c0 + cl < C && p < HW) dst[((int64_t)n * C + c0 + cl) * HW + p] = tile[tx][cl];
  }
}

}  // namespace ast

using namespace ast;

// dw_tiled.cu: shared-memory-tiled stride-1 kernels (AST_E_SHAPE = no tiling fits, use the direct kernel)
int dw_tiled_forward(const void* x, const float* w, const float* bias, void* out, float* pool, int N, int C, int H,
                     int W, int k, int up2, int act, cudaStream_t s);
bool dw_force_direct() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AST_DW_DIRECT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

extern "C" int ast_dw_conv(const void* x, const float* w, const float* bias, void* out, float* pool, int N,
                           int C, int H, int W, int k, int stride, int up2, int act, void* stream) {
  if (!x || !w || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if ((k != 3 && k != 5) || (stride != 1 && stride != 2) || C % 8 != 0 || C / 8 > kM) return AST_E_SHAPE;
  const int Hin = up2 ? 2 * H : H, Win = up2 ? 2 * W : W, pad = (k - 1) / 2;
  if (Hin <= pad || Win <= pad) return AST_E_SHAPE;   // reflect padding needs pad < size
  const int Ho = (Hin + 2 * pad - k) / stride + 1, Wo = (Win + 2 * pad - k) / stride + 1;
  if (N > 65535) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (pool) AST_CUDA(cudaMemsetAsync(pool, 0, sizeof(float) * (size_t)N * C, s));
  if (stride == 1 && !dw_force_direct()) {
    const int r = dw_tiled_forward(x, w, bias, out, pool, N, C, H, W, k, up2, act, s);
    if (r != AST_E_SHAPE) return r;
  }
  int64_t chunks = (4 * 148 + N - 1) / N;
  const int64_t npix = (int64_t)Ho * Wo;
  if (chunks > npix / 8) chunks = npix / 8;
  if (chunks < 1) chunks = 1;
  const int groups = kM / (C / 8);
  const size_t smem = pool ? (size_t)groups * C * sizeof(float) : 0;
  if (smem > 48 * 1024) return AST_E_SHAPE;
  AST_ACT_DISPATCH(dw_conv_kernel<AT><<<dim3((unsigned)chunks, N), kM, smem, s>>>(
      reinterpret_cast<const AT*>(x), w, bias, reinterpret_cast<AT*>(out), pool, C, H, W,
      Ho, Wo, k, stride, up2, act, (int)chunks));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_se_fc(const float* pool, float inv_hw, const float* w1, const float* b1, const float* w2,
                         const float* b2, float* scale, float* hid_out, float* pre_out, int N, int C, int S,
                         void* stream) {
  if (!pool || !w1 || !b1 || !w2 || !b2 || !scale || N <= 0 || C <= 0 || S <= 0) return AST_E_BADARG;
  const size_t smem = (size_t)(C + S) * sizeof(float);
  if (smem > 48 * 1024) return AST_E_SHAPE;
  se_fc_kernel<<<N, kSe, smem, (cudaStream_t)stream>>>(pool, inv_hw, w1, b1, w2, b2, scale, hid_out,
                                                               pre_out, C, S);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_scale_weights(const float* w, const float* se, void* out, int N, int Cout, int Cin,
                                 void* stream) {
  if (!w || !out || N <= 0 || Cout <= 0 || Cin <= 0) return AST_E_BADARG;
  const int64_t total = (int64_t)N * Cout * Cin;
  int64_t nb = (total + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  AST_ACT_DISPATCH(scale_weights_kernel<AT><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(
      w, se, reinterpret_cast<AT*>(out), N, Cout, Cin));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_stem_conv(const float* img, const float* w, void* out, void* out_raw, int N, int H, int W,
                             int Cout, void* stream) {
  if (!img || !w || !out || N <= 0 || H < 2 || W < 2) return AST_E_BADARG;
  if (Cout % 8 != 0 || Cout > kStemMaxCout) return AST_E_SHAPE;
  const int64_t nb = ((int64_t)N * H * W + 127) / 128;
  if (nb >= 0x7fffffffLL) return AST_E_SHAPE;
  AST_ACT_DISPATCH(stem_conv_kernel<AT><<<(unsigned)nb, 128, 0, (cudaStream_t)stream>>>(
      img, w, reinterpret_cast<AT*>(out), reinterpret_cast<AT*>(out_raw), N, H, W, Cout));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_head_conv(const void* x, const float* w, const float* bias, float* out, int N, int H, int W,
                             int Cin, int Cout, int clamp01, void* stream) {
  if (!x || !w || !out || N <= 0 || H < 2 || W < 2) return AST_E_BADARG;
  if (Cin % 8 != 0 || Cin > 32 || Cout < 1 || Cout > 4) return AST_E_SHAPE;
  const int64_t nb = ((int64_t)N * H * W + 127) / 128;
  if (nb >= 0x7fffffffLL) return AST_E_SHAPE;
  AST_ACT_DISPATCH(head_conv_kernel<AT><<<(unsigned)nb, 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const AT*>(x), w, bias, out, N, H, W, Cin, Cout, clamp01));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_nhwc_to_nchw(const void* x, int ld, float* out, int N, int C, int64_t HW, int dtype, void* stream) {
  if (!x || !out || N <= 0 || C <= 0 || HW <= 0 || ld < C || (dtype != AST_DT_BF16 && dtype != AST_DT_F16)) return AST_E_BADARG;
  if (N > 65535 || (C + 31) / 32 > 65535 || (HW + 31) / 32 >= 0x7fffffffLL) return AST_E_SHAPE;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N);
  nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint16_t*>(x), ld, out,
                                                              C, HW, dtype == AST_DT_F16);
  AST_CHECK_LAUNCH();
  return 0;
}
