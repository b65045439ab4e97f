// Native-layout (bf16 [N][H+2][W+2][C], one-pixel halo) helpers:
//   * converters to / from the reference's NCHW fp32 tensors (the drop-in boundary),
//   * K1n: AdaIN (models.py:43-51 + alpha blend models.py:471) evaluated directly on the native
//     layout between the encoder's last conv and the decoder's first, same fp32 Welford / Chan
//     arithmetic as K1 (adain.cu).
#include "common.cuh"

namespace ast {

__device__ __forceinline__ int halo_targets(int x, int X, bool reflect, int (&t)[3]) {
  int n = 0;
  t[n++] = x;
  if (reflect) {
    if (x == 1) t[n++] = -1;
    if (x == X - 2) t[n++] = X;
  }
  return n;
}

// ---- NCHW fp32 -> native ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_native_kernel(const float* __restrict__ src,
                                                             __nv_bfloat16* __restrict__ dst, int N,
                                                             int C, int H, int W, int reflect) {
  __shared__ float tile[32][33];  // [c][w]
  const int ctiles = (C + 31) / 32;
  const int n = blockIdx.z / ctiles, c0 = (blockIdx.z % ctiles) * 32;
  const int h = blockIdx.y, w0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int c = ty; c < 32; c += 8) {
    const int cc = c0 + c, ww = w0 + tx;
    tile[c][tx] = (cc < C && ww < W) ? src[(((int64_t)n * C + cc) * H + h) * W + ww] : 0.f;
  }
  __syncthreads();
  int rows[3];
  const int nr = halo_targets(h, H, reflect != 0, rows);
  for (int wl = ty; wl < 32; wl += 8) {
    const int ww = w0 + wl, cc = c0 + tx;
    if (ww >= W || cc >= C) continue;
    const __nv_bfloat16 v = __float2bfloat16_rn(tile[tx][wl]);
    int cols[3];
    const int nc = halo_targets(ww, W, reflect != 0, cols);
    for (int ri = 0; ri < nr; ++ri)
      for (int ci = 0; ci < nc; ++ci)
        dst[(((int64_t)n * (H + 2) + rows[ri] + 1) * (W + 2) + cols[ci] + 1) * C + cc] = v;
  }
}

// ---- NCHW fp32 -> channels [0,C) of a wider / deeper-halo native tensor ---------------------------
__global__ void __launch_bounds__(256) nchw_to_native_ex_kernel(const float* __restrict__ src,
                                                                __nv_bfloat16* __restrict__ dst, int N,
                                                                int C, int H, int W, int Cdst, int halo) {
  __shared__ float tile[32][33];  // [c][w]
  const int ctiles = (C + 31) / 32;
  const int n = blockIdx.z / ctiles, c0 = (blockIdx.z % ctiles) * 32;
  const int h = blockIdx.y, w0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int c = ty; c < 32; c += 8) {
    const int cc = c0 + c, ww = w0 + tx;
    tile[c][tx] = (cc < C && ww < W) ? src[(((int64_t)n * C + cc) * H + h) * W + ww] : 0.f;
  }
  __syncthreads();
  for (int wl = ty; wl < 32; wl += 8) {
    const int ww = w0 + wl, cc = c0 + tx;
    if (ww >= W || cc >= C) continue;
    dst[(((int64_t)n * (H + 2 * halo) + h + halo) * (W + 2 * halo) + ww + halo) * Cdst + cc] =
        __float2bfloat16_rn(tile[tx][wl]);
  }
}

__global__ void __launch_bounds__(256) native_to_nchw_ex_kernel(const __nv_bfloat16* __restrict__ src,
                                                                float* __restrict__ dst, int N, int C,
                                                                int H, int W, int halo) {
  __shared__ float tile[32][33];  // [w][c]
  const int ctiles = (C + 31) / 32;
  const int n = blockIdx.z / ctiles, c0 = (blockIdx.z % ctiles) * 32;
  const int h = blockIdx.y, w0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int wl = ty; wl < 32; wl += 8) {
    const int ww = w0 + wl, cc = c0 + tx;
    tile[wl][tx] = (ww < W && cc < C)
        ? __bfloat162float(src[(((int64_t)n * (H + 2 * halo) + h + halo) * (W + 2 * halo) + ww + halo) * C + cc])
        : 0.f;
  }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) {
    const int cc = c0 + c, ww = w0 + tx;
    if (cc < C && ww < W) dst[(((int64_t)n * C + cc) * H + h) * W + ww] = tile[tx][c];
  }
}

// ---- native -> NCHW fp32 ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) native_to_nchw_kernel(const __nv_bfloat16* __restrict__ src,
                                                             float* __restrict__ dst, int N, int C,
                                                             int H, int W) {
  __shared__ float tile[32][33];  // [w][c]
  const int ctiles = (C + 31) / 32;
  const int n = blockIdx.z / ctiles, c0 = (blockIdx.z % ctiles) * 32;
  const int h = blockIdx.y, w0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int wl = ty; wl < 32; wl += 8) {
    const int ww = w0 + wl, cc = c0 + tx;
    tile[wl][tx] = (ww < W && cc < C)
                       ? __bfloat162float(src[(((int64_t)n * (H + 2) + h + 1) * (W + 2) + ww + 1) * C + cc])
                       : 0.f;
  }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) {
    const int cc = c0 + c, ww = w0 + tx;
    if (cc < C && ww < W) dst[(((int64_t)n * C + cc) * H + h) * W + ww] = tile[tx][c];
  }
}

// ---- f3 input path: uint8 HWC images <-> the reference's NCHW fp32 tensors ---------------------------
// The loader's transforms.ToTensor() (data_loader.py:114, 132) turns a PIL uint8 HWC image into float CHW / 255;
// transforms.ToPILImage() (train.py:18) turns a float CHW image back into uint8 with pic.mul(255).byte() (truncation).
// Doing both on the device lets images cross PCIe as bytes: 4x less host traffic per stylised image.
// One thread handles 4 consecutive pixels of one image row: 12 bytes in (3 x u32 when aligned), 3 x float4 out.
__global__ void __launch_bounds__(256) u8hwc_to_nchw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                            int64_t npix_total, int64_t hw) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // group of 4 pixels
  const int64_t p0 = q * 4;
  if (p0 >= npix_total) return;
  const int64_t n = p0 / hw, r = p0 - n * hw;
  if (r + 4 <= hw && (hw & 3) == 0) {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + p0 * 3);   // p0 % 4 == 0 -> 12-byte groups, 4-aligned
    const uint32_t a = __ldg(s32), b = __ldg(s32 + 1), c = __ldg(s32 + 2);
    const uint8_t by[12] = {(uint8_t)a, (uint8_t)(a >> 8), (uint8_t)(a >> 16), (uint8_t)(a >> 24),
                            (uint8_t)b, (uint8_t)(b >> 8), (uint8_t)(b >> 16), (uint8_t)(b >> 24),
                            (uint8_t)c, (uint8_t)(c >> 8), (uint8_t)(c >> 16), (uint8_t)(c >> 24)};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float4 v = make_float4(__fdiv_rn((float)by[ch], 255.f), __fdiv_rn((float)by[3 + ch], 255.f),
                             __fdiv_rn((float)by[6 + ch], 255.f), __fdiv_rn((float)by[9 + ch], 255.f));
      *reinterpret_cast<float4*>(dst + (n * 3 + ch) * hw + r) = v;
    }
  } else {
    for (int64_t p = p0; p < p0 + 4 && p < npix_total; ++p) {
      const int64_t nn = p / hw, rr = p - nn * hw;
      for (int ch = 0; ch < 3; ++ch) dst[(nn * 3 + ch) * hw + rr] = __fdiv_rn((float)src[p * 3 + ch], 255.f);
    }
  }
}

__global__ void __launch_bounds__(256) nchw_to_u8hwc_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst,
                                                            int64_t npix_total, int64_t hw) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p0 = q * 4;
  if (p0 >= npix_total) return;
  const int64_t n = p0 / hw, r = p0 - n * hw;
  auto q8 = [](float x) -> uint32_t {     // Hardtanh(0,1) (models.py:315-316), then ToPILImage: mul(255).byte()
    x = fminf(fmaxf(x, 0.f), 1.f);        // NaN -> 0 (fmaxf returns the non-NaN operand)
    return (uint32_t)(x * 255.f);
  };
  if (r + 4 <= hw && (hw & 3) == 0) {
    uint32_t by[12];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float4 v = *reinterpret_cast<const float4*>(src + (n * 3 + ch) * hw + r);
      by[ch] = q8(v.x); by[3 + ch] = q8(v.y); by[6 + ch] = q8(v.z); by[9 + ch] = q8(v.w);
    }
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + p0 * 3);
    d32[0] = by[0] | (by[1] << 8) | (by[2] << 16) | (by[3] << 24);
    d32[1] = by[4] | (by[5] << 8) | (by[6] << 16) | (by[7] << 24);
    d32[2] = by[8] | (by[9] << 8) | (by[10] << 16) | (by[11] << 24);
  } else {
    for (int64_t p = p0; p < p0 + 4 && p < npix_total; ++p) {
      const int64_t nn = p / hw, rr = p - nn * hw;
      for (int ch = 0; ch < 3; ++ch) dst[p * 3 + ch] = (uint8_t)q8(src[(nn * 3 + ch) * hw + rr]);
    }
  }
}

// ---- K1n: AdaIN on the native layout -------------------------------------------------------------
constexpr int kNThreads = 256;
constexpr int kMaxChunks = 32;

// Stage 1: per (map q, image n, pixel chunk) moments of every channel.
// grid = (chunks, N, 1 + K); thread -> 8 consecutive channels of one pixel group.  Per-channel SHIFTED sums
// (d = x - x_first, sum d and sum d^2 with packed FADD2 / FFMA2: 1.5 instructions per element; the streaming Welford
// update this replaces made the kernel issue-bound at 69 % issue-slot utilisation and 40 % of the DRAM rate), pixel
// coordinates advanced by carries (no 64-bit divisions in the loop), four pixels' loads in flight per thread.
struct NativeStatArgs {
  const __nv_bfloat16* maps[1 + AST_MAX_STYLES];
  int H[1 + AST_MAX_STYLES], W[1 + AST_MAX_STYLES];
  float* partial;  // [(1+K)][N][chunks][C][3]
  int N, C, chunks;
};

__global__ void __launch_bounds__(kNThreads) native_stats_kernel(const NativeStatArgs a) {
  extern __shared__ float s_part[];  // [groups][C][3]
  const int q = blockIdx.z, n = blockIdx.y, chunk = blockIdx.x;
  const int cvecs = a.C / 8;
  const int groups = kNThreads / cvecs;
  const int g = threadIdx.x / cvecs, v = threadIdx.x % cvecs;
  const int H = a.H[q], W = a.W[q];
  const int npix = H * W;
  const int per = (npix + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per, p1 = (p0 + per < npix) ? p0 + per : npix;
  float cnt = 0.f;
  float2 sh[4], sm[4], sq[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[j] = sm[j] = sq[j] = make_float2(0.f, 0.f); }
  if (g < groups && p0 + g < p1) {
    const __nv_bfloat16* base = a.maps[q] + ((int64_t)n * (H + 2) + 1) * (W + 2) * a.C + a.C + v * 8;   // pixel (0,0)
    const int row = (W + 2) * a.C;                    // elements per padded row
    int p = p0 + g;
    int h = p / W, w = p - h * W;
    const int dh = groups / W, dw = groups - dh * W;  // one pixel-group step in (h, w)
    {
      float x0[8];
      Vec16<true>::unpack(ld_stream_u4(base + (int64_t)h * row + w * a.C), x0);
#pragma unroll
      for (int j = 0; j < 4; ++j) sh[j] = make_float2(-x0[2 * j], -x0[2 * j + 1]);
    }
    auto acc = [&](const uint4& u) {
      float x[8];
      Vec16<true>::unpack(u, x);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 d = __fadd2_rn(make_float2(x[2 * j], x[2 * j + 1]), sh[j]);
        sm[j] = __fadd2_rn(sm[j], d);
        sq[j] = __ffma2_rn(d, d, sq[j]);
      }
      cnt += 1.f;
    };
    auto step = [&]() {
      p += groups; w += dw; h += dh;
      if (w >= W) { w -= W; ++h; }
    };
    while (p + 3 * groups < p1) {
      uint4 u[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        u[jj] = ld_stream_u4(base + (int64_t)h * row + w * a.C);
        step();
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc(u[jj]);
    }
    while (p < p1) {
      acc(ld_stream_u4(base + (int64_t)h * row + w * a.C));
      step();
    }
  }
  if (g < groups) {
    const float rn = cnt > 0.f ? 1.f / cnt : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sj = (j & 1) ? sm[j >> 1].y : sm[j >> 1].x, qj = (j & 1) ? sq[j >> 1].y : sq[j >> 1].x;
      const float shj = (j & 1) ? sh[j >> 1].y : sh[j >> 1].x;
      float* sp = s_part + ((int64_t)g * a.C + v * 8 + j) * 3;
      sp[0] = cnt; sp[1] = sj * rn - shj; sp[2] = fmaxf(fmaf(-sj, sj * rn, qj), 0.f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += kNThreads) {
    Moments m{s_part[c * 3], s_part[c * 3 + 1], s_part[c * 3 + 2]};
    for (int gg = 1; gg < groups; ++gg) {
      const float* sp = s_part + ((int64_t)gg * a.C + c) * 3;
      m = moments_merge(m, Moments{sp[0], sp[1], sp[2]});
    }
    float* o = a.partial + ((((int64_t)q * a.N + n) * a.chunks + chunk) * a.C + c) * 3;
    o[0] = m.n; o[1] = m.mean; o[2] = m.m2;
  }
}

// Stage 2: merge chunks, build the per-(n,c) affine (mu_c, 1/sigma_c, A, B).
// grid = (C / kCoefCh, N); a CTA of kCoefCh x 32 threads: thread (ch lane, channel) loads the partial of one chunk
// and the 32 chunk lanes are merged by a fixed-order tree in shared memory (deterministic).  The serial loop over
// (1 + K) x 32 partials per thread this replaces took 40 us on 64 CTAs -- as long as the statistics pass itself.
constexpr int kCoefCh = 32;
struct NativeCoefArgs {
  const float* partial;
  float4* coef;  // [N][C]
  float style_w[AST_MAX_STYLES];
  int N, C, chunks, K;
  float eps;
  unsigned flags;
};

__global__ void __launch_bounds__(kCoefCh * kMaxChunks) native_coef_kernel(const NativeCoefArgs a) {
  __shared__ float s_m[kMaxChunks][kCoefCh][3];
  const int n = blockIdx.y;
  const int cl = threadIdx.x % kCoefCh, ch = threadIdx.x / kCoefCh;
  const int c = blockIdx.x * kCoefCh + cl;
  float mu = 0.f, rsig = 0.f, A = 0.f, B = 0.f;
  for (int q = 0; q <= a.K; ++q) {
    Moments m{0.f, 0.f, 0.f};
    if (ch < a.chunks && c < a.C) {
      const float* pp = a.partial + ((((int64_t)q * a.N + n) * a.chunks + ch) * a.C + c) * 3;
      m = Moments{pp[0], pp[1], pp[2]};
    }
    __syncthreads();
    s_m[ch][cl][0] = m.n; s_m[ch][cl][1] = m.mean; s_m[ch][cl][2] = m.m2;
    __syncthreads();
    for (int off = kMaxChunks / 2; off > 0; off >>= 1) {
      if (ch < off) {
        const Moments o{s_m[ch + off][cl][0], s_m[ch + off][cl][1], s_m[ch + off][cl][2]};
        m = moments_merge(m, o);
        s_m[ch][cl][0] = m.n; s_m[ch][cl][1] = m.mean; s_m[ch][cl][2] = m.m2;
      }
      __syncthreads();
    }
    if (ch == 0) {
      const float denom = (a.flags & AST_F_BIASED) ? m.n : m.n - 1.f;
      const float sig = sqrtf(m.m2 / denom + a.eps);
      if (q == 0) {
        mu = m.mean;
        rsig = 1.f / sig;
      } else if (a.flags & AST_F_CANONICAL) {
        A = fmaf(a.style_w[q - 1], sig, A);
        B = fmaf(a.style_w[q - 1], m.mean, B);
      } else {  // models.py:44: style_std := mean(style), style_mean := std(style)
        A = fmaf(a.style_w[q - 1], m.mean, A);
        B = fmaf(a.style_w[q - 1], sig, B);
      }
    }
  }
  if (ch == 0 && c < a.C) a.coef[(int64_t)n * a.C + c] = make_float4(mu, rsig, A, B);
}

// Stage 3: apply + alpha blend, write interior and (optionally) the reflection halo.
// grid = (pixel chunks, N); a thread owns 8 consecutive channels (its 8 coefficient quadruples stay
// in registers) and walks the pixels of its chunk: one 16-byte load and store per pixel, coordinates by carries.
__global__ void __launch_bounds__(kNThreads)
native_apply_kernel(const __nv_bfloat16* __restrict__ content, const float4* __restrict__ coef,
                    __nv_bfloat16* __restrict__ out, int N, int C, int H, int W, float alpha,
                    int reflect, int chunks) {
  const int cvecs = C / 8;
  const int groups = kNThreads / cvecs;
  const int g = threadIdx.x / cvecs, v = threadIdx.x % cvecs;
  if (g >= groups) return;
  const int n = blockIdx.y;
  const bool blend = alpha != 1.f;
  // y = (x - mu) * rsig * A + B, blended: y' = alpha * y + (1 - alpha) * x  ==  x * ka + kb
  float ka[8], kb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 k = __ldg(coef + (int64_t)n * C + v * 8 + j);
    const float sc = k.y * k.z;                       // rsig * A            (models.py:47, 50)
    const float of = fmaf(-k.x, sc, k.w);             // B - mu * rsig * A
    ka[j] = blend ? fmaf(alpha, sc, 1.f - alpha) : sc;    // models.py:471
    kb[j] = blend ? alpha * of : of;
  }
  const int npix = H * W;
  const int per = (npix + chunks - 1) / chunks;
  const int p0 = blockIdx.x * per, p1 = (p0 + per < npix) ? p0 + per : npix;
  if (p0 + g >= p1) return;
  const int row = (W + 2) * C;
  const int64_t img0 = ((int64_t)n * (H + 2) + 1) * row + C + v * 8;      // pixel (0,0), this thread's channels
  int p = p0 + g;
  int h = p / W, w = p - h * W;
  const int dh = groups / W, dw = groups - dh * W;
  while (p < p1) {
    // up to four pixels per iteration: their loads are issued before the first one is used
    uint4 ub[4];
    int hh[4], ww[4], cnt = 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (p < p1) {
        hh[jj] = h; ww[jj] = w;
        ub[jj] = ld_stream_u4(content + img0 + (int64_t)h * row + w * C);
        ++cnt;
        p += groups; w += dw; h += dh;
        if (w >= W) { w -= W; ++h; }
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (jj >= cnt) break;
      float x[8];
      Vec16<true>::unpack(ub[jj], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], ka[j], kb[j]);
      const uint4 o = Vec16<true>::pack(x);
      int rows[3], cols[3];
      const int nr = halo_targets(hh[jj], H, reflect != 0, rows);
      const int nc = halo_targets(ww[jj], W, reflect != 0, cols);
      for (int ri = 0; ri < nr; ++ri)
        for (int ci = 0; ci < nc; ++ci)
          *reinterpret_cast<uint4*>(out + img0 + (int64_t)rows[ri] * row + cols[ci] * C) = o;
    }
  }
}

static int native_chunks(int N, int64_t hw) {
  // ~8 CTAs per SM: these are streaming kernels (2 per SM left them latency-bound at a third of the HBM rate)
  int64_t c = (8 * 148 + N - 1) / N;
  if (c > kMaxChunks) c = kMaxChunks;
  if (c > hw / 32) c = hw / 32;
  if (c < 1) c = 1;
  return (int)c;
}

}  // namespace ast

using namespace ast;

extern "C" int ast_nchw_to_native(const float* nchw, void* native, int N, int C, int H, int W,
                                  int halo, void* stream) {
  if (!nchw || !native || N <= 0 || C <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if (halo == AST_HALO_REFLECT && (H < 2 || W < 2)) return AST_E_SHAPE;
  const int ctiles = (C + 31) / 32;
  if ((int64_t)N * ctiles > 65535 || H > 65535) return AST_E_SHAPE;
  dim3 grid((W + 31) / 32, H, N * ctiles);
  nchw_to_native_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      nchw, reinterpret_cast<__nv_bfloat16*>(native), N, C, H, W, halo == AST_HALO_REFLECT);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_native_to_nchw(const void* native, float* nchw, int N, int C, int H, int W,
                                  void* stream) {
  if (!nchw || !native || N <= 0 || C <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  const int ctiles = (C + 31) / 32;
  if ((int64_t)N * ctiles > 65535 || H > 65535) return AST_E_SHAPE;
  dim3 grid((W + 31) / 32, H, N * ctiles);
  native_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(native), nchw, N, C, H, W);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_nchw_to_native_ex(const float* nchw, void* native, int N, int C, int H, int W,
                                     int Cdst, int dst_halo, void* stream) {
  if (!nchw || !native || N <= 0 || C <= 0 || H <= 0 || W <= 0 || Cdst < C) return AST_E_BADARG;
  if (dst_halo < 1 || dst_halo > 2) return AST_E_BADARG;
  const int ctiles = (C + 31) / 32;
  if ((int64_t)N * ctiles > 65535 || H > 65535) return AST_E_SHAPE;
  dim3 grid((W + 31) / 32, H, N * ctiles);
  nchw_to_native_ex_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      nchw, reinterpret_cast<__nv_bfloat16*>(native), N, C, H, W, Cdst, dst_halo);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_native_to_nchw_ex(const void* native, float* nchw, int N, int C, int H, int W,
                                     int src_halo, void* stream) {
  if (!nchw || !native || N <= 0 || C <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if (src_halo < 1 || src_halo > 2) return AST_E_BADARG;
  const int ctiles = (C + 31) / 32;
  if ((int64_t)N * ctiles > 65535 || H > 65535) return AST_E_SHAPE;
  dim3 grid((W + 31) / 32, H, N * ctiles);
  native_to_nchw_ex_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(native), nchw, N, C, H, W, src_halo);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_u8hwc_to_nchw(const void* u8_nhwc, float* nchw, int N, int H, int W, void* stream) {
  if (!u8_nhwc || !nchw || N <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(u8_nhwc) & 3u) || !aligned16(nchw)) return AST_E_ALIGN;
  const int64_t hw = (int64_t)H * W, total = hw * N, groups = (total + 3) / 4;
  u8hwc_to_nchw_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(u8_nhwc), nchw, total, hw);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_nchw_to_u8hwc(const float* nchw, void* u8_nhwc, int N, int H, int W, void* stream) {
  if (!u8_nhwc || !nchw || N <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(u8_nhwc) & 3u) || !aligned16(nchw)) return AST_E_ALIGN;
  const int64_t hw = (int64_t)H * W, total = hw * N, groups = (total + 3) / 4;
  nchw_to_u8hwc_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      nchw, reinterpret_cast<uint8_t*>(u8_nhwc), total, hw);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t ast_adain_native_ws_bytes(int N, int C, int K) {
  if (N <= 0 || C <= 0 || K < 0) return 0;
  return (size_t)(1 + K) * N * kMaxChunks * C * 3 * sizeof(float) + (size_t)N * C * sizeof(float4);
}

extern "C" int ast_adain_native_fwd(const void* content, const void* const* styles,
                                    const float* style_w, int K, void* out, int N, int C, int H,
                                    int W, const int* Hs, const int* Ws, float alpha, float eps, unsigned flags,
                                    int halo, void* ws, size_t ws_bytes, void* stream) {
  if (!content || !out || !ws || N <= 0 || C <= 0 || H <= 0 || W <= 0 || K < 1) return AST_E_BADARG;
  if (!styles || !style_w || !Hs || !Ws) return AST_E_BADARG;
  if (K > AST_MAX_STYLES) return AST_E_TOOMANY;
  if (C % 8 != 0 || C / 8 > kNThreads) return AST_E_SHAPE;
  if (halo == AST_HALO_REFLECT && (H < 2 || W < 2)) return AST_E_SHAPE;
  if (ws_bytes < ast_adain_native_ws_bytes(N, C, K)) return AST_E_WORKSPACE;
  if (!aligned16(content) || !aligned16(out) || !aligned16(ws)) return AST_E_ALIGN;
  if (N > 65535 || (int64_t)H * W >= 0x7fffffffLL) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  int64_t hw_min = (int64_t)H * W;
  for (int k = 0; k < K; ++k) {
    if (Hs[k] <= 0 || Ws[k] <= 0) return AST_E_BADARG;
    if ((int64_t)Hs[k] * Ws[k] < hw_min) hw_min = (int64_t)Hs[k] * Ws[k];
  }
  const int chunks = native_chunks(N, hw_min);

  float* partial = reinterpret_cast<float*>(ws);
  float4* coef = reinterpret_cast<float4*>(reinterpret_cast<char*>(ws) +
                                           (size_t)(1 + K) * N * kMaxChunks * C * 3 * sizeof(float));
  NativeStatArgs sa = {};
  sa.maps[0] = reinterpret_cast<const __nv_bfloat16*>(content);
  sa.H[0] = H; sa.W[0] = W;
  for (int k = 0; k < K; ++k) {
    if (!styles[k] || !aligned16(styles[k])) return AST_E_BADARG;
    sa.maps[1 + k] = reinterpret_cast<const __nv_bfloat16*>(styles[k]);
    sa.H[1 + k] = Hs[k]; sa.W[1 + k] = Ws[k];
  }
  sa.partial = partial; sa.N = N; sa.C = C; sa.chunks = chunks;
  const int groups = kNThreads / (C / 8);
  const size_t smem = (size_t)groups * C * 3 * sizeof(float);
  if (smem > 48 * 1024) return AST_E_SHAPE;
  native_stats_kernel<<<dim3(chunks, N, 1 + K), kNThreads, smem, s>>>(sa);
  AST_CHECK_LAUNCH();

  NativeCoefArgs ca = {};
  ca.partial = partial; ca.coef = coef; ca.N = N; ca.C = C; ca.chunks = chunks; ca.K = K;
  ca.eps = eps; ca.flags = flags;
  for (int k = 0; k < K; ++k) ca.style_w[k] = style_w[k];
  native_coef_kernel<<<dim3((C + kCoefCh - 1) / kCoefCh, N), kCoefCh * kMaxChunks, 0, s>>>(ca);
  AST_CHECK_LAUNCH();

  int64_t achunks = (8 * 148 + N - 1) / N;
  if (achunks > (int64_t)H * W / 16) achunks = (int64_t)H * W / 16;
  if (achunks < 1) achunks = 1;
  native_apply_kernel<<<dim3((unsigned)achunks, N), kNThreads, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(content), coef, reinterpret_cast<__nv_bfloat16*>(out), N,
      C, H, W, alpha, halo == AST_HALO_REFLECT, (int)achunks);
  AST_CHECK_LAUNCH();
  return 0;
}
