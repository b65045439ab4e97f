// K3: loss kernels -- Huber (content loss) and Gram matrix, forward + backward.
//
// Reference arithmetic replaced (paths relative to /root/reference):
//   compute_content_loss  losses.py:124-126   F.huber_loss(inp, tgt): delta 1, mean
//   gram_matrix           losses.py:105-109   bmm(X, X^T) / (C*H*W)
//   compute_style_loss    losses.py:128-139   composes the two with channel_stats (adain.cu)
#include "common.cuh"

namespace ast {

constexpr int kLThreads = 256;
constexpr int kHuberMaxBlocks = 1024;

__device__ __forceinline__ float huber1(float d) {
  float a = fabsf(d);
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}

// stage 1: per-block partial sums (fixed assignment -> deterministic)
__global__ void __launch_bounds__(kLThreads) huber_partial_kernel(const float* __restrict__ inp,
                                                                  const float* __restrict__ tgt,
                                                                  float* __restrict__ partial,
                                                                  int64_t n) {
  __shared__ float s_red[kLThreads / 32];
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * kLThreads;
  const bool vec = aligned16(inp) && aligned16(tgt);
  const int64_t nvec = vec ? n / 4 : 0;
  for (int64_t i = (int64_t)blockIdx.x * kLThreads + threadIdx.x; i < nvec; i += stride) {
    float4 a = __ldg(reinterpret_cast<const float4*>(inp) + i);
    float4 b = __ldg(reinterpret_cast<const float4*>(tgt) + i);
    acc += huber1(a.x - b.x) + huber1(a.y - b.y) + huber1(a.z - b.z) + huber1(a.w - b.w);
  }
  for (int64_t i = nvec * 4 + (int64_t)blockIdx.x * kLThreads + threadIdx.x; i < n; i += stride)
    acc += huber1(inp[i] - tgt[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < kLThreads / 32; ++w) r += s_red[w];
    partial[blockIdx.x] = r;
  }
}

// stage 2: one block folds the partials in a fixed order
__global__ void __launch_bounds__(kLThreads) huber_final_kernel(const float* __restrict__ partial,
                                                                int nblocks, float* loss,
                                                                float scale_over_n) {
  __shared__ float s_red[kLThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += kLThreads) acc += partial[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < kLThreads / 32; ++w) r += s_red[w];
    loss[0] = r * scale_over_n;
  }
}

__global__ void __launch_bounds__(kLThreads) huber_bwd_kernel(const float* __restrict__ inp,
                                                              const float* __restrict__ tgt,
                                                              const float* __restrict__ g_loss,
                                                              float* __restrict__ g_inp, int64_t n,
                                                              float scale_over_n) {
  const float g = g_loss[0] * scale_over_n;
  const int64_t stride = (int64_t)gridDim.x * kLThreads;
  for (int64_t i = (int64_t)blockIdx.x * kLThreads + threadIdx.x; i < n; i += stride) {
    float d = inp[i] - tgt[i];
    g_inp[i] = g * fminf(fmaxf(d, -1.f), 1.f);
  }
}

// ---- total variation: sum (x[..,w]-x[..,w+1])^2 + sum (x[..,h,:]-x[..,h+1,:])^2   (losses.py:90-103) ------
// planes = N*C images of H x W fp32; same deterministic two-stage reduction as the Huber loss.
__global__ void __launch_bounds__(kLThreads) tv_partial_kernel(const float* __restrict__ img,
                                                               float* __restrict__ partial, int64_t planes,
                                                               int H, int W) {
  __shared__ float s_red[kLThreads / 32];
  float acc = 0.f;
  const int64_t n = planes * H * W;
  const int64_t stride = (int64_t)gridDim.x * kLThreads;
  for (int64_t i = (int64_t)blockIdx.x * kLThreads + threadIdx.x; i < n; i += stride) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const float x = img[i];
    if (w + 1 < W) { const float d = x - img[i + 1]; acc = fmaf(d, d, acc); }
    if (h + 1 < H) { const float d = x - img[i + W]; acc = fmaf(d, d, acc); }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < kLThreads / 32; ++w) r += s_red[w];
    partial[blockIdx.x] = r;
  }
}

// d loss / d x[h][w] = 2 g [ (x - right) - (left - x) + (x - below) - (above - x) ] with absent neighbours dropped
__global__ void __launch_bounds__(kLThreads) tv_bwd_kernel(const float* __restrict__ img,
                                                           const float* __restrict__ g_loss,
                                                           float* __restrict__ g_img, int64_t planes, int H, int W) {
  const float g2 = 2.f * g_loss[0];
  const int64_t n = planes * H * W;
  const int64_t stride = (int64_t)gridDim.x * kLThreads;
  for (int64_t i = (int64_t)blockIdx.x * kLThreads + threadIdx.x; i < n; i += stride) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const float x = img[i];
    float d = 0.f;
    if (w + 1 < W) d += x - img[i + 1];
    if (w > 0) d -= img[i - 1] - x;
    if (h + 1 < H) d += x - img[i + W];
    if (h > 0) d -= img[i - W] - x;
    g_img[i] = g2 * d;
  }
}

// ---- Gram: G[b] = X X^T / (C*HW) ---------------------------------------------------------------
// 64x64 output tile per CTA, 4x4 per thread, K chunks of 32 through shared memory, split-K over HW
// with fp32 atomics into a pre-zeroed G.  Only tiles with tj >= ti are computed; the mirror is
// written in the same pass.
constexpr int GT = 64, GK = 32;

__global__ void __launch_bounds__(256) gram_fwd_kernel(const float* __restrict__ x,
                                                       float* __restrict__ g, int C, int64_t HW,
                                                       int64_t k_per_split, float scale) {
  __shared__ float sA[GK][GT + 4];
  __shared__ float sB[GK][GT + 4];
  const int b = blockIdx.z;
  const int tiles = (C + GT - 1) / GT;
  // decode upper-triangular tile index
  int t = blockIdx.x, ti = 0;
  while (t >= tiles - ti) { t -= tiles - ti; ++ti; }
  const int tj = ti + t;
  const int64_t k0 = (int64_t)blockIdx.y * k_per_split;
  const int64_t k1 = (k0 + k_per_split < HW) ? k0 + k_per_split : HW;
  const float* xb = x + (int64_t)b * C * HW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int64_t kk = k0; kk < k1; kk += GK) {
    // 64 rows x 32 k per operand; 256 threads -> 8 elements each
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      int r = e / GK, k = e % GK;
      int64_t kg = kk + k;
      int ra = ti * GT + r, rb = tj * GT + r;
      sA[k][r] = (ra < C && kg < k1) ? xb[(int64_t)ra * HW + kg] : 0.f;
      sB[k][r] = (rb < C && kg < k1) ? xb[(int64_t)rb * HW + kg] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; bb[i] = sB[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* gb = g + (int64_t)b * C * C;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = ti * GT + ty * 4 + i, c = tj * GT + tx * 4 + j;
      if (r < C && c < C) {
        float v = acc[i][j] * scale;
        atomicAdd(&gb[(int64_t)r * C + c], v);
        if (ti != tj) atomicAdd(&gb[(int64_t)c * C + r], v);
      }
    }
}

// gx[b] = S X * scale, S = gG + gG^T (C x C), X (C x HW).  Tile = 64 rows (c) x 256 columns (q) per
// CTA, 8 x 8 outputs per thread, K chunks of 16 through shared memory: four 128-bit shared loads per
// 64 FMAs (the earlier 4 x 4 micro-tile was shared-memory-load bound).
constexpr int GB_R = 64, GB_Q = 256, GB_K = 16;

__global__ void __launch_bounds__(256) gram_bwd_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ gg,
                                                       float* __restrict__ gx, int C, int64_t HW,
                                                       float scale) {
  __shared__ __align__(16) float sS[GB_K][GB_R];   // S^T chunk: [k][row]
  __shared__ __align__(16) float sX[GB_K][GB_Q];   // X chunk:   [k][col]
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * GB_R;
  const int64_t c0 = (int64_t)blockIdx.x * GB_Q;
  const float* xb = x + (int64_t)b * C * HW;
  const float* gb = gg + (int64_t)b * C * C;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // tx -> 8 columns, ty -> 8 rows
  const bool vec_ok = (HW % 4 == 0) && aligned16(xb);
  float acc[8][8] = {};
  for (int kk = 0; kk < C; kk += GB_K) {
    for (int e = threadIdx.x; e < GB_K * GB_R; e += 256) {
      const int k = e / GB_R, r = e % GB_R;
      const int rg = r0 + r, kg = kk + k;
      sS[k][r] = (rg < C && kg < C) ? gb[(int64_t)rg * C + kg] + gb[(int64_t)kg * C + rg] : 0.f;
    }
    for (int e = threadIdx.x; e < GB_K * GB_Q / 4; e += 256) {
      const int k = e / (GB_Q / 4), c4 = (e % (GB_Q / 4)) * 4;
      const int kg = kk + k;
      const int64_t cg = c0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kg < C) {
        if (vec_ok && cg + 3 < HW) {
          v = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)kg * HW + cg));
        } else {
          if (cg < HW) v.x = xb[(int64_t)kg * HW + cg];
          if (cg + 1 < HW) v.y = xb[(int64_t)kg * HW + cg + 1];
          if (cg + 2 < HW) v.z = xb[(int64_t)kg * HW + cg + 2];
          if (cg + 3 < HW) v.w = xb[(int64_t)kg * HW + cg + 3];
        }
      }
      *reinterpret_cast<float4*>(&sX[k][c4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sS[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sS[k][ty * 8 + 4]);
      // columns owned by this thread: tx*4 .. tx*4+3 and 128 + tx*4 .. (conflict-free 128-bit loads)
      const float4 b0 = *reinterpret_cast<const float4*>(&sX[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sX[k][128 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* ob = gx + (int64_t)b * C * HW;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty * 8 + i;
    if (r >= C) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t c = c0 + half * 128 + tx * 4;
      float* o = ob + (int64_t)r * HW + c;
      if (vec_ok && aligned16(ob) && c + 3 < HW) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[i][half * 4] * scale, acc[i][half * 4 + 1] * scale,
                                                    acc[i][half * 4 + 2] * scale, acc[i][half * 4 + 3] * scale);
      } else {
        for (int j = 0; j < 4; ++j)
          if (c + j < HW) o[j] = acc[i][half * 4 + j] * scale;
      }
    }
  }
}

}  // namespace ast

using namespace ast;

static int huber_blocks(int64_t n) {
  int64_t b = (n + (int64_t)kLThreads * 16 - 1) / ((int64_t)kLThreads * 16);
  if (b < 1) b = 1;
  if (b > kHuberMaxBlocks) b = kHuberMaxBlocks;
  return (int)b;
}

extern "C" size_t ast_huber_ws_bytes(int64_t) { return kHuberMaxBlocks * sizeof(float); }

extern "C" int ast_huber_fwd(const float* inp, const float* tgt, float* loss, int64_t n,
                             float scale, void* ws, size_t ws_bytes, void* stream) {
  if (!inp || !tgt || !loss || !ws || n <= 0) return AST_E_BADARG;
  if (ws_bytes < kHuberMaxBlocks * sizeof(float)) return AST_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = huber_blocks(n);
  huber_partial_kernel<<<nb, kLThreads, 0, s>>>(inp, tgt, (float*)ws, n);
  AST_CHECK_LAUNCH();
  huber_final_kernel<<<1, kLThreads, 0, s>>>((const float*)ws, nb, loss, scale / (float)n);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_tv_fwd(const float* img, float* loss, int64_t planes, int H, int W, void* ws, size_t ws_bytes,
                          void* stream) {
  if (!img || !loss || !ws || planes <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if (ws_bytes < kHuberMaxBlocks * sizeof(float)) return AST_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = huber_blocks(planes * H * W);
  tv_partial_kernel<<<nb, kLThreads, 0, s>>>(img, (float*)ws, planes, H, W);
  AST_CHECK_LAUNCH();
  huber_final_kernel<<<1, kLThreads, 0, s>>>((const float*)ws, nb, loss, 1.f);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_tv_bwd(const float* img, const float* g_loss, float* g_img, int64_t planes, int H, int W,
                          void* stream) {
  if (!img || !g_loss || !g_img || planes <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  int64_t nb = (planes * H * W + kLThreads - 1) / kLThreads;
  if (nb > 148 * 16) nb = 148 * 16;
  tv_bwd_kernel<<<(unsigned)nb, kLThreads, 0, (cudaStream_t)stream>>>(img, g_loss, g_img, planes, H, W);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_huber_bwd(const float* inp, const float* tgt, const float* g_loss, float* g_inp,
                             int64_t n, float scale, void* stream) {
  if (!inp || !tgt || !g_loss || !g_inp || n <= 0) return AST_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  int64_t nb = (n + kLThreads * 4 - 1) / (kLThreads * 4);
  if (nb > 148 * 16) nb = 148 * 16;
  huber_bwd_kernel<<<(unsigned)nb, kLThreads, 0, s>>>(inp, tgt, g_loss, g_inp, n, scale / (float)n);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_gram_fwd(const float* x, float* g, int B, int C, int64_t HW, void* stream) {
  if (!x || !g || B <= 0 || C <= 0 || HW <= 0) return AST_E_BADARG;
  if (B > 65535) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  AST_CUDA(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)B * C * C, s));
  const int tiles = (C + GT - 1) / GT;
  const int ntri = tiles * (tiles + 1) / 2;
  // split K so that the grid fills the machine (~4 CTAs per SM), chunks multiple of GK
  int64_t want = (4 * 148 + (int64_t)ntri * B - 1) / ((int64_t)ntri * B);
  int64_t chunks = (HW + GK - 1) / GK;
  if (want > chunks) want = chunks;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  int64_t k_per_split = ((chunks + want - 1) / want) * GK;
  int splits = (int)((HW + k_per_split - 1) / k_per_split);
  dim3 grid(ntri, splits, B);
  gram_fwd_kernel<<<grid, 256, 0, s>>>(x, g, C, HW, k_per_split, 1.f / ((float)C * (float)HW));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_gram_bwd(const float* x, const float* gg, float* gx, int B, int C, int64_t HW,
                            void* stream) {
  if (!x || !gg || !gx || B <= 0 || C <= 0 || HW <= 0) return AST_E_BADARG;
  if (B > 65535 || (C + GB_R - 1) / GB_R > 65535) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)((HW + GB_Q - 1) / GB_Q), (C + GB_R - 1) / GB_R, B);
  gram_bwd_kernel<<<grid, 256, 0, s>>>(x, gg, gx, C, HW, 1.f / ((float)C * (float)HW));
  AST_CHECK_LAUNCH();
  return 0;
}
