// K3h: compute_hist_loss (losses.py:8-87) -- squared earth mover's distance between the soft 256-bin histograms of
// two image batches -- without the (B, 256, C*H*W) tensors the reference materialises twice (1.6 GB each at
// 8 x 3 x 256 x 256).
//
// Reference: hist_k = (1/N) sum_i [sigma((x_i - mu_k + L/2)/W) - sigma((x_i - mu_k - L/2)/W)], K = 256, L = 1/K,
// W = L/2.5, mu_k = L (k + 1/2), N = x.size(1) * x.size(2) (= C*H, the reference's normaliser; losses.py:42-56);
// cdf = hist @ triu(ones) (losses.py:14-20); loss = mean_b sum_t (cdf_x[t] - cdf_y[t])^2.
//
// With the bin edges e_j = j L the histogram telescopes: phi_k = sigma((x - e_k)/W) - sigma((x - e_{k+1})/W), so
//     cdf_t = (G_0 - G_{t+1}) / N,      G_j = sum_i sigma((x_i - e_j)/W),   j = 0..256,
// and the whole loss needs 257 sums per image.  In units of L an element sits at t = 256 x, the edges at the
// integers and sigma's argument is 2.5 (t - j): edges more than 6.8 below t contribute 1 - e^-17 (= 1 in fp32),
// edges more than 6.8 above contribute < 4.2e-8 and are dropped, so an element costs <= 15 sigmoids instead of 512.
// Accumulation is in 32.32 FIXED POINT with integer atomics (shared, then global): order-independent, hence
// bit-deterministic, and exact to 2^-32 per element where a float accumulator would carry 2^-24 of the running sum.
// Backward is a gather: d loss / d x_i = (g / (N W)) [sigma'(2.5 t) S - sum_j gd_{j-1} sigma'(2.5 (t - j))] with
// gd_t = d loss / d cdf_t (saved by the forward pass) and S = sum_t gd_t; no atomics.
#include "common.cuh"

namespace ast {

constexpr int kHB = 256;                 // bins
constexpr int kHE = kHB + 1;             // edges
constexpr float kHCut = 6.8f;            // 17 W in units of L
constexpr int kHT = 256;

__global__ void __launch_bounds__(kHT) hist_acc_kernel(const float* __restrict__ x, int64_t n,
                                                        unsigned long long* __restrict__ acc, int* __restrict__ bad) {
  __shared__ unsigned long long soft[kHE];
  __shared__ unsigned int cnt[kHE];
  const int b = blockIdx.y;
  for (int j = threadIdx.x; j < kHE; j += kHT) { soft[j] = 0ull; cnt[j] = 0u; }
  __syncthreads();
  const float* xb = x + (int64_t)b * n;
  for (int64_t i = (int64_t)blockIdx.x * kHT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kHT) {
    const float v = xb[i];
    if (!isfinite(v)) { *bad = 1; continue; }
    const float t = fminf(fmaxf(v * (float)kHB, -16.f), (float)kHB + 16.f);
    int jlo = (int)floorf(t - kHCut);                 // edges <= jlo: sigma = 1 in fp32
    if (jlo > kHB) jlo = kHB;
    if (jlo >= 0) atomicAdd(&cnt[jlo], 1u);
    int jhi = (int)ceilf(t + kHCut);
    if (jhi > kHB) jhi = kHB;
    for (int j = jlo < 0 ? 0 : jlo + 1; j <= jhi; ++j) {
      const float s = 1.f / (1.f + __expf(-2.5f * (t - (float)j)));
      atomicAdd(&soft[j], __float2ull_rn(s * 4294967296.f));
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < kHE; j += kHT) {
    unsigned long long c = 0ull;                       // elements with jlo >= j count 1.0 for edge j
    for (int jj = j; jj < kHE; ++jj) c += cnt[jj];
    const unsigned long long tot = soft[j] + (c << 32);
    if (tot) atomicAdd(&acc[(int64_t)b * kHE + j], tot);
  }
}

// one CTA: per image the 256 cdf differences, gd = d loss / d cdf_x, loss = mean_b sum_t diff^2
__global__ void __launch_bounds__(kHB) hist_emd_kernel(const unsigned long long* __restrict__ ax,
                                                        const unsigned long long* __restrict__ ay, double norm_x,
                                                        double norm_y, int B, const int* __restrict__ bad,
                                                        float* __restrict__ gd, float* __restrict__ loss) {
  __shared__ double red[kHB / 32];
  const int t = threadIdx.x;
  double total = 0.0;
  for (int b = 0; b < B; ++b) {
    const unsigned long long* px = ax + (int64_t)b * kHE;
    const unsigned long long* py = ay + (int64_t)b * kHE;
    const double cx = (double)(long long)(px[0] - px[t + 1]) / 4294967296.0 / norm_x;
    const double cy = (double)(long long)(py[0] - py[t + 1]) / 4294967296.0 / norm_y;
    const double d = cx - cy;
    gd[(int64_t)b * kHB + t] = (float)(2.0 * d / (double)B);
    double v = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = v;
    __syncthreads();
    if (t == 0) {
      double s = 0.0;
      for (int w = 0; w < kHB / 32; ++w) s += red[w];
      total += s;
    }
  }
  if (t == 0) loss[0] = *bad ? __int_as_float(0x7fc00000) : (float)(total / (double)B);
}

__global__ void __launch_bounds__(kHT) hist_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gd,
                                                        const float* __restrict__ g_loss, float coef, int64_t n,
                                                        float* __restrict__ gx) {
  __shared__ float s_gd[kHB];
  __shared__ float s_red[kHT / 32];
  __shared__ float s_S;
  const int b = blockIdx.y;
  const float gv = gd[(int64_t)b * kHB + threadIdx.x];     // kHT == kHB
  s_gd[threadIdx.x] = gv;
  float v = gv;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kHT / 32; ++w) s += s_red[w];
    s_S = s;
  }
  __syncthreads();
  const float S = s_S;
  const float k = coef * g_loss[0];
  const float* xb = x + (int64_t)b * n;
  float* gb = gx + (int64_t)b * n;
  for (int64_t i = (int64_t)blockIdx.x * kHT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kHT) {
    const float xv = xb[i];
    const float t = fminf(fmaxf(xv * (float)kHB, -16.f), (float)kHB + 16.f);
    float a = 0.f;
    {
      const float s0 = 1.f / (1.f + __expf(-2.5f * t));
      a = s0 * (1.f - s0) * S;                              // edge 0 enters every cdf_t with +
    }
    int jlo = (int)floorf(t - kHCut) + 1;
    if (jlo < 1) jlo = 1;
    int jhi = (int)ceilf(t + kHCut);
    if (jhi > kHB) jhi = kHB;
    for (int j = jlo; j <= jhi; ++j) {
      const float s = 1.f / (1.f + __expf(-2.5f * (t - (float)j)));
      a = fmaf(-s_gd[j - 1], s * (1.f - s), a);
    }
    gb[i] = k * a;
  }
}

}  // namespace ast

using namespace ast;

extern "C" size_t ast_hist_ws_bytes(int B) {
  return (size_t)2 * (B > 0 ? B : 0) * kHE * sizeof(unsigned long long) + 256;
}

extern "C" int ast_hist_loss_fwd(const float* x, const float* y, int B, int64_t nx, int64_t ny, float norm_x,
                                 float norm_y, float* loss, float* gd, void* ws, size_t ws_bytes, void* stream) {
  if (!x || !y || !loss || !gd || !ws || B <= 0 || nx <= 0 || ny <= 0 || !(norm_x > 0.f) || !(norm_y > 0.f))
    return AST_E_BADARG;
  if (B > 65535) return AST_E_SHAPE;
  if (ws_bytes < ast_hist_ws_bytes(B)) return AST_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(ws) & 7u) != 0) return AST_E_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  AST_CUDA(cudaMemsetAsync(ws, 0, ast_hist_ws_bytes(B), s));
  unsigned long long* ax = reinterpret_cast<unsigned long long*>(ws);
  unsigned long long* ay = ax + (size_t)B * kHE;
  int* bad = reinterpret_cast<int*>(ay + (size_t)B * kHE);
  auto chunks = [&](int64_t n) {
    int64_t c = (n + (int64_t)kHT * 16 - 1) / ((int64_t)kHT * 16);     // >= 16 elements per thread
    const int64_t cap = (148 * 8 + B - 1) / B;
    if (c > cap) c = cap;
    return (unsigned)(c < 1 ? 1 : c);
  };
  hist_acc_kernel<<<dim3(chunks(nx), B), kHT, 0, s>>>(x, nx, ax, bad);
  AST_CHECK_LAUNCH();
  hist_acc_kernel<<<dim3(chunks(ny), B), kHT, 0, s>>>(y, ny, ay, bad);
  AST_CHECK_LAUNCH();
  hist_emd_kernel<<<1, kHB, 0, s>>>(ax, ay, (double)norm_x, (double)norm_y, B, bad, gd, loss);
  AST_CHECK_LAUNCH();
  return 0;
}

// gx = sign * g_loss[0] * d loss / d x for the argument whose forward normaliser was `norm` (sign = +1 for the first
// argument of compute_hist_loss, -1 for the second).
extern "C" int ast_hist_loss_bwd(const float* x, const float* gd, const float* g_loss, float sign, float norm,
                                 float* gx, int B, int64_t n, void* stream) {
  if (!x || !gd || !g_loss || !gx || B <= 0 || n <= 0 || !(norm > 0.f)) return AST_E_BADARG;
  if (B > 65535) return AST_E_SHAPE;
  int64_t c = (n + (int64_t)kHT * 8 - 1) / ((int64_t)kHT * 8);
  const int64_t cap = (148 * 8 + B - 1) / B;
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  const float coef = sign * 2.5f * (float)kHB / norm;      // 1 / (N W), W = 1 / (2.5 K)
  hist_bwd_kernel<<<dim3((unsigned)c, B), kHT, 0, (cudaStream_t)stream>>>(x, gd, g_loss, coef, n, gx);
  AST_CHECK_LAUNCH();
  return 0;
}
