// K4t: training-mode pieces of the MobileNet-style blocks (train_autoencoder.py:111-148 over
// models.py:140-338 / mobilenetv2.py:63-181): BatchNorm with batch statistics (forward + backward),
// Hardswish / SE / residual element-wise passes, depthwise data and weight gradients, SE backward,
// stem / head gradients.  Everything is HBM-bound streaming over plain NHWC 16-bit [pixel][channel]
// matrices (row stride `ld`): ACTIVATIONS are fp16 (act_t), GRADIENTS bf16 (grad_t) -- the pointer type picks the
// conversion (common.cuh) --, fp32 arithmetic, fp64 for the cross-CTA BatchNorm accumulators.
//
// Thread mapping shared by the streaming kernels: a CTA of 256 threads covers `groups` pixels at a time,
// thread -> (pixel group g, 8-channel vector v); one 16-byte load per tensor per pixel, consecutive
// threads read consecutive 16-byte pieces of a pixel row and then the next pixel (coalesced).
// grid = (pixel chunks, N).
#include "common.cuh"

namespace ast {

constexpr int kT = 256;

__device__ __forceinline__ float hsw(float x) { return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f); }
// d Hardswish / dx as ATen's CPU hardswish_backward (what the reference runs): 0 for x <= -3, x/3 + 1/2 on the
// OPEN interval (-3, 3), 1 for x >= 3.  bf16-stored pre-activations hit +-3.0 exactly about once per
// thousand elements, so the convention at the kinks is measurable.
__device__ __forceinline__ float hsw_grad(float x) {
  return x <= -3.f ? 0.f : (x < 3.f ? fmaf(x, 1.f / 3.f, 0.5f) : 1.f);
}
__device__ __forceinline__ int reflect_idx(int p, int X) {
  p = p < 0 ? -p : p;
  return p >= X ? 2 * X - 2 - p : p;
}

struct RowMap {
  int cv, groups, g, v;
  bool on;
};
__device__ __forceinline__ RowMap row_map(int C) {
  RowMap m;
  m.cv = C >> 3;
  m.groups = kT / m.cv;
  m.g = threadIdx.x / m.cv;
  m.v = threadIdx.x - m.g * m.cv;
  m.on = m.g < m.groups;
  return m;
}
__device__ __forceinline__ void chunk_range(int64_t HW, int64_t& p0, int64_t& p1) {
  const int64_t per = (HW + gridDim.x - 1) / gridDim.x;
  p0 = (int64_t)blockIdx.x * per;
  p1 = p0 + per < HW ? p0 + per : HW;
}
__device__ __forceinline__ void ldf8(const float* p, float (&x)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void fill8(float (&x)[8], float v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = v;
}

// Sum Q per-channel quantities held per thread (q[Q][8]) over the CTA's pixel groups and add them to
// dst[(k * C) + c] (dst already offset to the image / tensor).  smem: groups * Q * C floats.
template <int Q, typename AccT>
__device__ __forceinline__ void block_channel_reduce(const RowMap& m, int C, float (&q)[Q][8], float* s_red,
                                                     AccT* dst) {
  if (m.on) {
#pragma unroll
    for (int k = 0; k < Q; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[(m.g * Q + k) * C + m.v * 8 + j] = q[k][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Q * C; i += kT) {
    float s = 0.f;
    for (int gg = 0; gg < m.groups; ++gg) s += s_red[gg * Q * C + i];
    atomicAdd(dst + i, (AccT)s);
  }
}

// ---- BatchNorm2d batch statistics: sums[0][c] = sum x, sums[1][c] = sum x^2 (fp64 accumulators) ----
template <typename AT>
__global__ void __launch_bounds__(kT)
bn_stats_kernel(const AT* __restrict__ x, int ld, double* __restrict__ sums, int C, int64_t HW) {
  extern __shared__ float s_red[];
  const RowMap m = row_map(C);
  int64_t p0, p1;
  chunk_range(HW, p0, p1);
  float q[2][8];
  fill8(q[0], 0.f); fill8(q[1], 0.f);
  if (m.on) {
    const AT* xp = x + (int64_t)blockIdx.y * HW * ld + m.v * 8;
    for (int64_t p = p0 + m.g; p < p1; p += m.groups) {
      float a[8];
      ld8(xp + p * ld, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) { q[0][j] += a[j]; q[1][j] = fmaf(a[j], a[j], q[1][j]); }
    }
  }
  block_channel_reduce<2, double>(m, C, q, s_red, sums);
}

// mean / biased variance -> (mean, invstd, scale = gamma * invstd, shift = beta - mean * scale), and the
// running-statistics update of nn.BatchNorm2d (momentum, UNBIASED variance).  stat = [4][C].
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps,
                                   float* __restrict__ stat, int C, long long* __restrict__ num_batches_tracked) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;   // nn.BatchNorm2d's counter (one thread: no race)
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  stat[c] = (float)mean;
  stat[C + c] = invstd;
  stat[2 * C + c] = sc;
  stat[3 * C + c] = beta[c] - (float)mean * sc;
  if (running_mean) {
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// ---- y = act(x * sc[c] + sh[c]) [* se[n][c]] [+ res]; optional per-(n,c) pool of act(..) ------------
template <typename AT>
__global__ void __launch_bounds__(kT)
affine_act_kernel(const AT* __restrict__ x, int ld_x, const float* __restrict__ sc,
                  const float* __restrict__ sh, int act, const float* __restrict__ se,
                  const AT* __restrict__ res, int ld_res, AT* __restrict__ out, int ld_out,
                  float* __restrict__ pool, int C, int64_t HW) {
  extern __shared__ float s_red[];
  const RowMap m = row_map(C);
  const int n = blockIdx.y;
  int64_t p0, p1;
  chunk_range(HW, p0, p1);
  float q[1][8];
  fill8(q[0], 0.f);
  if (m.on) {
    float a[8], b[8], s[8];
    fill8(a, 1.f); fill8(b, 0.f); fill8(s, 1.f);
    if (sc) { ldf8(sc + m.v * 8, a); ldf8(sh + m.v * 8, b); }
    if (se) ldf8(se + (int64_t)n * C + m.v * 8, s);
    const int64_t row0 = (int64_t)n * HW;
    for (int64_t p = p0 + m.g; p < p1; p += m.groups) {
      float v[8];
      ld8(x + (row0 + p) * ld_x + m.v * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = fmaf(v[j], a[j], b[j]);
        if (act) v[j] = hsw(v[j]);
        q[0][j] += v[j];
        v[j] *= s[j];
      }
      if (res) {
        float r[8];
        ld8(res + (row0 + p) * ld_res + m.v * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      }
      if (out) st8(out + (row0 + p) * ld_out + m.v * 8, v);
    }
  }
  if (pool) block_channel_reduce<1, float>(m, C, q, s_red, pool + (int64_t)n * C);
}

// ---- backward through u = Hardswish(z) * s, z = a * sc + sh:  per-(n,c) sums over pixels -------------
//   T0 = sum du * h          (-> gradient of the SE scale)            h = Hardswish(z)
//   T1 = sum du * h'(z)      T2 = sum h'(z)      T3 = sum du * h'(z) * ahat      T4 = sum h'(z) * ahat
// T1..T4 feed the BatchNorm backward (only when mean != null); out = [N][5][C].
template <typename AT, bool NORM>
__global__ void __launch_bounds__(kT)
dw_bwd_reduce_kernel(const grad_t* __restrict__ du, const AT* __restrict__ a,
                     const float* __restrict__ stat /*[4][C] (NORM only)*/, float* __restrict__ out, int C,
                     int64_t HW) {
  extern __shared__ float s_red[];
  const RowMap m = row_map(C);
  const int n = blockIdx.y;
  int64_t p0, p1;
  chunk_range(HW, p0, p1);
  constexpr int Q = NORM ? 5 : 1;   // without a norm only T0 is needed: a small, high-occupancy kernel
  float q[Q][8];
#pragma unroll
  for (int k = 0; k < Q; ++k) fill8(q[k], 0.f);
  if (m.on) {
    float mu[8], is[8], sc[8], sh[8];
    if (NORM) {
      ldf8(stat + m.v * 8, mu); ldf8(stat + C + m.v * 8, is);
      ldf8(stat + 2 * C + m.v * 8, sc); ldf8(stat + 3 * C + m.v * 8, sh);
    }
    const int64_t row0 = (int64_t)n * HW;
    for (int64_t p = p0 + m.g; p < p1; p += 2 * m.groups) {
      const bool two = p + m.groups < p1;
      float g[2][8], av[2][8];
      ld8(du + (row0 + p) * C + m.v * 8, g[0]);
      ld8(a + (row0 + p) * C + m.v * 8, av[0]);
      if (two) {
        ld8(du + (row0 + p + m.groups) * C + m.v * 8, g[1]);
        ld8(a + (row0 + p + m.groups) * C + m.v * 8, av[1]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (NORM) {
            const float z = fmaf(av[u][j], sc[j], sh[j]);
            const float hp = hsw_grad(z);
            const float ah = (av[u][j] - mu[j]) * is[j];
            q[0][j] = fmaf(g[u][j], hsw(z), q[0][j]);
            q[NORM ? 1 : 0][j] = fmaf(g[u][j], hp, q[NORM ? 1 : 0][j]);
            q[NORM ? 2 : 0][j] += hp;
            q[NORM ? 3 : 0][j] = fmaf(g[u][j] * hp, ah, q[NORM ? 3 : 0][j]);
            q[NORM ? 4 : 0][j] = fmaf(hp, ah, q[NORM ? 4 : 0][j]);
          } else {
            q[0][j] = fmaf(g[u][j], hsw(av[u][j]), q[0][j]);
          }
        }
      }
    }
  }
  block_channel_reduce<Q, float>(m, C, q, s_red, out + (int64_t)n * 5 * C);
}

// BatchNorm backward coefficients of the dw conv's norm from the per-sample sums above:
//   dz = (du * s + g) * h'(z);  dbeta = sum dz = sum_n s T1 + g T2;  dgamma = sum dz * ahat = sum_n s T3 + g T4
// coef[0][c] = dbeta / M, coef[1][c] = dgamma / M.
__global__ void se_bn_combine_kernel(const float* __restrict__ T, const float* __restrict__ s,
                                     const float* __restrict__ g, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ coef, int N, int C,
                                     float inv_count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float db = 0.f, dg = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* t = T + (int64_t)n * 5 * C;
    const float sv = s[(int64_t)n * C + c], gv = g[(int64_t)n * C + c];
    db += sv * t[C + c] + gv * t[2 * C + c];
    dg += sv * t[3 * C + c] + gv * t[4 * C + c];
  }
  dgamma[c] = dg;
  dbeta[c] = db;
  coef[c] = db * inv_count;
  coef[C + c] = dg * inv_count;
}

// da = ((du * s[n][c] + g[n][c]) * h'(z) - coef0 - ahat * coef1) * sc      (stat == null: da = (..) * h'(a))
template <typename AT, bool NORM>
__global__ void __launch_bounds__(kT)
dw_bwd_apply_kernel(const grad_t* __restrict__ du, const AT* __restrict__ a,
                    const float* __restrict__ s, const float* __restrict__ g, const float* __restrict__ stat,
                    const float* __restrict__ coef, grad_t* __restrict__ da, int C, int64_t HW) {
  const RowMap m = row_map(C);
  if (!m.on) return;
  const int n = blockIdx.y;
  int64_t p0, p1;
  chunk_range(HW, p0, p1);
  float mu[8], is[8], sc[8], sh[8], c0[8], c1[8], sv[8], gv[8];
  if (NORM) {
    ldf8(stat + m.v * 8, mu); ldf8(stat + C + m.v * 8, is);
    ldf8(stat + 2 * C + m.v * 8, sc); ldf8(stat + 3 * C + m.v * 8, sh);
    ldf8(coef + m.v * 8, c0); ldf8(coef + C + m.v * 8, c1);
  }
  ldf8(s + (int64_t)n * C + m.v * 8, sv);
  ldf8(g + (int64_t)n * C + m.v * 8, gv);
  const int64_t row0 = (int64_t)n * HW;
  for (int64_t p = p0 + m.g; p < p1; p += 2 * m.groups) {   // two pixels per iteration: 4 loads in flight
    const bool two = p + m.groups < p1;
    float d[2][8], av[2][8];
    ld8(du + (row0 + p) * C + m.v * 8, d[0]);
    ld8(a + (row0 + p) * C + m.v * 8, av[0]);
    if (two) {
      ld8(du + (row0 + p + m.groups) * C + m.v * 8, d[1]);
      ld8(a + (row0 + p + m.groups) * C + m.v * 8, av[1]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = NORM ? fmaf(av[u][j], sc[j], sh[j]) : av[u][j];
        const float dz = fmaf(d[u][j], sv[j], gv[j]) * hsw_grad(z);
        d[u][j] = NORM ? (dz - c0[j] - (av[u][j] - mu[j]) * is[j] * c1[j]) * sc[j] : dz;
      }
      st8(da + (row0 + p + u * m.groups) * C + m.v * 8, d[u]);
    }
  }
}

// ---- generic BatchNorm backward: sums[0][c] = sum dy, sums[1][c] = sum dy * ahat (fp64) -------------
template <typename AT>
__global__ void __launch_bounds__(kT)
bn_bwd_reduce_kernel(const grad_t* __restrict__ dy, int ld_dy, const AT* __restrict__ a,
                     int ld_a, const float* __restrict__ stat, double* __restrict__ sums, int C, int64_t HW) {
  extern __shared__ float s_red[];
  const RowMap m = row_map(C);
  int64_t p0, p1;
  chunk_range(HW, p0, p1);
  float q[2][8];
  fill8(q[0], 0.f); fill8(q[1], 0.f);
  if (m.on) {
    float mu[8], is[8];
    ldf8(stat + m.v * 8, mu); ldf8(stat + C + m.v * 8, is);
    const int64_t row0 = (int64_t)blockIdx.y * HW;
    for (int64_t p = p0 + m.g; p < p1; p += m.groups) {
      float d[8], av[8];
      ld8(dy + (row0 + p) * ld_dy + m.v * 8, d);
      ld8(a + (row0 + p) * ld_a + m.v * 8, av);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        q[0][j] += d[j];
        q[1][j] = fmaf(d[j], (av[j] - mu[j]) * is[j], q[1][j]);
      }
    }
  }
  block_channel_reduce<2, double>(m, C, q, s_red, sums);
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double inv_count, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (float)sums[c];
  dgamma[c] = (float)sums[C + c];
  coef[c] = (float)(sums[c] * inv_count);
  coef[C + c] = (float)(sums[C + c] * inv_count);
}

// da = (dy - coef0 - ahat * coef1) * sc
template <typename AT>
__global__ void __launch_bounds__(kT)
bn_bwd_apply_kernel(const grad_t* __restrict__ dy, int ld_dy, const AT* __restrict__ a, int ld_a,
                    const float* __restrict__ stat, const float* __restrict__ coef, grad_t* __restrict__ da,
                    int C, int64_t HW) {
  const RowMap m = row_map(C);
  if (!m.on) return;
  int64_t p0, p1;
  chunk_range(HW, p0, p1);
  float mu[8], is[8], sc[8], c0[8], c1[8];
  ldf8(stat + m.v * 8, mu); ldf8(stat + C + m.v * 8, is); ldf8(stat + 2 * C + m.v * 8, sc);
  ldf8(coef + m.v * 8, c0); ldf8(coef + C + m.v * 8, c1);
  const int64_t row0 = (int64_t)blockIdx.y * HW;
  for (int64_t p = p0 + m.g; p < p1; p += m.groups) {
    float d[8], av[8];
    ld8(dy + (row0 + p) * ld_dy + m.v * 8, d);
    ld8(a + (row0 + p) * ld_a + m.v * 8, av);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = (d[j] - c0[j] - (av[j] - mu[j]) * is[j] * c1[j]) * sc[j];
    st8(da + (row0 + p) * C + m.v * 8, d);
  }
}

// ---- depthwise conv: data gradient ------------------------------------------------------------------
// forward: y[o] = sum_k w[k] * xin[R(o*s + k - pad)] on the conv-input grid (Hin x Win; twice the stored
// tensor when up2), R = reflection.  Gather form: a conv-input position i receives from every padded
// position q with R(q) = i:  q = i, q = -i (1 <= i <= pad), q = 2(X-1) - i (X-1-pad <= i <= X-2); with up2
// the stored pixel sums its 2x2 conv-input positions.  Optional epilogue: * Hardswish'(a_pre * sc + sh)
// (the dw input was Hardswish(BN(a_pre))).
__device__ __forceinline__ int reflect_cands(int i, int X, int pad, int (&q)[3]) {
  int n = 0;
  q[n++] = i;
  if (i >= 1 && i <= pad) q[n++] = -i;
  const int r = 2 * (X - 1) - i;
  if (i <= X - 2 && r <= X - 1 + pad) q[n++] = r;
  return n;
}

template <typename AT>
__global__ void __launch_bounds__(kT)
dw_dgrad_kernel(const grad_t* __restrict__ dy, const float* __restrict__ w /*[k*k][C]*/,
                const AT* __restrict__ a_pre, const float* __restrict__ stat,
                const grad_t* __restrict__ dres, grad_t* __restrict__ dx, int C, int H, int W, int Ho, int Wo, int k, int stride, int up2) {
  const RowMap m = row_map(C);
  if (!m.on) return;
  const int n = blockIdx.y;
  const int pad = (k - 1) / 2;
  const int Hin = up2 ? 2 * H : H, Win = up2 ? 2 * W : W;
  int64_t p0, p1;
  chunk_range((int64_t)H * W, p0, p1);
  float sc[8], sh[8];
  fill8(sc, 1.f); fill8(sh, 0.f);
  if (stat) { ldf8(stat + 2 * C + m.v * 8, sc); ldf8(stat + 3 * C + m.v * 8, sh); }
  const grad_t* dyn = dy + (int64_t)n * Ho * Wo * C + m.v * 8;
  const float* wv = w + m.v * 8;
  const int reps = up2 ? 2 : 1;
  for (int64_t p = p0 + m.g; p < p1; p += m.groups) {
    const int ih = (int)(p / W), iw = (int)(p % W);
    float acc[8];
    fill8(acc, 0.f);
    for (int ry = 0; ry < reps; ++ry) {
      int qh[3];
      const int uy = up2 ? 2 * ih + ry : ih;
      const int nh = reflect_cands(uy, Hin, pad, qh);
      for (int rx = 0; rx < reps; ++rx) {
        int qw[3];
        const int ux = up2 ? 2 * iw + rx : iw;
        const int nw = reflect_cands(ux, Win, pad, qw);
        if (dres) {   // identity branch of the block (stride 1: Ho x Wo == Hin x Win)
          float r[8];
          ld8(dres + (((int64_t)n * Ho + uy) * Wo + ux) * C + m.v * 8, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += r[j];
        }
        for (int a = 0; a < nh; ++a)
          for (int kh = 0; kh < k; ++kh) {
            const int th = qh[a] + pad - kh;
            if (th < 0 || th % stride != 0 || th / stride >= Ho) continue;
            const int oh = th / stride;
            for (int b = 0; b < nw; ++b)
              for (int kw = 0; kw < k; ++kw) {
                const int tw = qw[b] + pad - kw;
                if (tw < 0 || tw % stride != 0 || tw / stride >= Wo) continue;
                const int ow = tw / stride;
                float g[8], ww[8];
                ld8(dyn + ((int64_t)oh * Wo + ow) * C, g);
                ldf8(wv + (int64_t)(kh * k + kw) * C, ww);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], ww[j], acc[j]);
              }
          }
      }
    }
    const int64_t row = (int64_t)n * H * W + p;
    if (a_pre) {
      float av[8];
      ld8(a_pre + row * C + m.v * 8, av);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= hsw_grad(fmaf(av[j], sc[j], sh[j]));
    }
    st8(dx + row * C + m.v * 8, acc);
  }
}

// ---- depthwise conv: weight gradient  dW[c][kh][kw] += sum_{n,o} dy[n,o,c] * xin[n, R(o*s+k-pad), c] ----
// blockIdx.z = kh; a thread keeps K accumulators x 8 channels.  Output in the parameter's own layout
// (C,1,K,K) fp32, accumulated with atomics (caller zero-fills).
template <typename AT, int K>
__global__ void __launch_bounds__(kT)
dw_wgrad_kernel(const grad_t* __restrict__ dy, const AT* __restrict__ x, float* __restrict__ dw,
                int C, int H, int W, int Ho, int Wo, int stride, int up2) {
  extern __shared__ float s_red[];
  const RowMap m = row_map(C);
  const int n = blockIdx.y, kh = blockIdx.z;
  constexpr int pad = (K - 1) / 2;
  const int Hin = up2 ? 2 * H : H, Win = up2 ? 2 * W : W;
  int64_t p0, p1;
  chunk_range((int64_t)Ho * Wo, p0, p1);
  float q[K][8];
#pragma unroll
  for (int t = 0; t < K; ++t) fill8(q[t], 0.f);
  if (m.on) {
    const AT* xn = x + (int64_t)n * H * W * C + m.v * 8;
    const grad_t* dyn = dy + (int64_t)n * Ho * Wo * C + m.v * 8;
    for (int64_t p = p0 + m.g; p < p1; p += m.groups) {
      const int oh = (int)(p / Wo), ow = (int)(p % Wo);
      float g[8];
      ld8(dyn + p * C, g);
      int ih = reflect_idx(oh * stride + kh - pad, Hin);
      if (up2) ih >>= 1;
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        int iw = reflect_idx(ow * stride + kw - pad, Win);
        if (up2) iw >>= 1;
        float xv[8];
        ld8(xn + ((int64_t)ih * W + iw) * C, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) q[kw][j] = fmaf(g[j], xv[j], q[kw][j]);
      }
    }
    for (int t = 0; t < K; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[(m.g * K + t) * C + m.v * 8 + j] = q[t][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += kT) {
    float s = 0.f;
    for (int gg = 0; gg < m.groups; ++gg) s += s_red[gg * K * C + i];
    const int t = i / C, c = i - t * C;
    atomicAdd(dw + (int64_t)c * K * K + kh * K + t, s);
  }
}

// ---- SELayer backward ----------------------------------------------------------------------------------
// per image: dpre = ds * 1[0 < pre < 1]; dhid = (W2^T dpre) * 1[hid > 0]; g = (W1^T dhid) / HW
constexpr int kSeB = 1024;
// out[o] = sum_r w[r * ldw + o] * x[r] for o < O, r < R, spread over all kSeB threads: thread -> (o, slice of r),
// coalesced along o, the slices combined through `part` in a fixed order (deterministic).  Requires O <= kSeB.
__device__ __forceinline__ void se_matvec_t(const float* __restrict__ w, int ldw, const float* x, int O, int R,
                                            float* part /*[kSeB]*/, float* out /*shared [O]*/) {
  const int Op = (O + 31) & ~31;
  const int groups = kSeB / Op;
  const int o = threadIdx.x % Op, gi = threadIdx.x / Op;
  float a = 0.f;
  if (gi < groups && o < O)
    for (int r = gi; r < R; r += groups) a = fmaf(w[(int64_t)r * ldw + o], x[r], a);
  part[threadIdx.x] = a;
  __syncthreads();
  for (int oo = threadIdx.x; oo < O; oo += kSeB) {
    float t = 0.f;
    for (int gg = 0; gg < groups; ++gg) t += part[gg * Op + oo];
    out[oo] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSeB)
se_bwd_sample_kernel(const float* __restrict__ ds, int ds_stride, const float* __restrict__ pre,
                     const float* __restrict__ hid, const float* __restrict__ w1, const float* __restrict__ w2,
                     float inv_hw, float* __restrict__ dpre, float* __restrict__ dhid, float* __restrict__ g,
                     int C, int S) {
  // One CTA per image, so the kernel is latency: both transposed mat-vecs use all 1024 threads (se_matvec_t)
  // instead of one serial dot product of length C per thread (measured 40 us per launch, independent of the batch).
  extern __shared__ float sm[];  // dpre[C] + dhid[S] + g[C] + part[kSeB]
  float* s_dp = sm;
  float* s_dh = sm + C;
  float* s_g = sm + C + S;
  float* part = sm + 2 * C + S;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += kSeB) {
    const float pv = pre[(int64_t)n * C + c];
    const float v = (pv > 0.f && pv < 1.f) ? ds[(int64_t)n * ds_stride + c] : 0.f;
    s_dp[c] = v;
    dpre[(int64_t)n * C + c] = v;
  }
  __syncthreads();
  se_matvec_t(w2, S, s_dp, S, C, part, s_dh);                  // dhid[j] = sum_c w2[c][j] dpre[c]
  for (int j = threadIdx.x; j < S; j += kSeB) {
    const float a = hid[(int64_t)n * S + j] > 0.f ? s_dh[j] : 0.f;
    s_dh[j] = a;
    dhid[(int64_t)n * S + j] = a;
  }
  __syncthreads();
  se_matvec_t(w1, C, s_dh, C, S, part, s_g);                   // g[c] = sum_j w1[j][c] dhid[j] / HW
  for (int c = threadIdx.x; c < C; c += kSeB) g[(int64_t)n * C + c] = s_g[c] * inv_hw;
}
// weight gradients: dW2[c][j] = sum_n dpre[n][c] hid[n][j]; db2 = sum_n dpre; dW1[j][c] = sum_n dhid[n][j] mean[n][c];
// db1 = sum_n dhid   (mean = pool * inv_hw)
__global__ void se_bwd_weights_kernel(const float* __restrict__ dpre, const float* __restrict__ dhid,
                                      const float* __restrict__ hid, const float* __restrict__ pool, float inv_hw,
                                      float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                      float* __restrict__ db2, int N, int C, int S) {
  const int64_t total = (int64_t)2 * C * S + C + S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    if (i < (int64_t)C * S) {                       // dW2[c][j]
      const int c = (int)(i / S), j = (int)(i % S);
      for (int n = 0; n < N; ++n) a = fmaf(dpre[(int64_t)n * C + c], hid[(int64_t)n * S + j], a);
      dw2[i] = a;
    } else if (i < (int64_t)2 * C * S) {            // dW1[j][c]
      const int64_t r = i - (int64_t)C * S;
      const int j = (int)(r / C), c = (int)(r % C);
      for (int n = 0; n < N; ++n) a = fmaf(dhid[(int64_t)n * S + j], pool[(int64_t)n * C + c] * inv_hw, a);
      dw1[r] = a;
    } else if (i < (int64_t)2 * C * S + C) {
      const int c = (int)(i - (int64_t)2 * C * S);
      for (int n = 0; n < N; ++n) a += dpre[(int64_t)n * C + c];
      db2[c] = a;
    } else {
      const int j = (int)(i - (int64_t)2 * C * S - C);
      for (int n = 0; n < N; ++n) a += dhid[(int64_t)n * S + j];
      db1[j] = a;
    }
  }
}

// ---- stem / head gradients ------------------------------------------------------------------------------
// Both are sums over pixels of outer products between a 16-channel NHWC bf16 vector and a 3-channel NCHW
// fp32 vector, one of them taken at the reflected tap position.  blockIdx.y = tap; one pixel per thread
// per iteration, 48 accumulators, warp-shuffle + shared reduce, atomics into the OIHW gradient.
__device__ __forceinline__ void reduce48_and_add(float (&acc)[48], float* s_part /*[8][48]*/, float* dst,
                                                 int tap, bool a_is_cout /*dst[a][b][tap] else dst[b][a][tap]*/) {
#pragma unroll
  for (int i = 0; i < 48; ++i) {
    float v = acc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    acc[i] = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 48; ++i) s_part[warp * 48 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 48) {
    float s = 0.f;
    for (int wq = 0; wq < kT / 32; ++wq) s += s_part[wq * 48 + threadIdx.x];
    const int a = threadIdx.x / 3, b = threadIdx.x % 3;   // acc index = a * 3 + b
    const int idx = a_is_cout ? (a * 3 + b) * 9 + tap : (b * 16 + a) * 9 + tap;
    atomicAdd(dst + idx, s);
  }
}

// stem: dW[co][ci][tap] = sum_p (dy[p][co] * Hardswish'(z[p][co])) * img[ci][R(p + tap)]   (Cout == 16)
template <typename AT>
__global__ void __launch_bounds__(kT)
stem_wgrad_kernel(const grad_t* __restrict__ dy, const AT* __restrict__ z,
                  const float* __restrict__ img, float* __restrict__ dw, int N, int H, int W) {
  __shared__ float s_part[(kT / 32) * 48];
  const int tap = blockIdx.y, kh = tap / 3, kw = tap % 3;
  const int64_t total = (int64_t)N * H * W;
  float acc[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  for (int64_t pix = (int64_t)blockIdx.x * kT + threadIdx.x; pix < total; pix += (int64_t)gridDim.x * kT) {
    const int xw = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
    float d[16], zz[16];
    ld8(dy + pix * 16, *reinterpret_cast<float(*)[8]>(d));
    ld8(dy + pix * 16 + 8, *reinterpret_cast<float(*)[8]>(d + 8));
    ld8(z + pix * 16, *reinterpret_cast<float(*)[8]>(zz));
    ld8(z + pix * 16 + 8, *reinterpret_cast<float(*)[8]>(zz + 8));
    const int ih = reflect_idx(h + kh - 1, H), iw = reflect_idx(xw + kw - 1, W);
    float iv[3];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) iv[ci] = __ldg(img + (((int64_t)n * 3 + ci) * H + ih) * W + iw);
#pragma unroll
    for (int co = 0; co < 16; ++co) {
      const float dz = d[co] * hsw_grad(zz[co]);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) acc[co * 3 + ci] = fmaf(dz, iv[ci], acc[co * 3 + ci]);
    }
  }
  reduce48_and_add(acc, s_part, dw, tap, true);
}

// head: dW[co][ci][tap] = sum_p dY[co][p] * x[R(p + tap)][ci];  db[co] = sum_p dY[co][p]   (Cin == 16, Cout == 3)
template <typename AT>
__global__ void __launch_bounds__(kT)
head_wgrad_kernel(const float* __restrict__ dY, const AT* __restrict__ x, float* __restrict__ dw,
                  float* __restrict__ db, int N, int H, int W) {
  __shared__ float s_part[(kT / 32) * 48];
  const int tap = blockIdx.y, kh = tap / 3, kw = tap % 3;
  const int64_t total = (int64_t)N * H * W;
  float acc[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  float bs[3] = {0.f, 0.f, 0.f};
  for (int64_t pix = (int64_t)blockIdx.x * kT + threadIdx.x; pix < total; pix += (int64_t)gridDim.x * kT) {
    const int xw = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
    float g[3];
#pragma unroll
    for (int co = 0; co < 3; ++co) g[co] = __ldg(dY + (((int64_t)n * 3 + co) * H + h) * W + xw);
    const int ih = reflect_idx(h + kh - 1, H), iw = reflect_idx(xw + kw - 1, W);
    const AT* xp = x + (((int64_t)n * H + ih) * W + iw) * 16;
    float xv[16];
    ld8(xp, *reinterpret_cast<float(*)[8]>(xv));
    ld8(xp + 8, *reinterpret_cast<float(*)[8]>(xv + 8));
#pragma unroll
    for (int ci = 0; ci < 16; ++ci)
#pragma unroll
      for (int co = 0; co < 3; ++co) acc[ci * 3 + co] = fmaf(xv[ci], g[co], acc[ci * 3 + co]);
    if (tap == 0) { bs[0] += g[0]; bs[1] += g[1]; bs[2] += g[2]; }
  }
  reduce48_and_add(acc, s_part, dw, tap, false);
  if (tap == 0 && db) {
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      float v = bs[co];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31) == 0) atomicAdd(db + co, v);
    }
  }
}

// head data gradient: dx[p][ci] = sum_{q in Q(p)} sum_{tap, co} W[co][ci][tap] * dY[co][q - tap + 1]
// (q runs over the padded-grid positions that reflect onto p).  NCHW fp32 in, NHWC bf16 out, Cin = 16.
__global__ void __launch_bounds__(128)
head_dgrad_kernel(const float* __restrict__ dY, const float* __restrict__ w /*OIHW [3][16][3][3]*/,
                  grad_t* __restrict__ dx, int N, int H, int W) {
  __shared__ float s_w[9][3][16];
  for (int i = threadIdx.x; i < 9 * 3 * 16; i += 128) {
    const int ci = i % 16, co = (i / 16) % 3, t = i / 48;
    s_w[t][co][ci] = w[((int64_t)co * 16 + ci) * 9 + t];
  }
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (pix >= (int64_t)N * H * W) return;
  const int xw = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  int qh[3], qw[3];
  const int nh = reflect_cands(h, H, 1, qh), nw = reflect_cands(xw, W, 1, qw);
  for (int a = 0; a < nh; ++a)
    for (int kh = 0; kh < 3; ++kh) {
      const int oh = qh[a] + 1 - kh;
      if (oh < 0 || oh >= H) continue;
      for (int b = 0; b < nw; ++b)
        for (int kw = 0; kw < 3; ++kw) {
          const int ow = qw[b] + 1 - kw;
          if (ow < 0 || ow >= W) continue;
#pragma unroll
          for (int co = 0; co < 3; ++co) {
            const float g = __ldg(dY + (((int64_t)n * 3 + co) * H + oh) * W + ow);
#pragma unroll
            for (int ci = 0; ci < 16; ++ci) acc[ci] = fmaf(g, s_w[kh * 3 + kw][co][ci], acc[ci]);
          }
        }
    }
  float lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { lo[i] = acc[i]; hi[i] = acc[8 + i]; }
  st8(dx + pix * 16, lo);
  st8(dx + pix * 16 + 8, hi);
}

// stem data gradient: dimg[n][ci][p] = sum_{q in Q(p)} sum_{tap, co} W[co][ci][tap] * (dy * Hardswish'(z))[q - tap + 1][co]
// (q runs over the padded-grid positions that reflect onto p; conv_3x3_bn, mobilenetv2.py:38-43, Cout == 16).
// NHWC gradient / activation in, NCHW fp32 out.  Only flows that differentiate through the Encoder's INPUT need it.
template <typename AT>
__global__ void __launch_bounds__(128)
stem_dgrad_kernel(const grad_t* __restrict__ dy, const AT* __restrict__ z, const float* __restrict__ w /*OIHW [16][3][3][3]*/,
                  float* __restrict__ dimg, int N, int H, int W) {
  __shared__ float s_w[9][3][16];
  for (int i = threadIdx.x; i < 9 * 3 * 16; i += 128) {
    const int co = i % 16, ci = (i / 16) % 3, t = i / 48;
    s_w[t][ci][co] = w[((int64_t)co * 3 + ci) * 9 + t];
  }
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (pix >= (int64_t)N * H * W) return;
  const int xw = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
  float acc[3] = {0.f, 0.f, 0.f};
  int qh[3], qw[3];
  const int nh = reflect_cands(h, H, 1, qh), nw = reflect_cands(xw, W, 1, qw);
  for (int a = 0; a < nh; ++a)
    for (int kh = 0; kh < 3; ++kh) {
      const int oh = qh[a] + 1 - kh;
      if (oh < 0 || oh >= H) continue;
      for (int b = 0; b < nw; ++b)
        for (int kw = 0; kw < 3; ++kw) {
          const int ow = qw[b] + 1 - kw;
          if (ow < 0 || ow >= W) continue;
          const int64_t o = (((int64_t)n * H + oh) * W + ow) * 16;
          float d[16], zz[16];
          ld8(dy + o, *reinterpret_cast<float(*)[8]>(d));
          ld8(dy + o + 8, *reinterpret_cast<float(*)[8]>(d + 8));
          ld8(z + o, *reinterpret_cast<float(*)[8]>(zz));
          ld8(z + o + 8, *reinterpret_cast<float(*)[8]>(zz + 8));
#pragma unroll
          for (int co = 0; co < 16; ++co) {
            const float dz = d[co] * hsw_grad(zz[co]);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) acc[ci] = fmaf(dz, s_w[kh * 3 + kw][ci][co], acc[ci]);
          }
        }
    }
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) dimg[(((int64_t)n * 3 + ci) * H + h) * W + xw] = acc[ci];
}

// Hardtanh(0,1) backward of the exporting decoder head (models.py:304, 315-316): dx = dy where 0 < y < 1 (y = the
// clamped output), else 0.
__global__ void hardtanh01_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                      int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = y[i];
    dx[i] = (v > 0.f && v < 1.f) ? dy[i] : 0.f;
  }
}

// ---- weight preparation: fp32 [R][Cc] -> 0: bf16, 1: bf16 transposed, 2: fp32 transposed, 3: fp16 (forward GEMMs) ----
template <typename AT>
__global__ void prep_weight_kernel(const float* __restrict__ w, void* __restrict__ out, int R, int Cc, int mode) {
  const int64_t total = (int64_t)R * Cc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Cc), c = (int)(i % Cc);
    const float v = w[i];
    if (mode == 0) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
    else if (mode == 1) reinterpret_cast<__nv_bfloat16*>(out)[(int64_t)c * R + r] = __float2bfloat16_rn(v);
    else if (mode == 2) reinterpret_cast<float*>(out)[(int64_t)c * R + r] = v;
    else H16<AT>::store1(out, i, v);
  }
}

// ---- all GEMM / depthwise forms of ONE DepthWiseConv block's weights in one launch (5 prep_weight launches -> 1):
// pointwise weights w1 [hid][inp], w2 [oup][hid]: forward form (activation format, same layout) + data-gradient form
// (bf16, transposed); depthwise weight wd [hid][k*k]: fp32 transposed [k*k][hid].
struct PrepSeg {
  const float* w;
  void* out_f;      // pointwise: activation-format copy; depthwise: fp32 transposed
  void* out_t;      // pointwise: bf16 transposed; depthwise: unused
  int R, C;
};
struct PrepBlockArgs {
  PrepSeg seg[3];   // w1 (may be empty: R = 0), wd, w2
  int64_t off[4];
};
template <typename AT>
__global__ void prep_block_weights_kernel(const PrepBlockArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.off[3]; i += (int64_t)gridDim.x * blockDim.x) {
    const int sgi = i >= a.off[2] ? 2 : (i >= a.off[1] ? 1 : 0);
    const PrepSeg& g = a.seg[sgi];
    const int64_t j = i - a.off[sgi];
    const int r = (int)(j / g.C), c = (int)(j - (int64_t)r * g.C);
    const float v = g.w[j];
    if (sgi == 1) {
      reinterpret_cast<float*>(g.out_f)[(int64_t)c * g.R + r] = v;
    } else {
      H16<AT>::store1(g.out_f, j, v);
      reinterpret_cast<__nv_bfloat16*>(g.out_t)[(int64_t)c * g.R + r] = __float2bfloat16_rn(v);
    }
  }
}

// ---- NCHW fp32 -> NHWC bf16 (row stride ld) --------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src,
                                                           uint16_t* __restrict__ dst, int ld, int C,
                                                           int64_t HW, int f16) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int cl = ty; cl < 32; cl += 8) {
    const int64_t p = p0 + tx;
    tile[cl][tx] = (c0 + cl < C && p < HW) ? src[((int64_t)n * C + c0 + cl) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int pl = ty; pl < 32; pl += 8) {
    const int64_t p = p0 + pl;
    if (p < HW && c0 + tx < C) dst[((int64_t)n * HW + p) * ld + c0 + tx] = (uint16_t)(pk2_dt(tile[tx][pl], 0.f, f16) & 0xffffu);
  }
}

// ---- fp16 activation -> bf16 copy: the operand of a weight-gradient GEMM (its other operand is a bf16 gradient and
// the tensor cores take one format per instruction).  rows x C elements, row strides ld_x / ld_out.
template <typename AT>
__global__ void __launch_bounds__(256) cvt_act_to_grad_kernel(const AT* __restrict__ x, int64_t ld_x,
                                                              grad_t* __restrict__ out, int64_t ld_out, int64_t rows,
                                                              int cv) {
  const int64_t total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int v = (int)(i - r * cv);
    float f[8];
    H16<AT>::unpack(ld_stream_u4(x + r * ld_x + v * 8), f);
    st8(out + r * ld_out + v * 8, f);
  }
}

static int pick_chunks(int N, int64_t HW, int C) {
  const int groups = kT / (C / 8);
  int64_t chunks = (8 * 148 + N - 1) / N;
  const int64_t cap = HW / ((int64_t)groups * 2);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}
static bool chan_ok(int C) { return C > 0 && C % 8 == 0 && C / 8 <= kT; }

}  // namespace ast

using namespace ast;
typedef __nv_bfloat16 bf16;
// dw_tiled.cu
int dw_tiled_dgrad(const void* dy, const float* w, const void* a_pre, const float* stat, const void* dres, void* dx,
                   int N, int C, int H, int W, int k, cudaStream_t s);
int dw_tiled_wgrad(const void* dy, const void* x, float* dw, int N, int C, int H, int W, int k, int up2,
                   cudaStream_t s);
bool dw_force_direct();
#define GR(p) reinterpret_cast<grad_t*>(p)
#define CGR(p) reinterpret_cast<const grad_t*>(p)
#define AC(p) reinterpret_cast<AT*>(p)          /* inside AST_ACT_DISPATCH */
#define CAC(p) reinterpret_cast<const AT*>(p)

extern "C" int ast_bn_stats(const void* x, int ld, double* sums, int N, int C, int64_t HW, void* stream) {
  if (!x || !sums || N <= 0 || HW <= 0) return AST_E_BADARG;
  if (!chan_ok(C) || ld % 8 != 0 || ld < C || N > 65535) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  AST_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s));
  const int groups = kT / (C / 8);
  AST_ACT_DISPATCH(bn_stats_kernel<AT><<<dim3(pick_chunks(N, HW, C), N), kT, (size_t)groups * 2 * C * 4, s>>>(CAC(x), ld, sums, C, HW));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, float* stat,
                               int C, long long* num_batches_tracked, void* stream) {
  if (!sums || !gamma || !beta || !stat || C <= 0 || count <= 0) return AST_E_BADARG;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, count, gamma, beta, running_mean,
                                                                        running_var, momentum, eps, stat, C, num_batches_tracked);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_affine_act(const void* x, int ld_x, const float* sc, const float* sh, int act, const float* se,
                              const void* res, int ld_res, void* out, int ld_out, float* pool, int N, int C,
                              int64_t HW, void* stream) {
  if (!x || (!out && !pool) || N <= 0 || HW <= 0 || (sc && !sh)) return AST_E_BADARG;
  if (!chan_ok(C) || ld_x % 8 != 0 || ld_x < C || (out && (ld_out % 8 != 0 || ld_out < C)) ||
      (res && (ld_res % 8 != 0 || ld_res < C)) || N > 65535)
    return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (pool) AST_CUDA(cudaMemsetAsync(pool, 0, sizeof(float) * (size_t)N * C, s));
  const int groups = kT / (C / 8);
  AST_ACT_DISPATCH(affine_act_kernel<AT><<<dim3(pick_chunks(N, HW, C), N), kT, pool ? (size_t)groups * C * 4 : 0, s>>>(
      CAC(x), ld_x, sc, sh, act, se, CAC(res), ld_res, AC(out), ld_out, pool, C, HW));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_dw_bwd_reduce(const void* du, const void* a, const float* stat, float* out, int N, int C,
                                 int64_t HW, void* stream) {
  if (!du || !a || !out || N <= 0 || HW <= 0) return AST_E_BADARG;
  if (!chan_ok(C) || N > 65535) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  AST_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * 5 * C, s));
  const int groups = kT / (C / 8);
  const size_t smem = (size_t)groups * 5 * C * 4;   // <= 40 KB
  if (stat)
    AST_ACT_DISPATCH(dw_bwd_reduce_kernel<AT, true><<<dim3(pick_chunks(N, HW, C), N), kT, smem, s>>>(CGR(du), CAC(a), stat, out, C, HW));
  else
    AST_ACT_DISPATCH(dw_bwd_reduce_kernel<AT, false><<<dim3(pick_chunks(N, HW, C), N), kT, smem / 5, s>>>(CGR(du), CAC(a), stat, out, C, HW));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_se_bn_combine(const float* T, const float* s, const float* g, float* dgamma, float* dbeta,
                                 float* coef, int N, int C, double count, void* stream) {
  if (!T || !s || !g || !dgamma || !dbeta || !coef || N <= 0 || C <= 0 || count <= 0) return AST_E_BADARG;
  se_bn_combine_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(T, s, g, dgamma, dbeta, coef, N, C,
                                                                          (float)(1.0 / count));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_dw_bwd_apply(const void* du, const void* a, const float* s, const float* g, const float* stat,
                                const float* coef, void* da, int N, int C, int64_t HW, void* stream) {
  if (!du || !a || !s || !g || !da || N <= 0 || HW <= 0 || (stat && !coef)) return AST_E_BADARG;
  if (!chan_ok(C) || N > 65535) return AST_E_SHAPE;
  if (stat)
    AST_ACT_DISPATCH(dw_bwd_apply_kernel<AT, true><<<dim3(pick_chunks(N, HW, C), N), kT, 0, (cudaStream_t)stream>>>(
        CGR(du), CAC(a), s, g, stat, coef, GR(da), C, HW));
  else
    AST_ACT_DISPATCH(dw_bwd_apply_kernel<AT, false><<<dim3(pick_chunks(N, HW, C), N), kT, 0, (cudaStream_t)stream>>>(
        CGR(du), CAC(a), s, g, stat, coef, GR(da), C, HW));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_bn_bwd_reduce(const void* dy, int ld_dy, const void* a, int ld_a, const float* stat,
                                 double* sums, int N, int C, int64_t HW, void* stream) {
  if (!dy || !a || !stat || !sums || N <= 0 || HW <= 0) return AST_E_BADARG;
  if (!chan_ok(C) || ld_dy % 8 != 0 || ld_dy < C || ld_a % 8 != 0 || ld_a < C || N > 65535) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  AST_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s));
  const int groups = kT / (C / 8);
  AST_ACT_DISPATCH(bn_bwd_reduce_kernel<AT><<<dim3(pick_chunks(N, HW, C), N), kT, (size_t)groups * 2 * C * 4, s>>>(
      CGR(dy), ld_dy, CAC(a), ld_a, stat, sums, C, HW));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_bn_bwd_finalize(const double* sums, double count, float* dgamma, float* dbeta, float* coef,
                                   int C, void* stream) {
  if (!sums || !dgamma || !dbeta || !coef || C <= 0 || count <= 0) return AST_E_BADARG;
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, 1.0 / count, dgamma, dbeta, coef,
                                                                            C);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_bn_bwd_apply(const void* dy, int ld_dy, const void* a, int ld_a, const float* stat,
                                const float* coef, void* da, int N, int C, int64_t HW, void* stream) {
  if (!dy || !a || !stat || !coef || !da || N <= 0 || HW <= 0) return AST_E_BADARG;
  if (!chan_ok(C) || ld_dy % 8 != 0 || ld_dy < C || ld_a % 8 != 0 || ld_a < C || N > 65535) return AST_E_SHAPE;
  AST_ACT_DISPATCH(bn_bwd_apply_kernel<AT><<<dim3(pick_chunks(N, HW, C), N), kT, 0, (cudaStream_t)stream>>>(
      CGR(dy), ld_dy, CAC(a), ld_a, stat, coef, GR(da), C, HW));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_dw_conv_dgrad(const void* dy, const float* w, const void* a_pre, const float* stat,
                                 const void* dres, void* dx, int N, int C, int H, int W, int k, int stride,
                                 int up2, void* stream) {
  if (!dy || !w || !dx || N <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if (!chan_ok(C) || (k != 3 && k != 5) || (stride != 1 && stride != 2) || N > 65535) return AST_E_SHAPE;
  const int Hin = up2 ? 2 * H : H, Win = up2 ? 2 * W : W, pad = (k - 1) / 2;
  if (Hin <= pad || Win <= pad) return AST_E_SHAPE;
  const int Ho = (Hin + 2 * pad - k) / stride + 1, Wo = (Win + 2 * pad - k) / stride + 1;
  if (dres && (stride != 1 || a_pre)) return AST_E_SHAPE;
  if (stride == 1 && !up2 && !dw_force_direct()) {
    const int r = dw_tiled_dgrad(dy, w, a_pre, stat, dres, dx, N, C, H, W, k, (cudaStream_t)stream);
    if (r != AST_E_SHAPE) return r;
  }
  AST_ACT_DISPATCH(dw_dgrad_kernel<AT><<<dim3(pick_chunks(N, (int64_t)H * W, C), N), kT, 0, (cudaStream_t)stream>>>(
      CGR(dy), w, CAC(a_pre), stat, CGR(dres), GR(dx), C, H, W, Ho, Wo, k, stride, up2));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_dw_conv_wgrad(const void* dy, const void* x, float* dw, int N, int C, int H, int W, int k,
                                 int stride, int up2, void* stream) {
  if (!dy || !x || !dw || N <= 0 || H <= 0 || W <= 0) return AST_E_BADARG;
  if (!chan_ok(C) || (k != 3 && k != 5) || (stride != 1 && stride != 2) || N > 65535) return AST_E_SHAPE;
  const int Hin = up2 ? 2 * H : H, Win = up2 ? 2 * W : W, pad = (k - 1) / 2;
  if (Hin <= pad || Win <= pad) return AST_E_SHAPE;
  const int Ho = (Hin + 2 * pad - k) / stride + 1, Wo = (Win + 2 * pad - k) / stride + 1;
  if (stride == 1 && !dw_force_direct()) {
    const int r = dw_tiled_wgrad(dy, x, dw, N, C, H, W, k, up2, (cudaStream_t)stream);
    if (r != AST_E_SHAPE) return r;
  }
  const int groups = kT / (C / 8);
  const size_t smem = (size_t)groups * k * C * 4;   // <= 40 KB
  int chunks = pick_chunks(N, (int64_t)Ho * Wo, C);
  chunks = (chunks + k - 1) / k;                     // blockIdx.z multiplies the grid by k
  if (chunks < 1) chunks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (k == 3)
    AST_ACT_DISPATCH(dw_wgrad_kernel<AT, 3><<<dim3(chunks, N, 3), kT, smem, s>>>(CGR(dy), CAC(x), dw, C, H, W, Ho, Wo, stride, up2));
  else
    AST_ACT_DISPATCH(dw_wgrad_kernel<AT, 5><<<dim3(chunks, N, 5), kT, smem, s>>>(CGR(dy), CAC(x), dw, C, H, W, Ho, Wo, stride, up2));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_se_bwd(const float* ds, int ds_stride, const float* pre, const float* hid, const float* pool,
                          float inv_hw, const float* w1, const float* w2, float* dpre, float* dhid, float* g,
                          float* dw1, float* db1, float* dw2, float* db2, int N, int C, int S, void* stream) {
  if (!ds || !pre || !hid || !pool || !w1 || !w2 || !dpre || !dhid || !g || !dw1 || !db1 || !dw2 || !db2 ||
      N <= 0 || C <= 0 || S <= 0)
    return AST_E_BADARG;
  const size_t smem = (size_t)(2 * C + S + kSeB) * sizeof(float);
  if (smem > 48 * 1024 || C > kSeB || S > kSeB) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  se_bwd_sample_kernel<<<N, kSeB, smem, s>>>(ds, ds_stride, pre, hid, w1, w2, inv_hw, dpre, dhid, g, C, S);
  AST_CHECK_LAUNCH();
  const int64_t total = (int64_t)2 * C * S + C + S;
  int64_t nb = (total + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  se_bwd_weights_kernel<<<(unsigned)nb, 256, 0, s>>>(dpre, dhid, hid, pool, inv_hw, dw1, db1, dw2, db2, N, C, S);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_stem_wgrad(const void* dy, const void* z, const float* img, float* dw, int N, int H, int W,
                              int Cout, void* stream) {
  if (!dy || !z || !img || !dw || N <= 0 || H < 2 || W < 2) return AST_E_BADARG;
  if (Cout != 16) return AST_E_SHAPE;
  int64_t nb = ((int64_t)N * H * W + kT - 1) / kT;
  if (nb > 148 * 2) nb = 148 * 2;
  AST_ACT_DISPATCH(stem_wgrad_kernel<AT><<<dim3((unsigned)nb, 9), kT, 0, (cudaStream_t)stream>>>(CGR(dy), CAC(z), img, dw, N, H, W));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_stem_dgrad(const void* dy, const void* z, const float* w, float* dimg, int N, int H, int W,
                              int Cout, void* stream) {
  if (!dy || !z || !w || !dimg || N <= 0 || H < 2 || W < 2) return AST_E_BADARG;
  if (Cout != 16) return AST_E_SHAPE;
  const int64_t nb = ((int64_t)N * H * W + 127) / 128;
  if (nb >= 0x7fffffffLL) return AST_E_SHAPE;
  AST_ACT_DISPATCH(stem_dgrad_kernel<AT><<<(unsigned)nb, 128, 0, (cudaStream_t)stream>>>(CGR(dy), CAC(z), w, dimg, N, H, W));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_hardtanh01_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  if (!dy || !y || !dx || n <= 0) return AST_E_BADARG;
  int64_t nb = (n + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  hardtanh01_bwd_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_head_wgrad(const float* dY, const void* x, float* dw, float* db, int N, int H, int W, int Cin,
                              int Cout, void* stream) {
  if (!dY || !x || !dw || N <= 0 || H < 2 || W < 2) return AST_E_BADARG;
  if (Cin != 16 || Cout != 3) return AST_E_SHAPE;
  int64_t nb = ((int64_t)N * H * W + kT - 1) / kT;
  if (nb > 148 * 2) nb = 148 * 2;
  AST_ACT_DISPATCH(head_wgrad_kernel<AT><<<dim3((unsigned)nb, 9), kT, 0, (cudaStream_t)stream>>>(dY, CAC(x), dw, db, N, H, W));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_head_dgrad(const float* dY, const float* w, void* dx, int N, int H, int W, int Cin, int Cout,
                              void* stream) {
  if (!dY || !w || !dx || N <= 0 || H < 2 || W < 2) return AST_E_BADARG;
  if (Cin != 16 || Cout != 3) return AST_E_SHAPE;
  const int64_t nb = ((int64_t)N * H * W + 127) / 128;
  if (nb >= 0x7fffffffLL) return AST_E_SHAPE;
  head_dgrad_kernel<<<(unsigned)nb, 128, 0, (cudaStream_t)stream>>>(dY, w, GR(dx), N, H, W);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_cvt_f16_to_bf16(const void* x, int64_t ld_x, void* out, int64_t ld_out, int64_t rows, int C,
                                   void* stream) {
  if (!x || !out || rows <= 0 || C <= 0) return AST_E_BADARG;
  if (C % 8 != 0 || ld_x % 8 != 0 || ld_out % 8 != 0 || ld_x < C || ld_out < C) return AST_E_SHAPE;
  if (!aligned16(x) || !aligned16(out)) return AST_E_ALIGN;
  const int64_t total = rows * (C / 8);
  int64_t nb = (total + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  if (ast::act_format() != AST_DT_F16) return AST_E_BADARG;   // bf16 activations need no conversion
  { using AT = __half; cvt_act_to_grad_kernel<AT><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(CAC(x), ld_x, GR(out), ld_out, rows, C / 8); }
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_prep_weight(const float* w, void* out, int R, int Cc, int mode, void* stream) {
  if (!w || !out || R <= 0 || Cc <= 0 || mode < 0 || mode > 3) return AST_E_BADARG;
  int64_t nb = ((int64_t)R * Cc + 255) / 256;
  if (nb > 148 * 4) nb = 148 * 4;
  AST_ACT_DISPATCH(prep_weight_kernel<AT><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(w, out, R, Cc, mode));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_prep_block_weights(const float* w1, int R1, int C1, void* w1_f, void* w1_t, const float* wd, int Rd,
                                      int Cd, float* wd_t, const float* w2, int R2, int C2, void* w2_f, void* w2_t,
                                      void* stream) {
  if (!wd || !wd_t || !w2 || !w2_f || !w2_t || Rd <= 0 || Cd <= 0 || R2 <= 0 || C2 <= 0) return AST_E_BADARG;
  if (w1 && (!w1_f || !w1_t || R1 <= 0 || C1 <= 0)) return AST_E_BADARG;
  PrepBlockArgs a = {};
  a.seg[0] = PrepSeg{w1, w1_f, w1_t, w1 ? R1 : 0, w1 ? C1 : 1};
  a.seg[1] = PrepSeg{wd, wd_t, nullptr, Rd, Cd};
  a.seg[2] = PrepSeg{w2, w2_f, w2_t, R2, C2};
  a.off[0] = 0;
  a.off[1] = w1 ? (int64_t)R1 * C1 : 0;
  a.off[2] = a.off[1] + (int64_t)Rd * Cd;
  a.off[3] = a.off[2] + (int64_t)R2 * C2;
  int64_t nb = (a.off[3] + 255) / 256;
  if (nb > 148 * 4) nb = 148 * 4;
  AST_ACT_DISPATCH(prep_block_weights_kernel<AT><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(a));
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_nchw_to_nhwc(const float* x, void* out, int ld, int N, int C, int64_t HW, int dtype, void* stream) {
  if (!x || !out || N <= 0 || C <= 0 || HW <= 0 || ld < C || (dtype != AST_DT_BF16 && dtype != AST_DT_F16)) return AST_E_BADARG;
  if (N > 65535 || (C + 31) / 32 > 65535 || (HW + 31) / 32 >= 0x7fffffffLL) return AST_E_SHAPE;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<uint16_t*>(out), ld, C, HW,
                                                              dtype == AST_DT_F16);
  AST_CHECK_LAUNCH();
  return 0;
}
