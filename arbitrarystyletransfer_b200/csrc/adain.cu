// K1: fused AdaIN / channel statistics / mean-variance-norm kernels (HBM-bound family).
//
// Reference arithmetic being replaced (paths relative to /root/reference):
//   channel_stats            model_util.py:3-8   mean + UNBIASED std over HxW, no eps
//   AdaIN.forward            models.py:43-51     (c - mu_c)/sigma_c * mu_s + sigma_s  (swapped!)
//   alpha blend              models.py:471       alpha*t + (1-alpha)*content
//   calc_mean_std / MVN      models.py:54-68     sqrt(var_unbiased + 1e-5); (x - mu)/std
//
// Design (B200): one CTA -- or one thread-block cluster of up to 8 CTAs for long rows -- owns one
// (n,c) row.  The content segment is loaded once with 128-bit streaming loads and stays in
// registers while the statistics are reduced (per-thread Welford -> Chan merge over warp
// shuffles -> shared memory -> DSMEM across the cluster), then the affine + blend is applied from
// registers and streamed out.  HBM traffic = the algorithmic (1 + K + 1) passes.  Rows that do
// not fit (or are not 16-byte aligned) use a two-pass kernel whose second pass re-reads L2.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace ast {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / kWarp;
constexpr int kMaxQ = 1 + AST_MAX_STYLES;  // content + styles

struct AdainArgs {
  const void* content;
  void* out;
  float* stats;
  const void* styles[AST_MAX_STYLES];
  int64_t style_hw[AST_MAX_STYLES];
  float style_w[AST_MAX_STYLES];
  int K;
  int64_t HW;
  float alpha, eps;
  unsigned flags;
  int identity_affine;  // MVN mode: scale 1, shift 0, no styles
};

// Block-level merge of Q moment sets.  Result valid in every thread.
template <int MAXQ>
__device__ __forceinline__ void block_merge(Moments (&m)[MAXQ], int Q, Moments (*smem)[MAXQ]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    if (q < Q) {
      Moments r = moments_warp_reduce(m[q]);
      if (lane == 0) smem[warp][q] = r;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    if (q < Q) {
      Moments r = smem[0][q];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) r = moments_merge(r, smem[w][q]);
      m[q] = r;
    }
  }
}

// Streaming moments over vectors [v0, v1) of a 16B-aligned row, strided by the block (shifted sums, see
// ShiftedLanes; the shift is the segment's first element, one broadcast load).
template <bool BF16>
__device__ __forceinline__ Moments stream_moments_vec(const void* row, int64_t v0, int64_t v1) {
  using VT = Vec16<BF16>;
  constexpr int V = VT::V;
  constexpr int U = 8;   // 8 x 16 B in flight per thread: the long-row paths are latency-bound otherwise
  ShiftedLanes<V> w;
  w.init(VT::load1(row, v0 * V));
  const uint4* p = reinterpret_cast<const uint4*>(row);
  int64_t i = v0 + threadIdx.x;
  for (; i + (U - 1) * kThreads < v1; i += U * kThreads) {
    uint4 u[U];
#pragma unroll
    for (int j = 0; j < U; ++j) u[j] = ld_stream_u4(p + i + j * kThreads);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      float x[V];
      VT::unpack(u[j], x);
      w.push(x);
    }
  }
  for (; i < v1; i += kThreads) {
    float x[V];
    VT::unpack(ld_stream_u4(p + i), x);
    w.push(x);
  }
  return w.fold();
}

template <bool BF16>
__device__ __forceinline__ Moments stream_moments_scalar(const void* base, int64_t e0, int64_t e1) {
  WelfordLanes<1> w;
  w.init();
  for (int64_t i = e0 + threadIdx.x; i < e1; i += kThreads) {
    float x[1] = {Vec16<BF16>::load1(base, i)};
    w.push(x);
  }
  return w.fold();
}

__device__ __forceinline__ float std_from(const Moments& m, float eps, unsigned flags) {
  float denom = (flags & AST_F_BIASED) ? m.n : (m.n - 1.f);
  return sqrtf(m.m2 / denom + eps);  // n == 1, unbiased: 0/0 = NaN, as torch.std
}

// Turn merged moments into the per-row affine (mu_c, 1/sigma_c, A, B) and optionally dump stats.
__device__ __forceinline__ void finish_stats(const AdainArgs& a, const Moments* m,
                                             int64_t row, bool writer, float& mu, float& rsig,
                                             float& A, float& B) {
  mu = m[0].mean;
  float sig = std_from(m[0], a.eps, a.flags);
  rsig = 1.f / sig;
  A = 0.f;
  B = 0.f;
  float* st = a.stats ? a.stats + row * (2 + 2 * a.K) : nullptr;
  if (st && writer) { st[0] = mu; st[1] = sig; }
  if (a.identity_affine) { A = 1.f; B = 0.f; return; }
  for (int k = 0; k < a.K; ++k) {
    float smu = m[1 + k].mean;
    float ssig = std_from(m[1 + k], a.eps, a.flags);
    if (st && writer) { st[2 + 2 * k] = smu; st[3 + 2 * k] = ssig; }
    if (a.flags & AST_F_CANONICAL) {
      A = fmaf(a.style_w[k], ssig, A);
      B = fmaf(a.style_w[k], smu, B);
    } else {  // reference: models.py:44 binds style_std := mean(style), style_mean := std(style)
      A = fmaf(a.style_w[k], smu, A);
      B = fmaf(a.style_w[k], ssig, B);
    }
  }
}

__device__ __forceinline__ float apply_affine(float x, float mu, float rsig, float A, float B,
                                              float alpha, bool blend) {
  float t = (x - mu) * rsig;       // models.py:47
  float y = fmaf(t, A, B);         // models.py:50
  if (blend) y = fmaf(alpha, y, (1.f - alpha) * x);  // models.py:471
  return y;
}

// ---- register-cached, cluster-split kernel -------------------------------------------------
// grid = rows * CS CTAs, cluster = CS.  Each CTA owns up to R*kThreads 16-byte vectors of the row.
// Register budget is what bounds HBM throughput here (bytes in flight per SM = resident CTAs x
// 2 x R x 4 KB), so the merged moments of the 1 + K quantities live in shared memory, not registers,
// and R <= 4 is compiled for four resident CTAs per SM.

// CTA-level Chan merge of one quantity: result lands in *dst (shared) for all threads to read
// after the next __syncthreads().
__device__ __forceinline__ void cta_merge_to(Moments m, Moments* s_warp, Moments* dst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  m = moments_warp_reduce(m);
  __syncthreads();  // s_warp may still be read from the previous quantity
  if (lane == 0) s_warp[warp] = m;
  __syncthreads();
  if (warp == 0) {
    Moments r = lane < kWarps ? s_warp[lane] : Moments{0.f, 0.f, 0.f};
#pragma unroll
    for (int off = kWarps / 2; off > 0; off >>= 1) {
      Moments o;
      o.n = __shfl_xor_sync(0xffffffffu, r.n, off);
      o.mean = __shfl_xor_sync(0xffffffffu, r.mean, off);
      o.m2 = __shfl_xor_sync(0xffffffffu, r.m2, off);
      r = moments_merge(r, o);
    }
    if (lane == 0) *dst = r;
  }
}

template <bool BF16, int R>
__global__ void __launch_bounds__(kThreads, (R <= 2 ? 4 : (R <= 4 ? 3 : 4))) adain_cached_kernel(const AdainArgs a) {
  using VT = Vec16<BF16>;
  constexpr int V = VT::V;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned CS = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const int64_t row = blockIdx.x / CS;

  __shared__ Moments s_warp[kWarps];
  __shared__ Moments s_q[kMaxQ];    // this CTA's moments per quantity
  __shared__ Moments s_fin[kMaxQ];  // cluster-merged

  const int Q = a.identity_affine ? 1 : 1 + a.K;
  const int64_t nvec = a.HW / V;
  const int64_t seg = (nvec + CS - 1) / CS;
  const int64_t v0 = rank * seg;
  const int64_t v1 = (v0 + seg < nvec) ? v0 + seg : nvec;

  const uint4* crow =
      reinterpret_cast<const uint4*>(reinterpret_cast<const typename VT::elem*>(a.content) + row * a.HW);
  // R <= 4: the content segment lives in registers.  R == 8 (long rows, cluster-split): it is staged in shared
  // memory with asynchronous 16-byte copies instead -- no registers are held while the style rows stream, which
  // doubles the CTAs per SM of this latency-bound path (each thread later reads back only what it copied).
  constexpr bool SM = (R == 8);
  extern __shared__ uint4 s_cache[];
  uint4 cache[SM ? 1 : R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
    if (i < v1) {
      if (SM) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(s_cache + threadIdx.x + j * kThreads);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(crow + i) : "memory");
      } else {
        cache[SM ? 0 : j] = ld_stream_u4(crow + i);
      }
    }
  }
  if (SM) asm volatile("cp.async.commit_group;" ::: "memory");
  uint4 scache[R <= 4 ? R : 1];
  if (R <= 4 && Q == 2 && a.style_hw[0] == a.HW) {
    const uint4* srow0 =
        reinterpret_cast<const uint4*>(reinterpret_cast<const typename VT::elem*>(a.styles[0]) + row * a.HW);
#pragma unroll
    for (int j = 0; j < (R <= 4 ? R : 1); ++j) {
      int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
      if (i < v1) scache[j] = ld_stream_u4(srow0 + i);
    }
  }
  if (!SM && !((R <= 4) && Q == 2 && a.style_hw[0] == a.HW && CS == 1)) {   // not the two-pass fast path below
    ShiftedLanes<V> w;
    w.init(VT::load1(a.content, row * a.HW));
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
      if (i < v1) {
        float x[V];
        VT::unpack(cache[SM ? 0 : j], x);
        w.push(x);
      }
    }
    cta_merge_to(w.fold(), s_warp, &s_q[0]);
  }
  // Common case (one style map of the content's size, both rows cached in <= 4 vectors per thread, no cluster
  // split): the style row's loads were issued together with the content row's, and the statistics are a TRUE
  // two-pass computation on the register-resident rows -- sum, then centred sum of squares -- with two cheap
  // two-value block reductions.  The streaming Welford + Chan merges this replaces cost ~1150 instructions per
  // thread and made the kernel issue-bound (ncu: issue slots 65 % busy at 62 % of the HBM copy peak).
  const bool style_prefetched = (R <= 4) && Q == 2 && a.style_hw[0] == a.HW;
  const bool two_pass = style_prefetched && CS == 1;
  if (two_pass) {
    __shared__ float s_sum[2][kWarps][2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
      if (i < v1) {
        float xs[V], ys[V];
        VT::unpack(cache[SM ? 0 : j], xs);
        VT::unpack(scache[j], ys);
#pragma unroll
        for (int e = 0; e < V; ++e) { t0 += xs[e]; t1 += ys[e]; }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, off);
      t1 += __shfl_xor_sync(0xffffffffu, t1, off);
    }
    if (lane == 0) { s_sum[0][wid][0] = t0; s_sum[0][wid][1] = t1; }
    __syncthreads();
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kWarps; ++w2) { m0 += s_sum[0][w2][0]; m1 += s_sum[0][w2][1]; }
    const float inv_n = 1.f / (float)a.HW;
    m0 *= inv_n; m1 *= inv_n;
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
      if (i < v1) {
        float xs[V], ys[V];
        VT::unpack(cache[SM ? 0 : j], xs);
        VT::unpack(scache[j], ys);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float dx = xs[e] - m0, dy = ys[e] - m1;
          q0 = fmaf(dx, dx, q0);
          q1 = fmaf(dy, dy, q1);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      q0 += __shfl_xor_sync(0xffffffffu, q0, off);
      q1 += __shfl_xor_sync(0xffffffffu, q1, off);
    }
    if (lane == 0) { s_sum[1][wid][0] = q0; s_sum[1][wid][1] = q1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float r0 = 0.f, r1 = 0.f;
      for (int w2 = 0; w2 < kWarps; ++w2) { r0 += s_sum[1][w2][0]; r1 += s_sum[1][w2][1]; }
      s_q[0] = Moments{(float)a.HW, m0, r0};
      s_q[1] = Moments{(float)a.HW, m1, r1};
    }
  } else if (style_prefetched) {
    ShiftedLanes<V> w;
    w.init(VT::load1(a.styles[0], row * a.HW));
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
      if (i < v1) {
        float x[V];
        VT::unpack(scache[j], x);
        w.push(x);
      }
    }
    cta_merge_to(w.fold(), s_warp, &s_q[1]);
  }
  for (int k = 0; k + 1 < Q && !style_prefetched && !SM; ++k) {
    const int64_t snvec = a.style_hw[k] / V;
    const int64_t sseg = (snvec + CS - 1) / CS;
    const int64_t s0 = rank * sseg;
    const int64_t s1 = (s0 + sseg < snvec) ? s0 + sseg : snvec;
    const void* srow = reinterpret_cast<const typename VT::elem*>(a.styles[k]) + row * a.style_hw[k];
    Moments m = (s0 < s1) ? stream_moments_vec<BF16>(srow, s0, s1) : Moments{0.f, 0.f, 0.f};
    cta_merge_to(m, s_warp, &s_q[1 + k]);
  }
  if (SM) {
    // Long rows (cluster-split, R == 8): plain sums about a COMMON per-row shift x0 (the row's first element, the same
    // in every thread and CTA of the cluster), so partial sums simply ADD -- across lanes, warps and CTAs -- and one
    // reduction with one pair of block barriers serves all 1 + K quantities.  The Chan merges this replaces ran in
    // every shuffle step of five separate block reductions: 61 thread-instructions per 16-byte vector, issue-bound
    // at 60 % of the HBM copy peak (profiles/r1_ncu_k1_long_rows_summary.csv).  mean = x0 + S/n,
    // m2 = Q - S^2/n: the cancellation costs (mean - x0)^2 / sigma^2 ulps, a handful for a shift taken from the row.
    __shared__ float s_red[kWarps][2 * kMaxQ];
    __shared__ float s_part[2 * kMaxQ];      // this CTA's (S, Q) per quantity; read by the cluster through DSMEM
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float accS[kMaxQ], accQ[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) { accS[q] = 0.f; accQ[q] = 0.f; }
#pragma unroll
    for (int k = 0; k < AST_MAX_STYLES; ++k) {
      if (k + 1 < Q) {
        const int64_t snvec = a.style_hw[k] / V;
        const int64_t sseg = (snvec + CS - 1) / CS;
        const int64_t s0 = rank * sseg;
        const int64_t s1 = (s0 + sseg < snvec) ? s0 + sseg : snvec;
        const typename VT::elem* sbase = reinterpret_cast<const typename VT::elem*>(a.styles[k]) + row * a.style_hw[k];
        const float x0 = VT::load1(sbase, 0);
        const float2 nk = make_float2(-x0, -x0);
        float2 ss[V / 2], qq[V / 2];
#pragma unroll
        for (int j = 0; j < V / 2; ++j) { ss[j] = make_float2(0.f, 0.f); qq[j] = make_float2(0.f, 0.f); }
        const uint4* sp = reinterpret_cast<const uint4*>(sbase);
        constexpr int U = 8;
        int64_t i = s0 + threadIdx.x;
        for (; i + (U - 1) * kThreads < s1; i += U * kThreads) {
          uint4 u[U];
#pragma unroll
          for (int j = 0; j < U; ++j) u[j] = ld_stream_u4(sp + i + j * kThreads);
#pragma unroll
          for (int j = 0; j < U; ++j) {
            float x[V];
            VT::unpack(u[j], x);
#pragma unroll
            for (int e = 0; e < V / 2; ++e) {
              const float2 d = __fadd2_rn(make_float2(x[2 * e], x[2 * e + 1]), nk);
              ss[e] = __fadd2_rn(ss[e], d);
              qq[e] = __ffma2_rn(d, d, qq[e]);
            }
          }
        }
        for (; i < s1; i += kThreads) {
          float x[V];
          VT::unpack(ld_stream_u4(sp + i), x);
#pragma unroll
          for (int e = 0; e < V / 2; ++e) {
            const float2 d = __fadd2_rn(make_float2(x[2 * e], x[2 * e + 1]), nk);
            ss[e] = __fadd2_rn(ss[e], d);
            qq[e] = __ffma2_rn(d, d, qq[e]);
          }
        }
        float S = 0.f, Qs = 0.f;
#pragma unroll
        for (int e = 0; e < V / 2; ++e) { S += ss[e].x + ss[e].y; Qs += qq[e].x + qq[e].y; }
        accS[1 + k] = S; accQ[1 + k] = Qs;
      }
    }
    // the content copies have had the whole style stream to land
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const float c0 = VT::load1(a.content, row * a.HW);
    {
      const float2 nk = make_float2(-c0, -c0);
      float2 ss[V / 2], qq[V / 2];
#pragma unroll
      for (int j = 0; j < V / 2; ++j) { ss[j] = make_float2(0.f, 0.f); qq[j] = make_float2(0.f, 0.f); }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
        if (i < v1) {
          float x[V];
          VT::unpack(s_cache[threadIdx.x + j * kThreads], x);
#pragma unroll
          for (int e = 0; e < V / 2; ++e) {
            const float2 d = __fadd2_rn(make_float2(x[2 * e], x[2 * e + 1]), nk);
            ss[e] = __fadd2_rn(ss[e], d);
            qq[e] = __ffma2_rn(d, d, qq[e]);
          }
        }
      }
      float S = 0.f, Qs = 0.f;
#pragma unroll
      for (int e = 0; e < V / 2; ++e) { S += ss[e].x + ss[e].y; Qs += qq[e].x + qq[e].y; }
      accS[0] = S; accQ[0] = Qs;
    }
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) {
      if (q < Q) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          accS[q] += __shfl_xor_sync(0xffffffffu, accS[q], off);
          accQ[q] += __shfl_xor_sync(0xffffffffu, accQ[q], off);
        }
        if (lane == 0) { s_red[wid][2 * q] = accS[q]; s_red[wid][2 * q + 1] = accQ[q]; }
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < 2 * Q) {
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kWarps; ++w2) t += s_red[w2][threadIdx.x];
      s_part[threadIdx.x] = t;
    }
    cluster.sync();   // every CTA's partial sums are complete and visible cluster-wide (CS == 1: a block barrier)
    if ((int)threadIdx.x < Q) {
      const int q = threadIdx.x;
      float S = 0.f, Qs = 0.f;
      for (unsigned c = 0; c < CS; ++c) {
        const float* pp = cluster.map_shared_rank(s_part, c);
        S += pp[2 * q];
        Qs += pp[2 * q + 1];
      }
      const float n = q == 0 ? (float)a.HW : (float)a.style_hw[q - 1];
      const float x0 = q == 0 ? c0 : VT::load1(reinterpret_cast<const typename VT::elem*>(a.styles[q - 1]) +
                                               row * a.style_hw[q - 1], 0);
      s_fin[q] = Moments{n, x0 + S / n, fmaxf(fmaf(-S, S / n, Qs), 0.f)};
    }
  }
  const Moments* fin = s_q;
  if (SM) {
    fin = s_fin;
  } else if (CS > 1) {
    cluster.sync();  // every CTA's s_q is complete and visible cluster-wide
    if ((int)threadIdx.x < Q) {
      Moments r = *cluster.map_shared_rank(&s_q[threadIdx.x], 0);
      for (unsigned c = 1; c < CS; ++c)
        r = moments_merge(r, *cluster.map_shared_rank(&s_q[threadIdx.x], c));
      s_fin[threadIdx.x] = r;
    }
    fin = s_fin;
  }
  __syncthreads();

  float mu, rsig, A, B;
  finish_stats(a, fin, row, rank == 0 && threadIdx.x == 0, mu, rsig, A, B);
  const bool blend = a.alpha != 1.f;

  uint4* orow = reinterpret_cast<uint4*>(reinterpret_cast<typename VT::elem*>(a.out) + row * a.HW);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    int64_t i = v0 + threadIdx.x + (int64_t)j * kThreads;
    if (i < v1) {
      float x[V];
      VT::unpack(SM ? s_cache[threadIdx.x + j * kThreads] : cache[SM ? 0 : j], x);
#pragma unroll
      for (int e = 0; e < V; ++e) x[e] = apply_affine(x[e], mu, rsig, A, B, a.alpha, blend);
      st_stream_u4(orow + i, VT::pack(x));
    }
  }
  if (CS > 1) cluster.sync();  // peers may still be reading s_q through DSMEM
}

// ---- generic two-pass kernel: one CTA per row, any length / alignment ----------------------
template <bool BF16>
__global__ void __launch_bounds__(kThreads) adain_generic_kernel(const AdainArgs a) {
  using VT = Vec16<BF16>;
  constexpr int V = VT::V;
  using E = typename VT::elem;
  const int64_t row = blockIdx.x;
  __shared__ Moments s_warp[kWarps][kMaxQ];

  const int Q = a.identity_affine ? 1 : 1 + a.K;
  Moments m[kMaxQ];
  const E* crow = reinterpret_cast<const E*>(a.content) + row * a.HW;
  E* orow = reinterpret_cast<E*>(a.out) + row * a.HW;
  const bool cvec = aligned16(crow) && aligned16(orow) && (a.HW % V == 0);
  m[0] = cvec ? stream_moments_vec<BF16>(crow, 0, a.HW / V)
              : stream_moments_scalar<BF16>(crow, 0, a.HW);
#pragma unroll
  for (int k = 0; k < AST_MAX_STYLES; ++k) {
    if (k + 1 >= Q) break;
    const E* srow = reinterpret_cast<const E*>(a.styles[k]) + row * a.style_hw[k];
    const bool svec = aligned16(srow) && (a.style_hw[k] % V == 0);
    m[1 + k] = svec ? stream_moments_vec<BF16>(srow, 0, a.style_hw[k] / V)
                    : stream_moments_scalar<BF16>(srow, 0, a.style_hw[k]);
  }
  block_merge<kMaxQ>(m, Q, s_warp);

  float mu, rsig, A, B;
  finish_stats(a, m, row, threadIdx.x == 0, mu, rsig, A, B);
  const bool blend = a.alpha != 1.f;
  if (a.out == nullptr) return;  // statistics only

  if (cvec) {
    const uint4* p = reinterpret_cast<const uint4*>(crow);
    uint4* o = reinterpret_cast<uint4*>(orow);
    const int64_t nvec = a.HW / V;
    for (int64_t i = threadIdx.x; i < nvec; i += kThreads) {
      float x[V];
      VT::unpack(__ldg(p + i), x);  // second pass: expected to hit L2
#pragma unroll
      for (int e = 0; e < V; ++e) x[e] = apply_affine(x[e], mu, rsig, A, B, a.alpha, blend);
      st_stream_u4(o + i, VT::pack(x));
    }
  } else {
    for (int64_t i = threadIdx.x; i < a.HW; i += kThreads) {
      float x = VT::load1(crow, i);
      VT::store1(orow, i, apply_affine(x, mu, rsig, A, B, a.alpha, blend));
    }
  }
}

// ---- statistics only: warp-per-row for short rows -----------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kThreads) stats_warp_rows_kernel(const void* x, float* mean,
                                                                   float* std_, int64_t rows,
                                                                   int64_t HW, float eps,
                                                                   unsigned flags) {
  using VT = Vec16<BF16>;
  constexpr int V = VT::V;
  using E = typename VT::elem;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const E* r = reinterpret_cast<const E*>(x) + row * HW;
  Moments m;
  if (aligned16(r) && HW % V == 0) {
    WelfordLanes<V> w;
    w.init();
    const uint4* p = reinterpret_cast<const uint4*>(r);
    for (int64_t i = lane; i < HW / V; i += 32) {
      float v[V];
      VT::unpack(ld_stream_u4(p + i), v);
      w.push(v);
    }
    m = w.fold();
  } else {
    WelfordLanes<1> w;
    w.init();
    for (int64_t i = lane; i < HW; i += 32) {
      float v[1] = {VT::load1(r, i)};
      w.push(v);
    }
    m = w.fold();
  }
  m = moments_warp_reduce(m);
  if (lane == 0) {
    mean[row] = m.mean;
    std_[row] = std_from(m, eps, flags);
  }
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads) stats_block_rows_kernel(const void* x, float* mean,
                                                                    float* std_, int64_t HW,
                                                                    float eps, unsigned flags) {
  using VT = Vec16<BF16>;
  constexpr int V = VT::V;
  using E = typename VT::elem;
  __shared__ Moments s_warp[kWarps][1];
  const int64_t row = blockIdx.x;
  const E* r = reinterpret_cast<const E*>(x) + row * HW;
  Moments m[1];
  m[0] = (aligned16(r) && HW % V == 0) ? stream_moments_vec<BF16>(r, 0, HW / V)
                                       : stream_moments_scalar<BF16>(r, 0, HW);
  block_merge<1>(m, 1, s_warp);
  if (threadIdx.x == 0) {
    mean[row] = m[0].mean;
    std_[row] = std_from(m[0], eps, flags);
  }
}

// ---- backward kernels ----------------------------------------------------------------------
// channel_stats backward: gx = g_mean/HW + g_std * (x - mean) / (denom * std)
template <bool BF16>
__global__ void __launch_bounds__(kThreads) stats_bwd_kernel(const void* x, const float* mean,
                                                             const float* std_, const float* gm,
                                                             const float* gs, void* gx, int64_t HW,
                                                             unsigned flags) {
  using VT = Vec16<BF16>;
  constexpr int V = VT::V;
  using E = typename VT::elem;
  const int64_t row = blockIdx.x;
  const E* r = reinterpret_cast<const E*>(x) + row * HW;
  E* o = reinterpret_cast<E*>(gx) + row * HW;
  const float n = (float)HW;
  const float denom = (flags & AST_F_BIASED) ? n : n - 1.f;
  const float mu = mean[row];
  const float c0 = gm ? gm[row] / n : 0.f;
  const float c1 = gs ? gs[row] / (denom * std_[row]) : 0.f;
  if (aligned16(r) && aligned16(o) && HW % V == 0) {
    for (int64_t i = threadIdx.x; i < HW / V; i += kThreads) {
      float v[V];
      VT::unpack(ld_stream_u4(reinterpret_cast<const uint4*>(r) + i), v);
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = fmaf(c1, v[e] - mu, c0);
      st_stream_u4(reinterpret_cast<uint4*>(o) + i, VT::pack(v));
    }
  } else {
    for (int64_t i = threadIdx.x; i < HW; i += kThreads)
      VT::store1(o, i, fmaf(c1, VT::load1(r, i) - mu, c0));
  }
}

__device__ __forceinline__ float block_sum(float v, float* smem /*[kWarps]*/) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();  // protect smem reuse between calls
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) r += smem[w];
  return r;
}

// MVN backward: gx = (gy - mean(gy))/std - y * sum(gy*y) / (denom * std)
template <bool BF16>
__global__ void __launch_bounds__(kThreads) mvn_bwd_kernel(const void* x, const void* gy,
                                                           const float* stats, void* gx,
                                                           int64_t HW, unsigned flags) {
  using VT = Vec16<BF16>;
  using E = typename VT::elem;
  __shared__ float s_red[kWarps];
  const int64_t row = blockIdx.x;
  const E* xr = reinterpret_cast<const E*>(x) + row * HW;
  const E* gr = reinterpret_cast<const E*>(gy) + row * HW;
  E* o = reinterpret_cast<E*>(gx) + row * HW;
  const float n = (float)HW;
  const float denom = (flags & AST_F_BIASED) ? n : n - 1.f;
  const float mu = stats[2 * row];
  const float rs = 1.f / stats[2 * row + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = threadIdx.x; i < HW; i += kThreads) {
    float g = VT::load1(gr, i);
    float y = (VT::load1(xr, i) - mu) * rs;
    s1 += g;
    s2 = fmaf(g, y, s2);
  }
  s1 = block_sum(s1, s_red);
  s2 = block_sum(s2, s_red);
  const float gmean = s1 / n;
  const float c = s2 / denom;
  for (int64_t i = threadIdx.x; i < HW; i += kThreads) {
    float g = VT::load1(gr, i);
    float y = (VT::load1(xr, i) - mu) * rs;
    VT::store1(o, i, ((g - gmean) - y * c) * rs);
  }
}

struct AdainBwdW { float w[AST_MAX_STYLES]; };

// AdaIN backward with respect to the content map, plus everything the style gradients need, in one kernel.
// Forward (models.py:43-51, :471): y = alpha * (z A + B) + (1 - alpha) c with z = (c - mu) / sigma (unbiased sigma, no
// eps) and, per row, (A, B) = sum_k w_k (mean_k, std_k) of the style maps [reference binding, models.py:44] or
// sum_k w_k (std_k, mean_k) [canonical].  With g = dL/dy:
//   gc = alpha A / sigma * (g - mean(g) - z sum(g z) / (HW - 1)) + (1 - alpha) g
//   dA = alpha sum(g z), dB = alpha sum(g)  ->  per style k: g_mean_k, g_std_k = w_k (dA, dB) [or (dB, dA)]
// aux[k] = (mean_k, std_k, g_mean_k, g_std_k), each [rows], contiguous: the arguments of ast_channel_stats_bwd.
template <bool BF16>
__global__ void __launch_bounds__(kThreads) adain_bwd_kernel(const void* c, const void* gy, const float* stats, int K,
                                                            const AdainBwdW ww, float alpha, void* gc, float* aux,
                                                            int64_t rows, int64_t HW, unsigned flags) {
  const float* w = ww.w;
  using VT = Vec16<BF16>;
  using E = typename VT::elem;
  __shared__ float s_red[kWarps];
  const int64_t row = blockIdx.x;
  const E* xr = reinterpret_cast<const E*>(c) + row * HW;
  const E* gr = reinterpret_cast<const E*>(gy) + row * HW;
  E* o = reinterpret_cast<E*>(gc) + row * HW;
  const float* st = stats + row * (2 + 2 * K);
  const float n = (float)HW;
  const float mu = st[0], rs = 1.f / st[1];
  float A = 0.f;
  for (int k = 0; k < K; ++k) A = fmaf(w[k], st[2 + 2 * k + ((flags & AST_F_CANONICAL) ? 1 : 0)], A);
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = threadIdx.x; i < HW; i += kThreads) {
    const float g = VT::load1(gr, i);
    const float z = (VT::load1(xr, i) - mu) * rs;
    s1 += g;
    s2 = fmaf(g, z, s2);
  }
  s1 = block_sum(s1, s_red);
  s2 = block_sum(s2, s_red);
  const float gmean = s1 / n, cz = s2 / (n - 1.f);
  const float sc = alpha * A * rs, direct = 1.f - alpha;
  for (int64_t i = threadIdx.x; i < HW; i += kThreads) {
    const float g = VT::load1(gr, i);
    const float z = (VT::load1(xr, i) - mu) * rs;
    VT::store1(o, i, fmaf(sc, (g - gmean) - z * cz, direct * g));
  }
  if (threadIdx.x == 0 && aux) {
    const float dA = alpha * s2, dB = alpha * s1;
    for (int k = 0; k < K; ++k) {
      float* a = aux + (int64_t)k * 4 * rows;
      a[row] = st[2 + 2 * k];
      a[rows + row] = st[3 + 2 * k];
      const bool canon = (flags & AST_F_CANONICAL) != 0;
      a[2 * rows + row] = w[k] * (canon ? dB : dA);     // dL/d mean_k
      a[3 * rows + row] = w[k] * (canon ? dA : dB);     // dL/d std_k
    }
  }
}

// ---- host dispatch -------------------------------------------------------------------------
template <bool BF16, int R>
static int launch_cached(const AdainArgs& a, int64_t rows, unsigned CS, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(rows * CS));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (R == 8) ? (size_t)R * kThreads * sizeof(uint4) : 0;   // 32 KB content segment
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AST_CUDA(cudaLaunchKernelEx(&cfg, adain_cached_kernel<BF16, R>, a));
  return 0;
}

template <bool BF16>
static int adain_dispatch(const AdainArgs& a, int64_t rows, cudaStream_t s) {
  constexpr int V = Vec16<BF16>::V;
  bool vec_ok = a.out && aligned16(a.content) && aligned16(a.out) && (a.HW % V == 0);
  const int Q = a.identity_affine ? 0 : a.K;
  for (int k = 0; k < Q; ++k)
    vec_ok = vec_ok && aligned16(a.styles[k]) && (a.style_hw[k] % V == 0);
  const int64_t nvec = a.HW / V;
  const int64_t kMaxSeg = 8 * kThreads;  // vectors per CTA at R = 8
  if (vec_ok && nvec <= 8 * kMaxSeg && rows * 8 < 0x7fffffffLL) {
    unsigned CS = 1;
    while ((int64_t)CS * kMaxSeg < nvec) CS <<= 1;
    // prefer more, smaller CTAs when there are few rows (fills 148 SMs at cfg 1 / cfg 5)
    while (CS < 8 && rows * CS < 2 * 148 && nvec / (2 * CS) >= kThreads) CS <<= 1;
    // (Tried: twice as wide clusters -- 16 CTAs, non-portable -- to cache 4 instead of 8 vectors per thread and
    // fit 3 CTAs per SM on the long-row path: config 5 went from 283 to 461 us, cluster scheduling costs more
    // than the occupancy buys.)
    const int64_t seg = (nvec + CS - 1) / CS;
    const int64_t r = (seg + kThreads - 1) / kThreads;
    if (r <= 1) return launch_cached<BF16, 1>(a, rows, CS, s);
    if (r <= 2) return launch_cached<BF16, 2>(a, rows, CS, s);
    if (r <= 4) return launch_cached<BF16, 4>(a, rows, CS, s);
    return launch_cached<BF16, 8>(a, rows, CS, s);
  }
  adain_generic_kernel<BF16><<<(unsigned)rows, kThreads, 0, s>>>(a);
  AST_CHECK_LAUNCH();
  return 0;
}

}  // namespace ast

using namespace ast;

extern "C" int ast_adain_fwd(const void* content, const void* const* styles,
                             const int64_t* style_hw, const float* style_w, int K, void* out,
                             float* stats, int N, int C, int64_t HW, float alpha, float eps,
                             unsigned flags, void* stream) {
  if (!content || !out || N <= 0 || C <= 0 || HW <= 0 || K < 0) return AST_E_BADARG;
  if (K > AST_MAX_STYLES) return AST_E_TOOMANY;
  if (K > 0 && (!styles || !style_hw || !style_w)) return AST_E_BADARG;
  AdainArgs a = {};
  a.content = content; a.out = out; a.stats = stats; a.K = K; a.HW = HW;
  a.alpha = alpha; a.eps = eps; a.flags = flags; a.identity_affine = 0;
  for (int k = 0; k < K; ++k) {
    if (!styles[k] || style_hw[k] <= 0) return AST_E_BADARG;
    a.styles[k] = styles[k]; a.style_hw[k] = style_hw[k]; a.style_w[k] = style_w[k];
  }
  const int64_t rows = (int64_t)N * C;
  if (rows >= 0x7fffffffLL) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  return (flags & AST_F_BF16) ? adain_dispatch<true>(a, rows, s) : adain_dispatch<false>(a, rows, s);
}

extern "C" int ast_channel_stats_fwd(const void* x, float* mean, float* std_, int64_t rows,
                                     int64_t HW, float eps, unsigned flags, void* stream) {
  if (!x || !mean || !std_ || rows <= 0 || HW <= 0) return AST_E_BADARG;
  if (rows >= 0x7fffffffLL) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  const bool bf = flags & AST_F_BF16;
  if (HW <= 2048) {
    unsigned grid = (unsigned)((rows + kWarps - 1) / kWarps);
    if (bf) stats_warp_rows_kernel<true><<<grid, kThreads, 0, s>>>(x, mean, std_, rows, HW, eps, flags);
    else stats_warp_rows_kernel<false><<<grid, kThreads, 0, s>>>(x, mean, std_, rows, HW, eps, flags);
  } else {
    if (bf) stats_block_rows_kernel<true><<<(unsigned)rows, kThreads, 0, s>>>(x, mean, std_, HW, eps, flags);
    else stats_block_rows_kernel<false><<<(unsigned)rows, kThreads, 0, s>>>(x, mean, std_, HW, eps, flags);
  }
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_channel_stats_bwd(const void* x, const float* mean, const float* std_,
                                     const float* g_mean, const float* g_std, void* gx,
                                     int64_t rows, int64_t HW, unsigned flags, void* stream) {
  if (!x || !mean || !std_ || !gx || rows <= 0 || HW <= 0) return AST_E_BADARG;
  if (rows >= 0x7fffffffLL) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & AST_F_BF16)
    stats_bwd_kernel<true><<<(unsigned)rows, kThreads, 0, s>>>(x, mean, std_, g_mean, g_std, gx, HW, flags);
  else
    stats_bwd_kernel<false><<<(unsigned)rows, kThreads, 0, s>>>(x, mean, std_, g_mean, g_std, gx, HW, flags);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_mvn_fwd(const void* x, void* y, float* stats, int64_t rows, int64_t HW,
                           float eps, unsigned flags, void* stream) {
  if (!x || !y || rows <= 0 || HW <= 0) return AST_E_BADARG;
  if (rows >= 0x7fffffffLL) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  AdainArgs a = {};
  a.content = x; a.out = y; a.stats = stats; a.K = 0; a.HW = HW;
  a.alpha = 1.f; a.eps = eps; a.flags = flags; a.identity_affine = 1;
  return (flags & AST_F_BF16) ? adain_dispatch<true>(a, rows, s) : adain_dispatch<false>(a, rows, s);
}

extern "C" int ast_mvn_bwd(const void* x, const void* gy, const float* stats, void* gx,
                           int64_t rows, int64_t HW, unsigned flags, void* stream) {
  if (!x || !gy || !stats || !gx || rows <= 0 || HW <= 0) return AST_E_BADARG;
  if (rows >= 0x7fffffffLL) return AST_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & AST_F_BF16)
    mvn_bwd_kernel<true><<<(unsigned)rows, kThreads, 0, s>>>(x, gy, stats, gx, HW, flags);
  else
    mvn_bwd_kernel<false><<<(unsigned)rows, kThreads, 0, s>>>(x, gy, stats, gx, HW, flags);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_adain_bwd(const void* content, const void* gy, const float* stats, const float* style_w, int K,
                             float alpha, void* g_content, float* aux, int64_t rows, int64_t HW, unsigned flags,
                             void* stream) {
  if (!content || !gy || !stats || !style_w || !g_content || K < 1 || K > AST_MAX_STYLES || rows <= 0 || HW <= 1)
    return AST_E_BADARG;
  if (rows >= 0x7fffffffLL) return AST_E_SHAPE;
  AdainBwdW ww = {};
  for (int k = 0; k < K; ++k) ww.w[k] = style_w[k];
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & AST_F_BF16)
    adain_bwd_kernel<true><<<(unsigned)rows, kThreads, 0, s>>>(content, gy, stats, K, ww, alpha, g_content, aux, rows, HW, flags);
  else
    adain_bwd_kernel<false><<<(unsigned)rows, kThreads, 0, s>>>(content, gy, stats, K, ww, alpha, g_content, aux, rows, HW, flags);
  AST_CHECK_LAUNCH();
  return 0;
}
