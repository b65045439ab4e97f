// K4d: shared-memory-tiled depthwise convolution kernels for stride 1 (the layers that carry the
// traffic: every decoder block and most encoder blocks of models.py:140-320).  Forward, data gradient
// and weight gradient of nn.Conv2d(C, C, k, 1, groups=C, padding_mode="reflect") on NHWC bf16.
//
// Why tiles: the direct kernels (mobile.cu / mobile_train.cu) issue k*k global loads with 64-bit index
// arithmetic and a reflection per tap and ran 15-20x above the HBM floor.  Here a CTA stages the
// (TH+k-1) x (TW+k-1) input patch of one channel block ONCE (reflection / zero fill / x2-upsample resolved
// during staging) and each thread produces a strip of 8 consecutive outputs x 4 channels from shared
// memory with packed FFMA2 (fma.rn.f32x2), fp32 accumulation.
//
// What bounds it (ncu, profiles/): a 5x5 depthwise conv does 25 FMA per 4 bytes of HBM traffic, which is
// the FP32 roof (128 FMA/clk/SM) rather than the HBM roof, and every FMA operand comes through the
// 128 B/clk shared-memory port: a strip of R outputs re-uses each LDS'ed input for up to k taps, so R = 8
// and 4 channels per thread (32 accumulators, ~80 registers, 3 CTAs per SM) balance the LDS port, the FMA
// pipe and the latency hiding.  The 3x3 layers get close to the HBM floor.
//
// Geometry is chosen on the host per channel count (pick_geom): hvn = 8-byte channel groups per CTA (a
// divisor of C/4), workers = 256 / hvn thread groups, tile = TH x TW with TH * TW/8 a multiple of the
// worker count so that every round of the strip loop is full.
#include <type_traits>
#include "common.cuh"

namespace ast {
namespace dwt {

constexpr int kThreads = 256;
constexpr int R = 8;   // outputs per strip
constexpr int CH = 4;  // channels per thread (one 8-byte vector)

__device__ __forceinline__ float hsw(float x) { return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f); }
__device__ __forceinline__ float hsw_grad(float x) {   // ATen CPU convention, see mobile_train.cu
  return x <= -3.f ? 0.f : (x < 3.f ? fmaf(x, 1.f / 3.f, 0.5f) : 1.f);
}
__device__ __forceinline__ int reflect_idx(int p, int X) {
  p = p < 0 ? -p : p;
  return p >= X ? 2 * X - 2 - p : p;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// 4 bf16 -> 2 float2 (channel pairs).  All FMAs are the packed FFMA2: on sm_100 the scalar 3-register FFMA
// issues at half rate per SM sub-partition, FFMA2 restores the full FP32 rate.
// T = act_t (fp16 activations) or grad_t (bf16 gradients): common.cuh
template <typename T>
__device__ __forceinline__ void unpack2x2(uint2 u, float2 (&x)[2]) {
  x[0] = H16<T>::un2(u.x);
  x[1] = H16<T>::un2(u.y);
}
template <typename T>
__device__ __forceinline__ void unpack4(uint2 u, float (&x)[4]) {
  const float2 a = H16<T>::un2(u.x), b = H16<T>::un2(u.y);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template <typename T>
__device__ __forceinline__ uint2 pack4(const float (&o)[4]) {
  return make_uint2(H16<T>::pk2(o[0], o[1]), H16<T>::pk2(o[2], o[3]));
}
__device__ __forceinline__ void ld_w2x2(const float* p, float2 (&w)[2]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  w[0] = make_float2(a.x, a.y);
  w[1] = make_float2(a.z, a.w);
}
__device__ __forceinline__ uint2 ldg_u2(const void* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
// Row pitch (in 8-byte units) of a staged tile with `cols` pixels of hvn vectors: padded so that
// pitch == hvn (mod 16).  Consecutive workers take consecutive ROWS of the same 8-pixel column group, so
// the (worker, vector) pairs of a half warp fall on 16 consecutive 8-byte bank pairs: conflict free for
// every hvn.
__host__ __device__ __forceinline__ int row_pitch(int cols, int hvn) {
  const int raw = cols * hvn;
  return raw + (((hvn - raw) % 16) + 16) % 16;
}

// Cooperative staging of a rows x cols pixel tile, `cpp` 16-byte chunks per pixel, into shared-memory rows of
// `pitch16` 16-byte units.  src(py, px) returns the pixel's channel-block address or nullptr for zero fill.
// 16-byte global loads whatever the compute granularity, and the (row, pixel, chunk) cursor advances by
// carries instead of divisions: staging used to cost more instructions than the convolution itself.
template <class SrcFn>
__device__ __forceinline__ void stage_tile(uint4* dst, int rows, int cols, int cpp, int pitch16, SrcFn src) {
  const int per_row = cols * cpp;
  int py = threadIdx.x / per_row;
  int rem = threadIdx.x - py * per_row;
  int px = rem / cpp;
  int cc = rem - px * cpp;
  const int dpy = kThreads / per_row;
  const int drem = kThreads - dpy * per_row;
  const int dpx = drem / cpp;
  const int dcc = drem - dpx * cpp;
  while (py < rows) {
    const uint16_t* sp = src(py, px);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (sp) val = __ldg(reinterpret_cast<const uint4*>(sp) + cc);
    dst[py * pitch16 + px * cpp + cc] = val;
    cc += dcc;
    if (cc >= cpp) { cc -= cpp; ++px; }
    px += dpx;
    if (px >= cols) { px -= cols; ++py; }
    py += dpy;
  }
}

// Same cursor, but the copy is an asynchronous 16-byte cp.async (zero fill when src returns nullptr): the
// caller commits the group, computes on the other buffer and waits before use.
__device__ __forceinline__ void cp_async16(uint4* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class SrcFn>
__device__ __forceinline__ void stage_tile_async(uint4* dst, int rows, int cols, int cpp, int pitch16,
                                                 const void* safe, SrcFn src) {
  const int per_row = cols * cpp;
  int py = threadIdx.x / per_row;
  int rem = threadIdx.x - py * per_row;
  int px = rem / cpp;
  int cc = rem - px * cpp;
  const int dpy = kThreads / per_row;
  const int drem = kThreads - dpy * per_row;
  const int dpx = drem / cpp;
  const int dcc = drem - dpx * cpp;
  while (py < rows) {
    const uint16_t* sp = src(py, px);
    cp_async16(dst + py * pitch16 + px * cpp + cc, sp ? (const void*)(reinterpret_cast<const uint4*>(sp) + cc) : safe,
               sp != nullptr);
    cc += dcc;
    if (cc >= cpp) { cc -= cpp; ++px; }
    px += dpx;
    if (px >= cols) { px -= cols; ++py; }
    py += dpy;
  }
}

struct Geom {
  int hvn, workers, TH, TW, tiles_h, tiles_w, cblocks;
};

struct Params {
  // 16-bit tensors as raw uint16_t: forward (MODE 0) in / out are fp16 activations; data gradient (MODE 1) in / out /
  // dres are bf16 gradients and a_pre is an fp16 activation
  const uint16_t* in;          // forward: x [N][H][W][C];  dgrad: dy [N][H][W][C]
  const float* w;              // fp32 [k*k][C]
  const float* bias;           // forward only, nullable
  uint16_t* out;               // [N][Hc][Wc][C]   (Hc x Wc = conv grid = 2H x 2W when up2)
  float* pool;                 // forward only, nullable: [N][C] sums
  const uint16_t* a_pre;       // dgrad only, nullable: multiply by Hardswish'(a_pre * sc + sh)
  const float* stat;           // dgrad only, nullable: [4][C]
  const uint16_t* dres;        // dgrad only, nullable: identity-branch gradient added
  int N, C, H, W, Hc, Wc, up2, act;
  Geom g;
};

// MODE 0: forward (reflect staging, bias / Hardswish / pool epilogue)
// MODE 1: data gradient (zero-fill staging, flipped weights, reflection fold at the borders)
template <typename AT, int K, int MODE>
__global__ void __launch_bounds__(kThreads, (MODE == 0 ? 2 : 3))
dw_tiled_kernel(const Params p) {
  extern __shared__ __align__(16) uint2 smem_u2[];
  constexpr int PAD = (K - 1) / 2;
  using IO = typename std::conditional<MODE == 0, AT, grad_t>::type;   // element type of in / out (AT = activation format)
  const Geom g = p.g;
  const int PH = g.TH + K - 1, PW = g.TW + K - 1;
  const int RP = row_pitch(PW, g.hvn);
  const int patch_elems = (PH * RP + 1) & ~1;                        // 16-byte aligned buffers
  float* s_w = reinterpret_cast<float*>(smem_u2 + 2 * patch_elems);  // [K*K][hvn*4], after the two patch buffers
  float* s_pool = s_w + K * K * g.hvn * CH;                          // [hvn*4]
  const int CB = g.hvn * CH;
  const int cb = blockIdx.y;
  const int c0 = cb * CB;
  const int tiles_per_img = g.tiles_h * g.tiles_w;
  const int64_t total_tiles = (int64_t)p.N * tiles_per_img;

  // ---- weights of this channel block, staged once (flipped for the data gradient) ----
  for (int i = threadIdx.x; i < K * K * CB; i += kThreads) {
    const int t = i / CB, c = i - t * CB;
    const int ts = MODE == 1 ? (K * K - 1 - t) : t;
    s_w[i] = __ldg(p.w + (int64_t)ts * p.C + c0 + c);
  }

  // ---- persistent, double-buffered tile loop: the cp.async copies of tile i+1 fly while tile i is consumed ----
  auto stage = [&](int64_t t, int b) {
    const int n = (int)(t / tiles_per_img);
    const int tile = (int)(t - (int64_t)n * tiles_per_img);
    const int th = tile / g.tiles_w, tw = tile - th * g.tiles_w;
    const int oh0 = th * g.TH, ow0 = tw * g.TW;
    const uint16_t* src = p.in + (int64_t)n * p.H * p.W * p.C + c0;
    stage_tile_async(reinterpret_cast<uint4*>(smem_u2 + b * patch_elems), PH, PW, g.hvn / 2, RP / 2, p.in,
                     [&](int py, int px) -> const uint16_t* {
      int y = oh0 - PAD + py, x = ow0 - PAD + px;      // position on the conv grid (may be outside)
      if (MODE == 0) {
        // reflection; positions only needed by masked outputs are clamped into range
        y = reflect_idx(clampi(y, -PAD, p.Hc - 1 + PAD), p.Hc);
        x = reflect_idx(clampi(x, -PAD, p.Wc - 1 + PAD), p.Wc);
        if (p.up2) { y >>= 1; x >>= 1; }
      } else if (y < 0 || y >= p.Hc || x < 0 || x >= p.Wc) {
        return nullptr;
      }
      return src + ((int64_t)y * p.W + x) * p.C;
    });
    cp_async_commit();
  };

  const int worker = threadIdx.x / g.hvn;
  const int v = threadIdx.x - worker * g.hvn;
  const int groups_w = g.TW / R;
  const int items = g.TH * groups_w;

  if ((int64_t)blockIdx.x < total_tiles) stage(blockIdx.x, 0);
  int buf = 0;
  for (int64_t tcur = blockIdx.x; tcur < total_tiles; tcur += gridDim.x, buf ^= 1) {
  const int64_t tnext = tcur + gridDim.x;
  if (tnext < total_tiles) {
    stage(tnext, buf ^ 1);   // that buffer was released by the barrier that ended the previous iteration
    cp_async_wait<1>();
  } else {
    cp_async_wait<0>();
  }
  if (MODE == 0 && p.pool && threadIdx.x < CB) s_pool[threadIdx.x] = 0.f;
  __syncthreads();
  const uint2* patch = smem_u2 + buf * patch_elems;
  const int n = (int)(tcur / tiles_per_img);
  const int tile = (int)(tcur - (int64_t)n * tiles_per_img);
  const int th = tile / g.tiles_w, tw = tile - th * g.tiles_w;
  const int oh0 = th * g.TH, ow0 = tw * g.TW;
  float psum[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) psum[j] = 0.f;

  if (worker < g.workers) {
    for (int item = worker; item < items; item += g.workers) {
      const int cg = item / g.TH, r = item - cg * g.TH;   // rows fastest: see row_pitch()
      const int oh = oh0 + r, owb = ow0 + cg * R;
      if (oh >= p.Hc || owb >= p.Wc) continue;
      float2 acc2[R][2];
#pragma unroll
      for (int rr = 0; rr < R; ++rr) { acc2[rr][0] = make_float2(0.f, 0.f); acc2[rr][1] = make_float2(0.f, 0.f); }

      // Plain correlation from the staged patch.  For the data gradient this is exact everywhere except on
      // the few border rows / columns that also receive reflected contributions; those pixels are
      // recomputed by the border pass below.
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        float2 wrow[K][2];
#pragma unroll
        for (int kw = 0; kw < K; ++kw) ld_w2x2(s_w + (kh * K + kw) * CB + v * CH, wrow[kw]);
        const uint2* prow = patch + (r + kh) * RP + cg * R * g.hvn + v;
#pragma unroll
        for (int c = 0; c < R + K - 1; ++c) {
          float2 xv[2];
          unpack2x2<IO>(prow[c * g.hvn], xv);
#pragma unroll
          for (int kw = 0; kw < K; ++kw) {
            const int rr = c - kw;
            if (rr >= 0 && rr < R) {
              acc2[rr][0] = __ffma2_rn(xv[0], wrow[kw][0], acc2[rr][0]);
              acc2[rr][1] = __ffma2_rn(xv[1], wrow[kw][1], acc2[rr][1]);
            }
          }
        }
      }

      // ---- epilogue ----
      const int64_t obase = (((int64_t)n * p.Hc + oh) * p.Wc) * p.C + c0 + v * CH;
      if (MODE == 0) {
        float b[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) b[j] = p.bias ? __ldg(p.bias + c0 + v * CH + j) : 0.f;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const int ow = owb + rr;
          if (ow >= p.Wc) continue;
          float o[CH] = {acc2[rr][0].x + b[0], acc2[rr][0].y + b[1], acc2[rr][1].x + b[2], acc2[rr][1].y + b[3]};
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < CH; ++j) o[j] = hsw(o[j]);
          }
          const uint2 ov = pack4<AT>(o);
          *reinterpret_cast<uint2*>(p.out + obase + (int64_t)ow * p.C) = ov;
          if (p.pool) {
            float rv[CH];
            unpack4<AT>(ov, rv);
#pragma unroll
            for (int j = 0; j < CH; ++j) psum[j] += (p.act == 2) ? hsw(rv[j]) : rv[j];
          }
        }
      } else {
        float sc[CH], sh[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
        if (p.stat) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            sc[j] = __ldg(p.stat + 2 * p.C + c0 + v * CH + j);
            sh[j] = __ldg(p.stat + 3 * p.C + c0 + v * CH + j);
          }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const int ow = owb + rr;
          if (ow >= p.Wc) continue;
          float o[CH] = {acc2[rr][0].x, acc2[rr][0].y, acc2[rr][1].x, acc2[rr][1].y};
          if (p.dres) {
            float r4[CH];
            unpack4<grad_t>(ldg_u2(p.dres + obase + (int64_t)ow * p.C), r4);
            o[0] += r4[0]; o[1] += r4[1]; o[2] += r4[2]; o[3] += r4[3];
          }
          if (p.a_pre) {
            float a4[CH];
            unpack4<AT>(ldg_u2(p.a_pre + obase + (int64_t)ow * p.C), a4);
#pragma unroll
            for (int j = 0; j < CH; ++j) o[j] *= hsw_grad(fmaf(a4[j], sc[j], sh[j]));
          }
          *reinterpret_cast<uint2*>(p.out + obase + (int64_t)ow * p.C) = pack4<grad_t>(o);
        }
      }
    }
  }
  if (MODE == 1) {
    // ---- border pass of the data gradient (reflection fold) ----
    // Position i collects the full correlation G[q] at every padded position q that reflects onto it:
    // q = i, q = -i (1 <= i <= PAD), q = 2(X-1)-i (X-1-PAD <= i <= X-2).  Only tiles that touch such rows /
    // columns do anything here; taps outside the staged patch are zeros by construction.
    const bool touches = (oh0 <= PAD) || (oh0 + g.TH - 1 >= p.Hc - 1 - PAD) || (ow0 <= PAD) ||
                         (ow0 + g.TW - 1 >= p.Wc - 1 - PAD);
    if (touches) {
      __syncthreads();   // the main loop's stores to these pixels are ordered before the updates below
      // Border rows / columns of this tile as (start, count) bands in tile-local coordinates; the affected pixels
      // are enumerated compactly (band rows x all columns, then the other rows x band columns) so that the
      // work spreads over all 256 threads instead of the few whose strips happen to lie on the border.
      const int ch = min(g.TH, p.Hc - oh0), cw = min(g.TW, p.Wc - ow0);
      const int rt0 = max(1, oh0), rt1 = min(PAD, oh0 + ch - 1);
      const int rb0 = max(p.Hc - 1 - PAD, oh0), rb1 = min(p.Hc - 2, oh0 + ch - 1);
      const int nrt = max(0, rt1 - rt0 + 1), nrb = max(0, rb1 - rb0 + 1), nr = nrt + nrb;
      const int ct0 = max(1, ow0), ct1 = min(PAD, ow0 + cw - 1);
      const int cb0 = max(p.Wc - 1 - PAD, ow0), cb1 = min(p.Wc - 2, ow0 + cw - 1);
      const int nct = max(0, ct1 - ct0 + 1), ncb = max(0, cb1 - cb0 + 1), nc = nct + ncb;
      const int nA = nr * cw, nB = (ch - nr) * nc;
      const int total = (nA + nB) * g.hvn;
      for (int i = threadIdx.x; i < total; i += kThreads) {
        const int vv = i % g.hvn;
        int e = i / g.hvn;
        int oh, ow;
        if (e < nA) {
          const int ri = e / cw;
          ow = ow0 + (e - ri * cw);
          oh = ri < nrt ? rt0 + ri : rb0 + (ri - nrt);
        } else {
          e -= nA;
          const int j = e / nc, ci = e - j * nc;
          ow = ci < nct ? ct0 + ci : cb0 + (ci - nct);
          oh = oh0 + j;                                   // j-th row of the tile that is NOT in a band
          if (nrt && oh >= rt0) oh += nrt;
          if (nrb && oh >= rb0) oh += nrb;
        }
        int qh[3], qw[3];
        int nh = 0, nw = 0;
        qh[nh++] = oh;
        if (oh >= 1 && oh <= PAD) qh[nh++] = -oh;
        if (oh <= p.Hc - 2 && 2 * (p.Hc - 1) - oh <= p.Hc - 1 + PAD) qh[nh++] = 2 * (p.Hc - 1) - oh;
        qw[nw++] = ow;
        if (ow >= 1 && ow <= PAD) qw[nw++] = -ow;
        if (ow <= p.Wc - 2 && 2 * (p.Wc - 1) - ow <= p.Wc - 1 + PAD) qw[nw++] = 2 * (p.Wc - 1) - ow;
        // the direct term (q = position itself) was produced by the main loop: add only the reflected ones
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
        for (int a = 0; a < nh; ++a)
          for (int b = (a == 0 ? 1 : 0); b < nw; ++b)
            for (int kh = 0; kh < K; ++kh) {
              const int pr = qh[a] - oh0 + kh;
              if (pr < 0 || pr >= PH) continue;
              for (int kw = 0; kw < K; ++kw) {
                const int pc = qw[b] - ow0 + kw;
                if (pc < 0 || pc >= PW) continue;
                float2 xv[2], wv[2];
                unpack2x2<IO>(patch[pr * RP + pc * g.hvn + vv], xv);
                ld_w2x2(s_w + (kh * K + kw) * CB + vv * CH, wv);
                a0 = __ffma2_rn(xv[0], wv[0], a0);
                a1 = __ffma2_rn(xv[1], wv[1], a1);
              }
            }
        float o[CH] = {a0.x, a0.y, a1.x, a1.y};
        const int64_t off = (((int64_t)n * p.Hc + oh) * p.Wc + ow) * p.C + c0 + vv * CH;
        if (p.a_pre) {
          float a4[CH];
          unpack4<AT>(ldg_u2(p.a_pre + off), a4);
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float scj = p.stat ? __ldg(p.stat + 2 * p.C + c0 + vv * CH + j) : 1.f;
            const float shj = p.stat ? __ldg(p.stat + 3 * p.C + c0 + vv * CH + j) : 0.f;
            o[j] *= hsw_grad(fmaf(a4[j], scj, shj));
          }
        }
        uint2* op = reinterpret_cast<uint2*>(p.out + off);
        float pv[CH];
        unpack4<grad_t>(*op, pv);
#pragma unroll
        for (int j = 0; j < CH; ++j) pv[j] += o[j];
        *op = pack4<grad_t>(pv);
      }
    }
  }
  if (MODE == 0 && p.pool) {
    if (worker < g.workers) {
#pragma unroll
      for (int j = 0; j < CH; ++j) atomicAdd(s_pool + v * CH + j, psum[j]);
    }
  }
  __syncthreads();   // tile consumed (its buffer may be refilled next iteration); s_pool complete
  if (MODE == 0 && p.pool && threadIdx.x < CB)
    atomicAdd(p.pool + (int64_t)n * p.C + c0 + threadIdx.x, s_pool[threadIdx.x]);
  }  // tile loop
}

// ---- weight gradient: dW[c][kh][kw] += sum_{n,o} dy[n,o,c] * x[n, R(o + k - pad), c] -------------------
// A thread owns 2 channels (one 4-byte bf16 pair) and ALL k*k taps: k*k float2 accumulators live in registers
// across a strided set of tiles, the dy tile and the reflected x patch are staged once per tile, and a strip
// of 8 outputs re-uses every LDS'ed x value for up to k taps.  One shared-memory reduction over the workers
// and one atomic per (channel, tap) per CTA at the end.  blockIdx.y = channel block.
struct WParams {
  const uint16_t* dy;   // [N][Hc][Wc][C]  bf16 gradient
  const uint16_t* x;    // [N][H][W][C]    fp16 activation
  float* dw;                 // (C,1,K,K) fp32, accumulated
  int N, C, H, W, Hc, Wc, up2;
  Geom g;                    // hvn = channel PAIRS per CTA here
};

// pitch in 4-byte units, == pn (mod 32): rows-fastest workers give one bank per lane
__host__ __device__ __forceinline__ int row_pitch32(int cols, int pn) {
  const int raw = cols * pn;
  return raw + (((pn - raw) % 32) + 32) % 32;
}

template <typename AT, int K>
__global__ void __launch_bounds__(kThreads, 3)
dw_wgrad_tiled_kernel(const WParams p) {
  extern __shared__ __align__(16) uint32_t smem_u1[];
  constexpr int PAD = (K - 1) / 2;
  const Geom g = p.g;
  const int pn = g.hvn;
  const int PH = g.TH + K - 1, PW = g.TW + K - 1;
  const int RPX = row_pitch32(PW, pn), RPD = row_pitch32(g.TW, pn);
  const int buf_words = PH * RPX + g.TH * RPD;   // one staged tile: x patch [PH] rows of RPX + dy [TH] rows of RPD
  const int CB = pn * 2;
  const int c0 = blockIdx.y * CB;
  const int worker = threadIdx.x / pn;
  const int v = threadIdx.x - worker * pn;
  const int groups_w = g.TW / R;
  const int items = g.TH * groups_w;
  const int tiles_per_img = g.tiles_h * g.tiles_w;
  const int64_t total_tiles = (int64_t)p.N * tiles_per_img;
  float2 acc[K][K];
#pragma unroll
  for (int kh = 0; kh < K; ++kh)
#pragma unroll
    for (int kw = 0; kw < K; ++kw) acc[kh][kw] = make_float2(0.f, 0.f);

  // Double-buffered tile loop: the cp.async copies of tile i+1 are in flight while tile i is consumed.
  auto stage = [&](int64_t t, int b) {
    const int n = (int)(t / tiles_per_img);
    const int tile = (int)(t - (int64_t)n * tiles_per_img);
    const int th = tile / g.tiles_w, tw = tile - th * g.tiles_w;
    const int oh0 = th * g.TH, ow0 = tw * g.TW;
    uint32_t* bx = smem_u1 + b * buf_words;
    uint32_t* bd = bx + PH * RPX;
    const uint16_t* src = p.x + (int64_t)n * p.H * p.W * p.C + c0;
    stage_tile_async(reinterpret_cast<uint4*>(bx), PH, PW, pn / 4, RPX / 4, p.x, [&](int py, int px) -> const uint16_t* {
      int y = reflect_idx(clampi(oh0 - PAD + py, -PAD, p.Hc - 1 + PAD), p.Hc);
      int x = reflect_idx(clampi(ow0 - PAD + px, -PAD, p.Wc - 1 + PAD), p.Wc);
      if (p.up2) { y >>= 1; x >>= 1; }
      return src + ((int64_t)y * p.W + x) * p.C;
    });
    const uint16_t* dsrc = p.dy + (int64_t)n * p.Hc * p.Wc * p.C + c0;
    stage_tile_async(reinterpret_cast<uint4*>(bd), g.TH, g.TW, pn / 4, RPD / 4, p.x, [&](int py, int px) -> const uint16_t* {
      const int y = oh0 + py, x = ow0 + px;
      if (y >= p.Hc || x >= p.Wc) return nullptr;   // outputs outside the image contribute nothing
      return dsrc + ((int64_t)y * p.Wc + x) * p.C;
    });
    cp_async_commit();
  };
  if ((int64_t)blockIdx.x < total_tiles) stage(blockIdx.x, 0);
  int buf = 0;
  for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, buf ^= 1) {
    const int64_t tn = t + gridDim.x;
    if (tn < total_tiles) {
      stage(tn, buf ^ 1);      // buffer buf^1 was released by the barrier that ended the previous iteration
      cp_async_wait<1>();      // everything but the group just committed has landed
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t* s_x = smem_u1 + buf * buf_words;
    const uint32_t* s_dy = s_x + PH * RPX;
    if (worker < g.workers) {
      for (int item = worker; item < items; item += g.workers) {
        const int cg = item / g.TH, r = item - cg * g.TH;   // rows fastest
        float2 d[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const uint32_t u = s_dy[r * RPD + (cg * R + rr) * pn + v];
          d[rr] = H16<grad_t>::un2(u);
        }
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
          const uint32_t* prow = s_x + (r + kh) * RPX + cg * R * pn + v;
#pragma unroll
          for (int c = 0; c < R + K - 1; ++c) {
            const uint32_t u = prow[c * pn];
            const float2 xv = H16<AT>::un2(u);
#pragma unroll
            for (int kw = 0; kw < K; ++kw) {
              const int rr = c - kw;
              if (rr >= 0 && rr < R) acc[kh][kw] = __ffma2_rn(d[rr], xv, acc[kh][kw]);
            }
          }
        }
      }
    }
    __syncthreads();   // tile consumed: its buffer may be refilled by the next iteration's prefetch
  }
  // ---- reduce over the workers of each channel pair, then one atomic per (channel, tap) ----
  __syncthreads();
  float* s_red = reinterpret_cast<float*>(smem_u1);   // [workers][K*K][CB]
  if (worker < g.workers) {
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        s_red[(worker * K * K + kh * K + kw) * CB + v * 2] = acc[kh][kw].x;
        s_red[(worker * K * K + kh * K + kw) * CB + v * 2 + 1] = acc[kh][kw].y;
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * CB; i += kThreads) {
    float s = 0.f;
    for (int wk = 0; wk < g.workers; ++wk) s += s_red[wk * K * K * CB + i];
    const int tap = i / CB, c = i - tap * CB;
    atomicAdd(p.dw + (int64_t)(c0 + c) * K * K + tap, s);
  }
}

// Host: tile geometry for C channels (C % 8 == 0).
static size_t fwd_smem(const Geom& g, int k) {
  return 2 * ((size_t)(g.TH + k - 1) * row_pitch(g.TW + k - 1, g.hvn) * 8 + 8) + (size_t)k * k * g.hvn * CH * 4 + g.hvn * CH * 4;
}
static size_t wgrad_smem(const Geom& g, int k) {   // g.hvn = channel pairs
  const size_t tiles = 2 * ((size_t)(g.TH + k - 1) * row_pitch32(g.TW + k - 1, g.hvn) +
                            (size_t)g.TH * row_pitch32(g.TW, g.hvn)) * 4;   // double buffered
  const size_t red = (size_t)g.workers * k * k * g.hvn * 2 * 4;
  return tiles > red ? tiles : red;
}
static bool pick_geom(int C, int Hc, int Wc, int k, size_t smem_limit, bool wgrad, Geom* out) {
  const int hv = wgrad ? C / 2 : C / CH;          // wgrad: channel pairs; else 4-channel vectors
  const int hvn_max = wgrad ? 64 : 24;
  double best = -1.0;
  Geom bg = {};
  for (int hvn = 1; hvn <= hvn_max && hvn <= hv; ++hvn) {
    if (hv % hvn) continue;
    if (hvn % (wgrad ? 4 : 2)) continue;          // whole 16-byte chunks per pixel (staging granularity)
    const int workers_max = kThreads / hvn;
    for (int slack = 0; slack <= 2; ++slack) {
      const int workers = workers_max - slack;
      if (workers < 1) break;
      for (int gw = 1; gw <= 6; ++gw) {            // strips per tile row: TW = 8 * gw
        for (int TH = 4; TH <= 32; ++TH) {
          const int items = TH * gw;
          if (items % workers) continue;
          if (items / workers > (wgrad ? 16 : 4)) continue;
          Geom g = {};
          g.hvn = hvn; g.workers = workers; g.TH = TH; g.TW = gw * R;
          g.tiles_h = (Hc + TH - 1) / TH; g.tiles_w = (Wc + g.TW - 1) / g.TW; g.cblocks = hv / hvn;
          if ((wgrad ? wgrad_smem(g, k) : fwd_smem(g, k)) > smem_limit) continue;
          const double thread_eff = (double)(workers * hvn) / kThreads;
          const double halo_eff = (double)(TH * g.TW) / ((TH + k - 1) * (g.TW + k - 1));
          const double edge_eff = (double)Hc * Wc / ((double)g.tiles_h * TH * g.tiles_w * g.TW);
          const int bytes = hvn * (wgrad ? 4 : 8);   // contiguous HBM bytes per pixel and CTA
          const double wide = bytes >= 64 ? 1.0 : (bytes >= 32 ? 0.95 : (bytes >= 16 ? 0.85 : 0.7));
          const double score = thread_eff * halo_eff * edge_eff * wide;
          if (score > best) { best = score; bg = g; }
        }
      }
    }
  }
  if (best < 0) return false;
  *out = bg;
  return true;
}

constexpr size_t kSmemLimit = 72 * 1024;      // data / weight gradient: 3 CTAs per SM, two patch buffers each
constexpr size_t kSmemLimitFwd = 100 * 1024;  // forward (heavier epilogue, 128 registers): 2 CTAs per SM, larger tiles

template <typename AT, int K, int MODE>
static int launch_tiled_t(const Params& p, int N, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(dw_tiled_kernel<AT, K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemLimitFwd);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  const Geom& g = p.g;
  const size_t smem = fwd_smem(g, K);
  const int64_t total_tiles = (int64_t)N * g.tiles_h * g.tiles_w;
  int64_t gx = (148 * (MODE == 0 ? 2 : 3)) / g.cblocks;   // never more CTAs than fit at once: a partial second wave doubles the time
  if (gx < 1) gx = 1;
  if (gx > total_tiles) gx = total_tiles;
  dim3 grid((unsigned)gx, g.cblocks, 1);
  dw_tiled_kernel<AT, K, MODE><<<grid, kThreads, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
template <int K, int MODE>
static int launch_tiled(const Params& p, int N, cudaStream_t s) {
  return act_format() == AST_DT_F16 ? launch_tiled_t<__half, K, MODE>(p, N, s) : launch_tiled_t<__nv_bfloat16, K, MODE>(p, N, s);
}

template <typename AT, int K>
static int launch_wgrad_t(const WParams& p, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(dw_wgrad_tiled_kernel<AT, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemLimit);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  const Geom& g = p.g;
  const size_t smem = wgrad_smem(g, K);
  const int64_t total_tiles = (int64_t)p.N * g.tiles_h * g.tiles_w;
  int64_t gx = (148 * 3) / g.cblocks;   // never more CTAs than fit at once: a partial second wave doubles the time
  if (gx < 1) gx = 1;
  if (gx < 1) gx = 1;
  if (gx > total_tiles) gx = total_tiles;
  dim3 grid((unsigned)gx, g.cblocks, 1);
  dw_wgrad_tiled_kernel<AT, K><<<grid, kThreads, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
template <int K>
static int launch_wgrad(const WParams& p, cudaStream_t s) {
  return act_format() == AST_DT_F16 ? launch_wgrad_t<__half, K>(p, s) : launch_wgrad_t<__nv_bfloat16, K>(p, s);
}

}  // namespace dwt
}  // namespace ast

using namespace ast;

// Entry points used by ast_dw_conv / ast_dw_conv_dgrad / ast_dw_conv_wgrad for stride 1.  Return
// AST_E_SHAPE when no tiling fits so that the caller can fall back to the direct kernels.
int dw_tiled_forward(const void* x, const float* w, const float* bias, void* out, float* pool, int N, int C, int H,
                     int W, int k, int up2, int act, cudaStream_t s) {
  dwt::Params p = {};
  p.in = reinterpret_cast<const uint16_t*>(x);
  p.w = w; p.bias = bias; p.out = reinterpret_cast<uint16_t*>(out); p.pool = pool;
  p.N = N; p.C = C; p.H = H; p.W = W; p.Hc = up2 ? 2 * H : H; p.Wc = up2 ? 2 * W : W; p.up2 = up2; p.act = act;
  if (N > 65535 || !dwt::pick_geom(C, p.Hc, p.Wc, k, dwt::kSmemLimitFwd, false, &p.g)) return AST_E_SHAPE;
  return k == 3 ? dwt::launch_tiled<3, 0>(p, N, s) : dwt::launch_tiled<5, 0>(p, N, s);
}

int dw_tiled_dgrad(const void* dy, const float* w, const void* a_pre, const float* stat, const void* dres, void* dx,
                   int N, int C, int H, int W, int k, cudaStream_t s) {
  dwt::Params p = {};
  p.in = reinterpret_cast<const uint16_t*>(dy);
  p.w = w; p.out = reinterpret_cast<uint16_t*>(dx);
  p.a_pre = reinterpret_cast<const uint16_t*>(a_pre); p.stat = stat;
  p.dres = reinterpret_cast<const uint16_t*>(dres);
  p.N = N; p.C = C; p.H = H; p.W = W; p.Hc = H; p.Wc = W;
  if (N > 65535 || H < 2 * k || W < 2 * k) return AST_E_SHAPE;   // tiny maps: the direct kernel handles them
  if (!dwt::pick_geom(C, H, W, k, dwt::kSmemLimit, false, &p.g)) return AST_E_SHAPE;
  return k == 3 ? dwt::launch_tiled<3, 1>(p, N, s) : dwt::launch_tiled<5, 1>(p, N, s);
}

int dw_tiled_wgrad(const void* dy, const void* x, float* dw, int N, int C, int H, int W, int k, int up2,
                   cudaStream_t s) {
  dwt::WParams p = {};
  p.dy = reinterpret_cast<const uint16_t*>(dy);
  p.x = reinterpret_cast<const uint16_t*>(x);
  p.dw = dw; p.N = N; p.C = C; p.H = H; p.W = W; p.Hc = up2 ? 2 * H : H; p.Wc = up2 ? 2 * W : W; p.up2 = up2;
  if (!dwt::pick_geom(C, p.Hc, p.Wc, k, dwt::kSmemLimit, true, &p.g)) return AST_E_SHAPE;
  return k == 3 ? dwt::launch_wgrad<3>(p, s) : dwt::launch_wgrad<5>(p, s);
}
