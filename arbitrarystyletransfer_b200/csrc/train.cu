// Backward-pass helpers on the native layout (bf16 [N][H+2h][W+2h][C], h = halo width) for the
// decoder training step (BASELINE config 2).  They replace what autograd derives for the
// reference modules: nn.ReLU / nn.MaxPool2d backward inside torchvision VGG-19 (models.py:186-240)
// and nn.ReLU / nn.Upsample(nearest x2) / nn.ReflectionPad2d(1) backward inside the classic decoder
// (models.py:598-628).  The convolution gradients themselves run on the tensor cores:
// data gradient = ast_conv3x3_fwd with flipped weights, weight gradient = ast_conv3x3_wgrad.
#include "common.cuh"

namespace ast {

constexpr int kT = 256;

__device__ __forceinline__ uint4 ldv(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// ---- 2x2/2 max pool forward (training forward keeps the un-pooled activation) ---------------------
__global__ void __launch_bounds__(kT) maxpool2_native_kernel(const __nv_bfloat16* __restrict__ in,
                                                             __nv_bfloat16* __restrict__ out, int N,
                                                             int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2, cv = C / 8;
  const int64_t total = (int64_t)N * Ho * Wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int v = (int)(i % cv);
    int64_t r = i / cv;
    const int w = (int)(r % Wo); r /= Wo;
    const int h = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float x[8];
        Vec16<true>::unpack(ldv(in + (((int64_t)n * (H + 2) + 2 * h + a + 1) * (W + 2) + 2 * w + b + 1) * C + v * 8), x);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], x[j]);
      }
    *reinterpret_cast<uint4*>(out + (((int64_t)n * (Ho + 2) + h + 1) * (Wo + 2) + w + 1) * C + v * 8) =
        Vec16<true>::pack(m);
  }
}

// ---- VGG layer backward prep ------------------------------------------------------------------------
// dZ = [Y > 0] * ( route(G) + tap_post ) + tap_pre      at the conv resolution H x W, where
// route(G)[h][w] = G[h][w]                              (no pool), or
//                = G[h/2][w/2] if Y[h][w] is the FIRST maximum of its 2x2 window (torch's rule)
// Y: un-pooled post-ReLU activation (1-halo); G: gradient w.r.t. the layer output (1-halo, pooled
// size when `pooled`), may be NULL; tap_post / tap_pre: gradients of fp32 taps already converted to
// the native layout (1-halo), may be NULL.  dZ has halo width `dz_halo` (its halo is never written).
__global__ void __launch_bounds__(kT)
vgg_bwd_prep_kernel(const __nv_bfloat16* __restrict__ Y, const __nv_bfloat16* __restrict__ G,
                    const __nv_bfloat16* __restrict__ tap_post, const __nv_bfloat16* __restrict__ tap_pre,
                    __nv_bfloat16* __restrict__ dZ, int N, int C, int H, int W, int pooled, int dz_halo) {
  const int cv = C / 8;
  const int Hg = pooled ? H / 2 : H, Wg = pooled ? W / 2 : W;
  const int64_t total = (int64_t)N * H * W * cv;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int v = (int)(i % cv);
    int64_t r = i / cv;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const int64_t yoff = (((int64_t)n * (H + 2) + h + 1) * (W + 2) + w + 1) * C + v * 8;
    float y[8], g[8];
    Vec16<true>::unpack(ldv(Y + yoff), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    if (G) {
      if (!pooled) {
        Vec16<true>::unpack(ldv(G + yoff), g);
      } else if ((h >> 1) < Hg && (w >> 1) < Wg) {
        float gg[8];
        Vec16<true>::unpack(ldv(G + (((int64_t)n * (Hg + 2) + (h >> 1) + 1) * (Wg + 2) + (w >> 1) + 1) * C + v * 8), gg);
        // first maximum in row-major window order wins
        const int h0 = h & ~1, w0 = w & ~1, me = (h & 1) * 2 + (w & 1);
        float best[8];
        int arg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float x[8];
          Vec16<true>::unpack(ldv(Y + (((int64_t)n * (H + 2) + h0 + (k >> 1) + 1) * (W + 2) + w0 + (k & 1) + 1) * C + v * 8), x);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (x[j] > best[j]) { best[j] = x[j]; arg[j] = k; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = (arg[j] == me) ? gg[j] : 0.f;
      }
    }
    if (tap_post) {
      float t[8];
      Vec16<true>::unpack(ldv(tap_post + yoff), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = y[j] > 0.f ? g[j] : 0.f;
    if (tap_pre) {
      float t[8];
      Vec16<true>::unpack(ldv(tap_pre + yoff), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
    const int64_t zoff = (((int64_t)n * (H + 2 * dz_halo) + h + dz_halo) * (W + 2 * dz_halo) + w + dz_halo) * C + v * 8;
    *reinterpret_cast<uint4*>(dZ + zoff) = Vec16<true>::pack(g);
  }
}

// ---- decoder layer backward fold ----------------------------------------------------------------------
// dXpad: gradient w.r.t. the reflection-PADDED input of decoder conv i, i.e. the output of the
//        data-gradient conv run over the (Hi+2) x (Wi+2) grid, stored 1-halo: [N][Hi+4][Wi+4][C].
// Fold the reflection (pad row -1 -> row 1, row Hi -> row Hi-2, same for columns), undo the nearest
// x2 upsample of the producing layer (sum of the 2x2 block) when `up`, apply that layer's ReLU mask
// (its stored output Xi > 0) and write dZ of the producing layer with a 2-pixel zero halo:
// [N][Hc+4][Wc+4][C], (Hc, Wc) = (Hi, Wi) / (up ? 2 : 1).
__device__ __forceinline__ int fold_sources(int u, int X, int (&s)[3]) {
  int n = 0;
  s[n++] = u;
  if (u == 1) s[n++] = -1;
  if (u == X - 2) s[n++] = X;
  return n;
}

__global__ void __launch_bounds__(kT)
dec_bwd_fold_kernel(const __nv_bfloat16* __restrict__ dXpad, const __nv_bfloat16* __restrict__ Xi,
                    __nv_bfloat16* __restrict__ dZ, int N, int C, int Hi, int Wi, int up, int relu) {
  const int cv = C / 8;
  const int Hc = up ? Hi / 2 : Hi, Wc = up ? Wi / 2 : Wi;
  const int64_t total = (int64_t)N * Hc * Wc * cv;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int v = (int)(i % cv);
    int64_t r = i / cv;
    const int wc = (int)(r % Wc); r /= Wc;
    const int hc = (int)(r % Hc);
    const int n = (int)(r / Hc);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int reps = up ? 2 : 1;
    for (int a = 0; a < reps; ++a)
      for (int b = 0; b < reps; ++b) {
        const int u = up ? 2 * hc + a : hc, w = up ? 2 * wc + b : wc;
        int su[3], sw[3];
        const int nu = fold_sources(u, Hi, su), nw = fold_sources(w, Wi, sw);
        for (int x = 0; x < nu; ++x)
          for (int y = 0; y < nw; ++y) {
            // padded coordinate p in [-1, Hi] sits at physical index p + 2 (pad grid +1, halo +1)
            float t[8];
            Vec16<true>::unpack(ldv(dXpad + (((int64_t)n * (Hi + 4) + su[x] + 2) * (Wi + 4) + sw[y] + 2) * C + v * 8), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += t[j];
          }
      }
    if (relu) {
      const int u = up ? 2 * hc : hc, w = up ? 2 * wc : wc;
      float m[8];
      Vec16<true>::unpack(ldv(Xi + (((int64_t)n * (Hi + 2) + u + 1) * (Wi + 2) + w + 1) * C + v * 8), m);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = m[j] > 0.f ? acc[j] : 0.f;
    }
    *reinterpret_cast<uint4*>(dZ + (((int64_t)n * (Hc + 4) + hc + 2) * (Wc + 4) + wc + 2) * C + v * 8) =
        Vec16<true>::pack(acc);
  }
}

// ---- native -> channel-planar (the K-major operands of the weight-gradient GEMM) ----------------------
// planar[s][c][q], q = (n*(H+2) + ph)*wp + pw over the 1-halo padded grid with the row pitch wp
// (a multiple of 8 pixels >= W+2, so that row shifts stay 16-byte aligned: TMA tiled loads need an
// aligned innermost coordinate).  Columns pw >= W+2 are zero.  The source has halo width src_halo;
// the planar halo is copied from the source when copy_halo, else zero.  nshift = 3 writes the three
// column-shifted copies planar[s][c][q] = x[c][q + s - 1] the kw = 0,1,2 taps read.
// Tile = 64 consecutive q x 64 channels (+ one q on each side for the shifted copies): 16-byte loads
// along the channels, transpose through shared memory, 16-byte stores along q (q0 is a multiple of
// 8, so the stores of all three copies are aligned; the +-1 shift is applied while reading the tile).
constexpr int PT_Q = 64, PT_C = 64;

__global__ void __launch_bounds__(256)
native_to_planar_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N,
                        int C, int H, int W, int src_halo, int copy_halo, int wp, int nshift) {
  __shared__ __align__(16) __nv_bfloat16 tile[PT_Q + 2][PT_C + 8];  // [q - q0 + 1][channel], padded rows
  const int Hp = H + 2;
  const int64_t ldq = (int64_t)N * Hp * wp;
  const int64_t q0 = (int64_t)blockIdx.x * PT_Q;
  const int c0 = blockIdx.y * PT_C;
  const int Ws = W + 2 * src_halo, Hs = H + 2 * src_halo;
  // load (PT_Q + 2) q-rows x 8 channel-vectors
  for (int i = threadIdx.x; i < (PT_Q + 2) * (PT_C / 8); i += 256) {
    const int r = i / (PT_C / 8), v = i % (PT_C / 8);
    const int64_t q = q0 + r - 1;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (q >= 0 && q < ldq && c0 + v * 8 < C) {
      const int pw = (int)(q % wp);
      const int ph = (int)((q / wp) % Hp);
      const int n = (int)(q / ((int64_t)wp * Hp));
      const bool halo = ph == 0 || pw == 0 || ph == Hp - 1 || pw == W + 1;
      if (pw < W + 2 && (!halo || copy_halo))
        val = ldv(src + (((int64_t)n * Hs + ph - 1 + src_halo) * Ws + pw - 1 + src_halo) * C + c0 + v * 8);
    }
    *reinterpret_cast<uint4*>(&tile[r][v * 8]) = val;
  }
  __syncthreads();
  // store: thread -> (channel, group of 8 consecutive q); copy s holds x[q + s - 1]
  for (int i = threadIdx.x; i < PT_C * (PT_Q / 8); i += 256) {
    const int c = i / (PT_Q / 8), g = i % (PT_Q / 8);
    const int64_t q = q0 + g * 8;
    if (c0 + c >= C || q >= ldq) continue;
    for (int sft = 0; sft < nshift; ++sft) {
      const int off = nshift == 1 ? 1 : sft;  // tile row of x[q + s - 1] is (q - q0) + s
      __align__(16) __nv_bfloat16 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = tile[g * 8 + j + off][c];
      __nv_bfloat16* d = dst + ((int64_t)(nshift == 1 ? 0 : sft) * C + c0 + c) * ldq + q;
      *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// packed fp32 gradient [9][Cout][Cin] -> OIHW fp32 [Cout][Cin][3][3]; also the bias gradient as the
// row sums of the planar dZ.
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwpk, float* __restrict__ g, int Cout,
                                    int Cin, int accumulate) {
  const int64_t total = (int64_t)Cout * Cin * 9;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % 9);
    const int64_t r = i / 9;
    const int ci = (int)(r % Cin), co = (int)(r / Cin);
    const float v = dwpk[((int64_t)t * Cout + co) * Cin + ci];
    g[i] = accumulate ? g[i] + v : v;
  }
}

__global__ void __launch_bounds__(256) planar_rowsum_kernel(const __nv_bfloat16* __restrict__ x,
                                                            float* __restrict__ out, int64_t ldq,
                                                            int accumulate) {
  __shared__ float s_red[8];
  const __nv_bfloat16* row = x + (int64_t)blockIdx.x * ldq;
  float acc = 0.f;
  const int64_t nv = ldq / 8;
  for (int64_t i = threadIdx.x; i < nv; i += 256) {
    float v[8];
    Vec16<true>::unpack(ldv(row + i * 8), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < 8; ++w) r += s_red[w];
    out[blockIdx.x] = accumulate ? out[blockIdx.x] + r : r;
  }
}

// Zero the halo ring (width hw) of a native tensor [N][H+2hw][W+2hw][C]: what a zero-padded conv reads around an
// interior that its producer rewrites completely (replaces a memset of the whole tensor).
__global__ void __launch_bounds__(kT) zero_halo_kernel(__nv_bfloat16* __restrict__ t, int N, int C, int H, int W, int hw) {
  const int cv = C >> 3;
  const int Hp = H + 2 * hw, Wp = W + 2 * hw;
  const int64_t ring = (int64_t)2 * hw * Wp + (int64_t)2 * hw * H;   // pixels of the ring: full top / bottom rows + side columns
  const int64_t total = (int64_t)N * ring * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    int64_t r = i / cv;
    const int n = (int)(r / ring);
    r -= (int64_t)n * ring;
    int y, x;
    if (r < (int64_t)2 * hw * Wp) {
      const int row = (int)(r / Wp);
      x = (int)(r - (int64_t)row * Wp);
      y = row < hw ? row : H + row;               // rows hw .. 2hw-1 of the ring are the bottom rows H+hw .. H+2hw-1
    } else {
      r -= (int64_t)2 * hw * Wp;
      const int row = (int)(r / (2 * hw));
      const int c = (int)(r - (int64_t)row * 2 * hw);
      y = hw + row;
      x = c < hw ? c : W + c;
    }
    reinterpret_cast<uint4*>(t + (((int64_t)n * Hp + y) * Wp + x) * C)[v] = make_uint4(0u, 0u, 0u, 0u);
  }
}

static unsigned grid_for(int64_t total) {
  int64_t nb = (total + kT - 1) / kT;
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  return (unsigned)nb;
}

}  // namespace ast

using namespace ast;

extern "C" int ast_maxpool2_native(const void* in, void* out, int N, int C, int H, int W, void* stream) {
  if (!in || !out || N <= 0 || C <= 0 || H < 2 || W < 2) return AST_E_BADARG;
  if (C % 8 != 0) return AST_E_SHAPE;
  maxpool2_native_kernel<<<grid_for((int64_t)N * (H / 2) * (W / 2) * (C / 8)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), N, C, H, W);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_zero_halo(void* native, int N, int C, int H, int W, int halo, void* stream) {
  if (!native || N <= 0 || C <= 0 || H <= 0 || W <= 0 || halo < 1 || halo > 2) return AST_E_BADARG;
  if (C % 8 != 0) return AST_E_SHAPE;
  if (!aligned16(native)) return AST_E_ALIGN;
  const int64_t ring = (int64_t)2 * halo * (W + 2 * halo) + (int64_t)2 * halo * H;
  zero_halo_kernel<<<grid_for((int64_t)N * ring * (C / 8)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(native), N, C, H, W, halo);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_vgg_bwd_prep(const void* Y, const void* G, const void* tap_post, const void* tap_pre,
                                void* dZ, int N, int C, int H, int W, int pooled, int dz_halo,
                                void* stream) {
  if (!Y || !dZ || N <= 0 || C <= 0 || H <= 0 || W <= 0 || dz_halo < 1 || dz_halo > 2) return AST_E_BADARG;
  if (C % 8 != 0) return AST_E_SHAPE;
  vgg_bwd_prep_kernel<<<grid_for((int64_t)N * H * W * (C / 8)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(Y), reinterpret_cast<const __nv_bfloat16*>(G),
      reinterpret_cast<const __nv_bfloat16*>(tap_post), reinterpret_cast<const __nv_bfloat16*>(tap_pre),
      reinterpret_cast<__nv_bfloat16*>(dZ), N, C, H, W, pooled, dz_halo);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_dec_bwd_fold(const void* dXpad, const void* Xi, void* dZ, int N, int C, int Hi, int Wi,
                                int up, int relu, void* stream) {
  if (!dXpad || !dZ || (relu && !Xi) || N <= 0 || C <= 0 || Hi < 2 || Wi < 2) return AST_E_BADARG;
  if (C % 8 != 0 || (up && ((Hi | Wi) & 1))) return AST_E_SHAPE;
  const int Hc = up ? Hi / 2 : Hi, Wc = up ? Wi / 2 : Wi;
  dec_bwd_fold_kernel<<<grid_for((int64_t)N * Hc * Wc * (C / 8)), kT, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dXpad), reinterpret_cast<const __nv_bfloat16*>(Xi),
      reinterpret_cast<__nv_bfloat16*>(dZ), N, C, Hi, Wi, up, relu);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_native_to_planar(const void* native, void* planar, int N, int C, int H, int W,
                                    int src_halo, int copy_halo, int wp, int nshift, void* stream) {
  if (!native || !planar || N <= 0 || C <= 0 || H <= 0 || W <= 0 || src_halo < 1 || src_halo > 2)
    return AST_E_BADARG;
  if (wp < W + 2 || wp % 8 != 0 || (nshift != 1 && nshift != 3)) return AST_E_SHAPE;
  if (C % 8 != 0 || !aligned16(native) || !aligned16(planar)) return AST_E_SHAPE;
  const int64_t ldq = (int64_t)N * (H + 2) * wp;
  const int64_t gx = (ldq + PT_Q - 1) / PT_Q;
  if (gx >= 0x7fffffffLL || (C + PT_C - 1) / PT_C > 65535) return AST_E_SHAPE;
  dim3 grid((unsigned)gx, (C + PT_C - 1) / PT_C);
  native_to_planar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(native), reinterpret_cast<__nv_bfloat16*>(planar), N, C, H, W,
      src_halo, copy_halo, wp, nshift);
  AST_CHECK_LAUNCH();
  return 0;
}

extern "C" int ast_unpack_wgrad(const float* dwpk, float* w_grad, const void* dz_planar, float* b_grad,
                                int Cout, int Cin, int64_t ldq, int accumulate, void* stream) {
  if (!dwpk || !w_grad || Cout <= 0 || Cin <= 0) return AST_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  unpack_wgrad_kernel<<<grid_for((int64_t)Cout * Cin * 9), kT, 0, s>>>(dwpk, w_grad, Cout, Cin, accumulate);
  AST_CHECK_LAUNCH();
  if (b_grad) {
    if (!dz_planar || ldq <= 0 || ldq % 8 != 0) return AST_E_BADARG;
    planar_rowsum_kernel<<<Cout, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dz_planar), b_grad, ldq,
                                              accumulate);
    AST_CHECK_LAUNCH();
  }
  return 0;
}
