// K2: 3x3 stride-1 convolution as an implicit GEMM on tcgen05 / TMEM, operands fed by TMA.
//
// Reference work replaced (paths relative to /root/reference):
//   VGG-19 features   models.py:186-240  nn.Conv2d(k3, zero pad 1, bias) + ReLU + MaxPool2d(2,2)
//   classic decoder   models.py:598-628  ReflectionPad2d(1) + Conv2d(k3) + ReLU + Upsample(x2)
//
// Data layout: activations are bf16 [N][H+2][W+2][C] with a one-pixel halo that already holds the
// padding of the consuming conv (zeros for VGG, the reflection for the decoder), so a 3x3 tap is a
// plain shifted box.  GEMM view: M = output pixels (tile = 8 rows x 16 cols = 128), N = Cout
// (BN = 64/128/256 per CTA), K = 9 taps x Cin in steps of 64 channels.
//   A (128 x 64)  one 4-D TMA box {64 ch, 16 w, 8 h, 1 n} at (cb*64, w0+kw, h0+kh, n): pixel rows of
//                 128 B, 128B-swizzled -> exactly the canonical K-major SW128 UMMA operand.
//   B (BN x 64)   one 3-D TMA box {64 ci, BN co, 1 tap} of the packed weights [9][Cout][Cin].
//   D (128 x BN)  fp32 in TMEM, double buffered (2*BN columns) so the epilogue of tile i overlaps
//                 the MMAs of tile i+1.
// Warp roles (256 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> bias -> ReLU
// -> {plain | 2x2 max-pool via shuffles | nearest x2 upsample} -> bf16 -> global, plus the
// reflection halo of the output when the next layer is a decoder conv).
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"

namespace ast {
namespace tc {

static EncodeTiledFn g_encode = nullptr;

EncodeTiledFn get_encode_tiled() {
  if (g_encode) return g_encode;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    (void)cudaGetLastError();
    return nullptr;
  }
  g_encode = (EncodeTiledFn)fn;
  return g_encode;
}

int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return AST_E_NODRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                   gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : AST_E_SHAPE;
}

constexpr int TILE_H = 8, TILE_W = 16, TILE_M = TILE_H * TILE_W, KBLK = 64;
constexpr int A_STAGE_BYTES = TILE_M * KBLK * 2;  // 16 KB
constexpr int kConvThreads = 256;

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * KBLK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // multiple of 1024
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);  // <= 192 KB of operands
  static constexpr int TMEM_COLS = 2 * BN;  // 32 / 128 / 256 / 512: powers of two >= 32
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + align slack
};

struct ConvParams {
  int N, H, W, Cin, Cout;
  int Ho, Wo;
  int relu, halo, tap_prerelu;
  int tiles_w, tiles_h, n_blocks, num_tiles;
  const float* bias;
  __nv_bfloat16* out;
  float* tap;
  float* out_nchw;   // EPI_NCHW32 only: fp32 [N][cout_real][H][W]
  int cout_real;     // EPI_NCHW32 only: channels actually stored (<= BN)
  int clamp01;       // EPI_NCHW32 only: Hardtanh(0,1) (models.py:304, 315)
  long long* dbg;    // optional [gridDim.x][8] wait-cycle counters per role (AST_CONV_DEBUG=1)
  uint32_t epi_sleep_ns, prod_sleep_ns;   // back-off of the waiting roles (mbar_wait_sleep); 0 = tight polling
  int wide_a;        // CTA-pair kernels: one {64, 10, 18} A box per (tile, block) instead of one {64, 8, 18} box per kw
  int pf_dist;       // CTA-pair kernel: L2 prefetch distance in tiles per CTA (0 = off)
  int tma_store;     // CTA-pair kernel, plain epilogue, BN <= 128: staging bytes reserved for the TMA-store epilogue (0 = off)
  int na, nbs;       // CTA-pair kernels: A ring slots and weight slots of this launch (pair_plan)
  int dbg_flags;     // AST_CONV_DBGFLAGS (bottleneck elimination, results are WRONG): 1 = operands loaded once per ring
                     // slot, never refreshed; 2 = epilogue without global stores; 4 = epilogue without TMEM loads
};

// Debug instrumentation (wait counters, bottleneck-elimination flags) is compiled in only with -DAST_KERNEL_DEBUG=1
// (AST_KERNEL_DEBUG=1 python -m arbitrarystyletransfer_b200._build): tested at run time, the flags cost the pointwise
// kernel 40 % and the fused conv1_1 + conv1_2 kernel 15 % (register pressure, lost unrolling, maybe-uninitialised
// values), so product builds fold them to constants.
#ifndef AST_KERNEL_DEBUG
#define AST_KERNEL_DEBUG 0
#endif
template <typename P> __device__ __forceinline__ int kdbg_flags(const P& p) { return AST_KERNEL_DEBUG ? p.dbg_flags : 0; }
template <typename P> __device__ __forceinline__ long long* kdbg_buf(const P& p) { return AST_KERNEL_DEBUG ? p.dbg : nullptr; }

constexpr int EPI_NCHW32 = 3;  // internal: last decoder layer, fp32 NCHW image out

// wait + optional accounting of the cycles spent waiting (debug instrumentation)
__device__ __forceinline__ void mbar_wait_acc(uint32_t bar, uint32_t parity, bool dbg, long long& acc,
                                              uint32_t sleep_ns = 0) {
  if (!dbg) {
    if (sleep_ns) mbar_wait_sleep(bar, parity, sleep_ns); else mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

// Persistent-loop tile cursor: tile -> (n-block, tile column, tile row, image), advanced by
// gridDim.x with mixed-radix carries instead of four integer divisions per tile (the epilogue warps
// are instruction-latency bound, every instruction per tile counts).
struct TileCursor {
  int nb, twi, thi, n;
  int g0, g1, g2, g3;
  int NB, TWc, THc, mult;
  // CTA pairs (mult = 2): the work index counts (spatial tile PAIR, n-block); this CTA's spatial tile is
  // 2 * pair + rank, so a carry out of the n-block digit advances the spatial digits by 2.
  __device__ __forceinline__ void init(const ConvParams& p, int tile, int stride, int mult_ = 1, int rank = 0) {
    NB = p.n_blocks; TWc = p.tiles_w; THc = p.tiles_h; mult = mult_;
    int t = tile;
    nb = t % NB; t = mult * (t / NB) + rank;
    twi = t % TWc; t /= TWc;
    thi = t % THc; n = t / THc;
    t = stride;
    g0 = t % NB; t = mult * (t / NB);
    g1 = t % TWc; t /= TWc;
    g2 = t % THc; g3 = t / THc;
  }
  __device__ __forceinline__ void next() {
    nb += g0;
    int c = 0;
    if (nb >= NB) { nb -= NB; c = mult; }
    twi += g1 + c;
    c = 0;
    while (twi >= TWc) { twi -= TWc; ++c; }
    thi += g2 + c;
    c = 0;
    while (thi >= THc) { thi -= THc; ++c; }
    n += g3 + c;
  }
};

// Output coordinates (unpadded grid, -1 and Xo are the halo) that conv coordinate x feeds.
// halo: AST_HALO_REFLECT = ReflectionPad2d(1) of the output grid (-1 <- 1, Xo <- Xo-2);
//       AST_HALO_CLAMP   = replicate (-1 <- 0, Xo <- Xo-1): what the reflection of the x2-upsampled grid is
//                          in low-resolution coordinates (consumer: conv3x3_fold_kernel).
template <int EPI>
__device__ __forceinline__ int out_targets(int x, int Xo, int halo, int (&t)[4]) {
  int n = 0;
  if (EPI == AST_EPI_PLAIN) {
    t[n++] = x;
  } else if (EPI == AST_EPI_UP2) {
    t[n++] = 2 * x;
    t[n++] = 2 * x + 1;
  } else {
    t[n++] = x >> 1;
  }
  if (halo != AST_HALO_KEEP) {
    const int lo = halo == AST_HALO_REFLECT ? 1 : 0, hi = halo == AST_HALO_REFLECT ? Xo - 2 : Xo - 1;
    const int m = n;
    for (int i = 0; i < m; ++i) {
      if (t[i] == lo) t[n++] = -1;
      if (t[i] == hi) t[n++] = Xo;
    }
  }
  return n;
}


// Epilogue warps (4 per CTA; warp e may touch TMEM lanes [32e, 32e+32) = tile rows 2e, 2e+1):
// tcgen05.ld -> +bias -> (tap) -> ReLU -> (tap) -> bf16 -> {plain | 2x2 max-pool | nearest x2} store
// with the optional reflection halo, or the fp32 NCHW image for the last decoder layer.
// The 4 * NG * TG epilogue warps form NG * TG groups of four (one warp per TMEM lane quarter).  NG groups split a
// tile's COLUMNS; TG sets of them work on DIFFERENT tiles at once (set tg takes the CTA's tiles tg, tg + TG, ... and
// so the accumulator stages (tg + k TG) mod NACC): a tile's epilogue is one warp's latency chain (barrier ->
// tcgen05.ld -> bias -> pack -> stores -> release), and with TG = 1 only one tile is in that chain at a time however
// many warps share it -- the limit for conv1_1 (whose work IS the epilogue) and, once CTA pairs had made the MMAs
// cheap, for the 64- and 128-channel layers.
// P2 (CTA pair, conv_pair.cuh): the work index counts tile pairs, this CTA takes spatial tile 2 * pair + rank, and the
// accumulator is handed back on the LEADER's tempty barrier (a remote arrive for rank 1).
// TS (TMA store, plain epilogue, BN <= 128, TG = 1): the 4 * NG warps pack the tile into a staging buffer of BN / 64
// boxes {64 ch, TW w, TH h} in the 128-byte-swizzle layout and one lane issues the cp.async.bulk.tensor stores (see
// epilogue_first_store); lanes that own border pixels still write the reflection / clamp halo copies themselves.
template <int BN, int EPI, int TW = TILE_W, int NG = 1, int NACC = 2, int TG = 1, bool P2 = false, bool TS = false>
__device__ __forceinline__ void epilogue_loop(const ConvParams& p, uint32_t tmem_base, int ew, int lane,
                                              uint32_t tfull_bar0, uint32_t tempty_bar0, int rank = 0,
                                              const CUtensorMap* tmOut = nullptr, uint32_t staging = 0) {
  static_assert(!TS || (EPI == AST_EPI_PLAIN && TG == 1 && BN % 64 == 0), "TMA store: plain tiles, one tile set");
  constexpr int CH = (TG > 1 && BN <= 64) ? 16 : ((BN >= 32 && BN / 32 >= NG) ? 32 : 16);  // columns per tcgen05.ld
  constexpr int TH = TILE_M / TW;         // tile = TH rows x TW cols of pixels, row-major in M
  constexpr int NCH = BN / CH;
  // warp ew owns TMEM lane quarter e = ew % 4 (hardware rule: a warp may only touch lanes [32*(warp%4), +32)) and,
  // inside its tile set, the column chunks g, g+NG, ...
  static_assert(NACC % TG == 0 && (NACC & (NACC - 1)) == 0, "stages are handed round-robin to the tile sets");
  const int e = ew & 3, g = (ew >> 2) % NG, tg = (ew >> 2) / NG;
  const uint32_t tempty_leader0 = P2 ? mapa_shared(tempty_bar0, 0) : 0u;
  const int units = P2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;          // CTAs or CTA pairs
  const int t0 = (P2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) + tg * units;
  const int tstride = TG * units;
  const int hl = (32 * e + lane) / TW;
  const int wl = (32 * e + lane) % TW;
  const int halo = p.halo;
  const bool wide_st = (reinterpret_cast<uintptr_t>(p.out) & 31u) == 0 && (p.Cout % 16) == 0;
  int as = tg % NACC;
  uint32_t aphase = 0;
  TileCursor cur;
  cur.init(p, t0, tstride, P2 ? 2 : 1, rank);
  int ts_buf = 0;
  long long dbg_wait = 0;
  const long long dbg_t0 = clock64();
  for (int tile = t0; tile < p.num_tiles; tile += tstride, cur.next()) {
    const int nb = cur.nb, twi = cur.twi, thi = cur.thi, n = cur.n;
    const int h = thi * TH + hl, w = twi * TW + wl;
    const bool in_img = (h < p.H) && (w < p.W) && (!P2 || n < p.N);   // an odd tile count leaves rank 1 a phantom tile

    int rows[4], cols[4], nr = 0, nc = 0;
    bool owner;
    if (EPI == AST_EPI_POOL2) {
      owner = ((hl & 1) == 0) && ((wl & 1) == 0) && ((h >> 1) < p.Ho) && ((w >> 1) < p.Wo) && (!P2 || n < p.N);
    } else {
      owner = in_img;
    }
    if (owner && EPI != EPI_NCHW32) {
      nr = out_targets<EPI>(h, p.Ho, halo, rows);
      nc = out_targets<EPI>(w, p.Wo, halo, cols);
    }

    mbar_wait_acc(tfull_bar0 + 8u * as, aphase, kdbg_buf(p) != nullptr, dbg_wait, p.epi_sleep_ns);
    tc_fence_after();
    if (TS) {   // two staging buffers alternate: the stores issued two tiles ago have finished reading this one
      if (ew == 0 && lane == 0) bulk_wait_group_read1();
      named_bar_sync(1, 128 * NG);
    }
    const uint32_t stg = TS ? staging + (ts_buf ? (uint32_t)(BN / 64) * (TILE_M * 128) : 0u) : 0u;
    const uint32_t trow = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(as * BN);
    uint32_t vnext[CH];
    const bool skip_ld = (kdbg_flags(p) & 4) != 0;
    if (g < NCH && !skip_ld) tmem_ld_cols(trow + g * CH, vnext);
#pragma unroll 1
    for (int chunk = g; chunk < NCH && !skip_ld; chunk += NG) {
      uint32_t v[CH];
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = vnext[i];
      if (chunk + NG < NCH) tmem_ld_cols(trow + (chunk + NG) * CH, vnext);  // prefetch next chunk
      const int ch0 = nb * BN + chunk * CH;
      if constexpr (EPI == EPI_NCHW32) {
        if (in_img) {
          for (int c = 0; c < p.cout_real; ++c) {
            float val = __uint_as_float(v[c]) + (p.bias ? __ldg(p.bias + c) : 0.f);
            if (p.relu) val = fmaxf(val, 0.f);
            if (p.clamp01) val = fminf(fmaxf(val, 0.f), 1.f);
            p.out_nchw[(((int64_t)n * p.cout_real + c) * p.H + h) * p.W + w] = val;
          }
        }
      } else {
        float f[CH];
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
          float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
          f[i + 0] = __uint_as_float(v[i + 0]) + b.x;
          f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
          f[i + 2] = __uint_as_float(v[i + 2]) + b.z;
          f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
        }
        if (p.tap && p.tap_prerelu && in_img) {
          float* tp = p.tap + (((int64_t)n * p.Cout + ch0) * p.H + h) * p.W + w;
#pragma unroll
          for (int i = 0; i < CH; ++i) tp[(int64_t)i * p.H * p.W] = f[i];
        }
        const bool post_tap = p.tap && !p.tap_prerelu;
        uint32_t pk[CH / 2];
        if (p.relu && !post_tap) {      // the common case: ReLU folded into the bf16 conversion
#pragma unroll
          for (int i = 0; i < CH / 2; ++i) pk[i] = pack_bf16_relu(f[2 * i], f[2 * i + 1]);
        } else {
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < CH; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (post_tap && in_img) {
            float* tp = p.tap + (((int64_t)n * p.Cout + ch0) * p.H + h) * p.W + w;
#pragma unroll
            for (int i = 0; i < CH; ++i) tp[(int64_t)i * p.H * p.W] = f[i];
          }
#pragma unroll
          for (int i = 0; i < CH / 2; ++i) pk[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
        }
        if (EPI == AST_EPI_POOL2) {
          // 2x2 max: partner along w is lane^1, along h is lane^TW (a warp holds 32/TW full tile rows).
          // max commutes with the (monotonic) bf16 rounding, so pool the packed values.
#pragma unroll
          for (int i = 0; i < CH / 2; ++i) {
            __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&pk[i]);
            uint32_t o1 = __shfl_xor_sync(0xffffffffu, pk[i], 1);
            a = __hmax2_nan(a, *reinterpret_cast<__nv_bfloat162*>(&o1));
            uint32_t cur = *reinterpret_cast<uint32_t*>(&a);
            uint32_t o2 = __shfl_xor_sync(0xffffffffu, cur, TW);
            a = __hmax2_nan(a, *reinterpret_cast<__nv_bfloat162*>(&o2));
            pk[i] = *reinterpret_cast<uint32_t*>(&a);
          }
        }
        if (TS) {
          const int m = 32 * e + lane, cl = chunk * CH;        // tile row, first channel of the chunk in the tile
          const uint32_t rowa = stg + (uint32_t)(cl >> 6) * (TILE_M * 128) + (uint32_t)m * 128u;
#pragma unroll
          for (int q = 0; q < CH / 8; ++q) {
            const uint32_t c16 = (uint32_t)(((cl & 63) >> 3) + q) ^ (uint32_t)(m & 7);
            st_shared_v4(rowa + c16 * 16u, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        if (p.out && !(kdbg_flags(p) & 2)) {
          for (int ri = 0; ri < nr; ++ri) {
            for (int ci = 0; ci < nc; ++ci) {
              if (TS && ri == 0 && ci == 0) continue;      // the pixel itself goes out with the TMA store
              __nv_bfloat16* o = p.out +
                  (((int64_t)n * (p.Ho + 2) + (rows[ri] + 1)) * (p.Wo + 2) + (cols[ci] + 1)) * p.Cout + ch0;
              if (wide_st) {
#pragma unroll
                for (int q = 0; q < CH / 16; ++q) st_global_v8(o + 16 * q, &pk[8 * q]);
              } else {
                uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
                for (int q = 0; q < CH / 8; ++q)
                  o4[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (P2) mbar_arrive_cluster(tempty_leader0 + 8u * as);
      else mbar_arrive(tempty_bar0 + 8u * as);
    }
    if (TS) {
      fence_proxy_async_smem();
      named_bar_sync(1, 128 * NG);
      if (ew == 0 && lane == 0 && !(kdbg_flags(p) & 2)) {
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_store_4d(tmOut, stg + (uint32_t)b * (TILE_M * 128), nb * BN + b * 64, twi * TW, thi * TH, n);
        bulk_commit_group();
      }
      ts_buf ^= 1;
    }
    as += TG;
    if (as >= NACC) { as -= NACC; aphase ^= 1u; }
  }
  if (TS && ew == 0 && lane == 0) bulk_wait_group0();
  if (kdbg_buf(p) && ew == 0 && lane == 0) {
    kdbg_buf(p)[blockIdx.x * 8 + 4] = dbg_wait;               // epilogue warp 0: waiting for an accumulator
    kdbg_buf(p)[blockIdx.x * 8 + 5] = clock64() - dbg_t0;     // epilogue warp 0: total loop time
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B operands need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int cblocks = p.Cin / KBLK;
  const int ksteps = 9 * cblocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int t = tile;
        const int nb = t % p.n_blocks; t /= p.n_blocks;
        const int twi = t % p.tiles_w; t /= p.tiles_w;
        const int thi = t % p.tiles_h;
        const int n = t / p.tiles_h;
        const int h0 = thi * TILE_H, w0 = twi * TILE_W;
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap - 3 * kh;
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
            const uint32_t a_dst = base + stage * C::STAGE_BYTES;
            tma_load_4d(a_dst, &tmA, full_bar(stage), cb * KBLK, w0 + kw, h0 + kh, n);
            tma_load_3d(a_dst + A_STAGE_BYTES, &tmB, full_bar(stage), cb * KBLK, nb * BN, tap);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = base + stage * C::STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < KBLK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle atom
            const uint64_t adesc = make_sdesc_k128(a_addr + k * 32);
            const uint64_t bdesc = make_sdesc_k128(b_addr + k * 32);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees this smem stage once the MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    epilogue_loop<BN, EPI>(p, tmem_base, warp - 4, lane, tfull_bar(0), tempty_bar(0));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// conv3x3_tc2_kernel: same GEMM, 2.7x less A traffic.  Layers with few channels are bound by
// L2 -> shared-memory bandwidth when every tap re-loads its own shifted 128-pixel box (9 x 16 KB
// per tile and 64-channel block).  Here the tile is 16 rows x 8 columns, and ONE box of
// {64 ch, 8 w, 18 h} per kw serves the three kh taps: the tap's operand starts kh * 8 pixels
// = kh * 1024 B further into the box, which keeps the 1024-byte alignment the 128B swizzle needs,
// so the UMMA descriptor is the standard one with a different start address.  A traffic drops
// from 144 KB to 54 KB per (tile, channel block).  A boxes and per-tap weight tiles ride separate
// mbarrier rings; when the layer has a single 64-channel block and one N block, all nine weight
// tiles are loaded once per CTA and stay resident.
constexpr int T2_W = 8, T2_H = 16, T2_BOX_H = T2_H + 2;
constexpr int A2_BYTES = T2_BOX_H * T2_W * KBLK * 2;  // 18 KB

template <int BN>
struct Cfg2 {
  static constexpr int B_BYTES = BN * KBLK * 2;
  static constexpr int NA = (BN >= 128) ? 4 : 8;
  static constexpr int NB = (BN == 256) ? 4 : 9;       // weight tap slots
  // BN <= 128: the nine tap slots form three groups (one per kw, holding kh = 0..2): one barrier
  // pair per group instead of per tap, or no barrier traffic at all when the weights stay resident.
  static constexpr bool GROUPED = BN <= 128;
  static constexpr int NACC = (BN <= 128) ? 4 : 2;  // TMEM accumulator stages (<= 512 columns)
  static constexpr int TMEM_COLS = (NACC * BN < 32) ? 32 : NACC * BN;
  static constexpr int NBAR = 2 * NA + 2 * NB + 2 * NACC;
  static constexpr int SMEM_BYTES = NA * A2_BYTES + NB * B_BYTES + NBAR * 8 + 16 + 1024;
};

// warps 0-3: TMA / MMA / TMEM alloc / idle; then 4 * NG epilogue warps.  N <= 128 uses 16 epilogue
// warps (4 per SM sub-partition): their per-tile work is short and latency bound, more resident
// warps is what hides it; N = 256 keeps 8 (tensor-bound, fewer registers spent).
template <int BN>
struct Epi2 {
  static constexpr int NG = (BN == 256) ? 2 : 4;
  static constexpr int THREADS = 128 + 128 * NG;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(Epi2<BN>::THREADS, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const ConvParams p) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + C::NA * A2_BYTES;
  const uint32_t bars = b_base + C::NB * C::B_BYTES;
  auto afull = [&](int s) { return bars + 8u * s; };
  auto aempty = [&](int s) { return bars + 8u * (C::NA + s); };
  auto bfull = [&](int s) { return bars + 8u * (2 * C::NA + s); };
  auto bempty = [&](int s) { return bars + 8u * (2 * C::NA + C::NB + s); };
  auto tfull = [&](int s) { return bars + 8u * (2 * C::NA + 2 * C::NB + s); };
  auto tempty = [&](int s) { return bars + 8u * (2 * C::NA + 2 * C::NB + C::NACC + s); };
  const uint32_t tmem_slot = bars + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + C::NA * A2_BYTES + C::NB * C::B_BYTES + 8 * C::NBAR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int cblocks = p.Cin / KBLK;
  const bool resident = (C::NB >= 9) && cblocks == 1 && p.n_blocks == 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < C::NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int s = 0; s < C::NACC; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 4 * Epi2<BN>::NG); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (resident) {
        // tap (kh,kw) lives in slot kw*3 + kh (group kw), loaded once per CTA
        for (int kw = 0; kw < 3; ++kw) {
          mbar_expect_tx(bfull(kw), 3 * C::B_BYTES);
          for (int kh = 0; kh < 3; ++kh)
            tma_load_3d(b_base + (kw * 3 + kh) * C::B_BYTES, &tmB, bfull(kw), 0, 0, kh * 3 + kw);
        }
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool a_filled = false, b_filled = false;   // dbg_flags & 1: every ring slot has been loaded once
      long long dbg_pa = 0, dbg_pb = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int t = tile;
        const int nb = t % p.n_blocks; t /= p.n_blocks;
        const int twi = t % p.tiles_w; t /= p.tiles_w;
        const int thi = t % p.tiles_h;
        const int n = t / p.tiles_h;
        const int h0 = thi * T2_H, w0 = twi * T2_W;
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int kw = 0; kw < 3; ++kw) {
            mbar_wait_acc(aempty(sa), pa ^ 1u, kdbg_buf(p) != nullptr, dbg_pa, p.prod_sleep_ns);
            if ((kdbg_flags(p) & 1) && a_filled) {
              mbar_arrive(afull(sa));
            } else {
              mbar_expect_tx(afull(sa), A2_BYTES);
              tma_load_4d(a_base + sa * A2_BYTES, &tmA, afull(sa), cb * KBLK, w0 + kw, h0, n);
            }
            if (sa + 1 == C::NA) a_filled = true;
            if (++sa == C::NA) { sa = 0; pa ^= 1u; }
            if (!resident) {
              if constexpr (C::GROUPED) {
                mbar_wait_acc(bempty(sb), pb ^ 1u, kdbg_buf(p) != nullptr, dbg_pb, p.prod_sleep_ns);
                if ((kdbg_flags(p) & 1) && b_filled) {
                  mbar_arrive(bfull(sb));
                } else {
                  mbar_expect_tx(bfull(sb), 3 * C::B_BYTES);
                  for (int kh = 0; kh < 3; ++kh)
                    tma_load_3d(b_base + (sb * 3 + kh) * C::B_BYTES, &tmB, bfull(sb), cb * KBLK, nb * BN,
                                kh * 3 + kw);
                }
                if (sb == 2) b_filled = true;
                if (++sb == 3) { sb = 0; pb ^= 1u; }
              } else {
                for (int kh = 0; kh < 3; ++kh) {
                  mbar_wait_acc(bempty(sb), pb ^ 1u, kdbg_buf(p) != nullptr, dbg_pb);
                  mbar_expect_tx(bfull(sb), C::B_BYTES);
                  tma_load_3d(b_base + sb * C::B_BYTES, &tmB, bfull(sb), cb * KBLK, nb * BN, kh * 3 + kw);
                  if (++sb == C::NB) { sb = 0; pb ^= 1u; }
                }
              }
            }
          }
        }
      }
      if (kdbg_buf(p)) { kdbg_buf(p)[blockIdx.x * 8 + 0] = dbg_pa; kdbg_buf(p)[blockIdx.x * 8 + 1] = dbg_pb; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp stays converged (all lanes wait on the barriers) and ONE elected lane issues:
    // with warp-uniform operands the tcgen05.mma / commit instructions take their descriptors from
    // uniform registers directly.  Descriptors advance by adding constants to one precomputed
    // 64-bit value (start-address field = bits [0,14) in 16-byte units; shared memory is < 256 KB so
    // the add never carries out of the field).
    {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN);
      const uint64_t a_desc0 = make_sdesc_k128(a_base);
      const uint64_t b_desc0 = make_sdesc_k128(b_base);
      constexpr uint64_t A_SLOT16 = A2_BYTES >> 4, B_SLOT16 = C::B_BYTES >> 4, KH16 = (T2_W * KBLK * 2) >> 4;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int as = 0;
      uint32_t aphase = 0;
      bool b_resident_ready = false;
      long long dbg_mt = 0, dbg_ma = 0, dbg_mb = 0;
      const long long dbg_m0 = clock64();
      const bool dbg = kdbg_buf(p) != nullptr;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait_acc(tempty(as), aphase ^ 1u, dbg, dbg_mt);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        uint32_t accum = 0;
        for (int cb = 0; cb < cblocks; ++cb) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            mbar_wait_acc(afull(sa), pa, dbg, dbg_ma);
            const uint64_t ad = a_desc0 + (uint64_t)sa * A_SLOT16;
            if constexpr (C::GROUPED) {
              // one weight group (kh = 0..2 of this kw) per A box: 12 MMAs between barrier operations
              int grp;
              if (resident) {
                grp = kw;
                if (!b_resident_ready) mbar_wait(bfull(kw), 0u);  // first tile only
              } else {
                grp = sb;
                mbar_wait_acc(bfull(sb), pb, dbg, dbg_mb);
              }
              tc_fence_after();
              const uint64_t bd = b_desc0 + (uint64_t)(grp * 3) * B_SLOT16;
              if (elect_one_sync()) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                  for (int k = 0; k < KBLK / 16; ++k) {
                    umma_bf16(d_tmem, ad + (uint64_t)(kh * KH16 + k * 2), bd + (uint64_t)(kh * B_SLOT16 + k * 2),
                              idesc, (kh | k) ? 1u : accum);
                  }
                }
                if (!resident) umma_commit(bempty(sb));
                umma_commit(aempty(sa));
              }
              __syncwarp();
              accum = 1u;
              if (!resident) {
                if (++sb == 3) { sb = 0; pb ^= 1u; }
              }
            } else {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                mbar_wait_acc(bfull(sb), pb, dbg, dbg_mb);
                tc_fence_after();
                const uint64_t bd = b_desc0 + (uint64_t)sb * B_SLOT16;
                if (elect_one_sync()) {
#pragma unroll
                  for (int k = 0; k < KBLK / 16; ++k)
                    umma_bf16(d_tmem, ad + (uint64_t)(kh * KH16 + k * 2), bd + (uint64_t)(k * 2), idesc,
                              k ? 1u : accum);
                  umma_commit(bempty(sb));
                  if (kh == 2) umma_commit(aempty(sa));
                }
                __syncwarp();
                accum = 1u;
                if (++sb == C::NB) { sb = 0; pb ^= 1u; }
              }
            }
            if (++sa == C::NA) { sa = 0; pa ^= 1u; }
          }
        }
        b_resident_ready = true;
        if (elect_one_sync()) umma_commit(tfull(as));
        __syncwarp();
        if (++as == C::NACC) { as = 0; aphase ^= 1u; }
      }
      if (kdbg_buf(p) && lane == 0) {
        kdbg_buf(p)[blockIdx.x * 8 + 2] = dbg_ma + dbg_mb;        // MMA warp: waiting for operands
        kdbg_buf(p)[blockIdx.x * 8 + 3] = dbg_mt;                 // MMA warp: waiting for a free accumulator
        kdbg_buf(p)[blockIdx.x * 8 + 6] = clock64() - dbg_m0;     // MMA warp: total loop time
      }
    }
  } else if (warp >= 4) {
    epilogue_loop<BN, EPI, T2_W, Epi2<BN>::NG, C::NACC>(p, tmem_base, warp - 4, lane, tfull(0), tempty(0));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int EPI>
static int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p,
                      int sm_count, cudaStream_t s) {
  using C = Cfg2<BN>;
  auto kern = conv3x3_tc2_kernel<BN, EPI>;
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  kern<<<grid, Epi2<BN>::THREADS, C::SMEM_BYTES, s>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}

template <int BN>
static int launch_tc2_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const ConvParams& p, int sm_count, cudaStream_t s) {
  switch (epi) {
    case AST_EPI_PLAIN: return launch_tc2<BN, AST_EPI_PLAIN>(tmA, tmB, p, sm_count, s);
    case AST_EPI_POOL2: return launch_tc2<BN, AST_EPI_POOL2>(tmA, tmB, p, sm_count, s);
    case AST_EPI_UP2: return launch_tc2<BN, AST_EPI_UP2>(tmA, tmB, p, sm_count, s);
  }
  return AST_E_BADARG;
}

template <int BN, int EPI>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p,
                     int sm_count, cudaStream_t s) {
  using C = Cfg<BN>;
  auto kern = conv3x3_tc_kernel<BN, EPI>;
  static bool attr_done = false;  // per (BN, EPI) instantiation
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  kern<<<grid, kConvThreads, C::SMEM_BYTES, s>>>(tmA, tmB, p);
  AST_CHECK_LAUNCH();
  return 0;
}

template <int BN>
static int launch_tc_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const ConvParams& p, int sm_count, cudaStream_t s) {
  switch (epi) {
    case AST_EPI_PLAIN: return launch_tc<BN, AST_EPI_PLAIN>(tmA, tmB, p, sm_count, s);
    case AST_EPI_POOL2: return launch_tc<BN, AST_EPI_POOL2>(tmA, tmB, p, sm_count, s);
    case AST_EPI_UP2: return launch_tc<BN, AST_EPI_UP2>(tmA, tmB, p, sm_count, s);
  }
  return AST_E_BADARG;
}

#include "conv_pair.cuh"
#include "conv_fold.cuh"

bool tc_supported(const ast_conv_desc* d) {
  return d->Cin % 64 == 0 && d->Cout % 64 == 0 && d->H >= 2 && d->W >= 2;
}

static int get_sm_count(int* out);

// Tensor maps for one launch.  kwbox = 1: A box {64, 8, 18, 1} (conv3x3_tc2_kernel);
// kwbox = 0: A box {64, 16, 8, 1} (conv3x3_tc_kernel).  Weights [9][rows][Cin], box {64, BN, 1}.
static int make_maps(CUtensorMap* tmA, CUtensorMap* tmB, const void* in, const void* wpk, int N, int H,
                     int W, int Cin, int wrows, int BN, int kwbox, int ntaps = 9, int wide_a = 0) {
  const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W + 2, (uint64_t)H + 2, (uint64_t)N};
  const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)(W + 2) * Cin * 2,
                           (uint64_t)(H + 2) * (W + 2) * Cin * 2};
  const uint32_t box_tap[4] = {KBLK, TILE_W, TILE_H, 1};
  const uint32_t box_kw[4] = {KBLK, (uint32_t)(wide_a ? T2_W + 2 : T2_W), T2_BOX_H, 1};
  int r = encode_bf16_map(tmA, in, 4, dims, str, kwbox ? box_kw : box_tap);
  if (r) return r;
  const uint64_t wdims[3] = {(uint64_t)Cin, (uint64_t)wrows, (uint64_t)ntaps};
  const uint64_t wstr[2] = {(uint64_t)Cin * 2, (uint64_t)wrows * Cin * 2};
  const uint32_t wbox[3] = {KBLK, (uint32_t)BN, 1};
  return encode_bf16_map(tmB, wpk, 3, wdims, wstr, wbox);
}

// impl decoding: AST_CONV_AUTO / AST_CONV_TC -> kw-box kernel, automatic N block;
// AST_CONV_TC_TAPBOX -> per-tap-box kernel; 64/128/256 force the N block of the kw-box kernel,
// 1064/1128/1256 of the per-tap-box kernel (tuning / A-B tests).
int conv3x3_tc(const ast_conv_desc* d, const void* in, const void* wpk, const float* bias, void* out,
               float* tap, cudaStream_t s) {
  if (!tc_supported(d)) return AST_E_SHAPE;
  if (!aligned16(in) || !aligned16(wpk) || (out && !aligned16(out)) || (bias && !aligned16(bias)))
    return AST_E_ALIGN;
  int sm_count = 0;
  int r = get_sm_count(&sm_count);
  if (r) return r;
  int impl = d->impl;
  int kwbox = 1;
  if (impl == AST_CONV_TC_TAPBOX) { kwbox = 0; impl = AST_CONV_TC; }
  bool pair_forced = false;
  if (impl >= 2000) { pair_forced = true; impl -= 2000; }
  else if (impl >= 1000) { kwbox = 0; impl -= 1000; }
  ConvParams p = {};
  static const int dbg_flags = getenv("AST_CONV_DBGFLAGS") ? atoi(getenv("AST_CONV_DBGFLAGS")) : 0;
  p.dbg_flags = dbg_flags;
  static const int epi_sleep = getenv("AST_CONV_EPI_SLEEP") ? atoi(getenv("AST_CONV_EPI_SLEEP")) : 0;
  static const int prod_sleep = getenv("AST_CONV_PROD_SLEEP") ? atoi(getenv("AST_CONV_PROD_SLEEP")) : 0;
  p.epi_sleep_ns = (uint32_t)epi_sleep; p.prod_sleep_ns = (uint32_t)prod_sleep;
  static const int pf_env = getenv("AST_CONV_PF") ? atoi(getenv("AST_CONV_PF")) : 0;
  p.pf_dist = pf_env;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  const bool up = d->epilogue == AST_EPI_UP2 || d->epilogue == AST_EPI_UPFOLD;
  p.Ho = d->epilogue == AST_EPI_POOL2 ? d->H / 2 : (up ? 2 * d->H : d->H);
  p.Wo = d->epilogue == AST_EPI_POOL2 ? d->W / 2 : (up ? 2 * d->W : d->W);
  p.relu = d->relu; p.halo = d->halo; p.tap_prerelu = d->tap_prerelu;
  if (d->epilogue == AST_EPI_UPFOLD) {
    // Upsample(x2) -> ReflectionPad2d(1) -> conv as four 2x2 convs on the low-res map (conv_fold.cuh):
    // d->H, d->W are the LOW-res input dims, wpk = ast_pack_conv_weight_fold's [16][Cout][Cin].
    if (!kwbox || tap || !out) return AST_E_BADARG;
    p.tiles_w = (d->W + T2_W - 1) / T2_W;
    p.tiles_h = (d->H + T2_H - 1) / T2_H;
    p.bias = bias; p.out = reinterpret_cast<__nv_bfloat16*>(out);
    const int64_t sp = (int64_t)d->N * p.tiles_h * p.tiles_w;
    int BN = 64;   // tiles per spatial tile: 4 * Cout / 256 for every BN, so the widest N block that divides Cout
    if (d->Cout % 256 == 0) BN = 256;
    else if (d->Cout % 128 == 0) BN = 128;
    if (impl >= 64 && impl <= 256 && d->Cout % impl == 0) BN = impl;
    if (BN != 64 && BN != 128 && BN != 256) return AST_E_SHAPE;
    p.n_blocks = (4 / (256 / BN)) * (d->Cout / BN);
    static const int pair_env_f = getenv("AST_CONV_PAIR") ? atoi(getenv("AST_CONV_PAIR")) : 1;
    const bool pair_f = (pair_forced || pair_env_f) && sp >= 2;
    const int64_t nt = (pair_f ? (sp + 1) / 2 : sp) * p.n_blocks;
    if (nt >= 0x7fffffffLL) return AST_E_SHAPE;
    p.num_tiles = (int)nt;
    CUtensorMap tmA, tmB;
    r = make_maps(&tmA, &tmB, in, wpk, d->N, d->H, d->W, d->Cin, d->Cout, pair_f ? BN / 2 : BN, 1, 16, pair_f ? 1 : 0);
    if (r) return r;
    if (pair_f) {
      switch (BN) {
        case 256: return launch_fold_pair<256>(tmA, tmB, p, sm_count, s);
        case 128: return launch_fold_pair<128>(tmA, tmB, p, sm_count, s);
        default: return launch_fold_pair<64>(tmA, tmB, p, sm_count, s);
      }
    }
    switch (BN) {
      case 256: return launch_fold<256>(tmA, tmB, p, sm_count, s);
      case 128: return launch_fold<128>(tmA, tmB, p, sm_count, s);
      default: return launch_fold<64>(tmA, tmB, p, sm_count, s);
    }
  }
  const int tw = kwbox ? T2_W : TILE_W, th = kwbox ? T2_H : TILE_H;
  p.tiles_w = (d->W + tw - 1) / tw;
  p.tiles_h = (d->H + th - 1) / th;
  p.bias = bias; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.tap = tap;

  // N-block: the widest that still gives every SM a tile (a wide N amortises the A-operand
  // shared-memory reads of each MMA); narrow blocks only when the problem is small.
  const int64_t sp_tiles = (int64_t)d->N * p.tiles_h * p.tiles_w;

  // CTA pairs (conv_pair.cuh): the default for the kw-box kernel once there are at least two spatial tiles.
  // AST_CONV_PAIR=0 keeps the one-CTA kernel (A/B reference); impl 2064 / 2128 / 2256 force the pair kernel's N block.
  static const int pair_env = getenv("AST_CONV_PAIR") ? atoi(getenv("AST_CONV_PAIR")) : 1;
  const bool pair = kwbox && sp_tiles >= 2 &&
                    (pair_forced || (pair_env && (impl == AST_CONV_AUTO || impl == AST_CONV_TC)));
  const int64_t sp_units = pair ? (sp_tiles + 1) / 2 : sp_tiles;      // work units per n-block: tile pairs or tiles
  const int units_wanted = pair ? sm_count / 2 : sm_count;
  int BN = 64;
  if (d->Cout % 256 == 0 && sp_units * (d->Cout / 256) >= units_wanted) BN = 256;
  else if (d->Cout % 128 == 0 && sp_units * (d->Cout / 128) >= units_wanted) BN = 128;
  if (impl >= 64 && impl <= 256 && d->Cout % impl == 0) BN = impl;  // tuning override
  if (BN != 64 && BN != 128 && BN != 256) return AST_E_SHAPE;
  p.n_blocks = d->Cout / BN;
  const int64_t nt = sp_units * p.n_blocks;
  if (nt >= 0x7fffffffLL) return AST_E_SHAPE;
  p.num_tiles = (int)nt;     // pair kernel: counts tile PAIRS x n-blocks

  CUtensorMap tmA, tmB;
  static const int wide_env = getenv("AST_CONV_WIDEA") ? atoi(getenv("AST_CONV_WIDEA")) : 1;
  p.wide_a = pair ? wide_env : 0;
  r = make_maps(&tmA, &tmB, in, wpk, d->N, d->H, d->W, d->Cin, d->Cout, pair ? BN / 2 : BN, kwbox, 9, p.wide_a);
  if (r) return r;
  static const bool dbg_on = getenv("AST_CONV_DEBUG") != nullptr;
  long long* dbg = nullptr;
  if (dbg_on && kwbox) {   // debug instrumentation only: allocates and synchronises
    AST_CUDA(cudaMalloc(&dbg, sizeof(long long) * 8 * sm_count));
    AST_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 8 * sm_count, s));
    p.dbg = dbg;
  }
  struct DbgDump {
    long long* d; int n; const ConvParams& p; int BN; cudaStream_t s; bool pair;
    ~DbgDump() {
      if (!d) return;
      cudaStreamSynchronize(s);
      long long* h = new long long[8 * n];
      cudaMemcpy(h, d, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost);
      double a[8] = {0};
      // pair kernel: one work unit per CTA pair; the MMA counters exist in the leader (even CTAs) only
      const int units = pair ? n / 2 : n;
      const int g = p.num_tiles < units ? p.num_tiles : units;
      const int ctas = pair ? 2 * g : g;
      for (int i = 0; i < ctas; ++i)
        for (int j = 0; j < 8; ++j) {
          const bool leader_only = pair && (j == 2 || j == 3 || j == 6);
          if (leader_only && (i & 1)) continue;
          a[j] += (double)h[i * 8 + j] / (leader_only ? g : ctas);
        }
      const double tiles = (double)p.num_tiles / g;
      fprintf(stderr, "[conv dbg] %sCin=%d Cout=%d H=%d BN=%d tiles/CTA=%.1f | per tile cycles: loop(mma)=%.0f loop(epi)=%.0f | "
              "producer wait A-empty=%.0f B-empty=%.0f | mma wait operands=%.0f accumulator=%.0f | epilogue wait tfull=%.0f\n",
              pair ? "PAIR " : "", p.Cin, p.Cout, p.H, BN, tiles, a[6] / tiles, a[5] / tiles, a[0] / tiles, a[1] / tiles, a[2] / tiles,
              a[3] / tiles, a[4] / tiles);
      delete[] h;
      cudaFree(d);
    }
  } dump{dbg, sm_count, p, BN, s, pair};
  if (pair) {
    switch (BN) {
      case 256: return launch_pair_epi<256>(d->epilogue, tmA, tmB, p, sm_count, s);
      case 128: return launch_pair_epi<128>(d->epilogue, tmA, tmB, p, sm_count, s);
      default: return launch_pair_epi<64>(d->epilogue, tmA, tmB, p, sm_count, s);
    }
  }
  if (kwbox) {
    switch (BN) {
      case 256: return launch_tc2_epi<256>(d->epilogue, tmA, tmB, p, sm_count, s);
      case 128: return launch_tc2_epi<128>(d->epilogue, tmA, tmB, p, sm_count, s);
      default: return launch_tc2_epi<64>(d->epilogue, tmA, tmB, p, sm_count, s);
    }
  }
  switch (BN) {
    case 256: return launch_tc_epi<256>(d->epilogue, tmA, tmB, p, sm_count, s);
    case 128: return launch_tc_epi<128>(d->epilogue, tmA, tmB, p, sm_count, s);
    default: return launch_tc_epi<64>(d->epilogue, tmA, tmB, p, sm_count, s);
  }
}


// ------------------------------------------------------------------------------------------------
// First VGG layer on the tensor cores: Normalization (models.py:129-131) + conv_1 (3 -> 64, zero pad)
// + relu_1 straight from the reference's NCHW fp32 image.  K = 27 is padded to 32; the A tile
// (128 pixels x 32) is built by four producer warps (one pixel per thread: gather 27 taps, normalise,
// round to bf16) directly in the no-swizzle canonical UMMA layout, so the layer costs two MMAs per
// tile and is bound by writing its 64-channel output.
// Warps 0-3 = im2col producers, warps 4-7 = epilogue, warp 8 = TMEM allocator + MMA issuer.
constexpr int kFirstThreads = 288;
constexpr int F_K = 32, F_N = 64, F_STAGES = 4;
constexpr int F_A_BYTES = TILE_M * F_K * 2;   // 8 KB
constexpr int F_B_BYTES = F_N * F_K * 2;      // 4 KB
constexpr int F_LBO = 128, F_SBO = (F_K / 8) * 128;

struct FirstParams {
  const float* img;   // [N][3][H][W]
  const float* w;     // OIHW fp32 [64][3][3][3]
  const float* bias;  // TMA variant only: folded into the GEMM through the K padding (see the kernel)
  float mean[3], rstd[3];
  int normalise;
};

constexpr int F_CTAS_PER_SM = 2;   // 41 KB shared memory, 128 TMEM columns and <= 75 registers per thread each
__global__ void __launch_bounds__(kFirstThreads, F_CTAS_PER_SM)
conv3x3_first_tc_kernel(const FirstParams fp, const ConvParams p) {
  __shared__ __align__(128) uint8_t s_a[F_STAGES][F_A_BYTES];
  __shared__ __align__(128) uint8_t s_b[F_B_BYTES];
  __shared__ __align__(8) uint64_t s_bar[2 * F_STAGES + 4];
  __shared__ uint32_t s_tmem;
  __shared__ float s_patch[2][3 * (TILE_H + 2) * (TILE_W + 2)];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t bars = smem_u32(s_bar);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (F_STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * F_STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * F_STAGES + 2 + s); };

  // weights -> bf16 [64][32] (k = ci*9 + kh*3 + kw, zero for k >= 27) in the canonical layout
  for (int i = threadIdx.x; i < F_N * F_K; i += kFirstThreads) {
    const int co = i / F_K, k = i % F_K;
    const float v = k < 27 ? fp.w[co * 27 + k] : 0.f;
    const uint32_t off = (uint32_t)(co >> 3) * F_SBO + (uint32_t)(k >> 3) * F_LBO + (co & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(s_b + off) = __float2bfloat16_rn(v);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < F_STAGES; ++s) {
      mbar_init(full_bar(s), 4);   // one arrival per producer warp
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<2 * F_N>(smem_u32(&s_tmem));
  fence_proxy_async_smem();   // s_b was written with generic stores, the MMA reads it via the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&s_tmem);

  if (warp < 4) {
    // ===================== im2col producers =====================
    // The 128 producer threads first stage the tile's (8+2) x (16+2) x 3 input patch -- normalised,
    // zero outside the image -- in shared memory with ~4 coalesced loads each (instead of 27 cached
    // loads per pixel), double buffered so the global loads of tile i+1 are in flight while tile i
    // is expanded; each thread then reads its pixel's 27 taps from the patch, rounds to bf16 and
    // stores its 64-byte row of the A tile.
    const int r = threadIdx.x;           // tile row = pixel
    const int hl = r >> 4, wl = r & 15;
    constexpr int PW = TILE_W + 2, PH = TILE_H + 2, PN = 3 * PH * PW;  // 540 floats
    constexpr int PER = (PN + 127) / 128;                                 // 5 loads per thread
    auto fetch = [&](int tile, float (&v)[PER]) {
      int t = tile;
      const int twi = t % p.tiles_w; t /= p.tiles_w;
      const int thi = t % p.tiles_h;
      const int n = t / p.tiles_h;
      const int h0 = thi * TILE_H - 1, w0 = twi * TILE_W - 1;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int e = r + j * 128;
        const int ci = e / (PH * PW), rem = e - ci * (PH * PW);
        const int ih = h0 + rem / PW, iw = w0 + rem % PW;
        const bool ok = e < PN && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
        const int ihc = min(max(ih, 0), p.H - 1), iwc = min(max(iw, 0), p.W - 1), cic = min(ci, 2);
        float x = __ldg(fp.img + (((int64_t)n * 3 + cic) * p.H + ihc) * p.W + iwc);
        if (fp.normalise) x = (x - fp.mean[cic]) * fp.rstd[cic];
        // zero padding applies to the NORMALISED image (models.py:131, then Conv2d padding=1)
        v[j] = ok ? x : 0.f;
      }
    };
    auto stash = [&](int buf, const float (&v)[PER]) {
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int e = r + j * 128;
        if (e < PN) s_patch[buf][e] = v[j];
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    int cur = 0;
    float pre[PER];
    int tile = blockIdx.x;
    if (tile < p.num_tiles) {
      fetch(tile, pre);
      stash(0, pre);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (; tile < p.num_tiles; tile += gridDim.x) {
      const int ntile = tile + gridDim.x;
      if (ntile < p.num_tiles) fetch(ntile, pre);
      uint32_t pk[16];
      {
        float v[28];
        v[27] = 0.f;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              v[ci * 9 + kh * 3 + kw] = s_patch[cur][(ci * PH + hl + kh) * PW + wl + kw];
#pragma unroll
        for (int i = 0; i < 14; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
        pk[14] = 0u;
        pk[15] = 0u;
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      uint8_t* row = &s_a[stage][0] + (uint32_t)(r >> 3) * F_SBO + (r & 7) * 16;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc)
        *reinterpret_cast<uint4*>(row + kc * F_LBO) =
            make_uint4(pk[4 * kc], pk[4 * kc + 1], pk[4 * kc + 2], pk[4 * kc + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      if (++stage == F_STAGES) { stage = 0; phase ^= 1u; }
      if (ntile < p.num_tiles) stash(cur ^ 1, pre);
      asm volatile("bar.sync 1, 128;" ::: "memory");  // patch[cur^1] complete, patch[cur] free
      cur ^= 1;
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, F_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t b_addr = smem_u32(s_b);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(&s_a[stage][0]);
#pragma unroll
        for (int k = 0; k < F_K / 16; ++k) {
          const uint64_t adesc = make_sdesc_k_noswizzle(a_addr + k * 2 * F_LBO, F_LBO, F_SBO);
          const uint64_t bdesc = make_sdesc_k_noswizzle(b_addr + k * 2 * F_LBO, F_LBO, F_SBO);
          umma_bf16(tmem_base + (uint32_t)(as * F_N), adesc, bdesc, idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
        if (++stage == F_STAGES) { stage = 0; phase ^= 1u; }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 4-7) =====================
    epilogue_loop<F_N, AST_EPI_PLAIN>(p, tmem_base, warp - 4, lane, tfull_bar(0), tempty_bar(0));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<2 * F_N>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// First layer, TMA-fed variant (default when W % 4 == 0): the (8+2) x (16+2) x 3 fp32 input patch of a tile is ONE
// 4-D TMA box {24 w, 10 h, 3 c, 1 n} of the NCHW image starting FOUR columns left of the tile (TMA loads never complete
// when the innermost coordinate is not 16-byte aligned -- the constraint measured for the wgrad kernel -- so the box
// starts at w0 - 4 instead of w0 - 1; out-of-image pixels arrive as zeros), prefetched four tiles
// ahead by a dedicated warp, so the four producer warps only expand -- normalise (zero padding applies to the
// NORMALISED image: out-of-image taps are forced to 0 in border tiles), round to bf16, store their A rows -- and
// are coupled to nothing but mbarriers: no block barrier, no global-load latency in their loop.
constexpr int F2_NG = 4;                   // epilogue groups of 4 warps; group g takes every 4th tile (accumulator stage g)
constexpr int F2_PSETS = 2;                // producer sets of 4 warps: set s expands the CTA's tiles s, s + 2, s + 4, ...
constexpr int F2_EPI_WARP0 = 4 * F2_PSETS;
constexpr int F2_MMA_WARP = F2_EPI_WARP0 + 4 * F2_NG, F2_TMA_WARP = F2_MMA_WARP + 1;
constexpr int F2_THREADS = 32 * (F2_TMA_WARP + 1);   // warps 0-7 producers, 8-23 epilogue, 24 MMA + TMEM, 25 TMA
constexpr int F2_PSTAGES = 4;
constexpr int F2_PW = 24, F2_PH = TILE_H + 2;                 // patch row = image columns [w0 - 4, w0 + 20)
constexpr int F2_X0 = 3;                                      // column of the tile's left halo pixel (w0 - 1) in it
constexpr int F2_PATCH_FLOATS = 3 * F2_PH * F2_PW;            // 720
constexpr int F2_PATCH_BYTES = F2_PATCH_FLOATS * 4;           // 2880

// Epilogue of conv1_1 with a TMA store.  The layer's work IS its 64-channel output, and written from registers it is
// one 32-byte sector per lane and store (a lane holds ONE pixel's channels): 512 sector transactions per tile through
// the LSU, which ncu showed 85 % busy.  Here the four warps of a tile set pack their 128 pixel rows (128 B each) into a
// 16 KB staging buffer in the 128-byte-swizzle layout (chunk c of row r at c ^ (r & 7): conflict-free 16-byte stores),
// and one lane issues ONE cp.async.bulk.tensor store of the {64 ch, 16 w, 8 h} box; TMA clips ragged tiles against the
// image.  Tile set tg owns accumulator stage tg and staging buffer tg, so four tiles are in flight per CTA.
__device__ __forceinline__ void epilogue_first_store(const ConvParams& p, const CUtensorMap* tmOut, uint32_t tmem_base,
                                                     int ew, int lane, uint32_t tfull_bar0, uint32_t tempty_bar0,
                                                     uint32_t staging) {
  const int e = ew & 3, tg = ew >> 2;
  const int m = 32 * e + lane;                 // pixel row of the tile = TMEM lane
  const int hl = m / TILE_W, wl = m % TILE_W;
  const uint32_t buf = staging + (uint32_t)tg * (TILE_M * 128);
  const uint32_t row = buf + (uint32_t)m * 128u;
  const uint32_t sw = (uint32_t)(m & 7);
  const bool issuer = (e == 0);
  uint32_t aphase = 0;
  TileCursor cur;
  cur.init(p, blockIdx.x + tg * gridDim.x, F2_NG * gridDim.x);
  long long dbg_wait = 0;
  const long long dbg_t0 = clock64();
  for (int tile = blockIdx.x + tg * gridDim.x; tile < p.num_tiles; tile += F2_NG * gridDim.x, cur.next()) {
    const int twi = cur.twi, thi = cur.thi, n = cur.n;
    mbar_wait_acc(tfull_bar0 + 8u * tg, aphase, kdbg_buf(p) != nullptr, dbg_wait, p.epi_sleep_ns);
    tc_fence_after();
    // the previous store of this tile set has finished reading the staging buffer
    if (issuer && lane == 0) bulk_wait_group_read0();
    named_bar_sync(1 + tg, 128);
    const uint32_t trow = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(tg * F_N);
    const int h = thi * TILE_H + hl, w = twi * TILE_W + wl;
    const bool in_img = (h < p.H) && (w < p.W);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld_32x32(trow + half * 32, v);
      tmem_ld_wait();
      if (p.tap && in_img) {      // fp32 NCHW tap (relu1_1 / conv1_1 for the losses): bias is already in the GEMM
        float* tp = p.tap + (((int64_t)n * p.Cout + half * 32) * p.H + h) * p.W + w;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float f = __uint_as_float(v[i]);
          tp[(int64_t)i * p.H * p.W] = p.tap_prerelu ? f : fmaxf(f, 0.f);
        }
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_bf16_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t chunk = (uint32_t)(half * 4 + c) ^ sw;
        st_shared_v4(row + chunk * 16u, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty_bar0 + 8u * tg);      // accumulator stage free: the MMA warp may refill it
    fence_proxy_async_smem();                               // staging writes -> visible to the TMA (async proxy)
    named_bar_sync(1 + tg, 128);
    if (issuer && lane == 0) {
      tma_store_4d(tmOut, buf, 0, twi * TILE_W, thi * TILE_H, n);
      bulk_commit_group();
    }
    aphase ^= 1u;
  }
  if (issuer && lane == 0) bulk_wait_group0();
  if (kdbg_buf(p) && ew == 0 && lane == 0) {
    kdbg_buf(p)[blockIdx.x * 8 + 4] = dbg_wait;
    kdbg_buf(p)[blockIdx.x * 8 + 5] = clock64() - dbg_t0;
  }
}

__global__ void __launch_bounds__(F2_THREADS, 1)
conv3x3_first_tma_kernel(const __grid_constant__ CUtensorMap tmImg, const __grid_constant__ CUtensorMap tmOut,
                         const FirstParams fp, const ConvParams p, const int tma_store) {
  extern __shared__ uint8_t f2_dyn[];          // tma_store: F2_NG staging buffers of 16 KB, 1024-byte aligned
  __shared__ __align__(128) uint8_t s_a[F_STAGES][F_A_BYTES];
  __shared__ __align__(128) uint8_t s_b[F_B_BYTES];
  __shared__ __align__(128) float s_patch[F2_PSTAGES][F2_PATCH_FLOATS + 16];  // +16: keeps every stage 128 B aligned
  __shared__ __align__(8) uint64_t s_bar[2 * F_STAGES + 2 * F2_NG + 2 * F2_PSTAGES];
  __shared__ uint32_t s_tmem;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t bars = smem_u32(s_bar);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (F_STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * F_STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * F_STAGES + F2_NG + s); };
  auto pfull_bar = [&](int s) { return bars + 8u * (2 * F_STAGES + 2 * F2_NG + s); };
  auto pempty_bar = [&](int s) { return bars + 8u * (2 * F_STAGES + 2 * F2_NG + F2_PSTAGES + s); };

  for (int i = threadIdx.x; i < F_N * F_K; i += F2_THREADS) {
    const int co = i / F_K, k = i % F_K;
    // k < 27: the taps.  k = 27, 28: the bias as a two-term bf16 split against two columns of ones in A -- K is
    // padded to 32 anyway, and it takes 16 broadcast loads + 64 adds per pixel out of an epilogue that ncu shows
    // bound by the L1/LSU pipe (86 %) and instruction issue (63 %).  hi + lo carries the bias to ~2^-17.
    float v = k < 27 ? fp.w[co * 27 + k] : 0.f;
    if (fp.bias && (k == 27 || k == 28)) {
      const float bv = fp.bias[co];
      const float hi = __bfloat162float(__float2bfloat16_rn(bv));
      v = k == 27 ? hi : bv - hi;
    }
    const uint32_t off = (uint32_t)(co >> 3) * F_SBO + (uint32_t)(k >> 3) * F_LBO + (co & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(s_b + off) = __float2bfloat16_rn(v);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < F_STAGES; ++s) { mbar_init(full_bar(s), 4); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < F2_NG; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    for (int s = 0; s < F2_PSTAGES; ++s) { mbar_init(pfull_bar(s), 1); mbar_init(pempty_bar(s), 4); }
    fence_barrier_init();
  }
  if (warp == F2_TMA_WARP && lane == 0) { tma_prefetch_desc(&tmImg); if (tma_store) tma_prefetch_desc(&tmOut); }
  if (warp == F2_MMA_WARP) tmem_alloc<F2_NG * F_N>(smem_u32(&s_tmem));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&s_tmem);

  if (warp == F2_TMA_WARP) {
    // ===================== TMA: input patches =====================
    if (lane == 0) {
      int ps = 0;
      uint32_t pphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int t = tile;
        const int twi = t % p.tiles_w; t /= p.tiles_w;
        const int thi = t % p.tiles_h;
        const int n = t / p.tiles_h;
        mbar_wait(pempty_bar(ps), pphase ^ 1u);
        mbar_expect_tx(pfull_bar(ps), F2_PATCH_BYTES);
        tma_load_4d(smem_u32(&s_patch[ps][0]), &tmImg, pfull_bar(ps), twi * TILE_W - 4, thi * TILE_H - 1, 0, n);
        if (++ps == F2_PSTAGES) { ps = 0; pphase ^= 1u; }
      }
    }
  } else if (warp < F2_EPI_WARP0) {
    // ===================== producers: expand the patch to the 128 x 32 bf16 A tile =====================
    // Two producer sets: set `pset` takes the CTA's tiles pset, pset + 2, ... (sequence index i), i.e. A stage i % 4 and
    // patch stage i % 4 -- with one set the expansion of a tile (a ~1 100-cycle dependent chain per thread) set the
    // pace of the whole kernel once the epilogue had become cheap.
    const int pset = warp >> 2;
    const int r = threadIdx.x & 127;     // tile row = pixel
    const int hl = r >> 4, wl = r & 15;
    float sc[3], sh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sc[c] = fp.normalise ? fp.rstd[c] : 1.f;
      sh[c] = fp.normalise ? -fp.mean[c] * fp.rstd[c] : 0.f;
    }
    static_assert(F_STAGES == 4 && F2_PSTAGES == 4 && F2_PSETS == 2, "a set alternates between two stages of each ring");
    int stage = pset, ps = pset;
    uint32_t phase = 0, pphase = 0;
    for (int tile = blockIdx.x + pset * gridDim.x; tile < p.num_tiles; tile += F2_PSETS * gridDim.x) {
      int t = tile;
      const int twi = t % p.tiles_w; t /= p.tiles_w;
      const int thi = t % p.tiles_h;
      const int h0 = thi * TILE_H - 1, w0 = twi * TILE_W - 1;
      const bool border = h0 < 0 || w0 < 0 || h0 + F2_PH > p.H || w0 + TILE_W + 2 > p.W;
      mbar_wait(pfull_bar(ps), pphase);
      const float* pt = &s_patch[ps][0];
      float v[28];
      v[27] = 0.f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            v[ci * 9 + kh * 3 + kw] = fmaf(pt[(ci * F2_PH + hl + kh) * F2_PW + F2_X0 + wl + kw], sc[ci], sh[ci]);
      __syncwarp();
      if (lane == 0) mbar_arrive(pempty_bar(ps));       // this warp has read its taps
      ps += F2_PSETS;
      if (ps >= F2_PSTAGES) { ps -= F2_PSTAGES; pphase ^= 1u; }
      if (border) {                                       // zero padding of the normalised image (models.py:131 + pad=1)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int ih = h0 + hl + kh, iw = w0 + wl + kw;
            if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) {
              v[kh * 3 + kw] = 0.f; v[9 + kh * 3 + kw] = 0.f; v[18 + kh * 3 + kw] = 0.f;
            }
          }
      }
      uint32_t pk[16];
#pragma unroll
      v[27] = 1.f;                                        // the ones that multiply the bias rows of B (k = 27, 28)
#pragma unroll
      for (int i = 0; i < 14; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
      pk[14] = pack_bf16(1.f, 0.f);
      pk[15] = 0u;
      mbar_wait(empty_bar(stage), phase ^ 1u);
      uint8_t* row = &s_a[stage][0] + (uint32_t)(r >> 3) * F_SBO + (r & 7) * 16;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc)
        *reinterpret_cast<uint4*>(row + kc * F_LBO) =
            make_uint4(pk[4 * kc], pk[4 * kc + 1], pk[4 * kc + 2], pk[4 * kc + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      stage += F2_PSETS;
      if (stage >= F_STAGES) { stage -= F_STAGES; phase ^= 1u; }
    }
  } else if (warp == F2_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, F_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t b_addr = smem_u32(s_b);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(&s_a[stage][0]);
#pragma unroll
        for (int k = 0; k < F_K / 16; ++k) {
          const uint64_t adesc = make_sdesc_k_noswizzle(a_addr + k * 2 * F_LBO, F_LBO, F_SBO);
          const uint64_t bdesc = make_sdesc_k_noswizzle(b_addr + k * 2 * F_LBO, F_LBO, F_SBO);
          umma_bf16(tmem_base + (uint32_t)(as * F_N), adesc, bdesc, idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
        if (++stage == F_STAGES) { stage = 0; phase ^= 1u; }
        if (++as == F2_NG) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps F2_EPI_WARP0 .. + 4 F2_NG) =====================
    if (tma_store) {
      const uint32_t staging = (smem_u32(f2_dyn) + 1023u) & ~1023u;
      epilogue_first_store(p, &tmOut, tmem_base, warp - F2_EPI_WARP0, lane, tfull_bar(0), tempty_bar(0), staging);
    } else {
      epilogue_loop<F_N, AST_EPI_PLAIN, TILE_W, 1, F2_NG, F2_NG>(p, tmem_base, warp - F2_EPI_WARP0, lane, tfull_bar(0), tempty_bar(0));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == F2_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<F2_NG * F_N>(tmem_base);
  }
}

static int get_sm_count(int* out) {
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    AST_CUDA(cudaGetDevice(&dev));
    AST_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  *out = sm_count;
  return 0;
}

int conv3x3_first_tc(const float* img, const float* w, const float* bias, const float* mean,
                     const float* std_, void* out, float* tap, int tap_prerelu, int N, int H, int W,
                     cudaStream_t s) {
  int sm_count = 0;
  int r = get_sm_count(&sm_count);
  if (r) return r;
  FirstParams fp = {};
  fp.img = img; fp.w = w; fp.normalise = (mean && std_) ? 1 : 0;
  for (int i = 0; i < 3; ++i) {
    fp.mean[i] = fp.normalise ? mean[i] : 0.f;
    fp.rstd[i] = fp.normalise ? 1.f / std_[i] : 1.f;
  }
  ConvParams p = {};
  p.N = N; p.H = H; p.W = W; p.Cin = 3; p.Cout = F_N; p.Ho = H; p.Wo = W;
  p.relu = 1; p.halo = AST_HALO_KEEP; p.tap_prerelu = tap_prerelu;
  p.tiles_w = (W + TILE_W - 1) / TILE_W;
  p.tiles_h = (H + TILE_H - 1) / TILE_H;
  p.n_blocks = 1;
  const int64_t nt = (int64_t)N * p.tiles_h * p.tiles_w;
  if (nt >= 0x7fffffffLL) return AST_E_SHAPE;
  p.num_tiles = (int)nt;
  p.bias = bias; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.tap = tap;
  const int grid = p.num_tiles < F_CTAS_PER_SM * sm_count ? p.num_tiles : F_CTAS_PER_SM * sm_count;
  static const bool dbg_on = getenv("AST_CONV_DEBUG") != nullptr;     // debug only: allocates and synchronises
  if (getenv("AST_FIRST_NOSTORE")) p.out = nullptr;                   // diagnostic: the layer without its output stores
  long long* dbg = nullptr;
  if (dbg_on) {
    AST_CUDA(cudaMalloc(&dbg, sizeof(long long) * 8 * grid));
    AST_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 8 * grid, s));
    p.dbg = dbg;
  }
  struct Dump {
    long long* d; int grid; int tiles; cudaStream_t s;
    ~Dump() {
      if (!d) return;
      cudaStreamSynchronize(s);
      long long* h = (long long*)malloc(sizeof(long long) * 8 * grid);
      cudaMemcpy(h, d, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost);
      double w = 0, t = 0;
      for (int i = 0; i < grid; ++i) { w += (double)h[i * 8 + 4]; t += (double)h[i * 8 + 5]; }
      const double per = (double)tiles / grid;
      fprintf(stderr, "[conv dbg] first layer: tiles/CTA=%.1f | epilogue warp 0 per tile: loop=%.0f cycles, of which waiting "
                      "for an accumulator=%.0f\n", per, t / grid / per, w / grid / per);
      free(h);
      cudaFree(d);
    }
  } dump{dbg, grid, p.num_tiles, s};   // dump.grid is corrected below when the one-CTA-per-SM kernel is launched
  static const bool no_tma = getenv("AST_FIRST_NO_TMA") != nullptr;   // A/B reference: the register-staged producer
  if (!no_tma && W % 4 == 0 && aligned16(img)) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return AST_E_NODRIVER;
    CUtensorMap tmImg;
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)N};
    const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * 12};
    const cuuint32_t bx[4] = {F2_PW, F2_PH, 3, 1}, es[4] = {1, 1, 1, 1};
    CUresult cr = enc(&tmImg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(img), gdim, gstr, bx, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return AST_E_SHAPE;
    const int grid1 = p.num_tiles < sm_count ? p.num_tiles : sm_count;
    dump.grid = grid1;
    fp.bias = bias;
    // TMA-store epilogue: needs the bias in the GEMM (always here), an output and a 16-byte aligned interior
    static const bool no_tma_store = getenv("AST_FIRST_NO_TMA_STORE") != nullptr;   // A/B reference: per-lane stores
    const bool tma_store = !no_tma_store && p.out != nullptr && bias != nullptr && aligned16(out);
    p.bias = nullptr;               // folded into the GEMM
    CUtensorMap tmOut = tmImg;
    if (tma_store) {
      // the INTERIOR of the padded NHWC buffer [N][H+2][W+2][64]: boxes are clipped at W and H, never touch the halo
      const uint64_t odims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)N};
      const uint64_t ostr[3] = {128, (uint64_t)(W + 2) * 128, (uint64_t)(H + 2) * (W + 2) * 128};
      const uint32_t obox[4] = {64, TILE_W, TILE_H, 1};
      const __nv_bfloat16* interior = reinterpret_cast<const __nv_bfloat16*>(out) + ((int64_t)(W + 2) + 1) * 64;
      r = encode_bf16_map(&tmOut, interior, 4, odims, ostr, obox);
      if (r) return r;
    }
    constexpr int F2_DYN = F2_NG * TILE_M * 128 + 1024;
    static bool attr_done = false;
    if (!attr_done) {
      AST_CUDA(cudaFuncSetAttribute(conv3x3_first_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_DYN));
      attr_done = true;
    }
    conv3x3_first_tma_kernel<<<grid1, F2_THREADS, tma_store ? F2_DYN : 0, s>>>(tmImg, tmOut, fp, p, tma_store ? 1 : 0);
    AST_CHECK_LAUNCH();
    return 0;
  }
  conv3x3_first_tc_kernel<<<grid, kFirstThreads, 0, s>>>(fp, p);
  AST_CHECK_LAUNCH();
  return 0;
}

#include "conv12_fused.cuh"

int conv3x3_last_tn(const void* in, const void* wpk16, const float* bias, float* out, int N, int H, int W, int Cout,
                    int clamp01, int sm_count, cudaStream_t s);   // conv_last_tn.cu

// Last decoder layer (Cin % 64 == 0, Cout <= 16).  Cin == 64, Cout <= 3 (the decoder's image layer and the image
// gradient of the VGG backward pass): the taps-in-N kernel of conv_last_tn.cu.  Otherwise -- or with
// AST_LAST_TAPS_IN_K=1 / the tap-box implementation, kept as A/B references -- the implicit-GEMM kernel with a
// 16-wide N block (weights zero-padded to 16 output channels) and the fp32 NCHW epilogue.
int conv3x3_last_tc(const void* in, const void* wpk16, const float* bias, float* out, int N, int H,
                    int W, int Cin, int Cout, int clamp01, int kwbox, cudaStream_t s) {
  if (Cin % 64 != 0 || Cout > 16 || H < 1 || W < 1) return AST_E_SHAPE;
  if (!aligned16(in) || !aligned16(wpk16)) return AST_E_ALIGN;
  int sm_count = 0;
  int r = get_sm_count(&sm_count);
  if (r) return r;
  static const bool taps_in_k = getenv("AST_LAST_TAPS_IN_K") != nullptr;
  if (kwbox && Cin == 64 && Cout <= 3 && !taps_in_k)
    return conv3x3_last_tn(in, wpk16, bias, out, N, H, W, Cout, clamp01, sm_count, s);
  ConvParams p = {};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = 16; p.Ho = H; p.Wo = W;
  p.relu = 0; p.halo = AST_HALO_KEEP;
  const int tw = kwbox ? T2_W : TILE_W, th = kwbox ? T2_H : TILE_H;
  p.tiles_w = (W + tw - 1) / tw;
  p.tiles_h = (H + th - 1) / th;
  p.n_blocks = 1;
  const int64_t nt = (int64_t)N * p.tiles_h * p.tiles_w;
  if (nt >= 0x7fffffffLL) return AST_E_SHAPE;
  p.num_tiles = (int)nt;
  p.bias = bias; p.out_nchw = out; p.cout_real = Cout; p.clamp01 = clamp01;
  CUtensorMap tmA, tmB;
  r = make_maps(&tmA, &tmB, in, wpk16, N, H, W, Cin, 16, 16, kwbox);
  if (r) return r;
  if (kwbox) return launch_tc2<16, EPI_NCHW32>(tmA, tmB, p, sm_count, s);
  return launch_tc<16, EPI_NCHW32>(tmA, tmB, p, sm_count, s);
}

}  // namespace tc
}  // namespace ast
