// K2u: Upsample(x2, nearest) -> ReflectionPad2d(1) -> Conv2d(3x3) as FOUR parity-specific 2x2 convolutions on the
// LOW-resolution map (included by conv_tc.cu; shares its tile geometry, barriers and descriptors).
//
// Reference work replaced (paths relative to /root/reference): the decoder's three
//   nn.Upsample(scale_factor=2, mode='nearest'), nn.ReflectionPad2d((1,1,1,1)), nn.Conv2d(c_in, c_out, (3,3))
// sequences, models.py:602-604, 616-618, 622-624.
//
// Output pixel (2h+py, 2w+px) of the 3x3 conv over the upsampled map reads hi-res rows 2h+py-1 .. 2h+py+1, i.e. the
// low-res rows {h-1, h, h} (py = 0) or {h, h, h+1} (py = 1): two distinct rows, with the weights of the coinciding taps
// pre-summed (ast_pack_conv_weight_fold).  Likewise for columns.  So
//   y[2h+py][2w+px][co] = b[co] + sum_{a,b in {0,1}} sum_ci Wf[py][px][a][b][co][ci] * x[h-1+py+a][w-1+px+b][ci]
// -- 16 tap-products per low-res pixel instead of 4 x 9 = 36 (2.25x fewer MMAs), and the upsampled tensor is never
// written or read.  ReflectionPad2d(1) of the upsampled map (index -1 -> 1, 2H -> 2H-2) lands on low-res rows 0 and
// H-1: a CLAMP (replicate) halo of the low-res input, which the producing layer writes (AST_HALO_CLAMP).
//
// GEMM view: M = 128 low-res pixels (16 rows x 8 cols, the kw-box tile of conv3x3_tc2_kernel: the SAME A boxes,
// {64 ch, 8 w, 18 h} per kw, serve all parities; a tap (kh, kw) starts kh * 1024 B into the box), N = BN output
// channels, K = 4 taps x Cin.  A tile carries NPAR = 256 / BN parities whose accumulators sit side by side in one
// 256-column TMEM stage (two stages = all 512 columns): BN = 64 -> all four parities share every A box (54 KB of A per
// 64-channel block for 16 tap-products), BN = 128 -> one output-row parity (py) per tile, BN = 256 -> one parity per
// tile (two A boxes).  Weight tiles ride a ring of two groups (one group = the taps one A box feeds, <= 64 KB); with
// Cin = Cout = 64 all sixteen tiles (128 KB) are loaded once per CTA and stay resident.
#pragma once

template <int BN>
struct CfgF {
  static constexpr int NPAR = 256 / BN;            // parities per tile
  static constexpr int NPG = 4 / NPAR;             // parity groups (tiles per spatial tile and cout block)
  static constexpr int B_BYTES = BN * KBLK * 2;
  static constexpr int NA = 4;                     // A boxes in flight
  static constexpr int GSLOTS = 2 * NPAR;          // weight tiles one A box can feed
  static constexpr int NBG = 2;                    // weight groups in the ring
  static constexpr int NACC = 2;
  static constexpr int ACC_COLS = 256;
  static constexpr int NBAR = 2 * NA + 2 * NBG + 2 * NACC;
  static constexpr int SMEM_BYTES = NA * A2_BYTES + NBG * GSLOTS * B_BYTES + NBAR * 8 + 16 + 1024;
};

// The parities of tile group `pg`: row parities [py0, py0 + npy), column parities [px0, px0 + npx).
template <int NPAR>
__device__ __forceinline__ void fold_parities(int pg, int& py0, int& npy, int& px0, int& npx) {
  if (NPAR == 4) { py0 = 0; npy = 2; px0 = 0; npx = 2; }
  else if (NPAR == 2) { py0 = pg; npy = 1; px0 = 0; npx = 2; }
  else { py0 = pg >> 1; npy = 1; px0 = pg & 1; npx = 1; }
}

// The four parity views of the folded layer's output (pixel (2h+py, 2w+px) of the padded NHWC buffer as a tensor over
// the LOW-res grid), for the TMA-store epilogue.
struct FoldOutMaps { CUtensorMap m[4]; };

// TS (TMA store, CTA-pair kernel, BN <= 128): warp (e, g) takes the 64 columns [64 g, 64 g + 64) of the 256-column tile
// -- one parity's 64 channels (BN = 64) or one 64-channel half of a parity (BN = 128) -- for its 32 pixels (tile rows
// 4e .. 4e+3), packs them into its OWN 4 KB staging buffer ([32 px][128 B], 128-byte swizzle) and lane 0 issues ONE
// cp.async.bulk.tensor store of the {64 ch, 8 w, 4 h} box against the parity's strided view: no barrier wider than a
// warp (a first version staged half tiles behind 512-thread named barriers: the staging alone took the layer from
// 195 to 310 us, profiles/r2_conv_fold_tma_store.txt).  Border lanes still write the halo copies themselves.
template <int BN, int NG, bool P2 = false, bool TS = false>
__device__ __forceinline__ void epilogue_fold(const ConvParams& p, uint32_t tmem_base, int ew, int lane,
                                              uint32_t tfull_bar0, uint32_t tempty_bar0, int rank = 0,
                                              const FoldOutMaps* om = nullptr, uint32_t staging = 0) {
  using C = CfgF<BN>;
  static_assert(!TS || (BN <= 128 && NG == 4), "TMA store: every warp owns one 32-column chunk per half tile");
  constexpr int CH = 32, NCH = C::ACC_COLS / CH;
  const int e = ew & 3, g = ew >> 2;
  const int hl = (32 * e + lane) / T2_W, wl = (32 * e + lane) % T2_W;
  const bool wide_st = (reinterpret_cast<uintptr_t>(p.out) & 31u) == 0 && (p.Cout % 16) == 0;
  int as = 0;
  uint32_t aphase = 0;
  const uint32_t tempty_leader0 = P2 ? mapa_shared(tempty_bar0, 0) : 0u;
  const int t0 = P2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstride = P2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  TileCursor cur;
  cur.init(p, t0, tstride, P2 ? 2 : 1, rank);
  for (int tile = t0; tile < p.num_tiles; tile += tstride, cur.next()) {
    const int cob = cur.nb / C::NPG, pg = cur.nb % C::NPG, n = cur.n;
    const int h = cur.thi * T2_H + hl, w = cur.twi * T2_W + wl;
    const bool in_img = (h < p.H) && (w < p.W) && (!P2 || n < p.N);
    int py0, npy, px0, npx;
    fold_parities<C::NPAR>(pg, py0, npy, px0, npx);
    if (p.epi_sleep_ns) mbar_wait_sleep(tfull_bar0 + 8u * as, aphase, p.epi_sleep_ns); else mbar_wait(tfull_bar0 + 8u * as, aphase);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(as * C::ACC_COLS);
    if constexpr (TS) {
      const uint32_t wbuf = staging + (uint32_t)ew * 4096u;
      const int jj = (64 * g) / BN, choff = (64 * g) % BN;            // parity slot and channel offset of this warp
      const int py = py0 + (npx == 2 ? (jj >> 1) : jj), px = px0 + (npx == 2 ? (jj & 1) : 0);
      uint32_t va[CH], vb[CH];
      tmem_ld_cols(trow + (2 * g) * CH, va);
      tmem_ld_cols(trow + (2 * g + 1) * CH, vb);
      if (lane == 0) bulk_wait_group_read0();      // this warp's previous store has finished reading its buffer
      __syncwarp();
      tmem_ld_wait();
      int rows[4], cols[4], nr = 0, nc = 0;
      if (in_img) {
        nr = out_targets<AST_EPI_PLAIN>(2 * h + py, p.Ho, p.halo, rows);
        nc = out_targets<AST_EPI_PLAIN>(2 * w + px, p.Wo, p.halo, cols);
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const uint32_t* v = cc ? vb : va;
        const int ch0 = cob * BN + choff + cc * CH;
        uint32_t pk[CH / 2];
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
          const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float f0 = __uint_as_float(v[i]) + b.x, f1 = __uint_as_float(v[i + 1]) + b.y;
          const float f2 = __uint_as_float(v[i + 2]) + b.z, f3 = __uint_as_float(v[i + 3]) + b.w;
          pk[i / 2] = p.relu ? pack_bf16_relu(f0, f1) : pack_bf16(f0, f1);
          pk[i / 2 + 1] = p.relu ? pack_bf16_relu(f2, f3) : pack_bf16(f2, f3);
        }
#pragma unroll
        for (int q = 0; q < CH / 8; ++q) {
          const uint32_t c16 = (uint32_t)(cc * (CH / 8) + q) ^ (uint32_t)(lane & 7);
          st_shared_v4(wbuf + (uint32_t)lane * 128u + c16 * 16u, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        if (!(kdbg_flags(p) & 2)) {
          for (int ri = 0; ri < nr; ++ri)
            for (int ci = 0; ci < nc; ++ci) {
              if (ri == 0 && ci == 0) continue;        // the pixel itself goes out with the TMA store
              __nv_bfloat16* o = p.out +
                  (((int64_t)n * (p.Ho + 2) + (rows[ri] + 1)) * (p.Wo + 2) + (cols[ci] + 1)) * p.Cout + ch0;
              uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
              for (int q = 0; q < CH / 8; ++q) o4[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (P2) mbar_arrive_cluster(tempty_leader0 + 8u * as);
        else mbar_arrive(tempty_bar0 + 8u * as);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && !(kdbg_flags(p) & 2)) {
        tma_store_4d(&om->m[py * 2 + px], wbuf, cob * BN + choff, cur.twi * T2_W, cur.thi * T2_H + 4 * e, n);
        bulk_commit_group();
      }
      if (++as == C::NACC) { as = 0; aphase ^= 1u; }
      continue;
    }
    uint32_t vnext[CH];
    const bool skip_ld = (kdbg_flags(p) & 4) != 0;
    if (g < NCH && !skip_ld) tmem_ld_cols(trow + g * CH, vnext);
#pragma unroll 1
    for (int chunk = g; chunk < NCH && !skip_ld; chunk += NG) {
      uint32_t v[CH];
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = vnext[i];
      if (chunk + NG < NCH) tmem_ld_cols(trow + (chunk + NG) * CH, vnext);
      const int j = (chunk * CH) / BN;                 // parity slot of this chunk
      const int ch0 = cob * BN + (chunk * CH) % BN;
      const int py = py0 + (npx == 2 ? (j >> 1) : j), px = px0 + (npx == 2 ? (j & 1) : 0);
      float f[CH];
#pragma unroll
      for (int i = 0; i < CH; i += 4) {
        float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[i + 0] = __uint_as_float(v[i + 0]) + b.x;
        f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
        f[i + 2] = __uint_as_float(v[i + 2]) + b.z;
        f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
      }
      uint32_t pk[CH / 2];
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) pk[i] = pack_bf16_relu(f[2 * i], f[2 * i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) pk[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
      }
      if (in_img && !(kdbg_flags(p) & 2)) {
        int rows[4], cols[4];
        const int nr = out_targets<AST_EPI_PLAIN>(2 * h + py, p.Ho, p.halo, rows);
        const int nc = out_targets<AST_EPI_PLAIN>(2 * w + px, p.Wo, p.halo, cols);
        for (int ri = 0; ri < nr; ++ri) {
          for (int ci = 0; ci < nc; ++ci) {
            __nv_bfloat16* o = p.out +
                (((int64_t)n * (p.Ho + 2) + (rows[ri] + 1)) * (p.Wo + 2) + (cols[ci] + 1)) * p.Cout + ch0;
            if (wide_st) {
#pragma unroll
              for (int q = 0; q < CH / 16; ++q) st_global_v8(o + 16 * q, &pk[8 * q]);
            } else {
              uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
              for (int q = 0; q < CH / 8; ++q) o4[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
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
    if (++as == C::NACC) { as = 0; aphase ^= 1u; }
  }
  if (TS && lane == 0) bulk_wait_group0();
}

template <int BN>
__global__ void __launch_bounds__(Epi2<BN>::THREADS, 1)
conv3x3_fold_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const ConvParams p) {
  using C = CfgF<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + C::NA * A2_BYTES;
  constexpr int B_REGION = C::NBG * C::GSLOTS * C::B_BYTES;
  const uint32_t bars = b_base + B_REGION;
  auto afull = [&](int s) { return bars + 8u * s; };
  auto aempty = [&](int s) { return bars + 8u * (C::NA + s); };
  auto bfull = [&](int s) { return bars + 8u * (2 * C::NA + s); };
  auto bempty = [&](int s) { return bars + 8u * (2 * C::NA + C::NBG + s); };
  auto tfull = [&](int s) { return bars + 8u * (2 * C::NA + 2 * C::NBG + s); };
  auto tempty = [&](int s) { return bars + 8u * (2 * C::NA + 2 * C::NBG + C::NACC + s); };
  const uint32_t tmem_slot = bars + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + C::NA * A2_BYTES + B_REGION + 8 * C::NBAR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int cblocks = p.Cin / KBLK;
  // all sixteen weight tiles fit the ring's 128 KB when BN = 64: keep them resident if one set serves every tile
  const bool resident = (BN == 64) && cblocks == 1 && p.Cout == BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < C::NBG; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int s = 0; s < C::NACC; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 4 * Epi2<BN>::NG); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (resident) {
        mbar_expect_tx(bfull(0), 16 * C::B_BYTES);
        for (int t = 0; t < 16; ++t) tma_load_3d(b_base + t * C::B_BYTES, &tmB, bfull(0), 0, 0, t);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool a_filled = false, b_filled = false;   // dbg_flags & 1 (bottleneck elimination): ring slots loaded once
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int t = tile;
        const int nb = t % p.n_blocks; t /= p.n_blocks;
        const int twi = t % p.tiles_w; t /= p.tiles_w;
        const int thi = t % p.tiles_h;
        const int n = t / p.tiles_h;
        const int cob = nb / C::NPG, pg = nb % C::NPG;
        int py0, npy, px0, npx;
        fold_parities<C::NPAR>(pg, py0, npy, px0, npx);
        const int h0 = thi * T2_H, w0 = twi * T2_W;
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int kw = px0; kw <= px0 + npx; ++kw) {
            if (p.prod_sleep_ns) mbar_wait_sleep(aempty(sa), pa ^ 1u, p.prod_sleep_ns); else mbar_wait(aempty(sa), pa ^ 1u);
            if ((kdbg_flags(p) & 1) && a_filled) {
              mbar_arrive(afull(sa));
            } else {
              mbar_expect_tx(afull(sa), A2_BYTES);
              tma_load_4d(a_base + sa * A2_BYTES, &tmA, afull(sa), cb * KBLK, w0 + kw, h0, n);
            }
            if (sa + 1 == C::NA) a_filled = true;
            if (++sa == C::NA) { sa = 0; pa ^= 1u; }
            if (!resident) {
              // the weight tiles this box feeds, in the order the MMA warp consumes them
              int cnt = 0;
              for (int py = py0; py < py0 + npy; ++py)
                for (int px = px0; px < px0 + npx; ++px)
                  if (kw - px >= 0 && kw - px <= 1) cnt += 2;
              if (p.prod_sleep_ns) mbar_wait_sleep(bempty(sb), pb ^ 1u, p.prod_sleep_ns); else mbar_wait(bempty(sb), pb ^ 1u);
              if ((kdbg_flags(p) & 1) && b_filled) {
                mbar_arrive(bfull(sb));
              } else {
                mbar_expect_tx(bfull(sb), cnt * C::B_BYTES);
                int slot = 0;
                for (int py = py0; py < py0 + npy; ++py)
                  for (int px = px0; px < px0 + npx; ++px) {
                    const int b = kw - px;
                    if (b < 0 || b > 1) continue;
                    for (int a = 0; a < 2; ++a) {
                      tma_load_3d(b_base + (sb * C::GSLOTS + slot) * C::B_BYTES, &tmB, bfull(sb), cb * KBLK, cob * BN,
                                  ((py * 2 + px) * 2 + a) * 2 + b);
                      ++slot;
                    }
                  }
              }
              if (sb + 1 == C::NBG) b_filled = true;
              if (++sb == C::NBG) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN);
    const uint64_t a_desc0 = make_sdesc_k128(a_base);
    const uint64_t b_desc0 = make_sdesc_k128(b_base);
    constexpr uint64_t A_SLOT16 = A2_BYTES >> 4, B_SLOT16 = C::B_BYTES >> 4, KH16 = (T2_W * KBLK * 2) >> 4;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int as = 0;
    uint32_t aphase = 0;
    bool b_ready = false;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nb = tile % p.n_blocks;
      const int pg = nb % C::NPG;
      int py0, npy, px0, npx;
      fold_parities<C::NPAR>(pg, py0, npy, px0, npx);
      mbar_wait(tempty(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tile = tmem_base + (uint32_t)(as * C::ACC_COLS);
      uint32_t started = 0;   // bit j: accumulator slot j already holds a partial sum
      for (int cb = 0; cb < cblocks; ++cb) {
        for (int kw = px0; kw <= px0 + npx; ++kw) {
          mbar_wait(afull(sa), pa);
          if (resident) {
            if (!b_ready) { mbar_wait(bfull(0), 0u); b_ready = true; }
          } else {
            mbar_wait(bfull(sb), pb);
          }
          tc_fence_after();
          const uint64_t ad = a_desc0 + (uint64_t)sa * A_SLOT16;
          if (elect_one_sync()) {
            int slot = 0;
            for (int iy = 0; iy < npy; ++iy) {
              for (int ix = 0; ix < npx; ++ix) {
                const int py = py0 + iy, px = px0 + ix;
                const int b = kw - px;
                if (b < 0 || b > 1) continue;
                const int j = iy * npx + ix;
                const uint32_t d_tmem = d_tile + (uint32_t)(j * BN);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                  const int kh = py + a;
                  const int bs = resident ? (((py * 2 + px) * 2 + a) * 2 + b) : (sb * C::GSLOTS + slot);
                  const uint64_t bd = b_desc0 + (uint64_t)bs * B_SLOT16;
#pragma unroll
                  for (int k = 0; k < KBLK / 16; ++k) {
                    umma_bf16(d_tmem, ad + (uint64_t)(kh * KH16 + k * 2), bd + (uint64_t)(k * 2), idesc,
                              (a | k) ? 1u : ((started >> j) & 1u));
                  }
                  ++slot;
                }
                started |= 1u << j;
              }
            }
            if (!resident) umma_commit(bempty(sb));
            umma_commit(aempty(sa));
          }
          __syncwarp();
          // `started` is only updated by the elected lane; recompute it warp-uniformly for the next box
          for (int iy = 0; iy < npy; ++iy)
            for (int ix = 0; ix < npx; ++ix)
              if (kw - (px0 + ix) >= 0 && kw - (px0 + ix) <= 1) started |= 1u << (iy * npx + ix);
          if (!resident) {
            if (++sb == C::NBG) { sb = 0; pb ^= 1u; }
          }
          if (++sa == C::NA) { sa = 0; pa ^= 1u; }
        }
      }
      if (elect_one_sync()) umma_commit(tfull(as));
      __syncwarp();
      if (++as == C::NACC) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4) {
    epilogue_fold<BN, Epi2<BN>::NG>(p, tmem_base, warp - 4, lane, tfull(0), tempty(0));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int BN>
static int launch_fold(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int sm_count,
                       cudaStream_t s) {
  using C = CfgF<BN>;
  auto kern = conv3x3_fold_kernel<BN>;
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

// ------------------------------------------------------------------------------------------------
// The same folded convolution on CTA PAIRS (tcgen05.mma.cta_group::2; protocol: conv_pair.cuh).  Each CTA of a pair
// takes one low-res spatial tile (16 x 8 pixels), both run the same cout block and parity group, and each holds
// half of every weight tile (BN/2 rows), so the weight ring holds twice as many groups in the same shared memory.
template <int BN>
struct CfgFP {
  static constexpr int NPAR = CfgF<BN>::NPAR, NPG = CfgF<BN>::NPG;
  static constexpr int HB = BN / 2;
  static constexpr int B_BYTES = HB * KBLK * 2;
  static constexpr int NA = 3;                     // single {64, 10, 18} boxes (conv_pair.cuh), one per (tile, block)
  static constexpr int GSLOTS = 2 * NPAR;
  static constexpr int NBG = 4;
  static constexpr int NACC = 2;
  static constexpr int ACC_COLS = 256;
  static constexpr int NBAR = 2 * NA + 2 * NBG + 2 * NACC;
  static constexpr int SMEM_BYTES = NA * AW_SLOT + NBG * GSLOTS * B_BYTES + NBAR * 8 + 16 + 1024;
};

template <int BN, bool TS>
__global__ void __launch_bounds__(Epi2<BN>::THREADS, 1)
conv3x3_fold_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ FoldOutMaps om, const ConvParams p) {
  using C = CfgFP<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  // [TMA-store staging: 2 x 32 KB] [A ring] [weights: ring of NBG groups, or p.nbs = 16 resident tiles] [barriers]
  const uint32_t a_base = base + (TS ? (uint32_t)p.tma_store : 0u);
  const uint32_t b_base = a_base + C::NA * AW_SLOT;
  const int B_REGION = p.nbs * C::B_BYTES;
  const uint32_t bars = b_base + B_REGION;
  auto afull = [&](int s) { return bars + 8u * s; };
  auto aempty = [&](int s) { return bars + 8u * (C::NA + s); };
  auto bfull = [&](int s) { return bars + 8u * (2 * C::NA + s); };
  auto bempty = [&](int s) { return bars + 8u * (2 * C::NA + C::NBG + s); };
  auto tfull = [&](int s) { return bars + 8u * (2 * C::NA + 2 * C::NBG + s); };
  auto tempty = [&](int s) { return bars + 8u * (2 * C::NA + 2 * C::NBG + C::NACC + s); };
  const uint32_t tmem_slot = bars + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (TS ? p.tma_store : 0) + C::NA * AW_SLOT + B_REGION + 8 * C::NBAR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair0 = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int cblocks = p.Cin / KBLK;
  // BN = 64, Cin = Cout = 64: all sixteen half weight tiles (64 KB) fit the ring's space and stay resident
  const bool resident = (BN == 64) && cblocks == 1 && p.Cout == BN;
  const int NBG = resident ? 1 : p.nbs / C::GSLOTS;      // weight groups in the ring (launch_fold_pair: 4, or 2 beside a staging buffer)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TS) { for (int i = 0; i < 4; ++i) tma_prefetch_desc(&om.m[i]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < C::NBG; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int s = 0; s < C::NACC; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 2 * 4 * Epi2<BN>::NG); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t afull_l = mapa_shared(afull(0), 0), bfull_l = mapa_shared(bfull(0), 0);
      if (resident) {
        if (rank == 0) mbar_expect_tx(bfull(0), 2 * 16 * C::B_BYTES);
        for (int t = 0; t < 16; ++t) tma_load_3d_2sm(b_base + t * C::B_BYTES, &tmB, bfull_l, 0, rank * C::HB, t);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int q = pair0; q < p.num_tiles; q += npairs) {
        int t = q;
        const int nb = t % p.n_blocks; t = 2 * (t / p.n_blocks) + rank;
        const int twi = t % p.tiles_w; t /= p.tiles_w;
        const int thi = t % p.tiles_h;
        const int n = t / p.tiles_h;
        const int cob = nb / C::NPG, pg = nb % C::NPG;
        int py0, npy, px0, npx;
        fold_parities<C::NPAR>(pg, py0, npy, px0, npx);
        const int h0 = thi * T2_H, w0 = twi * T2_W;
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int kw = px0; kw <= px0 + npx; ++kw) {
            if (kw == px0) {
              if (p.prod_sleep_ns) mbar_wait_sleep(aempty(sa), pa ^ 1u, p.prod_sleep_ns); else mbar_wait(aempty(sa), pa ^ 1u);
              if (rank == 0) mbar_expect_tx(afull(sa), 2 * AW_BYTES);
              tma_load_4d_2sm(a_base + sa * AW_SLOT, &tmA, afull_l + 8u * sa, cb * KBLK, w0, h0, n);
              if (++sa == C::NA) { sa = 0; pa ^= 1u; }
            }
            if (!resident) {
              int cnt = 0;
              for (int py = py0; py < py0 + npy; ++py)
                for (int px = px0; px < px0 + npx; ++px)
                  if (kw - px >= 0 && kw - px <= 1) cnt += 2;
              if (p.prod_sleep_ns) mbar_wait_sleep(bempty(sb), pb ^ 1u, p.prod_sleep_ns); else mbar_wait(bempty(sb), pb ^ 1u);
              if (rank == 0) mbar_expect_tx(bfull(sb), 2 * cnt * C::B_BYTES);
              int slot = 0;
              for (int py = py0; py < py0 + npy; ++py)
                for (int px = px0; px < px0 + npx; ++px) {
                  const int b = kw - px;
                  if (b < 0 || b > 1) continue;
                  for (int a = 0; a < 2; ++a) {
                    tma_load_3d_2sm(b_base + (sb * C::GSLOTS + slot) * C::B_BYTES, &tmB, bfull_l + 8u * sb, cb * KBLK,
                                    cob * BN + rank * C::HB, ((py * 2 + px) * 2 + a) * 2 + b);
                    ++slot;
                  }
                }
              if (++sb == NBG) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
      for (int i = 0; i < C::NA; ++i) {       // drain the multicast releases aimed at this CTA
        mbar_wait(aempty(sa), pa ^ 1u);
        if (++sa == C::NA) { sa = 0; pa ^= 1u; }
      }
      if (!resident) {
        for (int i = 0; i < NBG; ++i) {
          mbar_wait(bempty(sb), pb ^ 1u);
          if (++sb == NBG) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * TILE_M, BN);
      const uint64_t a_desc0 = make_sdesc_k128_sbo(a_base, WA_W * KBLK * 2, 0);
      const uint64_t b_desc0 = make_sdesc_k128(b_base);
      constexpr uint64_t A_SLOT16 = AW_SLOT >> 4, B_SLOT16 = C::B_BYTES >> 4, KH16 = (WA_W * KBLK * 2) >> 4;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int as = 0;
      uint32_t aphase = 0;
      bool b_ready = false;
      for (int q = pair0; q < p.num_tiles; q += npairs) {
        const int nb = q % p.n_blocks;
        const int pg = nb % C::NPG;
        int py0, npy, px0, npx;
        fold_parities<C::NPAR>(pg, py0, npy, px0, npx);
        mbar_wait(tempty(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tile = tmem_base + (uint32_t)(as * C::ACC_COLS);
        uint32_t started = 0;
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int kw = px0; kw <= px0 + npx; ++kw) {
            if (kw == px0) mbar_wait(afull(sa), pa);
            if (resident) {
              if (!b_ready) { mbar_wait(bfull(0), 0u); b_ready = true; }
            } else {
              mbar_wait(bfull(sb), pb);
            }
            tc_fence_after();
            const uint64_t ad = a_desc0 + (uint64_t)sa * A_SLOT16 + (uint64_t)(kw * 8);   // one pixel = 128 B per kw
            if (elect_one_sync()) {
              int slot = 0;
              for (int iy = 0; iy < npy; ++iy) {
                for (int ix = 0; ix < npx; ++ix) {
                  const int py = py0 + iy, px = px0 + ix;
                  const int b = kw - px;
                  if (b < 0 || b > 1) continue;
                  const int j = iy * npx + ix;
                  const uint32_t d_tmem = d_tile + (uint32_t)(j * BN);
#pragma unroll
                  for (int a = 0; a < 2; ++a) {
                    const int kh = py + a;
                    const int bs = resident ? (((py * 2 + px) * 2 + a) * 2 + b) : (sb * C::GSLOTS + slot);
                    const uint64_t bd = b_desc0 + (uint64_t)bs * B_SLOT16;
#pragma unroll
                    for (int k = 0; k < KBLK / 16; ++k) {
                      umma_bf16_2sm(d_tmem, ad + (uint64_t)(kh * KH16 + k * 2), bd + (uint64_t)(k * 2), idesc,
                                    (a | k) ? 1u : ((started >> j) & 1u));
                    }
                    ++slot;
                  }
                  started |= 1u << j;
                }
              }
              if (!resident) umma_commit_2sm(bempty(sb));
              if (kw == px0 + npx) umma_commit_2sm(aempty(sa));
            }
            __syncwarp();
            for (int iy = 0; iy < npy; ++iy)
              for (int ix = 0; ix < npx; ++ix)
                if (kw - (px0 + ix) >= 0 && kw - (px0 + ix) <= 1) started |= 1u << (iy * npx + ix);
            if (!resident) {
              if (++sb == NBG) { sb = 0; pb ^= 1u; }
            }
            if (kw == px0 + npx) {
              if (++sa == C::NA) { sa = 0; pa ^= 1u; }
            }
          }
        }
        if (elect_one_sync()) umma_commit_2sm(tfull(as));
        __syncwarp();
        if (++as == C::NACC) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    epilogue_fold<BN, Epi2<BN>::NG, true, TS>(p, tmem_base, warp - 4, lane, tfull(0), tempty(0), rank, &om, base);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

template <int BN, bool TS>
static int launch_fold_pair_ts(const CUtensorMap& tmA, const CUtensorMap& tmB, const FoldOutMaps& om, const ConvParams& p,
                               int sm_count, cudaStream_t s) {
  using C = CfgFP<BN>;
  auto kern = conv3x3_fold_pair_kernel<BN, TS>;
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_done = true;
  }
  const int smem_bytes = (TS ? p.tma_store : 0) + C::NA * AW_SLOT + p.nbs * C::B_BYTES + C::NBAR * 8 + 16 + 1024;
  const int max_pairs = sm_count / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(Epi2<BN>::THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AST_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, om, p));
  AST_CHECK_LAUNCH();
  return 0;
}

template <int BN>
static int launch_fold_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvParams p, int sm_count,
                            cudaStream_t s) {
  using C = CfgFP<BN>;
  const bool resident = (BN == 64) && p.Cin == KBLK && p.Cout == BN;
  p.nbs = resident ? 16 : C::NBG * C::GSLOTS;
  FoldOutMaps om = {};
  p.tma_store = 0;
  if constexpr (BN <= 128) {
    // TMA-store epilogue where the staging (64 KB) fits beside the operands: the resident 64 -> 64 layer by default
    // (store-bound); AST_CONV_TMA_STORE=0 / 1 forces it off / on where it fits
    static const int ts_env = getenv("AST_CONV_TMA_STORE") ? atoi(getenv("AST_CONV_TMA_STORE")) : -1;
    const int staging = 4 * TILE_M * 128;
    const bool want = ts_env < 0 ? resident : ts_env != 0;
    auto fits_with = [&](int nbs) { return staging + C::NA * AW_SLOT + nbs * C::B_BYTES + C::NBAR * 8 + 16 + 1024 <= 226 * 1024; };
    bool fits = fits_with(p.nbs);
    if (want && !fits && !resident && fits_with(2 * C::GSLOTS)) { p.nbs = 2 * C::GSLOTS; fits = true; }   // a shallower weight ring
    if (want && fits && p.out && aligned16(p.out) && p.Cout % 64 == 0) {
      bool ok = true;
      for (int py = 0; py < 2 && ok; ++py)
        for (int px = 0; px < 2 && ok; ++px) {
          // output pixel (2h+py, 2w+px) of the padded NHWC buffer [N][Ho+2][Wo+2][Cout] as a tensor over (c, w, h, n)
          const uint64_t odims[4] = {(uint64_t)p.Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
          const uint64_t ostr[3] = {(uint64_t)2 * p.Cout * 2, (uint64_t)2 * (p.Wo + 2) * p.Cout * 2,
                                    (uint64_t)(p.Ho + 2) * (p.Wo + 2) * p.Cout * 2};
          const uint32_t obox[4] = {64, T2_W, 4, 1};      // one warp's 32 pixels: 4 tile rows of 8
          const __nv_bfloat16* view = p.out + ((int64_t)(py + 1) * (p.Wo + 2) + (px + 1)) * p.Cout;
          ok = encode_bf16_map(&om.m[py * 2 + px], view, 4, odims, ostr, obox) == 0;
        }
      if (ok) p.tma_store = staging;
    }
    if (!p.tma_store && !resident) p.nbs = C::NBG * C::GSLOTS;
    if (p.tma_store) return launch_fold_pair_ts<BN, true>(tmA, tmB, om, p, sm_count, s);
  }
  return launch_fold_pair_ts<BN, false>(tmA, tmB, om, p, sm_count, s);
}
