// K2g: the first TWO VGG layers in one kernel -- Normalization + conv1_1 (3 -> 64, zero pad) + ReLU + conv1_2 (64 -> 64,
// zero pad) + ReLU + MaxPool2d(2, 2) -- on CTA pairs.  Included by conv_tc.cu.
//
// Reference work replaced (paths relative to /root/reference): models.py:129-131 (Normalization), models.py:198-224
// (vgg19.features[0..4]: conv_1, relu_1, conv_2, relu_2, pool_2) as used by PretrainedEncoder.forward models.py:230-240.
//
// Why: as two kernels the 64-channel full-resolution map is written by conv1_1 (1.07 GB at batch 32, 512x512) and read
// back by conv1_2, twice per stylisation; conv1_1 alone is an HBM-bound 0.24-0.28 ms launch.  Here that map exists
// only as (a) fp32 accumulators in TMEM and (b) the bf16 A operand of conv1_2 in shared memory.
//
// Per CTA and tile (16 x 8 output pixels of conv1_2 before the pool): conv1_1 is needed on the 18 x 10 pixels around it
// -- exactly the {64 ch, 10 w, 18 h} single A box of conv3x3_pair_kernel, so conv1_2's 36 MMAs and their descriptors are
// unchanged; only the box is WRITTEN BY WARPS instead of by TMA:
//   TMA warp        fp32 image patch {16 w, 20 h, 3 c} (from column w0 - 4: TMA needs 16-byte aligned inner coordinates)
//   8 producer warps (two sets alternating tiles)  im2col: 180 rows x K = 32 (27 taps, normalised, zero outside the image;
//                   k = 27, 28 hold ones against the bias rows of B1) as two 128-row no-swizzle operands
//   MMA warp (leader)  conv1_1: two M = 256 pair-MMAs of K = 32 per 128-row block into TMEM stage s (2 x 64 columns);
//                   then conv1_2's 36 pair-MMAs out of the A2 slot
//   8 mid warps     TMEM -> ReLU -> bf16 -> the A2 slot in the 128-byte-swizzle layout (row r at r * 128 B, chunk c at
//                   c ^ (r & 7)); rows whose pixel lies outside the image are written as ZEROS (conv1_2 zero-pads
//                   conv1_1's OUTPUT)
//   8 epilogue warps  epilogue_loop<64, POOL2>: + bias2, ReLU, 2x2 max-pool, bf16 NHWC stores
// Barrier protocol as in conv_pair.cuh: "full" barriers fed by both CTAs live in the leader, "empty" / accumulator-full
// barriers are signalled in both CTAs by multicast tcgen05.commit.
#pragma once

constexpr int G_WARP_PROD0 = 4, G_PSETS = 2;
constexpr int G_WARP_MID0 = G_WARP_PROD0 + 4 * G_PSETS;      // 12
constexpr int G_WARP_EPI0 = G_WARP_MID0 + 8;                 // 20
constexpr int G_EPI_NG = 2, G_EPI_TG = 1;                    // eight epilogue warps on one tile, 32 columns per tcgen05.ld
constexpr int G_THREADS = 32 * (G_WARP_EPI0 + 4 * G_EPI_NG * G_EPI_TG); // 896
constexpr int G_ROWS = T2_BOX_H * WA_W;                      // 180 conv1_1 pixels per tile
constexpr int G_PW = 16, G_PH = T2_H + 4, G_X0 = 2;          // patch: columns [w0 - 4, w0 + 12), rows [h0 - 2, h0 + 18)
constexpr int G_PATCH_FLOATS = 3 * G_PH * G_PW;              // 960
constexpr int G_PATCH_BYTES = G_PATCH_FLOATS * 4;            // 3 840
constexpr int G_PSTAGES = 4;
constexpr int G_A1_BLOCK = TILE_M * F_K * 2;                 // 8 KB: 128 rows x 32 k, no-swizzle canonical
constexpr int G_A1_STAGE = 2 * G_A1_BLOCK;                   // two row blocks (180 rows)
constexpr int G_B1_BYTES = 32 * F_K * 2;                     // this CTA's 32 of the 64 output channels
constexpr int G_NA2 = 3;
constexpr int G_B2_TILE = 32 * KBLK * 2;                     // 4 KB: this CTA's half of one tap's weight tile
constexpr int G_NACC = 4;
// dynamic shared memory (offsets from the 1024-byte aligned base)
constexpr int G_OFF_A2 = 0;
constexpr int G_OFF_B2 = G_OFF_A2 + G_NA2 * AW_SLOT;         // 70 656
constexpr int G_OFF_A1 = G_OFF_B2 + 9 * G_B2_TILE;           // + 36 864
constexpr int G_OFF_B1 = G_OFF_A1 + G_PSETS * G_A1_STAGE;    // + 32 768
constexpr int G_OFF_PATCH = G_OFF_B1 + G_B1_BYTES;           // + 2 048
constexpr int G_OFF_BAR = G_OFF_PATCH + G_PSTAGES * G_PATCH_BYTES;
constexpr int G_NBAR = 2 * G_PSTAGES + 4 * G_PSETS + 2 * G_NA2 + 1 + 2 * G_NACC;
constexpr int G_SMEM_BYTES = G_OFF_BAR + G_NBAR * 8 + 16 + 1024;

// V: bit 0 = per-role register budgets (setmaxnreg), bit 1 = mid warps work on 32 columns at a time; DBG: elimination
// flags and wait counters compiled in (AST_CONV_DEBUG / AST_CONV_DBGFLAGS)
template <int V, bool DBG>
__global__ void __launch_bounds__(G_THREADS, 1)
conv12_fused_pair_kernel(const __grid_constant__ CUtensorMap tmImg, const __grid_constant__ CUtensorMap tmB2,
                         const FirstParams fp, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a2_base = base + G_OFF_A2, b2_base = base + G_OFF_B2, a1_base = base + G_OFF_A1;
  const uint32_t b1_base = base + G_OFF_B1, bars = base + G_OFF_BAR;
  auto pfull = [&](int s) { return bars + 8u * s; };
  auto pempty = [&](int s) { return bars + 8u * (G_PSTAGES + s); };
  auto a1full = [&](int s) { return bars + 8u * (2 * G_PSTAGES + s); };                       // leader
  auto a1empty = [&](int s) { return bars + 8u * (2 * G_PSTAGES + G_PSETS + s); };            // both (multicast)
  auto t1full = [&](int s) { return bars + 8u * (2 * G_PSTAGES + 2 * G_PSETS + s); };         // both (multicast)
  auto t1empty = [&](int s) { return bars + 8u * (2 * G_PSTAGES + 3 * G_PSETS + s); };        // leader
  auto a2full = [&](int s) { return bars + 8u * (2 * G_PSTAGES + 4 * G_PSETS + s); };         // leader
  auto a2empty = [&](int s) { return bars + 8u * (2 * G_PSTAGES + 4 * G_PSETS + G_NA2 + s); };  // both (multicast)
  const uint32_t b2full = bars + 8u * (2 * G_PSTAGES + 4 * G_PSETS + 2 * G_NA2);              // leader
  auto tfull = [&](int s) { return b2full + 8u * (1 + s); };                                  // both (multicast)
  auto tempty = [&](int s) { return b2full + 8u * (1 + G_NACC + s); };                        // leader
  const uint32_t tmem_slot = bars + 8u * G_NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + G_OFF_BAR + 8 * G_NBAR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair0 = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int ntiles = pair0 < p.num_tiles ? (p.num_tiles - pair0 + npairs - 1) / npairs : 0;   // tiles of this CTA

  // conv1_1 weights -> bf16 [32 co of this CTA][32 k] (k = ci*9 + kh*3 + kw; k = 27, 28: the bias as a two-term bf16
  // split against the ones in A1; zero beyond) in the no-swizzle canonical layout
  for (int i = threadIdx.x; i < 32 * F_K; i += G_THREADS) {
    const int col = i / F_K, k = i % F_K;
    const int co = rank * 32 + col;
    float v = k < 27 ? fp.w[co * 27 + k] : 0.f;
    if (fp.bias && (k == 27 || k == 28)) {
      const float bv = fp.bias[co];
      const float hi = __bfloat162float(__float2bfloat16_rn(bv));
      v = k == 27 ? hi : bv - hi;
    }
    const uint32_t off = (uint32_t)(col >> 3) * F_SBO + (uint32_t)(k >> 3) * F_LBO + (col & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + G_OFF_B1 + off) = __float2bfloat16_rn(v);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmImg);
    tma_prefetch_desc(&tmB2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G_PSTAGES; ++s) { mbar_init(pfull(s), 1); mbar_init(pempty(s), 4); }
    for (int s = 0; s < G_PSETS; ++s) {
      mbar_init(a1full(s), 2 * 4);          // the four warps of producer set s, in both CTAs
      mbar_init(a1empty(s), 1);
      mbar_init(t1full(s), 1);
      mbar_init(t1empty(s), 2 * 8);         // the eight mid warps, in both CTAs
    }
    for (int s = 0; s < G_NA2; ++s) { mbar_init(a2full(s), 2 * 8); mbar_init(a2empty(s), 1); }
    mbar_init(b2full, 1);
    for (int s = 0; s < G_NACC; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 2 * 4 * G_EPI_NG); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  fence_proxy_async_smem();                // B1 was written with generic stores, the MMA reads it via the async proxy
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_t1 = tmem_base, tmem_t2 = tmem_base + 256u;    // conv1_1: 2 stages x 2 blocks x 64 columns; conv1_2: 4 x 64

  // this CTA's i-th tile is spatial tile 2 * (pair0 + i * npairs) + rank (TileCursor with mult = 2)

  // Register budget per warpgroup (896 threads x 72 at launch): the control warps and the mid warps hand registers to the
  // epilogue warps, whose spills would otherwise share the shared-memory / L1 port with the MMA's operand reads
  const int dflags = DBG ? kdbg_flags(p) : 0;
  if (warp < G_WARP_PROD0) {
    if (V & 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== TMA: resident conv1_2 weights, then the image patches =====================
    if (lane == 0) {
      const uint32_t b2full_l = mapa_shared(b2full, 0);
      if (rank == 0) mbar_expect_tx(b2full, 2 * 9 * G_B2_TILE);
      for (int kw = 0; kw < 3; ++kw)
        for (int kh = 0; kh < 3; ++kh)
          tma_load_3d_2sm(b2_base + (kw * 3 + kh) * G_B2_TILE, &tmB2, b2full_l, 0, rank * 32, kh * 3 + kw);
      int ps = 0;
      uint32_t pph = 0;
      TileCursor cur;
      cur.init(p, pair0, npairs, 2, rank);
      for (int i = 0; i < ntiles; ++i, cur.next()) {
        const int h0 = cur.thi * T2_H, w0 = cur.twi * T2_W, n = cur.n;
        mbar_wait(pempty(ps), pph ^ 1u);
        mbar_expect_tx(pfull(ps), G_PATCH_BYTES);
        tma_load_4d(base + G_OFF_PATCH + ps * G_PATCH_BYTES, &tmImg, pfull(ps), w0 - 4, h0 - 2, 0, n);
        if (++ps == G_PSTAGES) { ps = 0; pph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (rank == 0 && ntiles > 0) {
      constexpr uint32_t idesc1 = make_idesc_bf16(2 * TILE_M, F_N);
      constexpr uint32_t idesc2 = make_idesc_bf16(2 * TILE_M, 64);
      // A2 descriptors are assembled from two 32-bit halves: with the usual 64-bit `desc0 + offset` form ptxas 12.9
      // folded the slot multiply into a UIMAD.WIDE and DROPPED the constant's high word (SBO, version, layout type
      // all read as 0: the MMA walked the box as a no-swizzle operand) -- found by reading the SASS of this kernel.
      const uint32_t a2_lo0 = (uint32_t)(make_sdesc_k128_sbo(a2_base, WA_W * KBLK * 2, 0) & 0xffffffffu);
      constexpr uint32_t a2_hi = ((uint32_t)(WA_W * KBLK * 2) >> 4) | (1u << 14) | (2u << 29);
      auto a2_desc = [&](uint32_t lo) {
        uint64_t d;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(a2_hi));
        return d;
      };
      const uint64_t b2_desc0 = make_sdesc_k128(b2_base);
      constexpr uint32_t A2_SLOT16 = AW_SLOT >> 4, KH16 = (WA_W * KBLK * 2) >> 4;
      constexpr uint64_t B2_SLOT16 = G_B2_TILE >> 4;
      const bool dbg = DBG && kdbg_buf(p) != nullptr;
      long long w_a1 = 0, w_t1e = 0, w_a2 = 0, w_te = 0;
      auto issue_conv11 = [&](int i) {       // tile i of this pair: A1 stage / TMEM stage s = i & 1
        const int s = i & 1;
        const uint32_t ph = (uint32_t)(i >> 1) & 1u;
        mbar_wait_acc(a1full(s), ph, dbg, w_a1);
        mbar_wait_acc(t1empty(s), ph ^ 1u, dbg, w_t1e);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int k = 0; k < F_K / 16; ++k) {
              const uint64_t ad = make_sdesc_k_noswizzle(a1_base + s * G_A1_STAGE + j * G_A1_BLOCK + k * 2 * F_LBO, F_LBO, F_SBO);
              const uint64_t bd = make_sdesc_k_noswizzle(b1_base + k * 2 * F_LBO, F_LBO, F_SBO);
              umma_bf16_2sm(tmem_t1 + (uint32_t)(s * 128 + j * 64), ad, bd, idesc1, k ? 1u : 0u);
            }
          }
          umma_commit_2sm(a1empty(s));
          umma_commit_2sm(t1full(s));
        }
        __syncwarp();
      };
      const bool no_c12 = (dflags & 256) != 0;
      const long long t_begin = clock64();
      issue_conv11(0);
      int sa = 0, as = 0;
      uint32_t pa = 0, aphase = 0;
      mbar_wait(b2full, 0u);
      for (int i = 0; i < ntiles; ++i) {
        if (i + 1 < ntiles) issue_conv11(i + 1);      // its accumulators are packed while this tile's conv1_2 runs
        mbar_wait_acc(a2full(sa), pa, dbg, w_a2);
        mbar_wait_acc(tempty(as), aphase ^ 1u, dbg, w_te);
        tc_fence_after();
        const uint32_t d_tmem = tmem_t2 + (uint32_t)(as * 64);
        const uint32_t ad0 = a2_lo0 + (uint32_t)sa * A2_SLOT16;
        if (elect_one_sync()) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              if (no_c12 && (kw | kh)) continue;
#pragma unroll
              for (int k = 0; k < KBLK / 16; ++k) {
                umma_bf16_2sm(d_tmem, a2_desc(ad0 + (uint32_t)(kw * 8 + kh * KH16 + k * 2)),
                              b2_desc0 + (uint64_t)((kw * 3 + kh) * B2_SLOT16 + k * 2), idesc2, (kw | kh | k) ? 1u : 0u);
              }
            }
          }
          umma_commit_2sm(a2empty(sa));
          umma_commit_2sm(tfull(as));
        }
        __syncwarp();
        if (++sa == G_NA2) { sa = 0; pa ^= 1u; }
        if (++as == G_NACC) { as = 0; aphase ^= 1u; }
      }
      if (dbg && lane == 0) {
        long long* d = kdbg_buf(p) + 8 * gridDim.x + 16 * blockIdx.x;
        d[0] = w_a1; d[1] = w_t1e; d[2] = w_a2; d[3] = w_te; d[4] = clock64() - t_begin;
      }
    }
  }
  } else if (warp >= G_WARP_PROD0 && warp < G_WARP_MID0) {
    // ===================== im2col producers: patch -> A1 (two sets alternate tiles) =====================
    const int pset = (warp - G_WARP_PROD0) >> 2;
    const int r0 = (int)threadIdx.x - 32 * G_WARP_PROD0 - 128 * pset;      // 0 .. 127
    const uint32_t a1full_l = mapa_shared(a1full(pset), 0);
    float sc[3], sh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sc[c] = fp.normalise ? fp.rstd[c] : 1.f;
      sh[c] = fp.normalise ? -fp.mean[c] * fp.rstd[c] : 0.f;
    }
    uint32_t uses = 0;
    const bool dbg = DBG && kdbg_buf(p) != nullptr;
    long long w_pf = 0, w_a1e = 0;
    const long long t_begin = clock64();
    TileCursor cur;
    cur.init(p, pair0 + pset * npairs, G_PSETS * npairs, 2, rank);
    for (int i = pset; i < ntiles; i += G_PSETS, cur.next()) {
      const int h0 = cur.thi * T2_H, w0 = cur.twi * T2_W;
      const int ps = i & (G_PSTAGES - 1);
      const uint32_t pph = (uint32_t)(i / G_PSTAGES) & 1u;
      // taps outside the image must be 0 AFTER normalisation (models.py:131, then Conv2d padding = 1)
      const bool border = h0 - 2 < 0 || w0 - 2 < 0 || h0 + T2_H + 2 > p.H || w0 + T2_W + 2 > p.W;
      mbar_wait_acc(pfull(ps), pph, dbg, w_pf);
      const float* pt = reinterpret_cast<const float*>(smem + G_OFF_PATCH + ps * G_PATCH_BYTES);
      uint32_t pk[2][16];
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int r = r0 + 128 * pass;
        if (r < G_ROWS && !(dflags & 64)) {
          const int rr = r / WA_W, cc = r % WA_W;
          float v[28];
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
                v[ci * 9 + kh * 3 + kw] = fmaf(pt[(ci * G_PH + rr + kh) * G_PW + G_X0 + cc + kw], sc[ci], sh[ci]);
          if (border) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const int ih = h0 - 2 + rr + kh, iw = w0 - 2 + cc + kw;
                if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) {
                  v[kh * 3 + kw] = 0.f; v[9 + kh * 3 + kw] = 0.f; v[18 + kh * 3 + kw] = 0.f;
                }
              }
          }
          v[27] = 1.f;                                   // the ones that multiply the bias rows of B1 (k = 27, 28)
#pragma unroll
          for (int q = 0; q < 14; ++q) pk[pass][q] = pack_bf16(v[2 * q], v[2 * q + 1]);
          pk[pass][14] = pack_bf16(1.f, 0.f);
          pk[pass][15] = 0u;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pempty(ps));            // this warp has read its taps
      mbar_wait_acc(a1empty(pset), (uses & 1u) ^ 1u, dbg, w_a1e);
      ++uses;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int r = r0 + 128 * pass;
        if (r < G_ROWS) {
          const int rb = r & 127;
          uint8_t* row = smem + G_OFF_A1 + pset * G_A1_STAGE + (r >> 7) * G_A1_BLOCK + (uint32_t)(rb >> 3) * F_SBO + (rb & 7) * 16;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc)
            *reinterpret_cast<uint4*>(row + kc * F_LBO) =
                make_uint4(pk[pass][4 * kc], pk[pass][4 * kc + 1], pk[pass][4 * kc + 2], pk[pass][4 * kc + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(a1full_l);
    }
    if (dbg && r0 == 0 && pset == 0) {
      long long* d = kdbg_buf(p) + 8 * gridDim.x + 16 * blockIdx.x;
      d[5] = w_pf; d[6] = w_a1e; d[7] = clock64() - t_begin;
    }
    if (uses) mbar_wait(a1empty(pset), (uses - 1u) & 1u);      // the last multicast release has landed
  } else if (warp >= G_WARP_MID0 && warp < G_WARP_EPI0) {
    // ===================== mid warps: conv1_1 accumulators -> ReLU -> bf16 -> conv1_2's A slot =====================
    if (V & 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int mw = warp - G_WARP_MID0;
    const int e = mw & 3, j = mw >> 2;                   // TMEM lane quarter (= warp % 4), 128-row block
    const int r = j * 128 + 32 * e + lane;               // row of the 18 x 10 region
    const bool valid = r < G_ROWS;
    const int rr = r / WA_W, cc = r % WA_W;
    const uint32_t t1empty_l = mapa_shared(t1empty(0), 0), a2full_l = mapa_shared(a2full(0), 0);
    int sa = 0;
    uint32_t pa = 0;
    const bool dbg = DBG && kdbg_buf(p) != nullptr;
    long long w_t1f = 0, w_a2e = 0;
    const long long t_begin = clock64();
    TileCursor cur;
    cur.init(p, pair0, npairs, 2, rank);
    for (int i = 0; i < ntiles; ++i, cur.next()) {
      const int h0 = cur.thi * T2_H, w0 = cur.twi * T2_W, n = cur.n;
      const int s = i & 1;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      const int ih = h0 - 1 + rr, iw = w0 - 1 + cc;
      const bool inimg = valid && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && n < p.N;
      mbar_wait_acc(t1full(s), ph, dbg, w_t1f);
      tc_fence_after();
      const uint32_t trow = tmem_t1 + ((uint32_t)(e * 32) << 16) + (uint32_t)(s * 128 + j * 64);
      const bool skip_mid = (dflags & 128) != 0;
      const uint32_t row = a2_base + sa * AW_SLOT + (uint32_t)r * 128u;
      const uint32_t sw = (uint32_t)(r & 7);
      if constexpr (V & 2) {
        // 32 columns at a time: fewer live registers (896 threads share the register file)
        mbar_wait_acc(a2empty(sa), pa ^ 1u, dbg, w_a2e);   // three slots: free long before the accumulators are ready
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32], pk[16];
          if (!skip_mid) {
            tmem_ld_32x32(trow + 32 * half, v);
            tmem_ld_wait();
          }
#pragma unroll
          for (int q = 0; q < 16; ++q)
            pk[q] = inimg ? pack_bf16_relu(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])) : 0u;
          if (valid && !skip_mid) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              st_shared_v4(row + (((uint32_t)(4 * half + c) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(t1empty_l + 8u * s);    // TMEM stage free for conv1_1 of tile i + 2
      } else {
        // all 64 columns at once, the TMEM stage handed back before the pack
        uint32_t v0[32], v1[32];
        if (!skip_mid) {
          tmem_ld_32x32(trow, v0);
          tmem_ld_32x32(trow + 32, v1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(t1empty_l + 8u * s);
        uint32_t pk[32];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          pk[q] = inimg ? pack_bf16_relu(__uint_as_float(v0[2 * q]), __uint_as_float(v0[2 * q + 1])) : 0u;
          pk[16 + q] = inimg ? pack_bf16_relu(__uint_as_float(v1[2 * q]), __uint_as_float(v1[2 * q + 1])) : 0u;
        }
        mbar_wait_acc(a2empty(sa), pa ^ 1u, dbg, w_a2e);
        if (valid && !skip_mid) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            st_shared_v4(row + (((uint32_t)c ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(a2full_l + 8u * sa);
      if (++sa == G_NA2) { sa = 0; pa ^= 1u; }
    }
    if (dbg && mw == 0 && lane == 0) {
      long long* d = kdbg_buf(p) + 8 * gridDim.x + 16 * blockIdx.x;
      d[8] = w_t1f; d[9] = w_a2e; d[10] = clock64() - t_begin;
    }
    if (mw == 0) {
#pragma unroll
      for (int s = 0; s < G_NA2; ++s) {
        const int used = (ntiles - s + G_NA2 - 1) / G_NA2;         // tiles s, s + 3, ... of this CTA went through slot s
        if (ntiles > s) mbar_wait(a2empty(s), (uint32_t)(used - 1) & 1u);   // the last multicast releases have landed
      }
    }
  } else if (warp >= G_WARP_EPI0) {
    if (V & 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    epilogue_loop<64, AST_EPI_POOL2, T2_W, G_EPI_NG, G_NACC, G_EPI_TG, true>(p, tmem_t2, warp - G_WARP_EPI0, lane, tfull(0),
                                                                      tempty(0), rank);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// Normalization + conv1_1 + ReLU + conv1_2 + ReLU + 2x2 max-pool: fp32 NCHW image -> bf16 native [N][H/2+2][W/2+2][64].
int conv12_fused(const float* img, const float* w1, const float* b1, const float* mean, const float* std_,
                 const void* wpk2, const float* b2, void* out, int N, int H, int W, cudaStream_t s) {
  if (!img || !w1 || !b1 || !wpk2 || !out || N <= 0 || H < 2 || W < 4) return AST_E_BADARG;
  if (W % 4 != 0 || H % 2 != 0) return AST_E_SHAPE;       // TMA rows of 16-byte multiples; the 2x2 pool
  if (!aligned16(img) || !aligned16(wpk2) || !aligned16(out) || (b2 && !aligned16(b2))) return AST_E_ALIGN;
  int sm_count = 0;
  int r = get_sm_count(&sm_count);
  if (r) return r;
  FirstParams fp = {};
  fp.img = img; fp.w = w1; fp.bias = b1; fp.normalise = (mean && std_) ? 1 : 0;
  for (int i = 0; i < 3; ++i) {
    fp.mean[i] = fp.normalise ? mean[i] : 0.f;
    fp.rstd[i] = fp.normalise ? 1.f / std_[i] : 1.f;
  }
  ConvParams p = {};
  p.N = N; p.H = H; p.W = W; p.Cin = 64; p.Cout = 64; p.Ho = H / 2; p.Wo = W / 2;
  p.relu = 1; p.halo = AST_HALO_KEEP;
  p.tiles_w = (W + T2_W - 1) / T2_W;
  p.tiles_h = (H + T2_H - 1) / T2_H;
  p.n_blocks = 1;
  const int64_t sp = (int64_t)N * p.tiles_h * p.tiles_w;
  const int64_t pairs_total = (sp + 1) / 2;
  if (pairs_total >= 0x7fffffffLL) return AST_E_SHAPE;
  p.num_tiles = (int)pairs_total;
  p.bias = b2; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  // bottleneck elimination (tools/bench_conv12.py): 2 / 4 = epilogue without stores / TMEM loads (epilogue_loop),
  // 64 = producers skip the im2col arithmetic, 128 = mid warps skip TMEM load + pack + A2 stores, 256 = no conv1_2 MMAs
  static const int dbg_flags = getenv("AST_CONV_DBGFLAGS") ? atoi(getenv("AST_CONV_DBGFLAGS")) : 0;
  p.dbg_flags = dbg_flags;
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return AST_E_NODRIVER;
  CUtensorMap tmImg, tmB2;
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)N};
    const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * 12};
    const cuuint32_t bx[4] = {G_PW, G_PH, 3, 1}, es[4] = {1, 1, 1, 1};
    CUresult cr = enc(&tmImg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(img), gdim, gstr, bx, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return AST_E_SHAPE;
  }
  {
    const uint64_t wdims[3] = {64, 64, 9};
    const uint64_t wstr[2] = {128, 64 * 128};
    const uint32_t wbox[3] = {KBLK, 32, 1};
    r = encode_bf16_map(&tmB2, wpk2, 3, wdims, wstr, wbox);
    if (r) return r;
  }
  const int max_pairs = sm_count / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  static const bool dbg_on = getenv("AST_CONV_DEBUG") != nullptr;     // debug only: allocates and synchronises
  long long* dbg = nullptr;
  if (dbg_on) {
    AST_CUDA(cudaMalloc(&dbg, sizeof(long long) * 24 * 2 * pairs));
    AST_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 24 * 2 * pairs, s));
    p.dbg = dbg;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(G_THREADS, 1, 1);
  cfg.dynamicSmemBytes = G_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int variant = getenv("AST_CONV12_V") ? atoi(getenv("AST_CONV12_V")) & 3 : 3;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const FirstParams, const ConvParams);
  static const KernelFn kerns[5] = {conv12_fused_pair_kernel<0, false>, conv12_fused_pair_kernel<1, false>,
                                    conv12_fused_pair_kernel<2, false>, conv12_fused_pair_kernel<3, false>,
                                    conv12_fused_pair_kernel<3, AST_KERNEL_DEBUG != 0>};
  const int ki = (AST_KERNEL_DEBUG && (dbg_on || dbg_flags)) ? 4 : variant;
  static bool attr_done[5] = {false, false, false, false, false};
  if (!attr_done[ki]) {
    AST_CUDA(cudaFuncSetAttribute(kerns[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
    attr_done[ki] = true;
  }
  AST_CUDA(cudaLaunchKernelEx(&cfg, kerns[ki], tmImg, tmB2, fp, p));
  AST_CHECK_LAUNCH();
  if (dbg) {
    cudaStreamSynchronize(s);
    const int g = 2 * pairs;
    long long* h = new long long[24 * g];
    cudaMemcpy(h, dbg, sizeof(long long) * 24 * g, cudaMemcpyDeviceToHost);
    double a[16] = {0}, e[2] = {0};
    for (int b = 0; b < g; ++b) {
      for (int j = 0; j < 16; ++j) a[j] += (double)h[8 * g + 16 * b + j] / ((j < 5) ? pairs : g);
      e[0] += (double)h[8 * b + 4] / g; e[1] += (double)h[8 * b + 5] / g;
    }
    const double tiles = (double)p.num_tiles / pairs;
    fprintf(stderr, "[conv12 dbg] tiles/CTA=%.1f | per tile cycles: mma loop=%.0f wait a1full=%.0f t1empty=%.0f a2full=%.0f tempty=%.0f | "
            "producer set 0 (per its tile) loop=%.0f wait patch=%.0f a1empty=%.0f | mid loop=%.0f wait t1full=%.0f a2empty=%.0f | "
            "epilogue loop=%.0f wait tfull=%.0f\n", tiles, a[4] / tiles, a[0] / tiles, a[1] / tiles, a[2] / tiles, a[3] / tiles,
            a[7] / tiles * 2, a[5] / tiles * 2, a[6] / tiles * 2, a[10] / tiles, a[8] / tiles, a[9] / tiles, e[1] / tiles, e[0] / tiles);
    delete[] h;
    cudaFree(dbg);
  }
  return 0;
}
