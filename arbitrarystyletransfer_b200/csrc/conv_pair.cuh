// K2p: the kw-box implicit-GEMM conv of conv3x3_tc2_kernel on CTA PAIRS (tcgen05.mma.cta_group::2, M = 256).
// Included by conv_tc.cu; same tile geometry (16 rows x 8 columns of pixels per CTA, one {64 ch, 8 w, 18 h} A box per
// kw serving kh = 0..2), same epilogue (epilogue_loop<..., P2 = true>), same reference work replaced
// (/root/reference/models.py:186-240 VGG convs, models.py:598-628 decoder convs).
//
// Why: with both operands in shared memory an M128 x N x K16 instruction reads (128 + N) x 32 B; at N = 64 / 128
// that is 6 / 8 KB per 32 / 64 tensor cycles, i.e. the 128 B/clk shared-memory port is the limit (48 cycles measured at
// N = 64, tools/ubench/mma_rate.cu) and the TMA writes of the next operands compete for the same port.  A CTA pair
// runs ONE M = 256 instruction over the two SMs of a TPC: each CTA holds its own 128 pixel rows of A and only HALF of
// the weight tile (N/2 rows), so per SM and instruction 5 / 6 / 8 KB are read for N = 64 / 128 / 256, the weight
// TMA traffic per SM halves, and issue + commit costs are paid once per 256 rows.
//
// Protocol (leader = cluster rank 0):
//   A/B full barriers   live in the LEADER; both CTAs' TMA loads (cp.async.bulk.tensor.cta_group::2) complete their
//                       bytes there, the leader's producer arms them with the sum of both CTAs' bytes;
//   A/B empty barriers  live in BOTH CTAs; the leader's MMA warp releases a slot with a multicast tcgen05.commit;
//   tfull               in both CTAs (multicast commit); each CTA's epilogue warps read their own TMEM;
//   tempty              in the leader, count = both CTAs' epilogue warps (rank 1 arrives remotely).
// An odd number of spatial tiles leaves rank 1 of the last pair a phantom tile (image index N): its TMA boxes are out
// of bounds (zero fill) and its epilogue stores nothing.
#pragma once

template <int BN>
struct CfgP {
  static constexpr int HB = BN / 2;                       // weight rows per CTA
  static constexpr int B_BYTES = HB * KBLK * 2;           // per tap and CTA
  static constexpr int NA_MAX = 8;                        // A ring slots (p.na of them are used)
  static constexpr int NACC = (BN <= 128) ? 4 : 2;
  static constexpr int TMEM_COLS = NACC * BN;
  static constexpr int NBAR = 2 * NA_MAX + 2 * 3 + 2 * NACC;
  static constexpr int SMEM_BUDGET = 226 * 1024;          // of the 227 KB a CTA may have
  static constexpr int smem_bytes(int na, int nbs, int a_bytes = A2_BYTES) {
    return na * a_bytes + nbs * B_BYTES + NBAR * 8 + 16 + 1024;
  }
};

// Single A box (p.wide_a, the default): ONE box {64 ch, 10 w, 18 h} per (tile, 64-channel block) serves all nine taps.
// Its pixel rows have a pitch of 10 pixels = 1280 B, so tap (kh, kw) starts (kh * 10 + kw) * 128 B into it, the 8
// pixels of a tile row are 8 consecutive 128-byte rows, and the stride between 8-row groups (SBO) is 1280 B.  Neither
// the start nor the groups are aligned to the 1024-byte swizzle atom -- and they need not be: measured here
// (tests/test_gpu_conv.py bit-compares against the direct kernel), the tensor core applies the 128-byte swizzle to
// the ABSOLUTE shared-memory address of every 16-byte chunk (bits [4,7) ^= bits [7,10)), exactly as TMA does when it
// writes the box, so any 16-byte-aligned start and any SBO address the right data; the descriptor's base-offset
// field must stay 0 (with the phase (start >> 7) & 7 in it the results are wrong).
// 22.5 KB per (tile, block) instead of 3 x 18 KB of L2 -> shared-memory traffic (the 64- and 128-channel layers
// were bound by the chip-wide L2 bandwidth: an L2 PREFETCH of the next tile's boxes made them 17 % slower), one TMA
// operation and one barrier round instead of three.
constexpr int WA_W = T2_W + 2;
constexpr int AW_BYTES = T2_BOX_H * WA_W * KBLK * 2;             // 23 040 B landed per box
constexpr int AW_SLOT = (AW_BYTES + 1023) / 1024 * 1024;         // ring slots stay 1024-byte aligned (TMA swizzle)

// Shared-memory plan of one launch: weight slots (9 = a ring of three kw groups; 9 * Cin/64 = the layer's whole half
// weight tile stays RESIDENT, loaded once per CTA) and A ring depth.  Re-streaming the weights for every tile costs
// 9 * Cin * BN bytes of L2 -> SM traffic per tile and CTA; with Cin = 128 that made the 128-channel layers
// operand-supply bound (the chip-wide L2 cap is ~6300 B/clk = 42 B/clk per SM: guide B300_MICROARCH.md "LTS cap";
// measured here: the MMA warp waited for operands 53 % of the time at 42 B/clk per SM).
template <int BN>
static void pair_plan(int cblocks, int n_blocks, int wide, int staging, int* na, int* nbs) {
  using C = CfgP<BN>;
  const int a_bytes = wide ? AW_SLOT : A2_BYTES, a_min = wide ? 2 : 4;
  const int res_slots = 9 * cblocks;
  const int budget = C::SMEM_BUDGET - staging;
  int a = (budget - C::smem_bytes(0, res_slots)) / a_bytes;
  if (n_blocks == 1 && a >= a_min) {
    *nbs = res_slots;
    *na = a < C::NA_MAX ? a : C::NA_MAX;
    return;
  }
  *nbs = 9;
  a = (budget - C::smem_bytes(0, 9)) / a_bytes;
  *na = a < C::NA_MAX ? a : C::NA_MAX;
}

// TG: tile sets of the epilogue (epilogue_loop): the 4 * Epi2<BN>::NG epilogue warps work as TG sets on different
// tiles, Epi2<BN>::NG / TG column groups each.
template <int BN, int EPI, int TG, bool TS>
__global__ void __launch_bounds__(Epi2<BN>::THREADS, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const ConvParams p) {
  using C = CfgP<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int NA = p.na;
  const bool wide = p.wide_a != 0;
  const uint32_t a_slot_bytes = wide ? AW_SLOT : A2_BYTES, a_tx_bytes = wide ? AW_BYTES : A2_BYTES;
  // [TMA-store staging (1024-byte aligned, p.tma_store bytes)] [A ring] [weight slots] [barriers]
  const uint32_t a_base = base + (TS ? (uint32_t)p.tma_store : 0u);
  const uint32_t b_base = a_base + NA * a_slot_bytes;
  const uint32_t bars = b_base + p.nbs * C::B_BYTES;
  auto afull = [&](int s) { return bars + 8u * s; };
  auto aempty = [&](int s) { return bars + 8u * (C::NA_MAX + s); };
  auto bfull = [&](int s) { return bars + 8u * (2 * C::NA_MAX + s); };
  auto bempty = [&](int s) { return bars + 8u * (2 * C::NA_MAX + 3 + s); };
  auto tfull = [&](int s) { return bars + 8u * (2 * C::NA_MAX + 6 + s); };
  auto tempty = [&](int s) { return bars + 8u * (2 * C::NA_MAX + 6 + C::NACC + s); };
  const uint32_t tmem_slot = bars + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + (TS ? p.tma_store : 0) + NA * a_slot_bytes + p.nbs * C::B_BYTES + 8 * C::NBAR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair0 = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int cblocks = p.Cin / KBLK;
  const bool resident = p.nbs == 9 * cblocks && p.n_blocks == 1;   // pair_plan

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TS) tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < 3; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int s = 0; s < C::NACC; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 2 * 4 * (Epi2<BN>::NG / TG)); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own pixels, own half of the weight rows) =====================
    if (lane == 0) {
      const uint32_t afull_l = mapa_shared(afull(0), 0), bfull_l = mapa_shared(bfull(0), 0);
      if (resident) {
        // slot of (cb, kw, kh) = (cb * 3 + kw) * 3 + kh; one barrier for the lot
        if (rank == 0) mbar_expect_tx(bfull(0), 2 * p.nbs * C::B_BYTES);
        for (int cb = 0; cb < cblocks; ++cb)
          for (int kw = 0; kw < 3; ++kw)
            for (int kh = 0; kh < 3; ++kh)
              tma_load_3d_2sm(b_base + ((cb * 3 + kw) * 3 + kh) * C::B_BYTES, &tmB, bfull_l, cb * KBLK, rank * C::HB,
                              kh * 3 + kw);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      long long dbg_pa = 0, dbg_pb = 0;
      for (int q = pair0; q < p.num_tiles; q += npairs) {
        int t = q;
        const int nb = t % p.n_blocks; t = 2 * (t / p.n_blocks) + rank;
        const int twi = t % p.tiles_w; t /= p.tiles_w;
        const int thi = t % p.tiles_h;
        const int n = t / p.tiles_h;
        const int h0 = thi * T2_H, w0 = twi * T2_W;
        if (p.pf_dist > 0 && q + p.pf_dist * npairs < p.num_tiles) {
          // L2 prefetch of the A boxes of the tile this CTA takes pf_dist iterations from now (kw = 0 and 2 cover
          // the ten columns): a tile's first touch of its input comes from HBM
          int u = q + p.pf_dist * npairs;
          u = 2 * (u / p.n_blocks) + rank;
          const int ptw = u % p.tiles_w; u /= p.tiles_w;
          const int pth = u % p.tiles_h, pn = u / p.tiles_h;
          for (int cb = 0; cb < cblocks; ++cb) {
            tma_prefetch_l2_4d(&tmA, cb * KBLK, ptw * T2_W, pth * T2_H, pn);
            tma_prefetch_l2_4d(&tmA, cb * KBLK, ptw * T2_W + 2, pth * T2_H, pn);
          }
        }
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int kw = 0; kw < 3; ++kw) {
            if (!wide || kw == 0) {
              mbar_wait_acc(aempty(sa), pa ^ 1u, kdbg_buf(p) != nullptr, dbg_pa, p.prod_sleep_ns);
              if (rank == 0) mbar_expect_tx(afull(sa), 2 * a_tx_bytes);
              tma_load_4d_2sm(a_base + sa * a_slot_bytes, &tmA, afull_l + 8u * sa, cb * KBLK, w0 + kw, h0, n);
              if (++sa == NA) { sa = 0; pa ^= 1u; }
            }
            if (!resident) {
              mbar_wait_acc(bempty(sb), pb ^ 1u, kdbg_buf(p) != nullptr, dbg_pb, p.prod_sleep_ns);
              if (rank == 0) mbar_expect_tx(bfull(sb), 2 * 3 * C::B_BYTES);
              for (int kh = 0; kh < 3; ++kh)
                tma_load_3d_2sm(b_base + (sb * 3 + kh) * C::B_BYTES, &tmB, bfull_l + 8u * sb, cb * KBLK,
                                nb * BN + rank * C::HB, kh * 3 + kw);
              if (++sb == 3) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
      if (kdbg_buf(p)) { kdbg_buf(p)[blockIdx.x * 8 + 0] = dbg_pa; kdbg_buf(p)[blockIdx.x * 8 + 1] = dbg_pb; }
      // drain: every multicast release aimed at this CTA has landed before it may exit
      for (int i = 0; i < NA; ++i) {
        mbar_wait(aempty(sa), pa ^ 1u);
        if (++sa == NA) { sa = 0; pa ^= 1u; }
      }
      if (!resident) {
        for (int i = 0; i < 3; ++i) {
          mbar_wait(bempty(sb), pb ^ 1u);
          if (++sb == 3) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader's converged warp, one elected lane =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * TILE_M, BN);
      const uint64_t a_desc0 = wide ? make_sdesc_k128_sbo(a_base, WA_W * KBLK * 2, 0) : make_sdesc_k128(a_base);
      const uint64_t b_desc0 = make_sdesc_k128(b_base);
      constexpr uint64_t B_SLOT16 = C::B_BYTES >> 4;
      const uint64_t A_SLOT16 = a_slot_bytes >> 4, KH16 = ((wide ? WA_W : T2_W) * KBLK * 2) >> 4;
      const uint64_t KW_STEP = wide ? (uint64_t)8 : 0;      // single box: the kw shift is one pixel = 128 B
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int as = 0;
      uint32_t aphase = 0;
      bool b_resident_ready = false;
      long long dbg_mt = 0, dbg_ma = 0, dbg_mb = 0;
      const long long dbg_m0 = clock64();
      const bool dbg = kdbg_buf(p) != nullptr;
      for (int q = pair0; q < p.num_tiles; q += npairs) {
        mbar_wait_acc(tempty(as), aphase ^ 1u, dbg, dbg_mt);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        uint32_t accum = 0;
        for (int cb = 0; cb < cblocks; ++cb) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            if (!wide || kw == 0) mbar_wait_acc(afull(sa), pa, dbg, dbg_ma);
            const uint64_t ad = a_desc0 + (uint64_t)sa * A_SLOT16 + (uint64_t)kw * KW_STEP;
            int grp;
            if (resident) {
              grp = cb * 3 + kw;
              if (!b_resident_ready) { mbar_wait(bfull(0), 0u); b_resident_ready = true; }  // first tile only
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
                  umma_bf16_2sm(d_tmem, ad + (uint64_t)(kh * KH16 + k * 2), bd + (uint64_t)(kh * B_SLOT16 + k * 2),
                                idesc, (kh | k) ? 1u : accum);
                }
              }
              if (!resident) umma_commit_2sm(bempty(sb));
              if (!wide || kw == 2) umma_commit_2sm(aempty(sa));
            }
            __syncwarp();
            accum = 1u;
            if (!resident) {
              if (++sb == 3) { sb = 0; pb ^= 1u; }
            }
            if (!wide || kw == 2) {
              if (++sa == NA) { sa = 0; pa ^= 1u; }
            }
          }
        }
        if (elect_one_sync()) umma_commit_2sm(tfull(as));
        __syncwarp();
        if (++as == C::NACC) { as = 0; aphase ^= 1u; }
      }
      if (kdbg_buf(p) && lane == 0) {
        kdbg_buf(p)[blockIdx.x * 8 + 2] = dbg_ma + dbg_mb;
        kdbg_buf(p)[blockIdx.x * 8 + 3] = dbg_mt;
        kdbg_buf(p)[blockIdx.x * 8 + 6] = clock64() - dbg_m0;
      }
    }
  } else if (warp >= 4) {
    epilogue_loop<BN, EPI, T2_W, Epi2<BN>::NG / TG, C::NACC, TG, true, TS>(p, tmem_base, warp - 4, lane, tfull(0),
                                                                           tempty(0), rank, &tmOut, base);
  }

  tc_fence_before();
  cluster_sync_all();          // no CTA exits (or frees TMEM) while its partner may still touch it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int EPI, int TG, bool TS>
static int launch_pair_tg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const ConvParams& p,
                          int sm_count, cudaStream_t s) {
  using C = CfgP<BN>;
  auto kern = conv3x3_pair_kernel<BN, EPI, TG, TS>;
  static bool attr_done = false;
  if (!attr_done) {
    AST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BUDGET));
    attr_done = true;
  }
  const int smem_bytes = C::smem_bytes(p.na, p.nbs, p.wide_a ? AW_SLOT : A2_BYTES) + (TS ? p.tma_store : 0);
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
  AST_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, p));
  AST_CHECK_LAUNCH();
  return 0;
}

// AST_CONV_TG (tile sets of the epilogue, see epilogue_loop) was measured at 1 / 2 / 4: no gain for any layer
// (profiles/r2_conv_pair_tile_sets.txt), so only TG = 1 is instantiated.
template <int BN, int EPI>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const ConvParams& p,
                       int sm_count, cudaStream_t s) {
  if constexpr (EPI == AST_EPI_PLAIN && BN <= 128) {
    if (p.tma_store) return launch_pair_tg<BN, EPI, 1, true>(tmA, tmB, tmOut, p, sm_count, s);
  }
  return launch_pair_tg<BN, EPI, 1, false>(tmA, tmB, tmOut, p, sm_count, s);
}

template <int BN>
static int launch_pair_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, ConvParams p,
                           int sm_count, cudaStream_t s) {
  // TMA-store epilogue (plain tiles, BN <= 128, two staging buffers).  It pays where the epilogue's per-lane sector
  // stores are the limiter and costs where the operand rings need the shared memory more: measured enc_conv3 (Cin 64)
  // 230 -> 209 us, dec_conv7 (Cin 128) 242 -> 242 us, dec_conv5 (Cin 256) 211 -> 220 us
  // (profiles/r2_conv_pair_tma_store.txt; a single staging buffer was slower everywhere).  Default: Cin <= 64;
  // AST_CONV_TMA_STORE=0 / 1 forces it off / on (A/B).
  static const int ts_env = getenv("AST_CONV_TMA_STORE") ? atoi(getenv("AST_CONV_TMA_STORE")) : -1;
  const bool ts_want = ts_env < 0 ? p.Cin <= 64 : ts_env != 0;
  CUtensorMap tmOut = tmA;
  p.tma_store = 0;
  if (ts_want && epi == AST_EPI_PLAIN && BN <= 128 && p.out && !p.tap && aligned16(p.out) && p.Cout % 64 == 0) {
    // the INTERIOR of the padded NHWC output [N][H+2][W+2][Cout]: boxes are clipped at W and H, the halo is not touched
    const uint64_t odims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    const uint64_t ostr[3] = {(uint64_t)p.Cout * 2, (uint64_t)(p.Wo + 2) * p.Cout * 2,
                              (uint64_t)(p.Ho + 2) * (p.Wo + 2) * p.Cout * 2};
    const uint32_t obox[4] = {64, T2_W, T2_H, 1};
    const __nv_bfloat16* interior = p.out + ((int64_t)(p.Wo + 2) + 1) * p.Cout;
    if (encode_bf16_map(&tmOut, interior, 4, odims, ostr, obox) == 0) p.tma_store = 2 * (BN / 64) * TILE_M * 128;   // two buffers
  }
  pair_plan<BN>(p.Cin / KBLK, p.n_blocks, p.wide_a, p.tma_store, &p.na, &p.nbs);
  switch (epi) {
    case AST_EPI_PLAIN: return launch_pair<BN, AST_EPI_PLAIN>(tmA, tmB, tmOut, p, sm_count, s);
    case AST_EPI_POOL2: return launch_pair<BN, AST_EPI_POOL2>(tmA, tmB, tmOut, p, sm_count, s);
    case AST_EPI_UP2: return launch_pair<BN, AST_EPI_UP2>(tmA, tmB, tmOut, p, sm_count, s);
  }
  return AST_E_BADARG;
}
