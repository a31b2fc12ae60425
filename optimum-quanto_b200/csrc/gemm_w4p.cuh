// Large-M fused packed-int4 linear, second generation: CTA pairs (tcgen05 cta_group::2), swap-AB, the dequantised
// weight tile lives in TENSOR MEMORY and never touches shared memory.
//
//   out[m, n] = sum_k x[m, k] * dequant(W)[n, k]  (+ bias[n])          M > 128
//
// Why.  The first-generation kernel (gemm_tc.cuh, one CTA per 256 x 224/256 tile) stages the dequantised weight tile in
// shared memory.  Per 64-k stage its SM moves 148 KB through shared memory (TMA writes A 32 KB, staging writes B 28 KB,
// the tensor core reads A 32 KB + B 2 x 28 KB) in the 896 cycles the MMAs need: 165 B/cycle against the 128 B/cycle an
// SM's shared memory delivers -- that, and an epilogue that cannot overlap the next tile (the 256 x 256 fp32 accumulator
// fills all of TMEM), is why it stops at 0.70 of the cuBLAS bf16 figure with the tensor pipe 57 % busy.  Here:
//   * the roles of the operands are swapped: D^T[n, m] = W[n, :] . x[m, :].  The weight is the M-side ("A") operand of
//     tcgen05.mma, which may be read from tensor memory: the staging warps write the dequantised tile straight into TMEM
//     (tcgen05.st, 256 B/cycle) and shared memory only carries the activations;
//   * two CTAs share one 256-feature x 256-token tile (cta_group::2): each stages the 128 out-features of ITS 64 packed
//     rows into its own TMEM and loads only 128 of the 256 tokens; the pair's tensor cores read both halves.
//   Shared-memory traffic per SM and 128-k stage: 32 KB of TMA writes + 32 KB of operand reads in 1024 MMA cycles =
//   64 B/cycle, half of what the SM can deliver, instead of 1.3x more than it can.
//   * TMEM: 256 columns hold the 128 x 256 fp32 accumulator, the other 256 columns are four 128-k operand slots, so the
//     staging warps run up to four stages (4096 MMA cycles) ahead -- also across the tile boundary, which hides most of
//     the epilogue;
//   * the epilogue is a transpose for free: TMEM lane = out-feature, column = token, so a warp's 32 lanes hold 32
//     CONSECUTIVE out-features of one token: it writes 64-byte row segments into a [token][128 B] staging tile and one
//     thread hands 32-token x 64-feature blocks to the TMA store unit -- once per output buffer (this rank's, and its
//     peers' for the fused all-gather, gather.cuh).  The bias is a per-LANE scalar in this orientation.
//
// Roles per CTA (640 threads): warp 0 raw-weight TMA, warp 1 MMA issue (leader CTA only), warp 2 TMEM allocation +
// activation TMA, warp 3 idle, warps 4-7 epilogue (TMEM lane quarter = warp % 4), warps 8-19 staging: NG = 3 groups of
// 4 warps, group g converts the stages i == g (mod NG); thread = one TMEM lane = one out-feature (quarters 0,1: the low
// nibbles of packed rows 0-63, quarters 2,3: their high nibbles), 128 k per stage.
// Every wait is watchdog-bounded (common.cuh): a protocol bug traps instead of hanging the GPU.
#pragma once

#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"

namespace qb {

struct W4PParams {
  const void* scale;   // [N * K / group] weight dtype
  const void* shift;   // same, or uint8 zero-points
  const void* bias;    // [N] or nullptr
  GatherInfo g;        // output buffers (g.n_out >= 1) and the fused all-gather protocol
  int ld, col0;        // row pitch of the output buffers in elements, first column of this call's slab
  int M, N, K;         // N = out-features of this call (the local shard)
  int group, group_log2;
  int num_tok_blocks, num_feat_blocks;
  long long* trace;
};

template <typename WT_, bool ZP_>
struct W4PCfg {
  using WT = WT_;
  static constexpr bool ZP = ZP_;
  static constexpr int TOK = 256;          // tokens per pair tile = UMMA N (128 loaded by each CTA)
  static constexpr int FEAT = 128;         // out-features per CTA = TMEM lanes (64 packed rows); 256 per pair = UMMA M
  static constexpr int KS = 128;           // k per stage
  static constexpr int NSLOT = 4;          // operand slots: TMEM A (64 columns each) + shared-memory x stage
  static constexpr int NG = 3;             // staging groups
  static constexpr int RAW_STAGES = 8;
  static constexpr int RAW_BYTES = 64 * 128;        // 64 packed rows x 128 k
  static constexpr int X_PANEL = 128 * 128;         // 128 tokens x 64 k (bf16 / fp16)
  static constexpr int X_STAGE = 2 * X_PANEL;       // 128 k
  static constexpr int D_COLS = 256;
  static constexpr int A_COLS = 64;
  static constexpr int A_COL0 = D_COLS;
  static constexpr int EPI_BUF = 32 * 128;          // 32 tokens x 64 features (one nibble half), dense
  static constexpr int EPI_BYTES = 2 * 2 * EPI_BUF;  // [half][buffer]
  static constexpr int FIRST_CVT_WARP = 8;
  static constexpr int NTHREADS = (FIRST_CVT_WARP + NG * 4) * 32;
  static constexpr int SMEM_BYTES = NSLOT * X_STAGE + RAW_STAGES * RAW_BYTES + EPI_BYTES + 512;
  static_assert(A_COL0 + NSLOT * A_COLS <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// A operand from tensor memory, cta_group::2 (each CTA's TMEM holds the 128 rows it staged, same column address)
__device__ __forceinline__ void tc_mma_f16_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::NTHREADS, 1)
    gemm_w4p_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                    const __grid_constant__ StoreMaps smaps, const W4PParams p, const uint32_t idesc) {
  using WT = typename Cfg::WT;
  constexpr bool ZP = Cfg::ZP;
  constexpr int NSLOT = Cfg::NSLOT, RS = Cfg::RAW_STAGES, NG = Cfg::NG;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* x_ring = smem;                                   // identical offsets in both CTAs
  uint8_t* raw_ring = x_ring + NSLOT * Cfg::X_STAGE;
  uint8_t* epi_stage = raw_ring + RS * Cfg::RAW_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + Cfg::EPI_BYTES);  // [NSLOT], the leader's copy is used
  uint64_t* empty_bar = full_bar + NSLOT;                  // [NSLOT], multicast commit
  uint64_t* raw_full = empty_bar + NSLOT;                  // [RS]
  uint64_t* raw_empty = raw_full + RS;                     // [RS]
  uint64_t* tmem_full_bar = raw_empty + RS;                // [1]
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;            // [1], the leader's copy is used
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 2 && lane == 0) tma_prefetch_desc(&tmap_x);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&full_bar[s], 1 + 8);  // leader's expect_tx + the four warps of the stage's staging group in BOTH CTAs
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < RS; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], 4);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 8);  // 4 epilogue warps x 2 CTAs
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.num_tok_blocks * p.num_feat_blocks;
  const int ksteps = p.K / Cfg::KS;
  const int half_n = p.N / 2;
  const int my_tiles = (num_tiles > pair) ? (num_tiles - 1 - pair) / npairs + 1 : 0;
  const int total_it = my_tiles * ksteps;  // stages this CTA pair processes, in order: (tile, ks)

  if (warp == 0) {
    // ---------------------------------------------------------------- packed-weight TMA (this CTA's 64 packed rows)
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int fb = tile / p.num_tok_blocks;
        const int row0 = fb * 128 + static_cast<int>(rank) * 64;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&raw_empty[slot], phase ^ 1u);
          mbar_arrive_expect_tx(&raw_full[slot], Cfg::RAW_BYTES);
          tma_load_2d(raw_ring + slot * Cfg::RAW_BYTES, &tmap_w, &raw_full[slot], ks * Cfg::KS, row0);
          if (++slot == RS) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------------- activation TMA (this CTA's 128 of the 256 tokens)
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t leader_full = mapa_u32(smem_u32(full_bar), 0);
      gather_wait_start(p.g);  // the activation may be the gathered output of the previous linear
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int tb = tile % p.num_tok_blocks;
        const int tok0 = tb * Cfg::TOK + static_cast<int>(rank) * 128;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&empty_bar[slot], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[slot], 2u * Cfg::X_STAGE);
          const uint32_t dst = smem_u32(x_ring + slot * Cfg::X_STAGE);
          tma_load_2d_2cta(dst, &tmap_x, leader_full + slot * 8, ks * Cfg::KS, tok0);
          tma_load_2d_2cta(dst + Cfg::X_PANEL, &tmap_x, leader_full + slot * 8, ks * Cfg::KS + 64, tok0);
          if (++slot == NSLOT) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issue (leader CTA, one thread)
    if (rank == 0 && lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      uint32_t tile_it = 0;
#ifdef QB_DEVELOPER_KNOCKOUTS
      long long t_start = clock64(), w_full = 0, w_acc = 0, tq;
#define QB_TICK() tq = clock64()
#define QB_TOCK(acc) acc += clock64() - tq
#else
#define QB_TICK()
#define QB_TOCK(acc)
#endif
      for (int tile = pair; tile < num_tiles; tile += npairs, ++tile_it) {
        QB_TICK();
        mbar_wait(tmem_empty_bar, (tile_it & 1u) ^ 1u);  // both CTAs' epilogues have drained the accumulator
        QB_TOCK(w_acc);
        tc_fence_after();
        for (int ks = 0; ks < ksteps; ++ks) {
          QB_TICK();
          mbar_wait_cluster(&full_bar[slot], phase);
          QB_TOCK(w_full);
          tc_fence_after();
          const uint32_t a_tmem = tmem_base + Cfg::A_COL0 + slot * Cfg::A_COLS;
          const uint32_t x_addr = smem_u32(x_ring + slot * Cfg::X_STAGE);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            tc_mma_f16_ts_2cta(tmem_base, a_tmem + k * 8,
                               umma_desc_sw128_kmajor(x_addr + (k >> 2) * Cfg::X_PANEL + (k & 3) * 32), idesc,
                               (ks | k) != 0 ? 1u : 0u);
          }
          tc_commit_2cta(&empty_bar[slot], 3);  // TMEM slot + x stage reusable in both CTAs once these MMAs completed
          if (++slot == NSLOT) { slot = 0; phase ^= 1u; }
        }
        tc_commit_2cta(tmem_full_bar, 3);
      }
#ifdef QB_DEVELOPER_KNOCKOUTS
      if (p.trace != nullptr && pair < 8) {  // [pair][0..3]: MMA thread: total, wait(full), wait(accumulator free), tiles
        p.trace[pair * 16 + 0] = clock64() - t_start;
        p.trace[pair * 16 + 1] = w_full;
        p.trace[pair * 16 + 2] = w_acc;
        p.trace[pair * 16 + 3] = my_tiles;
      }
#endif
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- epilogue (this CTA's 128 out-features x 256 tokens)
    const int quarter = warp & 3;            // TMEM lane quarter
    const int hf = quarter >> 1;             // 0: low-nibble features, 1: high-nibble features (+N/2)
    const int sub = quarter & 1;             // which 32 of the half's 64 features
    const uint32_t leader_tmem_empty = mapa_u32(smem_u32(tmem_empty_bar), 0);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t stage0 = smem_u32(epi_stage) + static_cast<uint32_t>(hf) * 2u * Cfg::EPI_BUF;
    const uint32_t my_col = static_cast<uint32_t>(sub * 64 + lane * 2);
    const int bar_id = 2 + hf;               // named barrier of the half's two warps
    uint32_t tile_it = 0;
    int blk = 0;                              // running count of staged 32-token blocks (selects the staging buffer)
#ifdef QB_DEVELOPER_KNOCKOUTS
    long long e_wait = 0, e_busy = 0, tq;
#endif
    for (int tile = pair; tile < num_tiles; tile += npairs, ++tile_it) {
      const int tb = tile % p.num_tok_blocks, fb = tile / p.num_tok_blocks;
      const int prow0 = fb * 128 + static_cast<int>(rank) * 64;  // first packed row of this CTA's features
      const int n_feat0 = hf * half_n + prow0;                    // first out-feature of the half's 64-wide block
      const int n_mine = n_feat0 + sub * 32 + lane;
      const bool feat_ok = prow0 + sub * 32 + lane < half_n;
      const bool store_ok = prow0 < half_n;  // (N/2) % 64 == 0: a block is entirely inside or outside
      float bias_f = 0.f;
      const bool has_bias = p.bias != nullptr;
      if (has_bias && feat_ok) bias_f = to_float<WT>(static_cast<const WT*>(p.bias)[n_mine]);
      QB_TICK();
      mbar_wait(tmem_full_bar, tile_it & 1u);
      QB_TOCK(e_wait);
      tc_fence_after();
      QB_TICK();
#pragma unroll 1
      for (int tb32 = 0; tb32 < Cfg::TOK / 32; ++tb32, ++blk) {
        const uint32_t sbuf = stage0 + static_cast<uint32_t>(blk & 1) * Cfg::EPI_BUF;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + tb32 * 32, v);
        // the store that last read this staging buffer (two blocks ago) has consumed it
        if (sub == 0 && lane == 0) bulk_wait_group_read<1>();
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        tmem_ld_wait();
        if (tb32 == Cfg::TOK / 32 - 1) {
          // the tile's last accumulator columns are in registers: hand the accumulator back NOW, so that the next tile's
          // MMAs run under this block's conversion and stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_tmem_empty);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          WT r = from_float<WT>(__uint_as_float(v[j]));
          if (has_bias) r = from_float<WT>(__fadd_rn(to_float<WT>(r), bias_f));
          const uint16_t bits = *reinterpret_cast<const uint16_t*>(&r);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(sbuf + static_cast<uint32_t>(j) * 128u + my_col), "h"(bits) : "memory");
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        if (sub == 0 && lane == 0) {
          const int tok0 = tb * Cfg::TOK + tb32 * 32;
          if (store_ok && tok0 < p.M) {
#pragma unroll 1
            for (int q = 0; q < p.g.n_out; ++q) tma_store_2d(&smaps.m[q], sbuf, p.col0 + n_feat0, tok0);
          }
          bulk_commit_group();
        }
      }
      QB_TOCK(e_busy);
    }
#ifdef QB_DEVELOPER_KNOCKOUTS
    if (p.trace != nullptr && pair < 8 && rank == 0 && warp == 4 && lane == 0) {  // [4..5]: epilogue wait / busy
      p.trace[pair * 16 + 4] = e_wait;
      p.trace[pair * 16 + 5] = e_busy;
    }
#endif
    if (sub == 0 && lane == 0) bulk_wait_group_all();  // every output store of this half has been performed
  } else if (warp >= Cfg::FIRST_CVT_WARP) {
    // ---------------------------------------------------------------- staging: raw bytes -> exact dequant -> TMEM
    using D = Dq<WT>;
    const int grp = (warp - Cfg::FIRST_CVT_WARP) >> 2;
    const int quarter = warp & 3;
    const bool high_plane = quarter >= 2;
    const int r = (quarter & 1) * 32 + lane;   // packed row inside this CTA's 64-row block
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const WT* scale = static_cast<const WT*>(p.scale);
    const int groups_per_row = p.K / p.group;
    const int sets = (p.group >= 128) ? 1 : (p.group >= 64 ? 2 : 4);  // (scale, shift) pairs inside one 128-k stage
    const uint32_t raw0 = smem_u32(raw_ring) + static_cast<uint32_t>(r) * 128;
    const uint32_t a_taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + Cfg::A_COL0;
    const uint32_t raw_full0 = smem_u32(raw_full), raw_empty0 = smem_u32(raw_empty);
    const uint32_t empty0 = smem_u32(empty_bar);
    const uint32_t leader_full = mapa_u32(smem_u32(full_bar), 0);

    struct Pre {
      WT s[4];
      uint16_t z[4];
      bool ok;
    };
    // this group's stage sequence: it = grp, grp + NG, ...; (tile, ks) advance incrementally
    int f_it = grp;
    int f_ks = grp % ksteps;
    int f_tile = pair + (grp / ksteps) * npairs;
    auto fetch = [&](Pre& pr) {
      if (f_it >= total_it) return;
      const int rp = (f_tile / p.num_tok_blocks) * 128 + static_cast<int>(rank) * 64 + r;
      pr.ok = rp < half_n;
      if (pr.ok) {
        const int kk = f_ks * Cfg::KS;
        const int g0 = (p.group_log2 >= 0) ? (kk >> p.group_log2) : (kk / p.group);
        const size_t row = static_cast<size_t>(high_plane ? rp + half_n : rp) * groups_per_row + g0;
#pragma unroll
        for (int st = 0; st < 4; ++st) {
          if (st < sets) {
            pr.s[st] = __ldg(scale + row + st);
            pr.z[st] = ZP ? static_cast<uint16_t>(__ldg(static_cast<const uint8_t*>(p.shift) + row + st))
                          : __ldg(static_cast<const uint16_t*>(p.shift) + row + st);
          }
        }
      }
      f_it += NG;
      f_ks += NG;
      while (f_ks >= ksteps) { f_ks -= ksteps; f_tile += npairs; }
    };

    Pre cur, nxt;
    fetch(cur);
    int rslot = grp % RS;
    int aslot = grp % NSLOT;
    uint32_t rphase = 0, aphase = static_cast<uint32_t>(grp / NSLOT) & 1u;
#ifdef QB_DEVELOPER_KNOCKOUTS
    long long s_raw = 0, s_slot = 0, s_cvt = 0, tq;
#endif
    for (int it = grp; it < total_it; it += NG) {
      fetch(nxt);
      // Rows beyond N/2 (ragged last feature block) must stage zeros: folded into the coefficients (scale 0, shift 0
      // give an exact 0 for every nibble), so the conversion below has no branch.
      if (!cur.ok) {
#pragma unroll
        for (int st = 0; st < 4; ++st) { cur.s[st] = from_float<WT>(0.f); cur.z[st] = 0; }
      }
      typename D::Coef kc[4];
      kc[0] = D::make_raw(cur.s[0], cur.z[0], ZP);
      if (sets > 1) {
        // kc[j] = coefficients of the 32-k quarter j of the stage
        kc[1] = (sets >= 4) ? D::make_raw(cur.s[1], cur.z[1], ZP) : kc[0];
        kc[2] = D::make_raw(cur.s[sets >= 4 ? 2 : 1], cur.z[sets >= 4 ? 2 : 1], ZP);
        kc[3] = (sets >= 4) ? D::make_raw(cur.s[3], cur.z[3], ZP) : kc[2];
      } else {
        kc[1] = kc[0];
        kc[2] = kc[0];
        kc[3] = kc[0];
      }
      QB_TICK();
      mbar_wait_u32(raw_full0 + rslot * 8, rphase);
      QB_TOCK(s_raw);
      const uint32_t a_taddr = a_taddr0 + aslot * Cfg::A_COLS;
      QB_TICK();
      mbar_wait_u32(empty0 + aslot * 8, aphase ^ 1u);  // the pair's MMAs that read this TMEM slot have completed
      QB_TOCK(s_slot);
      tc_fence_after();
      QB_TICK();
      // The thread's 128 raw bytes are read in LOGICAL k order: 16-k chunk c sits at physical chunk c ^ sw of the row
      // (SWIZZLE_128B).  The row base has its low 7 bits clear, so the chunk address is (row | sw << 4) ^ (c << 4): one
      // LOP per chunk, all eight loads in flight before the first conversion.  The swizzle must stay on the LOAD side:
      // tcgen05.st is warp-collective and takes ONE (uniform) tensor-memory address, lane i writes TMEM lane i -- a
      // per-lane column offset (tried: physical load order, swizzled store column) is not expressible.  (The first
      // version kept a branch and a 32-register staging array per 64-k half: 432 instructions, 2070 cycles per stage.)
      {
        const uint32_t rbase = (raw0 + rslot * Cfg::RAW_BYTES) | (sw << 4);
        uint4 raw[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) raw[c] = ld_shared_v4(rbase ^ (static_cast<uint32_t>(c) << 4));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t o8[8];
          dequant16_plane<WT, ZP>(raw[c], high_plane, kc[c >> 1], o8);
          tmem_st_32x32b_x8(a_taddr + (static_cast<uint32_t>(c) << 3), o8);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_u32(raw_empty0 + rslot * 8);        // the stores above consumed the raw bytes (data dependency)
        mbar_arrive_cluster(leader_full + aslot * 8);   // this warp's quarter of the slot is in tensor memory
      }
      QB_TOCK(s_cvt);
      aslot += NG;
      if (aslot >= NSLOT) { aslot -= NSLOT; aphase ^= 1u; }
      rslot += NG;
      if (rslot >= RS) { rslot -= RS; rphase ^= 1u; }
      cur = nxt;
    }
#ifdef QB_DEVELOPER_KNOCKOUTS
    if (p.trace != nullptr && pair < 8 && rank == 0 && warp == Cfg::FIRST_CVT_WARP && lane == 0) {
      p.trace[pair * 16 + 6] = s_raw;   // staging group 0, warp 0: wait raw bytes / wait TMEM slot / convert + store
      p.trace[pair * 16 + 7] = s_slot;
      p.trace[pair * 16 + 8] = s_cvt;
      p.trace[pair * 16 + 9] = (total_it - grp + NG - 1) / NG;
    }
#endif
  }

  // Neither CTA may exit (or free TMEM) while its peer can still signal its barriers or read its operands.
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0 && rank == 0) gather_signal_end(p.g, gridDim.x >> 1);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace qb
