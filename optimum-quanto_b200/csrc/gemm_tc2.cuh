// CTA-pair (tcgen05 cta_group::2) GEMM for 8-bit x 8-bit operands: int8 x int8 -> s32 and fp8 x fp8 -> f32
// (quanto::qbytes_mm with quantized activations, library/qbytes_mm.py:36-52 / :55-70 in the reference).
//
// Why a pair: with one CTA per SM and a 128 x 256 tile every MAC pulls (128 + 256) rows of K through L2 -> shared
// memory per 128 x 256 outputs; measured, that stream (~9 TB/s at 1.5 POP/s) is what bounds the 1-CTA kernel, not
// the tensor pipe.  Two CTAs of a cluster share one 256 x BN tile: each loads its own 128 rows of A and only its
// half (BN/2 rows) of W, the tensor cores of both SMs read both halves, so the L2 -> SM stream per MAC drops by
// 1.5x and the shared-memory read per MMA from 12 KB to 8 KB per SM.
//
// Protocol (rank = %cluster_ctarank, leader = rank 0):
//   warp 0 (both CTAs)  TMA producer: waits its LOCAL empty barrier, issues its two loads with .cta_group::2 so the
//                       bytes are counted on the LEADER's full barrier (the leader arms it for both CTAs' bytes).
//   warp 1 (leader)     one thread issues tcgen05.mma.cta_group::2 (M = 256), commits with .multicast::cluster to the
//                       empty / tmem_full barriers of BOTH CTAs.
//   warps 2-5 (both)    epilogue of the CTA's own 128 accumulator rows (TMEM lanes), then a remote arrive on the
//                       leader's tmem_empty barrier (8 arrivals = 4 warps x 2 CTAs).
// Every wait is watchdog-bounded (common.cuh), a protocol bug traps instead of hanging the GPU.
#pragma once

#include "gemm_tc.cuh"

namespace qb {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of (possibly) the peer CTA.  Default semantics (release at CTA scope) are what is needed here:
// what the arrival publishes is either TMEM reads (ordered by tcgen05.fence::before_thread_sync) or this CTA's own
// shared-memory tile, made visible to the async proxy by fence.proxy.async and read by this SM's tensor core.  A
// cluster-scope release costs a device-level fence per arrival (measured ~1400 cycles in the staging loop).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait whose acquire covers writes released by threads of the peer CTA (staging warps of both CTAs arrive here)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  uint32_t probes = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++probes == 4096u) {
      uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > QB_WATCHDOG_NS) __trap();
      probes = 0;
    }
  }
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are counted on an mbarrier that may live in the peer CTA (cluster address)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
template <MmaKind K>
__device__ __forceinline__ void tc_mma_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  if constexpr (K == MmaKind::I8) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else if constexpr (K == MmaKind::F8F6F4) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// BSRC::TMA : both operands 8-bit, loaded by TMA (int8 x int8, fp8 x fp8)
// BSRC::INT4: bf16/fp16 activations by TMA; each CTA's staging warps dequantise the packed-int4 weights of ITS half of
//             the tile's out-features into its own shared memory (so the dequantisation work per CTA halves as well)
template <MmaKind KIND_, int BN_, BSrc BSRC_ = BSrc::TMA, typename WT_ = __nv_bfloat16, bool ZP_ = false>
struct PairCfg {
  static constexpr MmaKind KIND = KIND_;
  static constexpr BSrc BSRC = BSRC_;
  using WT = WT_;
  static constexpr bool ZP = ZP_;
  static constexpr int BN = BN_;          // UMMA N of the pair tile; each CTA stages BN/2 weight rows
  static constexpr int BM_PAIR = 256;     // UMMA M of the pair; 128 rows per CTA
  static constexpr int KBYTES = 128;
  static constexpr int A_TILE = 128 * KBYTES;
  static constexpr int B_HALF = (BN / 2) * KBYTES;
  static constexpr int STAGE = A_TILE + ((B_HALF + 1023) / 1024) * 1024;
  static constexpr int NSTAGES = (192 * 1024) / STAGE;
  static constexpr int ACC_COLS = 256;
  static constexpr int NACC = 2;
  static constexpr int TMEM_COLS = 512;
  // INT4: per pipeline slot one staging group of BN/4 threads (one per packed row of this CTA's quarter tile)
  static constexpr int CVT_GROUP_WARPS = (BSRC == BSrc::INT4) ? (BN / 4) / 32 : 0;
  static constexpr int NTHREADS = (6 + NSTAGES * CVT_GROUP_WARPS) * 32;
  static constexpr int FULL_ARRIVALS = 1 + 2 * CVT_GROUP_WARPS;  // leader's expect_tx + staging warps of both CTAs
  static constexpr int SMEM_BYTES = NSTAGES * STAGE + (2 * NSTAGES + 4) * 8 + 16 + 1024 + static_cast<int>(sizeof(EpiCols));
  static_assert(BN % 16 == 0 && BN <= 256 && (BN / 2) % 8 == 0, "UMMA M=256 needs N % 16 == 0");
  static_assert(BSRC != BSrc::INT4 || (BN % 128 == 0), "int4 pair tile: whole staging warps, 16-column chunks per nibble plane");
};

template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::NTHREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p, const uint32_t idesc) {
  constexpr int NSTAGES = Cfg::NSTAGES;
  constexpr int BN = Cfg::BN;

  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs (the MMA addresses the peer's operands through the leader's descriptors)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGES * Cfg::STAGE);
  uint64_t* empty_bar = full_bar + NSTAGES;
  uint64_t* tmem_full_bar = empty_bar + NSTAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  EpiCols* epi_cols = reinterpret_cast<EpiCols*>(tmem_ptr_smem + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    if constexpr (Cfg::BSRC == BSrc::TMA) tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(&full_bar[s], Cfg::FULL_ARRIVALS);  // leader's copy is the one in use
      mbar_init(&empty_bar[s], 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);  // 4 epilogue warps x 2 CTAs (leader's copy only)
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  constexpr int KELEMS = Cfg::KBYTES / ((Cfg::KIND == MmaKind::F16) ? 2 : 1);
  const int kblocks = (p.K + KELEMS - 1) / KELEMS;

  auto a_smem = [&](int s) { return smem + s * Cfg::STAGE; };
  auto b_smem = [&](int s) { return smem + s * Cfg::STAGE + Cfg::A_TILE; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      constexpr uint32_t tx_bytes = 2u * (Cfg::A_TILE + (Cfg::BSRC == BSrc::TMA ? Cfg::B_HALF : 0));
      const uint32_t leader_full = mapa_u32(smem_u32(full_bar), 0);
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m_blk = tile % p.num_m_blocks;
        const int n_blk = tile / p.num_m_blocks;
        const int a_row = m_blk * Cfg::BM_PAIR + static_cast<int>(rank) * 128;
        const int b_row = n_blk * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          const bool hot = QB_KO(p.dbg, 128);
          tma_load_2d_2cta(smem_u32(a_smem(stage)), &tmap_a, leader_full + stage * 8, hot ? 0 : kb * KELEMS,
                           hot ? static_cast<int>(rank) * 128 : a_row);
          if constexpr (Cfg::BSRC == BSrc::TMA)
            tma_load_2d_2cta(smem_u32(b_smem(stage)), &tmap_b, leader_full + stage * 8, hot ? 0 : kb * KELEMS,
                             hot ? static_cast<int>(rank) * (BN / 2) : b_row);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (rank == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      int tn = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs, ++acc_it) {
        const uint32_t acc = acc_it % Cfg::NACC;
        const uint32_t acc_phase = (acc_it / Cfg::NACC) & 1u;
        gemm_trace_evt(p, 2, tn);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        gemm_trace_evt(p, 2, tn);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          if constexpr (Cfg::BSRC == BSrc::TMA) {
            mbar_wait(&full_bar[stage], phase);
          } else {
            if (QB_KO(p.dbg, 512)) gemm_trace_evt(p, 2, tn);
            mbar_wait(&full_bar[stage], phase);
            if (QB_KO(p.dbg, 512)) gemm_trace_evt(p, 2, tn);
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_smem(stage));
          const uint32_t b_addr = smem_u32(b_smem(stage));
#pragma unroll
          for (int k = 0; k < Cfg::KBYTES / 32; ++k) {
            tc_mma_2cta<Cfg::KIND>(tmem_base + acc * Cfg::ACC_COLS, umma_desc_sw128_kmajor(a_addr + k * 32),
                                   umma_desc_sw128_kmajor(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_2cta(&empty_bar[stage], 3);  // both CTAs' slots are reusable once these MMAs have read them
          if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2cta(&tmem_full_bar[acc], 3);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue (4 warps = this CTA's 128 rows)
    const int quarter = warp & 3;
    uint32_t acc_it = 0;
    const uint32_t leader_tmem_empty = mapa_u32(smem_u32(tmem_empty_bar), 0);
    constexpr bool IS_INT = (Cfg::KIND == MmaKind::I8);
    int tn = 0;
    const bool tracer = (warp == 2 && lane == 0);
    for (int tile = pair; tile < num_tiles; tile += npairs, ++acc_it) {
      const int m_blk = tile % p.num_m_blocks;
      const int n_blk = tile / p.num_m_blocks;
      const uint32_t acc = acc_it % Cfg::NACC;
      const uint32_t acc_phase = (acc_it / Cfg::NACC) & 1u;
      constexpr int NCH = BN / 16;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * Cfg::ACC_COLS;
      const int row = QB_KO(p.dbg, 256) ? p.M : m_blk * Cfg::BM_PAIR + static_cast<int>(rank) * 128 + quarter * 32 + lane;
      const bool plain = (p.scales == nullptr) && (p.bias == nullptr) && !IS_INT;
      const int buf = static_cast<int>(acc_it & 1u);
      // tile column -> output feature.  int4: CTA r staged the packed rows  n_blk*BN/2 + r*BN/4 + [0, BN/4); its
      // columns [r*BN/2, +BN/4) are their low-nibble features (n < N/2), the next BN/4 the high-nibble ones (+N/2)
      auto col_first = [&](int c, int& n_limit) {
        if constexpr (Cfg::BSRC == BSrc::INT4) {
          const int r = c / (BN / 2), cc = c % (BN / 2);
          const int prow = n_blk * (BN / 2) + r * (BN / 4) + (cc % (BN / 4));
          if (cc < BN / 4) { n_limit = p.N / 2; return prow; }
          n_limit = p.N;
          return p.N / 2 + prow;
        } else {
          n_limit = p.N;
          return n_blk * BN + c;
        }
      };
      if (!plain) {
        // per-column scale / bias of this tile -> shared memory (fp32), before the accumulator is even ready
        epi_stage_cols(epi_cols, buf, p, threadIdx.x - 64, BN, [&](int c) {
          int lim;
          const int n = col_first(c, lim);
          return n < lim ? n : -1;
        });
      }
      if (tracer) gemm_trace_evt(p, 3, tn);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      if (tracer) gemm_trace_evt(p, 3, tn);
      tc_fence_after();
      uint32_t va[16], vb[16];
      tmem_ld_32x32b_x16(t_lane, va);
      tmem_ld_wait();
      auto do_chunk = [&](int chunk, const uint32_t (&v)[16]) {
        if (QB_KO(p.dbg, 64)) return;
        int n_limit;
        const int n_first = col_first(chunk * 16, n_limit);
        epilogue_chunk<IS_INT, (Cfg::KIND != MmaKind::F16)>(p, v, row, n_first, n_limit, plain, epi_cols, buf, chunk * 16);
      };
#pragma unroll 1
      for (int ch = 0; ch < NCH; ch += 2) {
        tmem_ld_32x32b_x16(t_lane + (ch + 1) * 16, vb);  // NCH is even
        do_chunk(ch, va);
        tmem_ld_wait();
        if (ch + 2 < NCH) tmem_ld_32x32b_x16(t_lane + (ch + 2) * 16, va);
        do_chunk(ch + 1, vb);
        tmem_ld_wait();
        if (tracer && ch < 4) gemm_trace_evt(p, 3, tn);
      }
      if (tracer) gemm_trace_evt(p, 3, tn);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty + acc * 8);
    }
  } else {
    // ------------------------------------------------------------------ weight staging (int4 -> WT half tile)
    if constexpr (Cfg::BSRC == BSrc::INT4) {
      using WT = typename Cfg::WT;
      using D = Dq<WT>;
      constexpr bool ZP = Cfg::ZP;
      constexpr int ROWP = BN / 4;  // packed rows this CTA stages per tile (both nibbles -> BN/2 operand rows)
      constexpr int GT = Cfg::CVT_GROUP_WARPS * 32;
      static_assert(GT == ROWP && ROWP % 8 == 0, "one staging thread per packed row");
      const int ct = threadIdx.x - 6 * 32;
      const int grp = ct / GT;  // staging group == pipeline slot it owns
      const int r = ct % GT;    // packed row inside this CTA's quarter tile
      const int half_n = p.N / 2;
      const int groups_per_row = p.K / p.group;
      const bool two_sets = p.group < 64;  // group size 32: two (scale, shift) pairs per 64-k stage
      const WT* scale = static_cast<const WT*>(p.wscale);
      const int my_tiles = (num_tiles > pair) ? (num_tiles - 1 - pair) / npairs + 1 : 0;
      const int total_it = my_tiles * kblocks;

      struct Pre {
        uint4 raw[4];
        WT s_lo[2], s_hi[2];
        uint16_t z_lo[2], z_hi[2];
        bool ok;
      };
      int f_it = grp;
      int f_kb = grp % kblocks;
      int f_tile = pair + (grp / kblocks) * npairs;
      auto load_pre = [&](Pre& pr) {
        if (f_it >= total_it) return;
        const int rp = (f_tile / p.num_m_blocks) * (BN / 2) + static_cast<int>(rank) * ROWP + r;
        const int kbase = f_kb * 64;
        pr.ok = rp < half_n && kbase < p.K;
        if (pr.ok) {
          const uint8_t* src = p.wq + static_cast<size_t>(rp) * p.K + kbase;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            pr.raw[v] = (kbase + v * 16 < p.K) ? __ldg(reinterpret_cast<const uint4*>(src + v * 16)) : make_uint4(0, 0, 0, 0);
          const int g0 = (p.group_log2 >= 0) ? (kbase >> p.group_log2) : (kbase / p.group);
#pragma unroll
          for (int st = 0; st < 2; ++st) {
            if (st == 1 && !two_sets) break;
            const size_t ilo = static_cast<size_t>(rp) * groups_per_row + g0 + st;
            const size_t ihi = ilo + static_cast<size_t>(half_n) * groups_per_row;
            pr.s_lo[st] = __ldg(scale + ilo);
            pr.s_hi[st] = __ldg(scale + ihi);
            if (ZP) {
              pr.z_lo[st] = __ldg(static_cast<const uint8_t*>(p.wshift) + ilo);
              pr.z_hi[st] = __ldg(static_cast<const uint8_t*>(p.wshift) + ihi);
            } else {
              pr.z_lo[st] = __ldg(static_cast<const uint16_t*>(p.wshift) + ilo);
              pr.z_hi[st] = __ldg(static_cast<const uint16_t*>(p.wshift) + ihi);
            }
          }
        }
        f_it += NSTAGES;
        f_kb += NSTAGES;
        while (f_kb >= kblocks) { f_kb -= kblocks; f_tile += npairs; }
      };

      const uint32_t sw = static_cast<uint32_t>(r) & 7;
      const uint32_t off_lo = (static_cast<uint32_t>(r) >> 3) * 1024 + (static_cast<uint32_t>(r) & 7) * 128;
      const uint32_t off_hi = off_lo + (ROWP / 8) * 1024;
      const uint32_t bt = smem_u32(smem) + grp * Cfg::STAGE + Cfg::A_TILE;  // this group's half B tile (own CTA)
      const uint32_t empty_addr = smem_u32(empty_bar) + grp * 8;
      const uint32_t leader_full = mapa_u32(smem_u32(full_bar), 0) + grp * 8;
      uint32_t phase = 0;
      int tn = 0;
      const bool tracer = (ct == 0) && QB_KO(p.dbg, 512);
      auto process = [&](const Pre& cur) {
        if (tracer) gemm_trace_evt(p, 4, tn);
        mbar_wait_u32(empty_addr, phase ^ 1u);  // multicast commit: the pair's MMAs have read this slot in both CTAs
        if (tracer) gemm_trace_evt(p, 4, tn);
        if (cur.ok) {
          typename D::Coef klo[2], khi[2];
          klo[0] = D::make_raw(cur.s_lo[0], cur.z_lo[0], ZP);
          khi[0] = D::make_raw(cur.s_hi[0], cur.z_hi[0], ZP);
          if (two_sets) {
            klo[1] = D::make_raw(cur.s_lo[1], cur.z_lo[1], ZP);
            khi[1] = D::make_raw(cur.s_hi[1], cur.z_hi[1], ZP);
          } else {
            klo[1] = klo[0];
            khi[1] = khi[0];
          }
          stage_rowpair_64k<WT, ZP>(cur.raw, klo, khi, bt + off_lo, bt + off_hi, sw);
        } else {
          zero_rowpair_64k(bt + off_lo, bt + off_hi);
        }
        if (tracer) gemm_trace_evt(p, 4, tn);
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_full);  // release at cluster scope, on the leader's barrier
        if (tracer) gemm_trace_evt(p, 4, tn);
        phase ^= 1u;
      };
      Pre pa, pb;
      load_pre(pa);
      for (int it = grp; it < total_it; it += 2 * NSTAGES) {
        load_pre(pb);
        process(pa);
        if (it + NSTAGES < total_it) {
          load_pre(pa);
          process(pb);
        }
      }
    }
  }

  // Neither CTA may exit (or free TMEM) while its peer can still signal its barriers or read its operands.
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace qb
