// Small-M ("decode") fused packed-int4 linear: HBM-bound, so the design goal is to keep every SM streaming packed
// weight bytes at full rate.
//
//   out[m, n] = sum_k x[m, k] * dequant(W)[n, k]          M <= 128
//
// Swap-AB: the dequantised weight tile is the M-side operand of tcgen05.mma (128 out-features per instruction), the
// activations are the N-side operand (MP = M padded to 16/32/64/128 columns), so the tensor core never idles on
// padding rows and the accumulator D[128 out-features, MP tokens] lives in TMEM.
//
// Work decomposition (stream-K): the weight is cut into "stages" of 64 packed rows x 128 k (8 KB of packed bytes =
// 128 out-features x 128 k, since one byte carries out-feature n and n + N/2).  The P*K/128 stages are dealt to the
// CTAs in equal contiguous spans, so all 148 SMs stream the same number of bytes whatever N is.  A span is cut into
// segments at out-feature-block boundaries; a block finished by several CTAs is reduced deterministically: every
// segment writes its fp32 partial to the workspace, an atomic ticket per block elects the last arriver, which sums
// the partials in segment order, rounds once, adds the bias and stores.
//
// Per-CTA pipeline:
//   warp 0        TMA: packed bytes  HBM -> 8-deep raw ring (64 KB in flight per SM, SWIZZLE_128B)
//   warps 8-23    staging: raw ring -> registers -> exact dequant (reference rounding order) -> bf16/fp16 operand tile
//   warp 6        TMA: activation tile for the same stage
//   warp 1        MMA issue (one thread), accumulators double-buffered in TMEM
//   warps 2-5     epilogue / split-K fix-up
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace qb {

struct DecodeParams {
  const void* scale;   // [N * K / group] weight dtype
  const void* shift;   // same, or uint8 zero-points
  const void* bias;    // [N] or nullptr
  void* out;           // [M, N]
  float* partials;     // workspace: [P][max_segs][M][128] fp32
  int* tickets;        // workspace: [P] zero-initialised once; the kernel leaves them zero
  int M, N, K;
  int group;
  int group_log2;      // log2(group) or -1
  int shift_is_int;
  int P;               // out-feature blocks = ceil((N/2) / 64)
  int SPB;             // stages per block = K / 128
  int span;            // stages per CTA
  int max_segs;
};

template <typename WT_, int MP_, bool ZP_ = false>
struct DecodeCfg {
  using WT = WT_;
  static constexpr bool ZP = ZP_;                 // shift is an integer zero-point
  static constexpr int MP = MP_;                  // padded token count = UMMA N
  static constexpr int RAW_STAGES = 8;
  static constexpr int RAW_BYTES = 64 * 128;      // 64 packed rows x 128 k
  static constexpr int A_PANEL = 128 * 128;       // 128 out-features x 64 k (bf16) = one SW128 panel
  static constexpr int A_BYTES = 2 * A_PANEL;     // 128 k
  static constexpr int X_PANEL = MP * 128;
  static constexpr int X_BYTES = 2 * X_PANEL;
  static constexpr int STAGE = A_BYTES + X_BYTES;
  static constexpr int NSTAGES = (MP <= 32) ? 3 : 2;
  static constexpr int TMEM_COLS = (2 * MP < 32) ? 32 : 2 * MP;
  static constexpr int NCVT_WARPS = 16;
  static constexpr int NCVT_THREADS = NCVT_WARPS * 32;
  static constexpr int FIRST_CVT_WARP = 8;
  static constexpr int NTHREADS = (FIRST_CVT_WARP + NCVT_WARPS) * 32;
  static constexpr int SMEM_BYTES = RAW_STAGES * RAW_BYTES + NSTAGES * STAGE + 1024 + 512;
  static_assert(MP % 16 == 0 && MP >= 16 && MP <= 128, "MP");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two");
};

// number of segments block `pb` is cut into when spans have `span` stages
__device__ __forceinline__ int decode_nsegs(int pb, int spb, int span) {
  return ((pb + 1) * spb - 1) / span - (pb * spb) / span + 1;
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, 1)
    gemm_w4_decode_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                          const DecodeParams p, const uint32_t idesc) {
  using WT = typename Cfg::WT;
  constexpr int MP = Cfg::MP;
  constexpr int RS = Cfg::RAW_STAGES;
  constexpr int NS = Cfg::NSTAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* raw_ring = smem;
  uint8_t* stage_ring = smem + RS * Cfg::RAW_BYTES;
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(stage_ring + NS * Cfg::STAGE);
  uint64_t* raw_empty = raw_full + RS;
  uint64_t* a_full = raw_empty + RS;
  uint64_t* a_empty = a_full + NS;
  uint64_t* tmem_full = a_empty + NS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int* ticket_smem = reinterpret_cast<int*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 6 && lane == 0) tma_prefetch_desc(&tmap_x);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < RS; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], Cfg::NCVT_WARPS);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&a_full[s], 1 + Cfg::NCVT_WARPS);
      mbar_init(&a_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total = p.P * p.SPB;
  const int s_begin = min(static_cast<int>(blockIdx.x) * p.span, total);
  const int s_end = min(s_begin + p.span, total);
  const int L = s_end - s_begin;
  const int half_n = p.N / 2;

  auto a_panel = [&](int slot, int panel) { return stage_ring + slot * Cfg::STAGE + panel * Cfg::A_PANEL; };
  auto x_panel = [&](int slot, int panel) { return stage_ring + slot * Cfg::STAGE + Cfg::A_BYTES + panel * Cfg::X_PANEL; };

  if (warp == 0) {
    // ---------------------------------------------------------------- packed-weight TMA producer
    if (lane == 0) {
      for (int i = 0; i < L; ++i) {
        const int s = s_begin + i;
        const int pb = s / p.SPB, ks = s - pb * p.SPB;
        const int slot = i % RS;
        mbar_wait(&raw_empty[slot], ((i / RS) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&raw_full[slot], Cfg::RAW_BYTES);
        tma_load_2d(raw_ring + slot * Cfg::RAW_BYTES, &tmap_w, &raw_full[slot], ks * 128, pb * 64);
      }
    }
  } else if (warp == 6) {
    // ---------------------------------------------------------------- activation TMA producer
    if (lane == 0) {
      for (int i = 0; i < L; ++i) {
        const int s = s_begin + i;
        const int ks = s % p.SPB;
        const int slot = i % NS;
        mbar_wait(&a_empty[slot], ((i / NS) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&a_full[slot], Cfg::X_BYTES);
        tma_load_2d(x_panel(slot, 0), &tmap_x, &a_full[slot], ks * 128, 0);
        tma_load_2d(x_panel(slot, 1), &tmap_x, &a_full[slot], ks * 128 + 64, 0);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      uint32_t seg = 0;
      bool seg_open = false;
      for (int i = 0; i < L; ++i) {
        const int s = s_begin + i;
        const int ks = s % p.SPB;
        const uint32_t acc = seg & 1u;
        if (!seg_open) {
          mbar_wait(&tmem_empty[acc], ((seg >> 1) & 1u) ^ 1u);
          tc_fence_after();
        }
        const int slot = i % NS;
        mbar_wait(&a_full[slot], (i / NS) & 1u);
        tc_fence_after();
#pragma unroll
        for (int panel = 0; panel < 2; ++panel) {
          const uint32_t a_addr = smem_u32(a_panel(slot, panel));
          const uint32_t x_addr = smem_u32(x_panel(slot, panel));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma<MmaKind::F16>(tmem_base + acc * MP, umma_desc_sw128_kmajor(a_addr + k * 32),
                                 umma_desc_sw128_kmajor(x_addr + k * 32), idesc,
                                 (seg_open || panel != 0 || k != 0) ? 1u : 0u);
          }
        }
        seg_open = true;
        tc_commit(&a_empty[slot]);
        const bool seg_end = (ks == p.SPB - 1) || (i == L - 1);
        if (seg_end) {
          tc_commit(&tmem_full[acc]);
          seg_open = false;
          ++seg;
        }
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ---------------------------------------------------------------- epilogue + split-K fix-up
    const int quarter = warp & 3;
    const int et = quarter * 32 + lane;  // 0..127 == TMEM lane == tile row
    const int etid = (warp - 2) * 32 + lane;
    uint32_t seg = 0;
    int i = 0;
    while (i < L) {
      const int s = s_begin + i;
      const int pb = s / p.SPB, ks = s - pb * p.SPB;
      const int seg_len = min(p.SPB - ks, L - i);
      const uint32_t acc = seg & 1u;
      const int nsegs = decode_nsegs(pb, p.SPB, p.span);
      const int seg_idx = static_cast<int>(blockIdx.x) - (pb * p.SPB) / p.span;
      // tile row -> out-feature
      const int rp = pb * 64 + (et & 63);
      const bool n_ok = rp < half_n;
      const int n = (et < 64) ? rp : half_n + rp;

      mbar_wait(&tmem_full[acc], (seg >> 1) & 1u);
      tc_fence_after();
      float* part = p.partials + (static_cast<size_t>(pb) * p.max_segs + seg_idx) * p.M * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < MP; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * MP + c0, v);
        tmem_ld_wait();
        if (nsegs == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = c0 + j;
            if (m < p.M && n_ok) {
              WT r = from_float<WT>(__uint_as_float(v[j]));
              if (p.bias != nullptr)
                r = from_float<WT>(__fadd_rn(to_float<WT>(r), to_float<WT>(static_cast<const WT*>(p.bias)[n])));
              static_cast<WT*>(p.out)[static_cast<size_t>(m) * p.N + n] = r;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = c0 + j;
            if (m < p.M) part[static_cast<size_t>(m) * 128 + et] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);  // TMEM buffer free: the MMA warp may start the next segment

      if (nsegs > 1) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) *ticket_smem = atomicAdd(&p.tickets[pb], 1);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool last = (*ticket_smem == nsegs - 1);
        if (last) {
          __threadfence();
          const float* base = p.partials + static_cast<size_t>(pb) * p.max_segs * p.M * 128;
          for (int m = 0; m < p.M; ++m) {
            float sum = 0.f;
            for (int sg = 0; sg < nsegs; ++sg)
              sum += __ldcg(base + (static_cast<size_t>(sg) * p.M + m) * 128 + et);
            if (n_ok) {
              WT r = from_float<WT>(sum);
              if (p.bias != nullptr)
                r = from_float<WT>(__fadd_rn(to_float<WT>(r), to_float<WT>(static_cast<const WT*>(p.bias)[n])));
              static_cast<WT*>(p.out)[static_cast<size_t>(m) * p.N + n] = r;
            }
          }
          if (etid == 0) p.tickets[pb] = 0;  // every other segment has already arrived: safe to recycle
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // ticket_smem reuse
      }
      i += seg_len;
      ++seg;
    }
  } else if (warp >= Cfg::FIRST_CVT_WARP) {
    // ---------------------------------------------------------------- staging: raw bytes -> exact dequant -> operand tile
    using D = Dq<WT>;
    constexpr bool ZP = Cfg::ZP;
    const int ct = threadIdx.x - Cfg::FIRST_CVT_WARP * 32;
    const int r = ct & 63;   // packed row inside the block: low out-feature r, high out-feature 64 + r
    const int q8 = ct >> 6;  // which 16-k slice of the 128-k stage (0..7)
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const WT* scale = static_cast<const WT*>(p.scale);
    const int groups_per_row = p.K / p.group;
    // source: raw chunk q8 of row r (SWIZZLE_128B); destination: panel q8/4, chunks (q8%4)*2, +1 of rows r and 64+r
    const uint32_t raw_off = static_cast<uint32_t>(r) * 128 + ((static_cast<uint32_t>(q8) ^ sw) << 4);
    const uint32_t c = static_cast<uint32_t>((q8 & 3) * 2);
    const uint32_t off_lo = static_cast<uint32_t>(q8 >> 2) * Cfg::A_PANEL + (static_cast<uint32_t>(r) >> 3) * 1024 +
                            (static_cast<uint32_t>(r) & 7) * 128;
    const uint32_t off_hi = off_lo + 8 * 1024;  // row 64 + r: same swizzle phase
    const uint32_t d0 = ((c + 0) ^ sw) << 4, d1 = ((c + 1) ^ sw) << 4;
    const uint32_t raw0 = smem_u32(raw_ring), stage0 = smem_u32(stage_ring);
    const uint32_t raw_full0 = smem_u32(raw_full), raw_empty0 = smem_u32(raw_empty);
    const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);

    // software prefetch ring for the per-group scale / shift (PF-1 stages ahead); indices advance incrementally
    constexpr int PF = 4;
    WT s_lo[PF], s_hi[PF];
    uint16_t z_lo[PF], z_hi[PF];  // raw 16-bit payload: WT bits, or the zero-point byte
    bool okv[PF];
    int f_left = L;
    int f_pb = s_begin / p.SPB;
    int f_ks = s_begin - f_pb * p.SPB;
    auto fetch = [&](int slot) {
      if (f_left <= 0) return;
      --f_left;
      const int rp = f_pb * 64 + r;
      okv[slot] = rp < half_n;
      if (okv[slot]) {
        const int kk = f_ks * 128 + q8 * 16;
        const int g = (p.group_log2 >= 0) ? (kk >> p.group_log2) : (kk / p.group);
        const size_t ilo = static_cast<size_t>(rp) * groups_per_row + g;
        const size_t ihi = ilo + static_cast<size_t>(half_n) * groups_per_row;
        s_lo[slot] = __ldg(scale + ilo);
        s_hi[slot] = __ldg(scale + ihi);
        if (ZP) {
          z_lo[slot] = __ldg(static_cast<const uint8_t*>(p.shift) + ilo);
          z_hi[slot] = __ldg(static_cast<const uint8_t*>(p.shift) + ihi);
        } else {
          z_lo[slot] = __ldg(static_cast<const uint16_t*>(p.shift) + ilo);
          z_hi[slot] = __ldg(static_cast<const uint16_t*>(p.shift) + ihi);
        }
      }
      if (++f_ks == p.SPB) { f_ks = 0; ++f_pb; }
    };
#pragma unroll
    for (int u = 0; u < PF - 1; ++u) fetch(u);

    int rslot = 0, aslot = 0;
    uint32_t rphase = 0, aphase = 0;
    for (int i0 = 0; i0 < L; i0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        if (i0 + u < L) {
          fetch((u + PF - 1) % PF);
          mbar_wait_u32(raw_full0 + rslot * 8, rphase);
          const uint4 raw = ld_shared_v4(raw0 + rslot * Cfg::RAW_BYTES + raw_off);
          uint32_t lo[8], hi[8];
          if (okv[u]) {
            const typename D::Coef klo = D::make_raw(s_lo[u], z_lo[u], ZP);
            const typename D::Coef khi = D::make_raw(s_hi[u], z_hi[u], ZP);
            dequant16<WT, ZP>(raw, klo, khi, lo, hi);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { lo[j] = 0u; hi[j] = 0u; }
          }
          // the dequant above consumed `raw` (data dependency => the smem read has completed): release the raw
          // slot with ONE arrive per warp -- per-thread arrives serialise in the shared-memory atomic unit
          __syncwarp();
          if (lane == 0) mbar_arrive_u32(raw_empty0 + rslot * 8);
          if (++rslot == RS) { rslot = 0; rphase ^= 1u; }

          mbar_wait_u32(a_empty0 + aslot * 8, aphase ^ 1u);
          const uint32_t pb_lo = stage0 + aslot * Cfg::STAGE + off_lo;
          const uint32_t pb_hi = stage0 + aslot * Cfg::STAGE + off_hi;
          st_shared_v4(pb_lo + d0, lo[0], lo[1], lo[2], lo[3]);
          st_shared_v4(pb_lo + d1, lo[4], lo[5], lo[6], lo[7]);
          st_shared_v4(pb_hi + d0, hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(pb_hi + d1, hi[4], hi[5], hi[6], hi[7]);
          fence_proxy_async_smem();  // each writer makes its own generic-proxy stores visible to the async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive_u32(a_full0 + aslot * 8);
          if (++aslot == NS) { aslot = 0; aphase ^= 1u; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace qb
