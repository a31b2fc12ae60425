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
// The dequantised weight tile never touches shared memory: the staging warps write it straight into TENSOR MEMORY
// (tcgen05.st) and the MMA reads its A operand from there (tcgen05.mma [d], [a_tmem], b_desc).  Measured on B200
// (tools/mma_probe.cu, tools/tmema_probe.cu): an M=128 MMA with a small N costs 88 cycles with A in shared memory
// but 46 with A in TMEM -- with A in shared memory the MMA, not HBM, would bound this kernel at ~45 % of the roofline.
//
// Per-CTA pipeline (NG = 4 operand slots of 64 TMEM columns = 128 k):
//   warp 0        TMA: packed bytes HBM -> 8-deep raw ring, one 8 KB box (64 rows x 128 k, SWIZZLE_128B) per stage
//   warps 8-23    staging: NG groups of 4 warps = 128 threads = the 128 TMEM lanes.  Lane quarter q = warp % 4:
//                 q 0,1 extract the low nibbles (out-features pb*64 + 0..63), q 2,3 the high nibbles (+N/2); both
//                 read the same raw bytes.  Group g converts the stages i == g (mod NG) into TMEM slot g: raw
//                 bytes -> registers -> exact dequant (reference rounding order) -> tcgen05.st.  Each thread
//                 amortises its per-stage overhead over one 128-k row.
//   warp 6        TMA: activation panels (MP tokens x 128 k) for the same stage into shared memory (B operand)
//   warp 1        MMA issue (one thread), accumulators double-buffered in TMEM columns [0, 2*MP)
//   warps 2-5     epilogue / split-K fix-up
#pragma once

#include "common.cuh"
#include "gather.cuh"
#include "gemm_tc.cuh"

namespace qb {

struct DecodeParams {
  const void* scale;   // [N * K / group] weight dtype
  const void* shift;   // same, or uint8 zero-points
  const void* bias;    // [N] or nullptr
  void* out;           // [M, ld] (ld = N, col0 = 0 for an ordinary call)
  GatherInfo g;        // fused all-gather of a column-parallel linear (gather.cuh); g.n_out == 1: ordinary call
  int ld, col0;
  float* partials;     // workspace: [P][max_segs][M][128] fp32
  int* tickets;        // workspace: [P] zero-initialised once; the kernel leaves them zero
  int M, N, K;
  int group;
  int group_log2;      // log2(group) or -1
  int shift_is_int;
  int P;               // out-feature blocks = ceil((N/2) / 64)
  int SPB;             // 128-k stages per block = K / 128
  int span;            // stages per CTA
  int max_segs;
  int dbg;             // developer flags (tools/trace_decode.py): 1 = skip the dequant math (timing experiments only)
  long long* trace;    // developer timeline (tools/trace_decode.py) or nullptr: [cta][role][event] clock64 stamps
};

// developer timeline: CTA `blockIdx.x` < 4 records up to 64 clock64 stamps per role (0 raw TMA, 1 x TMA, 2 MMA,
// 3 epilogue, 4 staging group 0 lane 0).  Costs one uniform branch per call when disabled.
__device__ __forceinline__ void trace_evt(const DecodeParams& p, int role, int& n) {
  if (p.trace != nullptr && blockIdx.x < 4 && n < 64) {
    p.trace[(static_cast<size_t>(blockIdx.x) * 5 + role) * 64 + n] = clock64();
    ++n;
  }
}

template <typename WT_, int MP_, bool ZP_ = false>
struct DecodeCfg {
  using WT = WT_;
  static constexpr bool ZP = ZP_;                 // shift is an integer zero-point
  static constexpr int MP = MP_;                  // padded token count = UMMA N
  static constexpr int NG = 4;                    // staging groups (4 warps each)
  static constexpr int D_COLS = (2 * MP < 64) ? 64 : 2 * MP;   // accumulators: TMEM columns [0, 2*MP)
  static constexpr int A_COLS = 64;               // TMEM columns of one operand slot (128 k, 2 elements / column)
  static constexpr int NSLOT = (512 - D_COLS) / A_COLS;  // operand slots: 7 (MP<=32), 6 (MP=64), 4 (MP=128)
  static constexpr int A_COL0 = D_COLS;           // operand slots live in TMEM columns [D_COLS, 512)
  static constexpr int RAW_STAGES = 8;
  static constexpr int RAW_BYTES = 64 * 128;      // 64 packed rows x 128 k
  static constexpr int X_PANEL = MP * 128;        // MP tokens x 64 k
  static constexpr int X_SLOT = 2 * X_PANEL;      // 128 k
  static constexpr int TMEM_COLS = 512;
  static constexpr int NCVT_WARPS = NG * 4;
  static constexpr int FIRST_CVT_WARP = 8;
  static constexpr int NTHREADS = (FIRST_CVT_WARP + NCVT_WARPS) * 32;
  static constexpr int SMEM_BYTES = RAW_STAGES * RAW_BYTES + NSLOT * X_SLOT + 1024 + 512;
  static_assert(MP % 16 == 0 && MP >= 16 && MP <= 128, "MP");
  static_assert(2 * MP <= A_COL0 && A_COL0 + NSLOT * A_COLS <= TMEM_COLS && NSLOT >= NG, "TMEM budget");
  static_assert(FIRST_CVT_WARP % 4 == 0, "staging warp w must own TMEM lane quarter w % 4");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
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
  constexpr int NG = Cfg::NG;
  constexpr int NSLOT = Cfg::NSLOT;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* raw_ring = smem;
  uint8_t* x_ring = smem + RS * Cfg::RAW_BYTES;
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(x_ring + NSLOT * Cfg::X_SLOT);
  uint64_t* raw_empty = raw_full + RS;
  uint64_t* a_full = raw_empty + RS;
  uint64_t* a_empty = a_full + NSLOT;
  uint64_t* tmem_full = a_empty + NSLOT;
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
      mbar_init(&raw_empty[s], 4);  // the four warps of the staging group that consumes the box
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&a_full[s], 1 + 4);  // activation TMA + the group's four warps
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
  pdl_launch_dependents();

  const int total = p.P * p.SPB;
  const int s_begin = min(static_cast<int>(blockIdx.x) * p.span, total);
  const int s_end = min(s_begin + p.span, total);
  const int L = s_end - s_begin;
  const int half_n = p.N / 2;

  if (warp == 0) {
    // ---------------------------------------------------------------- packed-weight TMA producer
    if (lane == 0) {
      int pb = s_begin / p.SPB, ks = s_begin - pb * p.SPB;
      int slot = 0;
      uint32_t phase = 0;
      int tn = 0;
      trace_evt(p, 0, tn);
      for (int i = 0; i < L; ++i) {
        mbar_wait(&raw_empty[slot], phase ^ 1u);
        mbar_arrive_expect_tx(&raw_full[slot], Cfg::RAW_BYTES);
        tma_load_2d(raw_ring + slot * Cfg::RAW_BYTES, &tmap_w, &raw_full[slot], ks * 128, pb * 64);
        trace_evt(p, 0, tn);
        if (++slot == RS) { slot = 0; phase ^= 1u; }
        if (++ks == p.SPB) { ks = 0; ++pb; }
      }
    }
  } else if (warp == 6) {
    // ---------------------------------------------------------------- activation TMA producer
    if (lane == 0) {
      int ks = s_begin % p.SPB;
      int slot = 0;
      uint32_t phase = 0;
      int tn = 0;
      trace_evt(p, 1, tn);
      pdl_wait();              // activations = the previous kernel's output (the weight stream does not wait)
      gather_wait_start(p.g);  // the activation may be the gathered output of the previous linear
      for (int i = 0; i < L; ++i) {
        mbar_wait(&a_empty[slot], phase ^ 1u);
        trace_evt(p, 1, tn);
        mbar_arrive_expect_tx(&a_full[slot], Cfg::X_SLOT);
        uint8_t* xs = x_ring + slot * Cfg::X_SLOT;
        tma_load_2d(xs, &tmap_x, &a_full[slot], ks * 128, 0);
        tma_load_2d(xs + Cfg::X_PANEL, &tmap_x, &a_full[slot], ks * 128 + 64, 0);
        if (++slot == NSLOT) { slot = 0; phase ^= 1u; }
        if (++ks == p.SPB) ks = 0;
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (A operand from TMEM)
    if (lane == 0) {
      uint32_t seg = 0;
      bool seg_open = false;
      int ks = s_begin % p.SPB;
      int slot = 0;
      uint32_t phase = 0;
      int tn = 0;
      trace_evt(p, 2, tn);
      for (int i = 0; i < L; ++i) {
        const uint32_t acc = seg & 1u;
        if (!seg_open) {
          mbar_wait(&tmem_empty[acc], ((seg >> 1) & 1u) ^ 1u);
          tc_fence_after();
        }
        mbar_wait(&a_full[slot], phase);
        tc_fence_after();
        trace_evt(p, 2, tn);
        const uint32_t a_tmem = tmem_base + Cfg::A_COL0 + slot * Cfg::A_COLS;
        const uint32_t x_addr = smem_u32(x_ring + slot * Cfg::X_SLOT);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          tc_mma_f16_ts(tmem_base + acc * MP, a_tmem + k * 8,
                        umma_desc_sw128_kmajor(x_addr + (k >> 2) * Cfg::X_PANEL + (k & 3) * 32), idesc,
                        (seg_open || k != 0) ? 1u : 0u);
        }
        seg_open = true;
        tc_commit(&a_empty[slot]);
        if (++slot == NSLOT) { slot = 0; phase ^= 1u; }
        const bool seg_end = (ks == p.SPB - 1) || (i == L - 1);
        if (++ks == p.SPB) ks = 0;
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
    int tn = 0;
    if (etid == 0) trace_evt(p, 3, tn);
    pdl_wait();  // workspace (tickets / partials) and output buffers are shared with the previous launch
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
      if (etid == 0) trace_evt(p, 3, tn);
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
              const size_t o_idx = static_cast<size_t>(m) * p.ld + p.col0 + n;
              static_cast<WT*>(p.out)[o_idx] = r;
              for (int q = 1; q < p.g.n_out; ++q) static_cast<WT*>(p.g.out_peer[q])[o_idx] = r;  // peers, over NVLink
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
      if (etid == 0) trace_evt(p, 3, tn);

      if (nsegs > 1) {
        // publish the partial (release), take a ticket, and let the last arriver of the block reduce (acquire)
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          int t;
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(t) : "l"(p.tickets + pb) : "memory");
          *ticket_smem = t;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool last = (*ticket_smem == nsegs - 1);
        if (last) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          const float* base = p.partials + static_cast<size_t>(pb) * p.max_segs * p.M * 128 + et;
          for (int m0 = 0; m0 < p.M; m0 += 4) {
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
            for (int sg0 = 0; sg0 < nsegs; sg0 += 4) {
              float v[4][4];
#pragma unroll
              for (int a = 0; a < 4; ++a)  // 16 independent L2 loads in flight, summed in segment order below
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  v[a][b] = (m0 + a < p.M && sg0 + b < nsegs)
                                ? __ldcg(base + (static_cast<size_t>(sg0 + b) * p.M + m0 + a) * 128)
                                : 0.f;
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  if (sg0 + b < nsegs) sum[a] += v[a][b];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const int m = m0 + a;
              if (m < p.M && n_ok) {
                WT r = from_float<WT>(sum[a]);
                if (p.bias != nullptr)
                  r = from_float<WT>(__fadd_rn(to_float<WT>(r), to_float<WT>(static_cast<const WT*>(p.bias)[n])));
                const size_t o_idx = static_cast<size_t>(m) * p.ld + p.col0 + n;
                static_cast<WT*>(p.out)[o_idx] = r;
                for (int q = 1; q < p.g.n_out; ++q) static_cast<WT*>(p.g.out_peer[q])[o_idx] = r;
              }
            }
          }
          if (etid == 0) p.tickets[pb] = 0;  // every other segment has already arrived: safe to recycle
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // ticket_smem reuse
      }
      if (etid == 0) trace_evt(p, 3, tn);
      i += seg_len;
      ++seg;
    }
  } else if (warp >= Cfg::FIRST_CVT_WARP) {
    // ---------------------------------------------------------------- staging: raw bytes -> exact dequant -> TMEM
    using D = Dq<WT>;
    constexpr bool ZP = Cfg::ZP;
    const int grp = (warp - Cfg::FIRST_CVT_WARP) >> 2;  // staging group: converts the stages i == grp (mod NG)
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may write
    const bool high_plane = quarter >= 2;                // lanes 64..127 hold the high-nibble out-features
    const int r = (quarter & 1) * 32 + lane;             // packed row inside the block
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const WT* scale = static_cast<const WT*>(p.scale);
    const int groups_per_row = p.K / p.group;
    const int sets = (p.group >= 128) ? 1 : (p.group >= 64 ? 2 : 4);  // (scale, shift) pairs inside one 128-k stage
    const uint32_t raw0 = smem_u32(raw_ring) + static_cast<uint32_t>(r) * 128;
    const uint32_t a_taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + Cfg::A_COL0;
    const uint32_t raw_full0 = smem_u32(raw_full), raw_empty0 = smem_u32(raw_empty);
    const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);

    // this group's stage sequence: i = grp, grp + NG, ...  ((pb, ks) advance incrementally)
    struct Pre {
      WT s[4];
      uint16_t z[4];
      bool ok;
    };
    int f_i = grp;
    int f_pb = (s_begin + grp) / p.SPB;
    int f_ks = (s_begin + grp) - f_pb * p.SPB;
    auto fetch = [&](Pre& pr) {
      if (f_i >= L) return;
      const int rp = f_pb * 64 + r;
      pr.ok = rp < half_n;
      if (pr.ok) {
        const int kk = f_ks * 128;
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
      f_i += NG;
      f_ks += NG;
      while (f_ks >= p.SPB) { f_ks -= p.SPB; ++f_pb; }
    };

    Pre cur, nxt;
    fetch(cur);
    int rslot = grp % RS;
    int aslot = grp % NSLOT;  // stage i lives in operand slot i % NSLOT (NSLOT >= NG: staging rarely waits for the MMA)
    uint32_t rphase = 0, aphase = static_cast<uint32_t>(grp / NSLOT) & 1u;
    int tn = 0;
    const bool tracer = (warp == Cfg::FIRST_CVT_WARP && lane == 0);
    if (tracer) trace_evt(p, 4, tn);
    for (int i = grp; i < L; i += NG) {
      fetch(nxt);  // next stage of this group is NG pipeline stages ahead
      // rows beyond N/2 (ragged last block) stage zeros: folded into the coefficients (scale 0, shift 0 -> exact 0), so
      // the conversion has no branch
      if (!cur.ok) {
#pragma unroll
        for (int st = 0; st < 4; ++st) { cur.s[st] = from_float<WT>(0.f); cur.z[st] = 0; }
      }
      typename D::Coef kc[4];
      kc[0] = D::make_raw(cur.s[0], cur.z[0], ZP);
      if (sets > 1) {
        // coefficients of the four 32-k quarters of the stage
        kc[1] = (sets >= 4) ? D::make_raw(cur.s[1], cur.z[1], ZP) : kc[0];
        kc[2] = D::make_raw(cur.s[sets >= 4 ? 2 : 1], cur.z[sets >= 4 ? 2 : 1], ZP);
        kc[3] = (sets >= 4) ? D::make_raw(cur.s[3], cur.z[3], ZP) : kc[2];
      } else {
        kc[1] = kc[0];
        kc[2] = kc[0];
        kc[3] = kc[0];
      }
      mbar_wait_u32(raw_full0 + rslot * 8, rphase);
      if (tracer) trace_evt(p, 4, tn);
      const uint32_t a_taddr = a_taddr0 + aslot * Cfg::A_COLS;
      mbar_wait_u32(a_empty0 + aslot * 8, aphase ^ 1u);  // the MMAs that read this TMEM slot have completed
      tc_fence_after();
      if (tracer) trace_evt(p, 4, tn);
      // LOGICAL k order: 16-k chunk c sits at physical chunk c ^ sw of the row (SWIZZLE_128B); the row base has its low 7
      // bits clear, so the address is (row | sw << 4) ^ (c << 4).  The swizzle stays on the load side: tcgen05.st takes
      // one warp-uniform tensor-memory address.
      {
        const uint32_t rbase = (raw0 + rslot * Cfg::RAW_BYTES) | (sw << 4);
        uint4 raw[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) raw[c] = ld_shared_v4(rbase ^ (static_cast<uint32_t>(c) << 4));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t o8[8];
          if (QB_KO(p.dbg, 1)) {
            o8[0] = raw[c].x; o8[1] = raw[c].y; o8[2] = raw[c].z; o8[3] = raw[c].w;
            o8[4] = raw[c].x; o8[5] = raw[c].y; o8[6] = raw[c].z; o8[7] = raw[c].w;
          } else {
            dequant16_plane<WT, ZP>(raw[c], high_plane, kc[c >> 1], o8);
          }
          tmem_st_32x32b_x8(a_taddr + (static_cast<uint32_t>(c) << 3), o8);
        }
      }
      tmem_st_wait();
      if (tracer) trace_evt(p, 4, tn);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_u32(raw_empty0 + rslot * 8);  // the stores above consumed the raw bytes (data dependency)
        mbar_arrive_u32(a_full0 + aslot * 8);
      }
      if (tracer) trace_evt(p, 4, tn);
      aslot += NG;
      if (aslot >= NSLOT) { aslot -= NSLOT; aphase ^= 1u; }
      rslot += NG;
      if (rslot >= RS) { rslot -= RS; rphase ^= 1u; }
      cur = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) gather_signal_end(p.g);  // fused all-gather: "this rank's slab has landed everywhere"
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace qb
