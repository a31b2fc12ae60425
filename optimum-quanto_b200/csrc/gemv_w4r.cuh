// Batch-1..16 decode fused packed-int4 linear, second generation of the TMA-ring gemv (gemv_w4s.cuh): same data flow,
// compile-time shaped inner loop.
//
//   out[m, n] = sum_k x[m, k] * dequant(W)[n, k]  (+ bias[n])          M <= 8 * TG
//
// What the first generation measured (ncu, profiles/r2_decode_m1_ncu_summary.json): 41 executed instructions per 256
// weights at 57 % issue-active -- the kernel is bound by instruction issue, not by HBM (0.28 of the roofline).  Of the
// ~105 instructions per 64-byte slab only 64 are the exact dequantisation + MMA (7 integer + 8 bf16x2 + 1 HMMA per 8
// weights and lane); the rest was address arithmetic, per-slab coefficient loads / re-packing behind a data-dependent
// branch, run-time loop control and re-materialised kernel parameters.  Here the shape of the loop is a template:
//   * SPW  slabs of 64 k-bytes per warp and stage (stage = 8 packed rows x SPW * 1024 k), fully unrolled: every
//          shared-memory address is `stage base + immediate`;
//   * GL   log2(group / 64): the (scale, shift) pairs of a lane change every 2^GL slabs at compile-time positions, two
//          adjacent groups come in with one 32-bit load per coefficient array;
//   * TG   token groups of 8: the dequantised A fragment of a k-step feeds TG mma.sync (M <= 16 runs here too);
//   * the activations are staged once per kernel, rows beyond M alias row M - 1 (their MMA columns are never stored), so
//          the loop carries no token predicate.
//   * a CTA owns WHOLE 8-row groups (the unit of the MMA tile), and its outputs are staged in shared memory as
//          [token][nibble plane][row] and leave with one cp.async.bulk store per (token, plane, destination buffer):
//          2 M stores of ~100 contiguous bytes per CTA and peer.  (The first generation wrote every output with a 2-byte
//          st.global from the reducer warp; measured on 2 GPUs the fused all-gather then cost 79 us at M = 1 and 1.1 ms
//          at M = 8 for a 13 us kernel: scattered sub-sector stores into peer memory are served one at a time.)
// Ownership, determinism and the protocols are otherwise unchanged: a CTA owns a contiguous range of packed rows over the whole K
// (no split-K, no workspace; the k-order of every output's sum is independent of how out-features are sharded, which is
// what makes the column-parallel result bit-identical to the single-GPU one), one producer warp streams packed rows and
// coefficient runs with cp.async.bulk (second generation: eight producer warps, one per packed row of a stage), one warp
// reduces the 16 per-warp partial tiles in a fixed order.
#pragma once

#include "gemv_w4s.cuh"

namespace qb {

struct GemvRParams {
  const uint8_t* wq;   // [N/2, K] packed bytes
  const void* scale;   // [N * K / group]
  const void* shift;   // same shape (weight dtype, or uint8 zero-points)
  const void* bias;    // [N] or nullptr
  const void* x;       // [M, K]
  void* out;           // [M, ld]
  GatherInfo g;        // fused all-gather (gather.cuh); g.n_out == 1: ordinary call
  int ld, col0;
  int M, N, K;
  int nxc;             // passes over K ("x chunks"): the activations of one pass sit in shared memory
  int kx;              // k per pass = K / nxc (a multiple of the stage width)
  int nkc;             // stages per row group and pass = kx / (SPW * 1024)
  int nstages;         // weight ring depth
  int cdepth;          // coefficient ring depth (row groups)
  int coef_arr;        // bytes of one coefficient array of a slot: 8 rows x (K / group) x 2
  int x_stride;        // bytes per token row in shared memory (kx * 2 + 16)
  int rc;              // rows of the output staging tile per (token, nibble plane): 8 * max row groups per CTA
  long long* trace;
};

// shared memory -> global (own buffer or a peer's over NVLink), 16-byte aligned on both sides, bytes % 16 == 0
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}

// developer timeline (make KNOCKOUTS=1 only): CTA < 4, role 0 compute warp 0, 1 producer, 2 reducer, 3 compute warp 15;
// up to 64 clock64 stamps each, [cta][role][64]
#ifdef QB_DEVELOPER_KNOCKOUTS
#define QB_RTRACE(role, cnt)                                                                       \
  do {                                                                                             \
    if (p.trace != nullptr && blockIdx.x < 4 && lane == 0 && (cnt) < 64)                           \
      p.trace[(blockIdx.x * 4 + (role)) * 64 + (cnt)++] = clock64();                               \
  } while (0)
#else
#define QB_RTRACE(role, cnt) do { } while (0)
#endif

constexpr int kGemvRWarps = 16;
constexpr int gemvr_threads(int np) { return (kGemvRWarps + np + 1) * 32; }

// MPASS: several passes over K (compiled out otherwise).  NP: producer warps (8: one per packed row of a stage; 4: the
// "lite" variant -- 672 threads, <= 48 registers, ~100 KB of shared memory, so that TWO CTAs fit an SM: consecutive decode
// kernels then overlap through programmatic dependent launch -- the next linear's CTAs are resident and stream its weights
// into their rings while this one still computes, instead of starting from an empty pipeline after every launch).
template <typename WT, bool ZP, int TG, int SPW, int GL, bool MPASS, int NP>
__global__ void __launch_bounds__(gemvr_threads(NP), NP == 4 ? 2 : 1) gemv_w4r_kernel(const GemvRParams p) {
  using D = Dq<WT>;
  constexpr int kGemvRProducers = NP;
  const int nxc = MPASS ? p.nxc : 1;
  static_assert(SPW >= (1 << GL), "a warp's run of slabs must start on a group boundary");
  constexpr int KC = SPW * kGemvRWarps * 64;     // k-bytes of a stage row
  constexpr int ROW_PITCH = KC + 64;             // +64: conflict-free LDS.128 across the 8 rows of a stage
  constexpr int STAGE_BYTES = 8 * ROW_PITCH;
  constexpr int RED_TILE = TG * 128;             // floats of one warp's partial tile(s)
  constexpr int RED_BYTES = 2 * kGemvRWarps * RED_TILE * 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t ring = smem_u32(smem);
  uint8_t* coef = smem + static_cast<size_t>(p.nstages) * STAGE_BYTES;                          // [cdepth][4][coef_arr]
  float* red = reinterpret_cast<float*>(coef + static_cast<size_t>(p.cdepth) * 4 * p.coef_arr);  // [2][16][RED_TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + RED_BYTES);
  // bars: full[nstages], empty[nstages], red_full[2], red_empty[2], coef_full[cdepth], coef_empty[cdepth]
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + p.nstages * 8;
  const uint32_t red_full0 = empty0 + p.nstages * 8, red_empty0 = red_full0 + 16;
  const uint32_t cfull0 = red_empty0 + 16, cempty0 = cfull0 + p.cdepth * 8;
  const uint32_t coef_addr = smem_u32(coef);
  uint32_t* magic_s = reinterpret_cast<uint32_t*>(bars + 2 * p.nstages + 4 + 2 * p.cdepth);  // 16-byte slot, see `magic` below
  uint8_t* outs = reinterpret_cast<uint8_t*>(magic_s) + 16;  // [M][2][rc] WT, 16-byte aligned
  float* part = reinterpret_cast<float*>(outs + static_cast<size_t>(p.M) * 2 * p.rc * sizeof(WT));  // [rc / 8][RED_TILE], nxc > 1
  uint8_t* xs = reinterpret_cast<uint8_t*>(part) + (nxc > 1 ? static_cast<size_t>(p.rc >> 3) * RED_TILE * 4 : 0);
  const uint32_t xs_addr = smem_u32(xs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half_n = p.N / 2;
  const int total_groups = half_n >> 3;  // (N / 2) % 8 == 0 on this path
  const int g_begin = static_cast<int>(static_cast<int64_t>(blockIdx.x) * total_groups / gridDim.x);
  const int g_end = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * total_groups / gridDim.x);
  const int r_begin = g_begin * 8, r_end = g_end * 8;
  const int ngroups = g_end - g_begin;
  const int gpr = (p.K >> 6) >> GL;  // groups per out-feature row

  if (threadIdx.x == 0) {
    *magic_s = D::MAGIC_BYTES;
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&bars[p.nstages + s], kGemvRWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[2 * p.nstages + b], kGemvRWarps);
      mbar_init(&bars[2 * p.nstages + 2 + b], 1);
    }
    for (int c = 0; c < p.cdepth; ++c) {
      mbar_init(&bars[2 * p.nstages + 4 + c], 1);
      mbar_init(&bars[2 * p.nstages + 4 + p.cdepth + c], kGemvRWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();  // the next kernel may start its weight stream on every SM this grid frees

  if (warp >= kGemvRWarps && warp < kGemvRWarps + kGemvRProducers) {
    // ------------------------------------------------------------------ TMA producers: one WARP per packed row
    // Measured (tools/trace_gemvr.py, tools/tma_probe.cu): a cp.async.bulk occupies its issuing warp for 200-400 cycles,
    // and the lanes of one warp issue theirs one after the other -- eight row copies from one warp took ~2000 cycles per
    // stage, i.e. 13 B/cycle/SM, less than the SM's share of HBM.  Issued from eight warps the copies overlap.
    if (lane != 0) return;
    const int pr = warp - kGemvRWarps;  // the packed row of every stage this warp copies
    const uint64_t pol = l2_policy_evict_first();
    int s = 0, pc = 0;
    uint32_t phase = 0, pcph = 0;
    [[maybe_unused]] int tn = 0;
    if (pr == 0) QB_RTRACE(1, tn);
    for (int pass = 0; pass < nxc; ++pass)
    for (int gi = 0; gi < ngroups; ++gi) {
      const int r0 = r_begin + gi * 8;
      const uint8_t* row_src = p.wq + static_cast<size_t>(r0 + pr) * p.K + static_cast<size_t>(pass) * p.kx;
      for (int kc = 0; kc < p.nkc; ++kc) {
        mbar_wait_u32(empty0 + s * 8, phase ^ 1u);
        if (pr == 0) mbar_arrive_expect_tx_u32(full0 + s * 8, 8u * KC);
#pragma unroll
        for (int rr = 0; rr < 8 / NP; ++rr)  // packed rows pr, pr + NP, ...
          bulk_load_1d(ring + s * STAGE_BYTES + (pr + rr * NP) * ROW_PITCH,
                       row_src + static_cast<size_t>(rr * NP) * p.K + static_cast<size_t>(kc) * KC, KC, full0 + s * 8, pol);
        if (pr == 0) QB_RTRACE(1, tn);
        if (++s == p.nstages) { s = 0; phase ^= 1u; }
        // scales / shifts of this row group: producer warps 0..3 copy one contiguous run each.  With several passes over
        // K every group keeps its own slot for the whole kernel (cdepth >= row groups of the CTA): loaded in pass 0 only.
        if (kc == 0 && pr < 4 && pass == 0) {
          const uint32_t sbytes = 8u * gpr * 2, zbytes = ZP ? sbytes / 2 : sbytes;
          mbar_wait_u32(cempty0 + pc * 8, pcph ^ 1u);
          if (pr == 0) mbar_arrive_expect_tx_u32(cfull0 + pc * 8, 2 * sbytes + 2 * zbytes);
          const size_t lo = static_cast<size_t>(r0) * gpr, hi = lo + static_cast<size_t>(half_n) * gpr;
          const uint8_t* sc = static_cast<const uint8_t*>(p.scale);
          const uint8_t* zs = static_cast<const uint8_t*>(p.shift);
          const uint8_t* src = (pr == 0) ? sc + lo * 2 : (pr == 1) ? sc + hi * 2
                               : (pr == 2) ? zs + lo * (ZP ? 1 : 2) : zs + hi * (ZP ? 1 : 2);
          bulk_load_1d(coef_addr + (pc * 4 + pr) * p.coef_arr, src, pr < 2 ? sbytes : zbytes, cfull0 + pc * 8, pol);
          if (++pc == p.cdepth) { pc = 0; pcph ^= 1u; }
        }
      }
    }
    return;
  }

  if (warp == kGemvRWarps + kGemvRProducers) {
    // ------------------------------------------------------------------ reducer: 16 partial tiles -> outputs
    WT* os = reinterpret_cast<WT*>(outs);
    [[maybe_unused]] int tn = 0;
    QB_RTRACE(2, tn);
    int it = 0;
    for (int pass = 0; pass < nxc; ++pass)
    for (int gi = 0; gi < ngroups; ++gi, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      const int r0 = r_begin + gi * 8;
      const bool first = pass == 0, last = pass == nxc - 1;
      mbar_wait_u32(red_full0 + b * 8, ph);
      QB_RTRACE(2, tn);
      const float* rb = red + b * (kGemvRWarps * RED_TILE);
#pragma unroll
      for (int q = 0; q < 4 * TG; ++q) {
        const int o = lane + 32 * q;       // o = token group * 128 + tile row * 8 + token
        const int tgi = o >> 7, trow = (o >> 3) & 15, tok = tgi * 8 + (o & 7);
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kGemvRWarps; ++w) sum += rb[w * RED_TILE + o];
        // several passes over K: the passes' sums are added in pass order (fixed, independent of the sharding)
        if (!first) sum = __fadd_rn(part[gi * RED_TILE + o], sum);
        if (!last) part[gi * RED_TILE + o] = sum;
        if (last && tok < p.M) {
          const int plane = trow >> 3;
          WT r = from_float<WT>(sum);
          if (p.bias != nullptr) {
            const int n = plane * half_n + r0 + (trow & 7);
            r = from_float<WT>(__fadd_rn(to_float<WT>(r), to_float<WT>(static_cast<const WT*>(p.bias)[n])));
          }
          os[(tok * 2 + plane) * p.rc + gi * 8 + (trow & 7)] = r;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_u32(red_empty0 + b * 8);
    }
    // the CTA's outputs: one bulk store per (token, nibble plane, destination buffer)
    fence_proxy_async_smem();
    __syncwarp();
    const uint32_t run_bytes = static_cast<uint32_t>(r_end - r_begin) * sizeof(WT);
    if (run_bytes != 0) {
      for (int i = lane; i < p.M * 2; i += 32) {
        const int tok = i >> 1, plane = i & 1;
        const size_t o_idx = static_cast<size_t>(tok) * p.ld + p.col0 + plane * half_n + r_begin;
        const uint32_t src = smem_u32(os + static_cast<size_t>(i) * p.rc);
        bulk_store_1d(static_cast<WT*>(p.out) + o_idx, src, run_bytes);
        for (int pq = 1; pq < p.g.n_out; ++pq) bulk_store_1d(static_cast<WT*>(p.g.out_peer[pq]) + o_idx, src, run_bytes);
      }
      bulk_commit_group();
      QB_RTRACE(2, tn);
      // fused gather: the writes must be PERFORMED before the flag is published.  An ordinary call only has to keep the
      // CTA (its shared memory) alive until the copy engine has read the staging tile: the writes complete before the
      // grid does, which is all a later kernel or copy can observe.
      if (p.g.n_out > 1) bulk_wait_group_all();
      else bulk_wait_group_read<0>();
    }
    QB_RTRACE(2, tn);
    __syncwarp();  // this warp made every output store of the CTA
    if (lane == 0) gather_signal_end(p.g);
    return;
  }

  // -------------------------------------------------------------------- compute warps
  [[maybe_unused]] int tn = 0;
  [[maybe_unused]] const int trole = (warp == 0) ? 0 : 3;
  [[maybe_unused]] const bool tracer = (warp == 0 || warp == kGemvRWarps - 1);
  if (tracer) QB_RTRACE(trole, tn);
  {
    const int ct = threadIdx.x;  // 0 .. 511
    pdl_wait();  // the activations (and the gather flags) are the previous kernel's output; the weight stream is running
    if (p.g.n_out > 1 && p.g.wait_start) {  // the activation is the gathered output of the previous linear
      if (ct == 0) gather_wait_start(p.g);
      asm volatile("bar.sync 1, %0;" ::"n"(kGemvRWarps * 32) : "memory");
    }
  }
  if (tracer) QB_RTRACE(trole, tn);
  const int g = lane >> 2;  // packed row inside the group (MMA rows g and g + 8), and the token column of B
  const int t = lane & 3;   // owns bytes [16t, 16t + 16) of every 64-byte slab
  // this lane's activations: token tg * 8 + g (tokens beyond M alias the last one; their MMA columns are never stored)
  uint32_t x_lane[TG];
#pragma unroll
  for (int tg = 0; tg < TG; ++tg)
    x_lane[tg] = xs_addr + min(tg * 8 + g, p.M - 1) * p.x_stride + (warp * SPW * 64 + t * 16) * 2;
  const uint32_t w_lane = ring + g * ROW_PITCH + warp * SPW * 64 + t * 16;
  // coefficient index of the warp's first slab inside a stage; the groups of a stage start at kc * (KC >> (6 + GL))
  const int qg_warp = (warp * SPW) >> GL;
  constexpr int NQ = SPW >> GL;        // groups covered by a warp's run of slabs
  constexpr int QSTEP = KC >> (6 + GL);  // groups per stage row

  // the magic exponent bytes live in a REGISTER the compiler cannot fold (read back from shared memory): PRMT then takes the byte selector as its
  // immediate.  (With both constant, ptxas keeps the selector in a register that every PRMT overwrites and re-creates
  // it with a move per PRMT: 64 extra instructions per stage.)
  const uint32_t magic = *reinterpret_cast<volatile uint32_t*>(magic_s);
  int s = 0;
  uint32_t phase = 0;
  int cs = 0;
  uint32_t cphase = 0;
  int it = 0;
  for (int pass = 0; pass < nxc; ++pass) {
  {  // the activations of this pass: columns [pass * kx, (pass + 1) * kx) of every token -> shared memory
    const int ct = threadIdx.x;  // 0 .. 511
    if (pass > 0) asm volatile("bar.sync 1, %0;" ::"n"(kGemvRWarps * 32) : "memory");  // every warp has finished the previous pass
    const int vec_per_row = p.kx / 8;  // 16-byte vectors per token row and pass
    for (int i = ct; i < p.M * vec_per_row; i += kGemvRWarps * 32) {
      const int m = i / vec_per_row, v = i - m * vec_per_row;
      // plain (coherent) load: with a gathered input these bytes were written by peers during the previous kernel
      const uint4 val = *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.x) +
                                                        (static_cast<size_t>(m) * p.K + static_cast<size_t>(pass) * p.kx + v * 8) * 2);
      *reinterpret_cast<uint4*>(xs + static_cast<size_t>(m) * p.x_stride + v * 16) = val;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kGemvRWarps * 32) : "memory");
    if (tracer) QB_RTRACE(trole, tn);
  }
  const int qg_pass = (pass * p.kx) >> (6 + GL);  // first group of this pass
  if (nxc > 1) { cs = 0; cphase = 0; }       // resident coefficient slots: slot = row group, filled once
  for (int gi = 0; gi < ngroups; ++gi, ++it) {
    float acc[TG][2][4];
#pragma unroll
    for (int tg = 0; tg < TG; ++tg)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[tg][a][j] = 0.f;
    // this lane's row in the four coefficient arrays of the slot (scale lo / hi, shift lo / hi)
    const uint32_t c_lane = coef_addr + cs * 4 * p.coef_arr + (g * gpr + qg_pass + qg_warp) * 2;
    const uint32_t z_lane = coef_addr + (cs * 4 + 2) * p.coef_arr + (g * gpr + qg_pass + qg_warp) * (ZP ? 1 : 2);
    mbar_wait_u32(cfull0 + cs * 8, cphase);
#pragma unroll 1
    for (int kc = 0; kc < p.nkc; ++kc) {
      // raw coefficients of the NQ groups this warp's slabs fall into (16-bit payloads, or zero-point bytes)
      uint16_t s_lo[NQ], s_hi[NQ], z_lo[NQ], z_hi[NQ];
      {
        const uint32_t ca = c_lane + kc * QSTEP * 2, za = z_lane + kc * QSTEP * (ZP ? 1 : 2);
        if constexpr (NQ == 2) {
          uint32_t a, b;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(ca));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b) : "r"(ca + p.coef_arr));
          s_lo[0] = static_cast<uint16_t>(a); s_lo[1] = static_cast<uint16_t>(a >> 16);
          s_hi[0] = static_cast<uint16_t>(b); s_hi[1] = static_cast<uint16_t>(b >> 16);
          if constexpr (ZP) {
            uint16_t c, d;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c) : "r"(za));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(d) : "r"(za + p.coef_arr));
            z_lo[0] = c & 0xFF; z_lo[1] = c >> 8;
            z_hi[0] = d & 0xFF; z_hi[1] = d >> 8;
          } else {
            uint32_t c, d;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(za));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"(za + p.coef_arr));
            z_lo[0] = static_cast<uint16_t>(c); z_lo[1] = static_cast<uint16_t>(c >> 16);
            z_hi[0] = static_cast<uint16_t>(d); z_hi[1] = static_cast<uint16_t>(d >> 16);
          }
        } else {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(s_lo[q]) : "r"(ca + q * 2));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(s_hi[q]) : "r"(ca + p.coef_arr + q * 2));
            if constexpr (ZP) {
              asm volatile("ld.shared.u8 %0, [%1];" : "=h"(z_lo[q]) : "r"(za + q));
              asm volatile("ld.shared.u8 %0, [%1];" : "=h"(z_hi[q]) : "r"(za + p.coef_arr + q));
            } else {
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(z_lo[q]) : "r"(za + q * 2));
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(z_hi[q]) : "r"(za + p.coef_arr + q * 2));
            }
          }
        }
      }
      mbar_wait_u32(full0 + s * 8, phase);
      if (tracer) QB_RTRACE(trole, tn);
      const uint32_t wb = w_lane + s * STAGE_BYTES;
      // every operand of the stage is loaded before the first conversion (SPW * (1 + 2 TG) LDS.128 in flight)
      uint4 w[SPW], xa[SPW][TG], xc[SPW][TG];
#pragma unroll
      for (int sl = 0; sl < SPW; ++sl) {
        w[sl] = ld_shared_v4(wb + sl * 64);
#pragma unroll
        for (int tg = 0; tg < TG; ++tg) {
          xa[sl][tg] = ld_shared_v4(x_lane[tg] + kc * (KC * 2) + sl * 128);
          xc[sl][tg] = ld_shared_v4(x_lane[tg] + kc * (KC * 2) + sl * 128 + 16);
        }
      }
      typename D::Coef c_lo, c_hi;
#pragma unroll
      for (int sl = 0; sl < SPW; ++sl) {
        if ((sl & ((1 << GL) - 1)) == 0) {  // compile-time: a new group starts at this slab
          c_lo = D::make_raw(*reinterpret_cast<const WT*>(&s_lo[sl >> GL]), z_lo[sl >> GL], ZP);
          c_hi = D::make_raw(*reinterpret_cast<const WT*>(&s_hi[sl >> GL]), z_hi[sl >> GL], ZP);
        }
        const uint32_t w4[4] = {w[sl].x, w[sl].y, w[sl].z, w[sl].w};
#pragma unroll
        for (int st = 0; st < 4; ++st) {  // k-step: bytes 4 st .. 4 st + 3 of the lane's 16
          const uint32_t l = w4[st] & 0x0F0F0F0Fu, h = (w4[st] >> 4) & 0x0F0F0F0Fu;
          uint32_t a[4];
          a[0] = D::cvt(__byte_perm(l, magic, 0x4140), c_lo, ZP);  // row g   (low nibble),  k slots 2t, 2t+1
          a[1] = D::cvt(__byte_perm(h, magic, 0x4140), c_hi, ZP);  // row g+8 (high nibble)
          a[2] = D::cvt(__byte_perm(l, magic, 0x4342), c_lo, ZP);  // row g,   k slots 2t+8, 2t+9
          a[3] = D::cvt(__byte_perm(h, magic, 0x4342), c_hi, ZP);
#pragma unroll
          for (int tg = 0; tg < TG; ++tg) {
            const uint32_t xb[8] = {xa[sl][tg].x, xa[sl][tg].y, xa[sl][tg].z, xa[sl][tg].w,
                                    xc[sl][tg].x, xc[sl][tg].y, xc[sl][tg].z, xc[sl][tg].w};
            mma_m16n8k16<WT>(acc[tg][st & 1], a, xb[2 * st], xb[2 * st + 1]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_u32(empty0 + s * 8);
      if (tracer) QB_RTRACE(trole, tn);
      if (++s == p.nstages) { s = 0; phase ^= 1u; }
    }
    __syncwarp();
    if (nxc == 1) {
      if (lane == 0) mbar_arrive_u32(cempty0 + cs * 8);
      if (++cs == p.cdepth) { cs = 0; cphase ^= 1u; }
    } else {
      ++cs;
    }
    // ---- this warp's partial 16 x 8 tile(s) -> reducer
    const int b = it & 1;
    mbar_wait_u32(red_empty0 + b * 8, ((it >> 1) & 1u) ^ 1u);
    float* rp = red + (b * kGemvRWarps + warp) * RED_TILE;
#pragma unroll
    for (int tg = 0; tg < TG; ++tg) {
      *reinterpret_cast<float2*>(rp + tg * 128 + g * 8 + 2 * t) =
          make_float2(acc[tg][0][0] + acc[tg][1][0], acc[tg][0][1] + acc[tg][1][1]);        // row g
      *reinterpret_cast<float2*>(rp + tg * 128 + (g + 8) * 8 + 2 * t) =
          make_float2(acc[tg][0][2] + acc[tg][1][2], acc[tg][0][3] + acc[tg][1][3]);        // row g + 8
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_u32(red_full0 + b * 8);
  }
  }
}

}  // namespace qb
