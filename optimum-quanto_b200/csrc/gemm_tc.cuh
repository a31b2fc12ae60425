// Persistent, warp-specialised tcgen05 GEMM for the quantized-linear forward.
//
//   out[m, n] = epilogue( sum_k A[m, k] * Wt[n, k] )
//
// One CTA per SM loops over output tiles.  Roles (one warp each unless noted):
//   warp 0      TMA producer   : A tiles (and W tiles when W is 8-bit and fed straight to the tensor core)
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma, accumulators live in TMEM
//   warps 2-5   epilogue       : tcgen05.ld TMEM -> registers -> scale / bias / round -> global
//   warps 6-13  weight staging : (int4 / mixed paths only) vector-load the packed uint8 weight bytes, unpack,
//                                apply the per-group scale/shift with the reference's exact rounding order and
//                                write the bf16/fp16 tile into shared memory in the 128B-swizzled K-major
//                                layout the tensor core reads -- no separate dequantise launch, the
//                                dequantised weight never exists in HBM.
//
// Shared-memory operand layout (both operands, K-major): a tile is `rows x 128 bytes`; row r lives at
// (r/8)*1024 + (r%8)*128 and its 16-byte chunk c is stored at chunk position c ^ (r%8)  (SWIZZLE_128B, the
// layout TMA produces and the UMMA descriptor in common.cuh describes).
#pragma once

#include "common.cuh"

namespace qb {

enum class BSrc { TMA, INT4 };

struct GemmParams {
  // epilogue
  const void* scales;  // [N] in the output dtype, applied in fp32 to the accumulator (8-bit paths) or nullptr
  const void* bias;    // [N] in the output dtype or nullptr; added after the result is rounded (reference order)
  void* out;           // [M, N]
  int out_dt;          // DT_F32 / DT_F16 / DT_BF16
  int M, N, K;         // K in elements
  int num_m_blocks, num_n_blocks;
  // int4 weight source (BSrc::INT4): canonical quanto storage viewed as [N/2, K] bytes,
  // low nibble = out-feature n, high nibble = out-feature n + N/2 (tensor/packed.py:45-69)
  const uint8_t* wq;
  const void* wscale;  // [N * K / group]
  const void* wshift;  // same shape; weight dtype, or uint8 zero-point when shift_is_int
  int group;
  int group_log2;      // log2(group) when group is a power of two, else -1
  int shift_is_int;
};

template <MmaKind KIND_, BSrc BSRC_, int MSUB_, int BN_, typename WT_, bool ZP_ = false>
struct GemmCfg {
  static constexpr bool ZP = ZP_;     // INT4 only: shift is an integer zero-point (compile-time: keeps the hot loop lean)
  static constexpr MmaKind KIND = KIND_;
  static constexpr BSrc BSRC = BSRC_;
  static constexpr int MSUB = MSUB_;  // 128-row A sub-tiles per CTA tile (B tile reused across them)
  static constexpr int BN = BN_;      // UMMA N
  using WT = WT_;                     // element type of the staged weight tile (bf16 / half), INT4 only
  static constexpr int BM = 128;      // UMMA M (cta_group::1)
  static constexpr int KBYTES = 128;  // K extent of one pipeline stage in bytes = one swizzle atom
  static constexpr int A_TILE = BM * KBYTES;
  static constexpr int B_TILE = BN * KBYTES;
  static constexpr int STAGE = MSUB * A_TILE + B_TILE;
  static constexpr int NSTAGES = (196 * 1024) / STAGE;
  static constexpr int ACC_COLS = MSUB * BN;
  static constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = 512;
  static constexpr int NCVT_WARPS = (BSRC == BSrc::TMA) ? 0 : 8;
  static constexpr int NCVT_THREADS = NCVT_WARPS * 32;
  static constexpr int NTHREADS = (6 + NCVT_WARPS) * 32;
  static constexpr int SMEM_BYTES = NSTAGES * STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static_assert(ACC_COLS * NACC <= 512, "TMEM overflow");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
};

// ------------------------------------------------------------------------------------------------
// int4 -> bf16/fp16 with the reference's rounding order (tensor/qbits.py:34-45)
//   float shift: d = rnd(rnd(s*q) - z)            int shift: d = rnd(s * (q - zp))
// q enters as the magic-number pattern (bf16: 0x4300|q = 128+q, fp16: 0x6400|q = 1024+q), two per register.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Dq;

template <>
struct Dq<__nv_bfloat16> {
  using V2 = __nv_bfloat162;
  static constexpr uint32_t MAGIC_BYTES = 0x43434343u;
  struct Coef { V2 s, c, z; };  // float shift: c = -128*s ; int shift: c = 128 + zp
  // `zraw`: the shift's 16-bit payload (bf16 bits) or, for integer shifts, the zero-point byte
  __device__ __forceinline__ static Coef make_raw(__nv_bfloat16 s, uint16_t zraw, bool is_int) {
    Coef k;
    k.s = __bfloat162bfloat162(s);
    if (is_int) {
      const int zp = static_cast<int8_t>(static_cast<uint8_t>(zraw));
      k.c = __bfloat162bfloat162(__float2bfloat16_rn(128.f + static_cast<float>(zp)));
      k.z = k.c;
    } else {
      k.c = __hmul2_rn(k.s, __bfloat162bfloat162(__float2bfloat16_rn(-128.f)));
      k.z = __bfloat162bfloat162(__ushort_as_bfloat16(zraw));
    }
    return k;
  }
  __device__ static Coef make(__nv_bfloat16 s, const void* shift_ptr, int64_t idx, bool is_int) {
    const uint16_t zraw = is_int ? static_cast<uint16_t>(static_cast<const uint8_t*>(shift_ptr)[idx])
                                 : static_cast<const uint16_t*>(shift_ptr)[idx];
    return make_raw(s, zraw, is_int);
  }
  __device__ __forceinline__ static uint32_t cvt(uint32_t m, const Coef& k, bool is_int) {
    V2 v = *reinterpret_cast<V2*>(&m);
    V2 d;
    // the *_rn intrinsics forbid ptxas from contracting mul+sub into one fma (which would drop a rounding step)
    if (is_int) d = __hmul2_rn(k.s, __hsub2_rn(v, k.c));      // (128+q)-(128+zp) exact, one rounding
    else d = __hsub2_rn(__hfma2(k.s, v, k.c), k.z);           // s*(128+q) - 128*s == s*q exactly -> rnd ; then - z -> rnd
    return *reinterpret_cast<uint32_t*>(&d);
  }
};

template <>
struct Dq<__half> {
  using V2 = __half2;
  static constexpr uint32_t MAGIC_BYTES = 0x64646464u;
  struct Coef { V2 s, c, z; };  // c = 1024 (float shift) or 1024 + zp (int shift)
  __device__ __forceinline__ static Coef make_raw(__half s, uint16_t zraw, bool is_int) {
    Coef k;
    k.s = __half2half2(s);
    if (is_int) {
      const int zp = static_cast<int8_t>(static_cast<uint8_t>(zraw));
      k.c = __half2half2(__float2half_rn(1024.f + static_cast<float>(zp)));
      k.z = k.c;
    } else {
      k.c = __half2half2(__float2half_rn(1024.f));
      k.z = __half2half2(__ushort_as_half(zraw));
    }
    return k;
  }
  __device__ static Coef make(__half s, const void* shift_ptr, int64_t idx, bool is_int) {
    const uint16_t zraw = is_int ? static_cast<uint16_t>(static_cast<const uint8_t*>(shift_ptr)[idx])
                                 : static_cast<const uint16_t*>(shift_ptr)[idx];
    return make_raw(s, zraw, is_int);
  }
  __device__ __forceinline__ static uint32_t cvt(uint32_t m, const Coef& k, bool is_int) {
    V2 v = *reinterpret_cast<V2*>(&m);
    V2 d;
    if (is_int) d = __hmul2_rn(k.s, __hsub2_rn(v, k.c));
    else d = __hsub2_rn(__hmul2_rn(k.s, __hsub2_rn(v, k.c)), k.z);  // (1024+q)-1024 exact ; s*q -> rnd ; - z -> rnd
    return *reinterpret_cast<uint32_t*>(&d);
  }
};

// 16 packed bytes (16 k of two out-features) -> 2 x 8 registers of bf16x2 / half2 (natural k order).
// 7 integer ops + 8 (bf16) half-precision ops per 8 weights.
template <typename WT, bool ZP>
__device__ __forceinline__ void dequant16(const uint4& raw, const typename Dq<WT>::Coef& klo,
                                          const typename Dq<WT>::Coef& khi, uint32_t (&lo)[8], uint32_t (&hi)[8]) {
  using D = Dq<WT>;
  const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t wl = w4[i] & 0x0F0F0F0Fu;
    const uint32_t wh = (w4[i] >> 4) & 0x0F0F0F0Fu;
    lo[2 * i + 0] = D::cvt(__byte_perm(wl, D::MAGIC_BYTES, 0x4140), klo, ZP);
    lo[2 * i + 1] = D::cvt(__byte_perm(wl, D::MAGIC_BYTES, 0x4342), klo, ZP);
    hi[2 * i + 0] = D::cvt(__byte_perm(wh, D::MAGIC_BYTES, 0x4140), khi, ZP);
    hi[2 * i + 1] = D::cvt(__byte_perm(wh, D::MAGIC_BYTES, 0x4342), khi, ZP);
  }
}

// Epilogue for one 16-column chunk held by one thread (= one output row).
template <typename OT, bool IS_INT_ACC>
__device__ __forceinline__ void epilogue_store16(const uint32_t (&v)[16], OT* __restrict__ out_row, int n_first,
                                                 int n_limit, bool row_ok, const OT* __restrict__ scales,
                                                 const OT* __restrict__ bias, bool vec_ok) {
  alignas(16) OT o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float f = IS_INT_ACC ? __int2float_rn(static_cast<int>(v[j])) : __uint_as_float(v[j]);
    const int n = n_first + j;
    if (scales != nullptr && n < n_limit) f = __fmul_rn(f, to_float<OT>(scales[n]));
    OT r = from_float<OT>(f);
    if (bias != nullptr && n < n_limit) r = from_float<OT>(__fadd_rn(to_float<OT>(r), to_float<OT>(bias[n])));
    o[j] = r;
  }
  if (!row_ok) return;
  if (vec_ok && n_first + 16 <= n_limit) {
    uint4* dst = reinterpret_cast<uint4*>(out_row + n_first);
    const uint4* src = reinterpret_cast<const uint4*>(o);
#pragma unroll
    for (int q = 0; q < static_cast<int>(sizeof(OT) * 16 / 16); ++q) dst[q] = src[q];
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n_first + j < n_limit) out_row[n_first + j] = o[j];
  }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const GemmParams p, const uint32_t idesc) {
  constexpr int NSTAGES = Cfg::NSTAGES;
  constexpr int MSUB = Cfg::MSUB;
  constexpr int BN = Cfg::BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGES * Cfg::STAGE);
  uint64_t* empty_bar = full_bar + NSTAGES;
  uint64_t* tmem_full_bar = empty_bar + NSTAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    if constexpr (Cfg::BSRC == BSrc::TMA) tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(&full_bar[s], 1 + Cfg::NCVT_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  constexpr int A_ELEM = (Cfg::KIND == MmaKind::F16) ? 2 : 1;
  constexpr int KELEMS = Cfg::KBYTES / A_ELEM;  // K elements per stage
  const int kblocks = (p.K + KELEMS - 1) / KELEMS;

  auto a_smem = [&](int s, int ms) { return smem + s * Cfg::STAGE + ms * Cfg::A_TILE; };
  auto b_smem = [&](int s) { return smem + s * Cfg::STAGE + MSUB * Cfg::A_TILE; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      constexpr uint32_t tx_bytes = MSUB * Cfg::A_TILE + (Cfg::BSRC == BSrc::TMA ? Cfg::B_TILE : 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.num_m_blocks;
        const int n_blk = tile / p.num_m_blocks;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
          for (int ms = 0; ms < MSUB; ++ms)
            tma_load_2d(a_smem(stage, ms), &tmap_a, &full_bar[stage], kb * KELEMS, (m_blk * MSUB + ms) * Cfg::BM);
          if constexpr (Cfg::BSRC == BSrc::TMA)
            tma_load_2d(b_smem(stage), &tmap_b, &full_bar[stage], kb * KELEMS, n_blk * BN);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
        const uint32_t acc = acc_it % Cfg::NACC;
        const uint32_t acc_phase = (acc_it / Cfg::NACC) & 1u;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_smem(stage));
#pragma unroll
          for (int k = 0; k < Cfg::KBYTES / 32; ++k) {
            const uint64_t b_desc = umma_desc_sw128_kmajor(b_addr + k * 32);
#pragma unroll
            for (int ms = 0; ms < MSUB; ++ms) {
              const uint64_t a_desc = umma_desc_sw128_kmajor(smem_u32(a_smem(stage, ms)) + k * 32);
              tc_mma<Cfg::KIND>(tmem_base + acc * Cfg::ACC_COLS + ms * BN, a_desc, b_desc, idesc,
                                (kb | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int quarter = warp & 3;
    uint32_t acc_it = 0;
    const bool int_split = (Cfg::BSRC == BSrc::INT4);
    const int half_n = p.N / 2;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
      const int m_blk = tile % p.num_m_blocks;
      const int n_blk = tile / p.num_m_blocks;
      const uint32_t acc = acc_it % Cfg::NACC;
      const uint32_t acc_phase = (acc_it / Cfg::NACC) & 1u;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ms = 0; ms < MSUB; ++ms) {
        const int row = (m_blk * MSUB + ms) * Cfg::BM + quarter * 32 + lane;
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 16; ++chunk) {
          uint32_t v[16];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>(acc * Cfg::ACC_COLS + ms * BN + chunk * 16);
          tmem_ld_32x32b_x16(taddr, v);
          tmem_ld_wait();
          int n_first, n_limit;
          if (int_split) {
            // tile columns [0, BN/2) are the low-nibble rows, [BN/2, BN) the high-nibble rows (+N/2)
            const int c = chunk * 16;
            if (c < BN / 2) { n_first = n_blk * (BN / 2) + c; n_limit = half_n; }
            else { n_first = half_n + n_blk * (BN / 2) + (c - BN / 2); n_limit = p.N; }
          } else {
            n_first = n_blk * BN + chunk * 16;
            n_limit = p.N;
          }
          constexpr bool IS_INT = (Cfg::KIND == MmaKind::I8);
          const size_t row_off = static_cast<size_t>(row_ok ? row : 0) * p.N;
          if (p.out_dt == DT_BF16) {
            epilogue_store16<__nv_bfloat16, IS_INT>(v, static_cast<__nv_bfloat16*>(p.out) + row_off, n_first, n_limit,
                                                   row_ok, static_cast<const __nv_bfloat16*>(p.scales),
                                                   static_cast<const __nv_bfloat16*>(p.bias), (p.N % 8) == 0);
          } else if (p.out_dt == DT_F16) {
            epilogue_store16<__half, IS_INT>(v, static_cast<__half*>(p.out) + row_off, n_first, n_limit, row_ok,
                                             static_cast<const __half*>(p.scales),
                                             static_cast<const __half*>(p.bias), (p.N % 8) == 0);
          } else {
            epilogue_store16<float, IS_INT>(v, static_cast<float*>(p.out) + row_off, n_first, n_limit, row_ok,
                                            static_cast<const float*>(p.scales), static_cast<const float*>(p.bias),
                                            (p.N % 4) == 0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  } else {
    // ------------------------------------------------------------------ weight staging (int4 -> WT tile)
    if constexpr (Cfg::BSRC == BSrc::INT4) {
      using WT = typename Cfg::WT;
      using D = Dq<WT>;
      constexpr bool ZP = Cfg::ZP;
      constexpr int ROWP = BN / 2;                             // row pairs (packed byte rows) per tile
      constexpr int KB_BYTES = 64;                             // packed bytes per row-pair per stage (64 k)
      constexpr int BPT = KB_BYTES * ROWP / Cfg::NCVT_THREADS;  // packed bytes per thread per stage
      constexpr int NV = BPT / 16;
      static_assert(BPT % 16 == 0 && NV >= 1, "staging split");
      constexpr int TPR = KB_BYTES / BPT;                       // threads per row pair
      static_assert(ROWP * TPR == Cfg::NCVT_THREADS, "thread map");
      static_assert(ROWP % 8 == 0, "both rows of a pair must share the swizzle phase");
      const int ct = threadIdx.x - 6 * 32;
      const int r = ct % ROWP;
      const int h = ct / ROWP;  // which BPT-byte slice of the 64-byte row
      const int half_n = p.N / 2;
      const int groups_per_row = p.K / p.group;
      const WT* scale = static_cast<const WT*>(p.wscale);

      const int my_tiles = (num_tiles > static_cast<int>(blockIdx.x))
                               ? (num_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1
                               : 0;
      const int total_it = my_tiles * kblocks;

      // Register prefetch ring, 2 stages ahead: packed bytes AND the group's scale / shift, so that neither the
      // L2 latency of the weights nor that of the (strided) scale reads sits on the staging critical path.
      // The host guarantees group % 32 == 0, so the BPT (<= 32) k handled by one thread share one group.
      // All indices advance incrementally: no integer division in the steady state.
      struct Pre {
        uint4 raw[NV];
        WT s_lo, s_hi;
        uint16_t z_lo, z_hi;
        bool ok;
      };
      constexpr int PF = 3;
      Pre ring[PF];
      int f_kb = 0, f_tile = blockIdx.x, f_left = total_it;
      int f_rp = (f_tile / p.num_m_blocks) * ROWP + r;
      auto load_pre = [&](Pre& pr) {
        if (f_left <= 0) return;
        --f_left;
        const int kbase = f_kb * KB_BYTES + h * BPT;
        pr.ok = f_rp < half_n && kbase < p.K;
        if (pr.ok) {
          const uint8_t* src = p.wq + static_cast<size_t>(f_rp) * p.K + kbase;
#pragma unroll
          for (int v = 0; v < NV; ++v) pr.raw[v] = __ldg(reinterpret_cast<const uint4*>(src + v * 16));
          const int g = (p.group_log2 >= 0) ? (kbase >> p.group_log2) : (kbase / p.group);
          const size_t ilo = static_cast<size_t>(f_rp) * groups_per_row + g;
          const size_t ihi = ilo + static_cast<size_t>(half_n) * groups_per_row;
          pr.s_lo = __ldg(scale + ilo);
          pr.s_hi = __ldg(scale + ihi);
          if (ZP) {
            pr.z_lo = __ldg(static_cast<const uint8_t*>(p.wshift) + ilo);
            pr.z_hi = __ldg(static_cast<const uint8_t*>(p.wshift) + ihi);
          } else {
            pr.z_lo = __ldg(static_cast<const uint16_t*>(p.wshift) + ilo);
            pr.z_hi = __ldg(static_cast<const uint16_t*>(p.wshift) + ihi);
          }
        }
        if (++f_kb == kblocks) {
          f_kb = 0;
          f_tile += gridDim.x;
          f_rp = (f_tile / p.num_m_blocks) * ROWP + r;
        }
      };
#pragma unroll
      for (int u = 0; u < PF - 1; ++u) load_pre(ring[u]);

      // destination offsets inside the B tile (constant per thread)
      const uint32_t row_lo = static_cast<uint32_t>(r);
      const uint32_t sw = row_lo & 7;
      const uint32_t off_lo = (row_lo >> 3) * 1024 + (row_lo & 7) * 128;
      const uint32_t off_hi = off_lo + (ROWP / 8) * 1024;
      uint32_t dst[2 * NV];  // 16-byte chunk byte offsets within a row (swizzled)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const uint32_t c = static_cast<uint32_t>((h * BPT + v * 16) >> 3);  // 16 k = 32 B = chunks c, c+1
        dst[2 * v + 0] = ((c + 0) ^ sw) << 4;
        dst[2 * v + 1] = ((c + 1) ^ sw) << 4;
      }
      const uint32_t b_base0 = smem_u32(smem) + MSUB * Cfg::A_TILE;
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);

      int stage = 0;
      uint32_t phase = 0;
      for (int it0 = 0; it0 < total_it; it0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          if (it0 + u < total_it) {
            load_pre(ring[(u + PF - 1) % PF]);
            const Pre& cur = ring[u];
            typename D::Coef klo, khi;
            if (cur.ok) {
              klo = D::make_raw(cur.s_lo, cur.z_lo, ZP);
              khi = D::make_raw(cur.s_hi, cur.z_hi, ZP);
            }
            mbar_wait_u32(empty0 + stage * 8, phase ^ 1u);
            const uint32_t bt = b_base0 + stage * Cfg::STAGE;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              uint32_t lo[8], hi[8];
              if (cur.ok) {
                dequant16<WT, ZP>(cur.raw[v], klo, khi, lo, hi);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) { lo[i] = 0u; hi[i] = 0u; }
              }
              st_shared_v4(bt + off_lo + dst[2 * v + 0], lo[0], lo[1], lo[2], lo[3]);
              st_shared_v4(bt + off_lo + dst[2 * v + 1], lo[4], lo[5], lo[6], lo[7]);
              st_shared_v4(bt + off_hi + dst[2 * v + 0], hi[0], hi[1], hi[2], hi[3]);
              st_shared_v4(bt + off_hi + dst[2 * v + 1], hi[4], hi[5], hi[6], hi[7]);
            }
            fence_proxy_async_smem();  // every writer: generic-proxy stores -> visible to the async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_u32(full0 + stage * 8);  // one arrive per warp (per-thread arrives serialise)
            if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
          }
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
