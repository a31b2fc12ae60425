// Weight-freeze and calibration kernels: the step BEFORE the quantized-linear hot path.
//
//   quantize_affine            -- quanto::quantize_affine, optimum/quanto/library/quantize.py:58-78 (3-5 ATen launches)
//   pack                       -- pack_weights, optimum/quanto/tensor/packed.py:24-69 (python loop of shift+or launches)
//   quantize_qbits_max         -- MaxOptimizer (tensor/optimizers/max_optimizer.py:26-37, affine_optimizer.py:52-63)
//                                 + quantize_affine + pack_weights for an axis-0 grouped weight as ONE launch:
//                                 reads the float weight once (2 B/weight), writes 0.5 B/weight + scale/shift
//   absmax                     -- calibrate.py:37-61 absmax_scale, per-tensor reduction
//   quantize_qbytes_absmax     -- AbsmaxOptimizer (tensor/optimizers/absmax_optimizer.py:29-36) + quantize_symmetric
//                                 (library/quantize.py:51-55) for an axis-0 8-bit weight as ONE launch
//
// All bit-exact with the reference's CPU arithmetic (see quantize_math.cuh); HBM-bound; launched on the caller's stream.
#include <type_traits>

#include "../../include/quanto_b200.h"
#include "common.cuh"
#include "quantize_math.cuh"

namespace qb {

int set_error(int code, const char* fmt, ...);  // api.cu: records the thread's last error text, returns `code`

constexpr int kFzThreads = 256;

static inline int fz_grid(int64_t threads_needed, int ctas_per_sm) {
  int64_t blocks = (threads_needed + kFzThreads - 1) / kFzThreads;
  const int64_t cap = static_cast<int64_t>(kNumSMsB200) * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ---------------------------------------------------------------------------------------------------------------
// quantize_affine (unfused): base viewed as [outer, inner] (already grouped); scale/shift index = row (axis_mode 1),
// column (axis_mode 2) or 0 (axis_mode 0).  shift: T, or uint8 zero-points when ZP.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool ZP, int VEC>
__global__ void __launch_bounds__(kFzThreads)
    quantize_affine_kernel(const T* __restrict__ base, const T* __restrict__ scale, const void* __restrict__ shift,
                           uint8_t* __restrict__ out, int64_t numel, int64_t inner, int axis_mode, float qmax,
                           int zp_signed) {
  const int64_t n_items = (numel + VEC - 1) / VEC;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // zero-points: the reference adds the VALUE of the integer tensor (library/quantize.py:74-76), so an int8 tensor may
  // carry negative zero-points (a group whose minimum is positive) and a uint8 tensor values up to 255
  auto load_z = [&](int64_t i) -> float {
    if constexpr (ZP) {
      const uint8_t b = static_cast<const uint8_t*>(shift)[i];
      return zp_signed ? static_cast<float>(static_cast<int8_t>(b)) : static_cast<float>(b);
    } else {
      return to_float<T>(static_cast<const T*>(shift)[i]);
    }
  };
  for (int64_t it = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; it < n_items; it += stride) {
    const int64_t e0 = it * VEC;
    float f[VEC];
    if constexpr (VEC == 8) load8_stream<T>(base + e0, f);  // host guarantees inner % 8 == 0: one row per vector
    else f[0] = to_float<T>(base[e0]);
    const int64_t i0 = (axis_mode == 1) ? (e0 / inner) : ((axis_mode == 2) ? (e0 % inner) : 0);
    float s = to_float<T>(scale[i0]);
    float z = load_z(i0);
    alignas(8) uint8_t q[VEC];
    bool done = false;
    if constexpr (VEC == 8) {
      if (axis_mode != 2) {  // one scale / shift for the whole vector
        uint32_t b8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (rcp_is_safe<T>(s)) affine_quantize8<T, ZP, true>(f, s, __frcp_rn(s), z, static_cast<int>(qmax), 0, b8);
        else affine_quantize8<T, ZP, false>(f, s, 0.f, z, static_cast<int>(qmax), 0, b8);
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = static_cast<uint8_t>(b8[j]);
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (axis_mode == 2 && j > 0) {
          s = to_float<T>(scale[i0 + j]);
          z = load_z(i0 + j);
        }
        q[j] = static_cast<uint8_t>(affine_quantize_one<T, ZP>(f[j], s, z, qmax));
      }
    }
    if constexpr (VEC == 8) __stcs(reinterpret_cast<uint2*>(out + e0), *reinterpret_cast<uint2*>(q));
    else out[e0] = q[0];
  }
}

template <typename T, bool ZP>
static int launch_qa_t(const void* base, const void* scale, const void* shift, uint8_t* out, int64_t numel,
                       int64_t inner, int axis_mode, float qmax, int zp_signed, cudaStream_t stream) {
  const bool vec = (inner % 8 == 0) && (reinterpret_cast<uintptr_t>(base) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 8 == 0);
  if (vec) {
    quantize_affine_kernel<T, ZP, 8><<<fz_grid(numel / 8, 16), kFzThreads, 0, stream>>>(
        static_cast<const T*>(base), static_cast<const T*>(scale), shift, out, numel, inner, axis_mode, qmax, zp_signed);
  } else {
    quantize_affine_kernel<T, ZP, 1><<<fz_grid(numel, 16), kFzThreads, 0, stream>>>(
        static_cast<const T*>(base), static_cast<const T*>(scale), shift, out, numel, inner, axis_mode, qmax, zp_signed);
  }
  return cudaGetLastError() == cudaSuccess ? OK : ERR_CUDA;
}

template <typename T>
static int launch_qa(const void* base, const void* scale, const void* shift, uint8_t* out, int64_t numel, int64_t inner,
                     int axis_mode, float qmax, int shift_is_int, cudaStream_t stream) {
  const int zp_signed = shift_is_int == 2 ? 1 : 0;
  return shift_is_int != 0 ? launch_qa_t<T, true>(base, scale, shift, out, numel, inner, axis_mode, qmax, zp_signed, stream)
                           : launch_qa_t<T, false>(base, scale, shift, out, numel, inner, axis_mode, qmax, 0, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// pack: out[i] = OR_p (in[p * plane + i] << bits*p) truncated to a byte, i < plane = R * cols, R = ceil(rows / planes);
// source bytes beyond rows*cols count as zero (the reference ORs a shorter last slice, packed.py:64-67).
// No masking of the inputs: like the reference, out-of-range values spill into the higher planes.
// ---------------------------------------------------------------------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(kFzThreads)
    pack_vec_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t plane_vecs, int64_t total_vecs) {
  constexpr int PLANES = 8 / BITS;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane_vecs; i += stride) {
    uint4 v[PLANES];
#pragma unroll
    for (int p = 0; p < PLANES; ++p) {
      const int64_t idx = static_cast<int64_t>(p) * plane_vecs + i;
      v[p] = (idx < total_vecs) ? __ldcs(in + idx) : make_uint4(0u, 0u, 0u, 0u);
    }
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int p = 0; p < PLANES; ++p) {
      const uint32_t keep = ((0xFFu << (BITS * p)) & 0xFFu) * 0x01010101u;  // per-byte wrap of the uint8 shift
      acc.x |= (v[p].x << (BITS * p)) & keep;
      acc.y |= (v[p].y << (BITS * p)) & keep;
      acc.z |= (v[p].z << (BITS * p)) & keep;
      acc.w |= (v[p].w << (BITS * p)) & keep;
    }
    __stcs(out + i, acc);
  }
}

template <int BITS>
__global__ void __launch_bounds__(kFzThreads)
    pack_scalar_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int64_t plane, int64_t total) {
  constexpr int PLANES = 8 / BITS;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane; i += stride) {
    uint32_t acc = 0;
#pragma unroll
    for (int p = 0; p < PLANES; ++p) {
      const int64_t idx = static_cast<int64_t>(p) * plane + i;
      if (idx < total) acc |= (static_cast<uint32_t>(in[idx]) << (BITS * p)) & 0xFFu;
    }
    out[i] = static_cast<uint8_t>(acc);
  }
}

static int launch_pack(const uint8_t* in, uint8_t* out, int64_t rows, int64_t cols, int bits, cudaStream_t stream) {
  const int planes = 8 / bits;
  const int64_t packed_rows = (rows + planes - 1) / planes;
  const int64_t plane = packed_rows * cols, total = rows * cols;
  if (plane == 0) return OK;
  const bool vec = (plane % 16 == 0) && (total % 16 == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if (vec) {
    const int grid = fz_grid(plane / 16, 16);
    if (bits == 4) pack_vec_kernel<4><<<grid, kFzThreads, 0, stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), plane / 16, total / 16);
    else pack_vec_kernel<2><<<grid, kFzThreads, 0, stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), plane / 16, total / 16);
  } else {
    const int grid = fz_grid(plane, 16);
    if (bits == 4) pack_scalar_kernel<4><<<grid, kFzThreads, 0, stream>>>(in, out, plane, total);
    else pack_scalar_kernel<2><<<grid, kFzThreads, 0, stream>>>(in, out, plane, total);
  }
  return cudaGetLastError() == cudaSuccess ? OK : ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// quantize_qbits_max: fused MaxOptimizer + quantize_affine + pack_weights, axis 0.
// The weight [N, K] viewed as grouped rows [rows = N*K/G, G] (tensor/grouped.py:17-30: a pure reshape for axis 0).
// Packed row pr holds grouped rows pr + p * packed_rows, p < 8/BITS, in its bit planes (tensor/packed.py:45-69).
// A team of TL lanes owns one packed row: each lane takes 8 consecutive columns of every plane's row (16-byte loads),
// the team reduces min / max with xor shuffles, every lane derives the row's scale / shift redundantly, quantises its
// 8 columns of each plane and stores 8 packed bytes.  TL = the power of two >= G/8 (lanes past G/8 idle).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct Raw8 {  // 8 elements of T as loaded: one 16-byte vector (fp16/bf16) or two (fp32)
  uint4 v[sizeof(T) == 2 ? 1 : 2];
};

template <typename T>
__device__ __forceinline__ Raw8<T> load_raw8(const T* __restrict__ p) {
  Raw8<T> r;
  r.v[0] = __ldcs(reinterpret_cast<const uint4*>(p));
  if constexpr (sizeof(T) == 4) r.v[1] = __ldcs(reinterpret_cast<const uint4*>(p) + 1);
  return r;
}

template <typename T>
__device__ __forceinline__ void widen8(const Raw8<T>& r, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    const T* e = reinterpret_cast<const T*>(&r.v[0]);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = to_float<T>(e[j]);
  } else {
    f[0] = __uint_as_float(r.v[0].x); f[1] = __uint_as_float(r.v[0].y); f[2] = __uint_as_float(r.v[0].z);
    f[3] = __uint_as_float(r.v[0].w); f[4] = __uint_as_float(r.v[1].x); f[5] = __uint_as_float(r.v[1].y);
    f[6] = __uint_as_float(r.v[1].z); f[7] = __uint_as_float(r.v[1].w);
  }
}

template <typename T, int BITS, bool ZP, int TL>
__global__ void __launch_bounds__(kFzThreads, BITS == 4 ? 4 : 2)
    quantize_qbits_max_kernel(const T* __restrict__ base, uint8_t* __restrict__ packed, T* __restrict__ scale,
                              void* __restrict__ shift, int64_t rows, int64_t packed_rows, int group) {
  constexpr int PLANES = 8 / BITS;
  constexpr float QMAX = static_cast<float>((1 << BITS) - 1);
  constexpr int TEAMS_PER_WARP = 32 / TL;
  const int lane = threadIdx.x & 31;
  const int team_in_warp = lane / TL;
  const int col0 = (lane % TL) * 8;
  const bool lane_active = col0 < group;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * (kFzThreads / 32) + (threadIdx.x >> 5);
  const int64_t step = static_cast<int64_t>(gridDim.x) * (kFzThreads / 32) * TEAMS_PER_WARP;

  // the rows of packed row `pr` this lane reads (zeros where the row does not exist)
  auto fetch = [&](int64_t pr, Raw8<T> (&raw)[PLANES]) {
#pragma unroll
    for (int p = 0; p < PLANES; ++p) {
      const int64_t row = pr + static_cast<int64_t>(p) * packed_rows;
      if (pr < packed_rows && row < rows && lane_active) {
        raw[p] = load_raw8<T>(base + row * group + col0);
      } else {
        raw[p].v[0] = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (sizeof(T) == 4) raw[p].v[1] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };

  // the loop bound is uniform per warp (full-mask shuffles inside); the next packed row's loads are issued before
  // the current one is processed so that each thread keeps 2 x PLANES x 16 B in flight
  int64_t p0 = warp_global * TEAMS_PER_WARP;
  Raw8<T> cur[PLANES];
  if (p0 < packed_rows) fetch(p0 + team_in_warp, cur);
  for (; p0 < packed_rows; p0 += step) {
    const int64_t pr = p0 + team_in_warp;
    const bool team_valid = pr < packed_rows;
    Raw8<T> nxt[PLANES];
    if (p0 + step < packed_rows) fetch(pr + step, nxt);
    constexpr bool PACKED_BF16 = std::is_same<T, __nv_bfloat16>::value;
    uint32_t bytes[8];    // generic path: one quantised value per element, planes or-ed in
    uint32_t acc[4];      // bf16 path: elements (2i, 2i+1) in bytes 0 and 2 of acc[i]
#pragma unroll
    for (int j = 0; j < 8; ++j) bytes[j] = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 0u;
#pragma unroll
    for (int p = 0; p < PLANES; ++p) {
      const int64_t row = pr + static_cast<int64_t>(p) * packed_rows;
      const bool row_valid = team_valid && row < rows;
      float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);  // +inf / -inf
      float f[8];
      if constexpr (PACKED_BF16) {
        // (min, -max) of the lane's 8 values in ONE bf16x2 register: the team reduction is then one shuffle and one
        // packed min per round instead of two of each (max(x) = -min(-x), exact)
        uint32_t red = 0x7F807F80u;  // (+inf, +inf): neutral for inactive lanes / missing rows
        if (row_valid && lane_active) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur[p].v[0]);
          const __nv_bfloat162 mn = __hmin2(__hmin2(h[0], h[1]), __hmin2(h[2], h[3]));
          const __nv_bfloat162 nx = __hneg2(__hmax2(__hmax2(h[0], h[1]), __hmax2(h[2], h[3])));
          const __nv_bfloat162 both = __hmin2(__halves2bfloat162(__low2bfloat16(mn), __low2bfloat16(nx)),
                                              __halves2bfloat162(__high2bfloat16(mn), __high2bfloat16(nx)));
          red = *reinterpret_cast<const uint32_t*>(&both);
        }
#pragma unroll
        for (int off = TL / 2; off > 0; off >>= 1) {
          const uint32_t other = __shfl_xor_sync(0xffffffffu, red, off);
          const __nv_bfloat162 m = __hmin2(*reinterpret_cast<const __nv_bfloat162*>(&red),
                                           *reinterpret_cast<const __nv_bfloat162*>(&other));
          red = *reinterpret_cast<const uint32_t*>(&m);
        }
        lo = __uint_as_float(red << 16);
        hi = -__uint_as_float(red & 0xFFFF0000u);
      } else {
        widen8<T>(cur[p], f);
        if (row_valid && lane_active) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            lo = fminf(lo, f[j]);
            hi = fmaxf(hi, f[j]);
          }
        }
#pragma unroll
        for (int off = TL / 2; off > 0; off >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
          hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        }
      }
      // scale = rnd(rnd(hi - lo) / (2^bits - 1)), shift = -lo          (max_optimizer.py:31-36)
      // bf16: rnd(d * rcp(qmax)) == rnd(d / qmax) for every normal quotient (checked exhaustively over all bf16 d for
      // qmax = 15 and 3; the general argument is in quantize_math.cuh), so the IEEE division runs only for tiny ranges
      const float d = rnd<T>(__fsub_rn(hi, lo));
      float s;
      if (PACKED_BF16 && d > 1e-30f) s = rnd<T>(__fmul_rn(d, 1.0f / QMAX));
      else s = rnd<T>(__fdiv_rn(d, QMAX));
      const bool fast = rcp_is_safe<T>(s);
      const float r = fast ? __frcp_rn(s) : 0.f;
      float z = -lo;
      if constexpr (ZP) {
        // shift = clamp(round(shift / scale), 0, 2^bits - 1) as uint8    (affine_optimizer.py:59-62); NaN -> 0
        const float zq = fast ? rnd<T>(__fmul_rn(z, r)) : rnd<T>(__fdiv_rn(z, s));
        z = fminf(fmaxf(rintf(zq), 0.f), QMAX);
      }
      if (row_valid) {
        if (lane_active) {
          if constexpr (PACKED_BF16) {
            if (fast) {
              // All in packed bf16 except the product with the fp32 reciprocal (quantize_math.cuh explains why each step
              // equals the reference's fp32-then-round sequence): b + z is one exact-then-rounded bf16 add; the quotient
              // is widened, multiplied, rounded back as a pair; clamp o rint = rint o clamp (integer bounds; max before
              // min so that NaN -> lower bound -> 0 like the reference's cast); rint rides on one bf16 add of 192: in
              // [128, 256) the bf16 ulp is 1, so the sum's 7 mantissa bits are 64 + rint(t).  The zero-point is added to
              // that integer afterwards (folding it into the 192 would flip half-to-even ties when zp is odd); both
              // 16-bit lanes at once: every lane holds 64 + n >= 64 - zp, so the subtraction never borrows.
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur[p].v[0]);
              const __nv_bfloat162 zadd = __float2bfloat162_rn(ZP ? 0.f : z);
              const __nv_bfloat162 cl_lo = __float2bfloat162_rn(ZP ? -z : 0.f);
              const __nv_bfloat162 cl_hi = __float2bfloat162_rn(ZP ? QMAX - z : QMAX);
              const __nv_bfloat162 magic = __float2bfloat162_rn(192.f);
              const uint32_t bias2 = (64u - (ZP ? static_cast<uint32_t>(z) : 0u)) * 0x00010001u;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 a = h[i];
                if constexpr (!ZP) a = __hadd2_rn(a, zadd);
                const __nv_bfloat162 t = __floats2bfloat162_rn(__fmul_rn(__low2float(a), r), __fmul_rn(__high2float(a), r));
                const __nv_bfloat162 m = __hadd2_rn(__hmin2(__hmax2(t, cl_lo), cl_hi), magic);
                acc[i] |= ((*reinterpret_cast<const uint32_t*>(&m) & 0x007F007Fu) - bias2) << (BITS * p);
              }
            } else {
              uint32_t tmp[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
              widen8<T>(cur[p], f);
              affine_quantize8<T, ZP, false>(f, s, 0.f, z, static_cast<int>(QMAX), BITS * p, tmp);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] |= tmp[2 * i] | (tmp[2 * i + 1] << 16);
            }
          } else {
            if (fast) affine_quantize8<T, ZP, true>(f, s, r, z, static_cast<int>(QMAX), BITS * p, bytes);
            else affine_quantize8<T, ZP, false>(f, s, 0.f, z, static_cast<int>(QMAX), BITS * p, bytes);
          }
        }
        if ((lane % TL) == 0) {
          scale[row] = from_float<T>(s);
          if constexpr (ZP) static_cast<uint8_t*>(shift)[row] = static_cast<uint8_t>(z);
          else static_cast<T*>(shift)[row] = from_float<T>(z);
        }
      }
    }
    if (team_valid && lane_active) {
      uint2 o;
      if constexpr (PACKED_BF16) {
        o.x = __byte_perm(acc[0], acc[1], 0x6420);
        o.y = __byte_perm(acc[2], acc[3], 0x6420);
      } else {
        o.x = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
        o.y = bytes[4] | (bytes[5] << 8) | (bytes[6] << 16) | (bytes[7] << 24);
      }
      __stcs(reinterpret_cast<uint2*>(packed + pr * group + col0), o);
    }
#pragma unroll
    for (int p = 0; p < PLANES; ++p) cur[p] = nxt[p];
  }
}

template <typename T, int BITS, bool ZP>
static int launch_qqm_tl(const void* base, uint8_t* packed, void* scale, void* shift, int64_t rows, int64_t packed_rows,
                         int group, cudaStream_t stream) {
  const int lanes = group / 8;
  int tl = 1;
  while (tl < lanes) tl <<= 1;
  const int grid = fz_grid(packed_rows * tl, 8);
  const T* b = static_cast<const T*>(base);
  T* s = static_cast<T*>(scale);
#define QB_QQM(TLV) \
  quantize_qbits_max_kernel<T, BITS, ZP, TLV><<<grid, kFzThreads, 0, stream>>>(b, packed, s, shift, rows, packed_rows, group)
  switch (tl) {
    case 1: QB_QQM(1); break;
    case 2: QB_QQM(2); break;
    case 4: QB_QQM(4); break;
    case 8: QB_QQM(8); break;
    case 16: QB_QQM(16); break;
    default: QB_QQM(32); break;
  }
#undef QB_QQM
  return cudaGetLastError() == cudaSuccess ? OK : ERR_CUDA;
}

template <typename T>
static int launch_qqm(const void* base, uint8_t* packed, void* scale, void* shift, int64_t rows, int64_t packed_rows,
                      int group, int bits, bool zp, cudaStream_t stream) {
  if (bits == 4) {
    return zp ? launch_qqm_tl<T, 4, true>(base, packed, scale, shift, rows, packed_rows, group, stream)
              : launch_qqm_tl<T, 4, false>(base, packed, scale, shift, rows, packed_rows, group, stream);
  }
  return zp ? launch_qqm_tl<T, 2, true>(base, packed, scale, shift, rows, packed_rows, group, stream)
            : launch_qqm_tl<T, 2, false>(base, packed, scale, shift, rows, packed_rows, group, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// absmax (per tensor): out[0] = max |base| as float32.  Non-negative floats order like their bit patterns, so the
// cross-CTA step is one atomicMax on the int view; `out` is zeroed on the stream by the launcher.
// ---------------------------------------------------------------------------------------------------------------
// max that PROPAGATES NaN (max.NaN.f32), as torch.max / amax do (calibrate.py:54-57, absmax_optimizer.py:29-36): a NaN
// activation or weight yields a NaN scale in the reference, and must here too (fmaxf would silently drop it)
__device__ __forceinline__ float nanmax(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

__device__ __forceinline__ float block_max_256(float m, float* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = nanmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  __syncthreads();  // red[] may still be read by the previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kFzThreads / 32; ++w) r = nanmax(r, red[w]);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kFzThreads)
    absmax_kernel(const T* __restrict__ base, T* __restrict__ out, int* __restrict__ scratch, int64_t numel, int vec_ok) {
  __shared__ float red[kFzThreads / 32];
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float m = 0.f;
  if (vec_ok) {
    const int64_t n_vec = numel / 8;
    int64_t v = tid;
    for (; v + stride < n_vec; v += 2 * stride) {  // two independent 16-byte loads in flight per thread
      float f[8], g[8];
      load8_stream<T>(base + v * 8, f);
      load8_stream<T>(base + (v + stride) * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = nanmax(m, nanmax(fabsf(f[j]), fabsf(g[j])));
    }
    if (v < n_vec) {
      float f[8];
      load8_stream<T>(base + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = nanmax(m, fabsf(f[j]));
    }
    for (int64_t i = n_vec * 8 + tid; i < numel; i += stride) m = nanmax(m, fabsf(to_float<T>(base[i])));
  } else {
    for (int64_t i = tid; i < numel; i += stride) m = nanmax(m, fabsf(to_float<T>(base[i])));
  }
  m = block_max_256(m, red);
  if (threadIdx.x == 0) {
    // scratch[0] = running maximum (bit pattern of a non-negative float; a NaN's pattern is the largest of all, so the
    // integer atomicMax propagates it), scratch[1] = CTAs done; the last CTA
    // publishes the result in T, so the caller needs no conversion launch
    atomicMax(scratch, __float_as_int(m));
    __threadfence();
    const int done = atomicAdd(scratch + 1, 1);
    if (done == static_cast<int>(gridDim.x) - 1) {
      __threadfence();
      out[0] = from_float<T>(__int_as_float(atomicMax(scratch, 0)));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// quantize_qbytes_absmax: per out-feature row, scale = rnd(max|w| / qmax), data = quantize_symmetric(w, scale).
// One CTA per row (grid-stride over rows); the second pass re-reads the row the CTA has just streamed (L1/L2 hit).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int OUT_DT>
__global__ void __launch_bounds__(kFzThreads)
    quantize_qbytes_absmax_kernel(const T* __restrict__ base, uint8_t* __restrict__ out, T* __restrict__ scale,
                                  int64_t n_rows, int64_t k, float qmax, int vec_ok) {
  __shared__ float red[kFzThreads / 32];
  const int tid = threadIdx.x;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* src = base + row * k;
    uint8_t* dst = out + row * k;
    float m = 0.f;
    if (vec_ok) {
      for (int64_t v = tid; v < k / 8; v += kFzThreads) {
        const uint4 r = *reinterpret_cast<const uint4*>(src + v * 8);  // default caching: read again below
        float f[8];
        if constexpr (sizeof(T) == 2) {
          const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = to_float<T>(e[j]);
        } else {
          const uint4 r2 = *(reinterpret_cast<const uint4*>(src + v * 8) + 1);
          f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
          f[4] = __uint_as_float(r2.x); f[5] = __uint_as_float(r2.y); f[6] = __uint_as_float(r2.z); f[7] = __uint_as_float(r2.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) m = nanmax(m, fabsf(f[j]));
      }
    } else {
      for (int64_t i = tid; i < k; i += kFzThreads) m = nanmax(m, fabsf(to_float<T>(src[i])));
    }
    m = block_max_256(m, red);
    const float s = rnd<T>(__fdiv_rn(m, qmax));  // absmax_optimizer.py:36: rmax / qtype.qmax, rounded to T
    if (tid == 0) scale[row] = from_float<T>(s);
    const bool fast = rcp_is_safe<T>(s);  // bf16: x * rcp(s) rounds to the same bf16 as x / s (quantize_math.cuh)
    const float r = fast ? __frcp_rn(s) : 0.f;
    if (vec_ok) {
      for (int64_t v = tid; v < k / 8; v += kFzThreads) {
        float f[8];
        load8_stream<T>(src + v * 8, f);
        alignas(8) uint8_t q[8];
        if (fast) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __fmul_rn(f[j], r);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __fdiv_rn(f[j], s);
        }
#pragma unroll
        for (int j = 0; j < 8; j += 2) rnd_pair<T>(f[j], f[j + 1]);
        if constexpr (sizeof(T) == 4) { /* rnd_pair is the identity for fp32 */ }
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = quantize_one<OUT_DT>(f[j]);
        __stcs(reinterpret_cast<uint2*>(dst + v * 8), *reinterpret_cast<uint2*>(q));
      }
    } else {
      for (int64_t i = tid; i < k; i += kFzThreads) dst[i] = quantize_one<OUT_DT>(rnd<T>(__fdiv_rn(to_float<T>(src[i]), s)));
    }
  }
}

template <typename T>
static int launch_qqa(const void* base, void* out, void* scale, int64_t n, int64_t k, int out_dt, cudaStream_t stream) {
  const int vec_ok = (k % 8 == 0) && (reinterpret_cast<uintptr_t>(base) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % 8 == 0);
  const int grid = static_cast<int>(n < static_cast<int64_t>(kNumSMsB200) * 8 ? n : static_cast<int64_t>(kNumSMsB200) * 8);
  const T* b = static_cast<const T*>(base);
  uint8_t* o = static_cast<uint8_t*>(out);
  T* s = static_cast<T*>(scale);
  switch (out_dt) {
    case DT_I8: quantize_qbytes_absmax_kernel<T, DT_I8><<<grid, kFzThreads, 0, stream>>>(b, o, s, n, k, 127.f, vec_ok); break;
    case DT_E4M3: quantize_qbytes_absmax_kernel<T, DT_E4M3><<<grid, kFzThreads, 0, stream>>>(b, o, s, n, k, 448.f, vec_ok); break;
    case DT_E5M2: quantize_qbytes_absmax_kernel<T, DT_E5M2><<<grid, kFzThreads, 0, stream>>>(b, o, s, n, k, 57344.f, vec_ok); break;
    default: return ERR_ARG;
  }
  return cudaGetLastError() == cudaSuccess ? OK : ERR_CUDA;
}

}  // namespace qb

using namespace qb;

extern "C" {

int qb200_quantize_affine(const void* base, const void* scale, const void* shift, uint8_t* out, int64_t outer,
                          int64_t inner, int axis_mode, int bits, int dtype, int shift_is_int, void* stream) {
  if (bits < 1 || bits > 8) return set_error(ERR_ARG, "quantize_affine: bits must be in 1..8, got %d", bits);
  if (outer < 0 || inner < 0 || axis_mode < 0 || axis_mode > 2)
    return set_error(ERR_ARG, "quantize_affine: bad shape / axis_mode (%lld x %lld, mode %d)", (long long)outer,
                     (long long)inner, axis_mode);
  const int64_t numel = outer * inner;
  if (numel == 0) return OK;
  if (!base || !scale || !shift || !out) return set_error(ERR_ARG, "quantize_affine: null buffer");
  const float qmax = static_cast<float>((1 << bits) - 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  switch (dtype) {
    case DT_F32: rc = launch_qa<float>(base, scale, shift, out, numel, inner, axis_mode, qmax, shift_is_int, st); break;
    case DT_F16: rc = launch_qa<__half>(base, scale, shift, out, numel, inner, axis_mode, qmax, shift_is_int, st); break;
    case DT_BF16: rc = launch_qa<__nv_bfloat16>(base, scale, shift, out, numel, inner, axis_mode, qmax, shift_is_int, st); break;
    default: return set_error(ERR_ARG, "quantize_affine: dtype %d not floating point", dtype);
  }
  return rc == OK ? OK : set_error(rc, "quantize_affine: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
}

int qb200_pack(const uint8_t* in, uint8_t* out, int64_t rows, int64_t cols, int bits, void* stream) {
  if (bits != 2 && bits != 4) return set_error(ERR_ARG, "pack: bits must be 2 or 4, got %d", bits);
  if (rows < 0 || cols < 0) return set_error(ERR_ARG, "pack: negative shape");
  if (rows * cols > 0 && (!in || !out)) return set_error(ERR_ARG, "pack: null buffer");
  int rc = launch_pack(in, out, rows, cols, bits, static_cast<cudaStream_t>(stream));
  return rc == OK ? OK : set_error(rc, "pack: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
}

int qb200_quantize_qbits_max(const void* base, uint8_t* packed, void* scale, void* shift, int64_t n, int64_t k,
                             int group, int bits, int dtype, int zeropoint, void* stream) {
  if (bits != 2 && bits != 4) return set_error(ERR_ARG, "quantize_qbits_max: bits must be 2 or 4, got %d", bits);
  if (n <= 0 || k <= 0 || group <= 0) return set_error(ERR_ARG, "quantize_qbits_max: bad shape");
  if (k % group != 0) return set_error(ERR_ARG, "quantize_qbits_max: group %d does not divide K=%lld", group, (long long)k);
  if (group % 8 != 0 || group > 256)
    return set_error(ERR_UNSUPPORTED, "quantize_qbits_max: group %d (needs a multiple of 8, <= 256)", group);
  if (!base || !packed || !scale || !shift) return set_error(ERR_ARG, "quantize_qbits_max: null buffer");
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0 || reinterpret_cast<uintptr_t>(packed) % 8 != 0)
    return set_error(ERR_UNSUPPORTED, "quantize_qbits_max: base must be 16-byte and packed 8-byte aligned");
  const int64_t rows = n * k / group;
  const int planes = 8 / bits;
  const int64_t packed_rows = (rows + planes - 1) / planes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  switch (dtype) {
    case DT_F32: rc = launch_qqm<float>(base, packed, scale, shift, rows, packed_rows, group, bits, zeropoint != 0, st); break;
    case DT_F16: rc = launch_qqm<__half>(base, packed, scale, shift, rows, packed_rows, group, bits, zeropoint != 0, st); break;
    case DT_BF16: rc = launch_qqm<__nv_bfloat16>(base, packed, scale, shift, rows, packed_rows, group, bits, zeropoint != 0, st); break;
    default: return set_error(ERR_ARG, "quantize_qbits_max: dtype %d not floating point", dtype);
  }
  return rc == OK ? OK : set_error(rc, "quantize_qbits_max: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
}

int qb200_absmax(const void* base, void* out, void* scratch, int64_t numel, int dtype, void* stream) {
  if (numel < 0 || !out || !scratch || (numel > 0 && !base)) return set_error(ERR_ARG, "absmax: bad buffer");
  if (reinterpret_cast<uintptr_t>(scratch) % 4 != 0) return set_error(ERR_ARG, "absmax: scratch must be 4-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(scratch, 0, 8, st) != cudaSuccess) return set_error(ERR_CUDA, "absmax: memset failed");
  const int vec_ok = reinterpret_cast<uintptr_t>(base) % 16 == 0;
  const int grid = numel == 0 ? 1 : fz_grid((numel + 15) / 16, 8);
  int* sc = static_cast<int*>(scratch);
  switch (dtype) {
    case DT_F32: absmax_kernel<float><<<grid, kFzThreads, 0, st>>>(static_cast<const float*>(base), static_cast<float*>(out), sc, numel, vec_ok); break;
    case DT_F16: absmax_kernel<__half><<<grid, kFzThreads, 0, st>>>(static_cast<const __half*>(base), static_cast<__half*>(out), sc, numel, vec_ok); break;
    case DT_BF16: absmax_kernel<__nv_bfloat16><<<grid, kFzThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(base), static_cast<__nv_bfloat16*>(out), sc, numel, vec_ok); break;
    default: return set_error(ERR_ARG, "absmax: dtype %d not floating point", dtype);
  }
  return cudaGetLastError() == cudaSuccess ? OK : set_error(ERR_CUDA, "absmax: launch failed");
}

int qb200_quantize_qbytes_absmax(const void* base, void* out, void* scale, int64_t n, int64_t k, int dtype,
                                 int out_dtype, void* stream) {
  if (n < 0 || k < 0) return set_error(ERR_ARG, "quantize_qbytes_absmax: negative shape");
  if (n == 0 || k == 0) return OK;
  if (!base || !out || !scale) return set_error(ERR_ARG, "quantize_qbytes_absmax: null buffer");
  if (out_dtype != DT_I8 && out_dtype != DT_E4M3 && out_dtype != DT_E5M2)
    return set_error(ERR_ARG, "quantize_qbytes_absmax: unsupported target dtype %d", out_dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  switch (dtype) {
    case DT_F32: rc = launch_qqa<float>(base, out, scale, n, k, out_dtype, st); break;
    case DT_F16: rc = launch_qqa<__half>(base, out, scale, n, k, out_dtype, st); break;
    case DT_BF16: rc = launch_qqa<__nv_bfloat16>(base, out, scale, n, k, out_dtype, st); break;
    default: return set_error(ERR_ARG, "quantize_qbytes_absmax: dtype %d not floating point", dtype);
  }
  return rc == OK ? OK : set_error(rc, "quantize_qbytes_absmax: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
}

}  // extern "C"
