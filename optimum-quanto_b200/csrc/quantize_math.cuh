// Exact (bit-for-bit with the reference's CPU arithmetic) scalar building blocks shared by the quantisation kernels
// of elementwise.cu and freeze.cu.  Every step uses the *_rn intrinsics so that ptxas cannot contract an add/mul
// pair into an fma or replace a division by a reciprocal multiply.
#pragma once

#include "common.cuh"

namespace qb {

// round a float to T (nearest even) and widen it again: the value an ATen op on dtype T hands to the next op
template <typename T>
__device__ __forceinline__ float rnd(float v) {
  return to_float<T>(from_float<T>(v));
}

// clamp + cast of quantize_symmetric (optimum/quanto/library/quantize.py:51-55); t is already rounded to the input dtype
template <int OUT_DT>
__device__ __forceinline__ uint8_t quantize_one(float t) {
  if constexpr (OUT_DT == DT_I8) {
    // cvt.rni.s32.f32: half-to-even like torch.round, saturating, NaN -> 0 (what the reference's NaN -> int8 cast gives
    // on x86 and on CUDA: an all-zero row has scale 0 and quotient 0/0); the clamp then runs on integers
    const int v = max(-128, min(127, __float2int_rn(t)));
    return static_cast<uint8_t>(static_cast<int8_t>(v));
  } else if constexpr (OUT_DT == DT_E4M3) {
    if (t != t) return static_cast<uint8_t>(0x7Fu | ((__float_as_uint(t) >> 24) & 0x80u));  // NaN stays NaN (torch)
    float c = fminf(fmaxf(t, -448.f), 448.f);
    return static_cast<uint8_t>(__nv_cvt_float_to_fp8(c, __NV_SATFINITE, __NV_E4M3));
  } else {
    if (t != t) return static_cast<uint8_t>(0x7Fu | ((__float_as_uint(t) >> 24) & 0x80u));
    float c = fminf(fmaxf(t, -57344.f), 57344.f);
    return static_cast<uint8_t>(__nv_cvt_float_to_fp8(c, __NV_SATFINITE, __NV_E5M2));
  }
}

// quanto::quantize_affine for one element (optimum/quanto/library/quantize.py:71-78), arithmetic in dtype T:
//   float shift : clamp(rint(rnd(rnd(b + z) / s)), 0, qmax)
//   zero-point  : clamp(rnd(rint(rnd(b / s)) + zp), 0, qmax)
// NaN (0/0 of a constant group) maps to 0, what the reference's NaN -> uint8 cast yields on x86.
template <typename T, bool ZP>
__device__ __forceinline__ uint32_t affine_quantize_one(float b, float s, float z, float qmax) {
  float r;
  if constexpr (ZP) r = rnd<T>(__fadd_rn(rintf(rnd<T>(__fdiv_rn(b, s))), z));
  else r = rintf(rnd<T>(__fdiv_rn(rnd<T>(__fadd_rn(b, z)), s)));
  r = fminf(fmaxf(r, 0.f), qmax);  // fmaxf returns the non-NaN operand
  return static_cast<uint32_t>(r);
}

// 8 consecutive elements of T (16-byte aligned address) widened to float, streaming load
template <typename T>
__device__ __forceinline__ void load8_stream(const T* __restrict__ p, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    const uint4 r = __ldcs(reinterpret_cast<const uint4*>(p));
    const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = to_float<T>(e[j]);
  } else {
    const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldcs(reinterpret_cast<const uint4*>(p) + 1);
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
  }
}

}  // namespace qb
