// Scalar numeric helpers shared by the tensor-core kernels and the bandwidth-bound kernels: the bf16 hi/lo split of
// the error-compensated operand mode, the MUFU approximations and the exact-erf GELU.  This header is also compiled
// by the host compiler for the CUDA-on-CPU emulation the tests use (tests/cuda_emu, -DEB_HOST_EMU), where the
// inline-PTX forms are replaced by bit-exact C++ (cvt.rn.bf16x2) or libm (ex2 / rcp) equivalents.
#pragma once
#ifdef EB_HOST_EMU
#include "cuda_emu.h"
#else
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace eb {

#ifndef EB_HOST_EMU
// ---------------------------------------------------------------------------------------------
// bf16 hi/lo split (error-compensated operands): x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return uint32_t(__bfloat16_as_ushort(a)) | (uint32_t(__bfloat16_as_ushort(b)) << 16);
}
// two floats -> packed bf16x2 (x0 in the low half), round to nearest even: one F2FP instruction
__device__ __forceinline__ uint32_t cvt_bf16x2(float x0, float x1) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x1), "f"(x0));
  return d;
}
// packed hi/lo split of two floats: 6 instructions per pair
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_bf16x2(x0, x1);
  lo = cvt_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2, 2 ulp
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, 1 ulp
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#else
// ---- host emulation of the helpers above (round-to-nearest-even bf16 conversion, as cvt.rn.bf16x2.f32)
inline uint32_t emu_bf16_bits(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;           // NaN
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
inline float emu_bf16_float(uint32_t b) {
  const uint32_t u = b << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}
inline uint32_t cvt_bf16x2(float x0, float x1) { return emu_bf16_bits(x0) | (emu_bf16_bits(x1) << 16); }
inline void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  const uint16_t h = uint16_t(emu_bf16_bits(x)), l = uint16_t(emu_bf16_bits(x - emu_bf16_float(h)));
  memcpy(&hi, &h, 2);
  memcpy(&lo, &l, 2);
}
inline uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  uint16_t x, y;
  memcpy(&x, &a, 2);
  memcpy(&y, &b, 2);
  return uint32_t(x) | (uint32_t(y) << 16);
}
inline void split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_bf16x2(x0, x1);
  lo = cvt_bf16x2(x0 - emu_bf16_float(hi & 0xffffu), x1 - emu_bf16_float(hi >> 16));
}
inline float ex2_approx(float x) { return exp2f(x); }
inline float rcp_approx(float x) { return 1.0f / x; }
#endif

// Exact-erf GELU, x * Phi(x), branch-free: Phi through the complementary error function of |x|/sqrt(2) in the
// Abramowitz-Stegun 7.1.26 form (|erf error| <= 1.5e-7, i.e. fp32 rounding level; measured max |gelu error| 3.3e-7
// over [-12, 12] against fp64, the same as an fp32 evaluation of 0.5*x*(1+erf(x/sqrt 2))).
//   gelu(x) = max(x, 0) - |x| * h(|x|),   h(a) = 0.5 erfc(a / sqrt 2) = (0.5 p(t)) t exp(-a^2 / 2),  t = 1 / (1 + 0.3275911 a / sqrt 2)
// with s = a sqrt(log2(e) / 2) so that exp(-a^2 / 2) = 2^(-s^2): 11 FMA-pipe instructions + 2 MUFU and no divergence (the first
// version spent 17 + 2: the GELU epilogue of the MLP-up GEMM is bound by its own instruction stream, profiles/r02e_epilogue_warps.md).
// Reference: ACT2FN['gelu'], modeling_vit.py:326.
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fabsf(x);
  const float s = a * 0.84932180028801904f;                       // sqrt(log2(e) / 2)
  const float t = rcp_approx(fmaf(0.27273748087922248f, s, 1.0f));   // 0.3275911 / sqrt(log2(e))
  float p = fmaf(0.5307027145f, t, -0.7265760135f);               // the 7.1.26 coefficients, halved
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float te = t * ex2_approx(-s * s);
  return fmaf(-a, p * te, fmaxf(x, 0.0f));
}

}  // namespace eb
