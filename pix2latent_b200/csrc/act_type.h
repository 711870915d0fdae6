// Storage / tensor-core operand type of activations, gradients and packed weights.
//   P2L_ACT_FP16 = 1 (default): IEEE fp16 — 10-bit mantissa, 8x less rounding error per operand than
//       bf16 at the same tcgen05 rate and the same bytes. Its narrow exponent range is handled by a
//       static power-of-two gradient scale (kGradScale) applied where the backward pass enters 16-bit
//       storage and removed where it leaves it (the backward pass is linear in the upstream gradient).
//   P2L_ACT_FP16 = 0: bfloat16.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#ifndef P2L_ACT_FP16
#define P2L_ACT_FP16 1
#endif

namespace p2l {

#if P2L_ACT_FP16
typedef __half act_t;
#define P2L_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
constexpr uint32_t kUmmaFmt = 0;          // tcgen05 instruction descriptor a/b format: F16
constexpr float kGradScale = 4096.f;
// Conversions to fp16 saturate to +-65504 on the device (one F2FP.SATFINITE) instead of producing inf:
// an out-of-range activation then costs accuracy in one element rather than poisoning the step.
__host__ __device__ inline act_t f2a(float x) {
#ifdef __CUDA_ARCH__
    unsigned short s;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(s) : "f"(x));
    return __ushort_as_half(s);
#else
    return __float2half_rn(x);
#endif
}
__device__ __forceinline__ float a2f(act_t x) { return __half2float(x); }
__device__ __forceinline__ float act_lo(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xFFFFu))); }
__device__ __forceinline__ float act_hi(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))); }
__device__ __forceinline__ uint32_t pack_act(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
#else
typedef __nv_bfloat16 act_t;
#define P2L_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
constexpr uint32_t kUmmaFmt = 1;          // BF16
constexpr float kGradScale = 1.f;
__host__ __device__ inline act_t f2a(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float a2f(act_t x) { return __bfloat162float(x); }
__device__ __forceinline__ float act_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float act_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_act(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
}
#endif

}  // namespace p2l
