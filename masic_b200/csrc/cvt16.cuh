// cvt16.cuh — the two 16-bit activation / operand formats of the library.
//
// MASIC_FMT_BF16 (0): bfloat16 — fp32's exponent range, 8 significant bits.  The TRAINING step uses it: gradients span
//                     far more than fp16's range.
// MASIC_FMT_F16  (1): IEEE half — 11 significant bits (8x finer rounding), range +-65504, conversions saturate instead
//                     of overflowing.  The INFERENCE engines use it: the codec's deviation from the fp32 reference is
//                     dominated by latent symbols that round the other way (|y - k - 0.5| below the encoder's noise),
//                     and fp16 operands cut that noise 8x at the same tensor-core rate (kind::f16 takes either format).
// MASIC_FMT_SPLIT (2, or-ed in): image producers write every pixel of <= 4 channels as [hi(c) | lo(c)] with
//                     lo = x - float(hi): the 3-channel images feeding g_a_conv1 then enter the tensor cores with ~2x the
//                     significant bits at no cost (the 8-channel pixel pitch has room, the packed weights repeat).
// Both are 16 bits wide, so buffers, pitches, TMA boxes and shared-memory layouts are identical; only the conversion
// instructions and the a/b format fields of the UMMA instruction descriptor differ.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/masic_b200.h"

namespace masic {

__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, int f16) {      // {lo, hi} -> one 32-bit word
  uint32_t r;
  if (f16 & 1) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint16_t pack16(float v, int f16) {
  return static_cast<uint16_t>(pack16x2(v, 0.0f, f16) & 0xFFFFu);
}
__device__ __forceinline__ float2 unpack16x2(uint32_t w, int f16) {
  if (f16 & 1) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ float unpack16(uint16_t h, int f16) {
  return unpack16x2(static_cast<uint32_t>(h), f16).x;
}
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b, int f16) {
  if (f16 & 1) {
    __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

}  // namespace masic
