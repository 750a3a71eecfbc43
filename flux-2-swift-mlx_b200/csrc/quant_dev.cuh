// quant_dev.cuh — device-side format conversions shared by the weight packers (quant.cu) and by every kernel that emits
// block-scaled activations (the standalone quantiser, LayerNorm + modulate, the SwiGLU GEMM epilogue).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_fp4.h>
#include <stdint.h>

namespace f2b {

// ---- scalar format conversions (integer / compare logic only: deterministic everywhere)
// round(log2(x)) for x > 0 without libm: exponent + (mantissa >= sqrt(2))
__device__ __forceinline__ int round_log2_pos(float x) {
  uint32_t u = __float_as_uint(x);
  int e = (int)((u >> 23) & 0xff);
  uint32_t m = u & 0x7fffff;
  if (e == 0) {  // subnormal: below 2^-126, clamps to -127 anyway
    return -127;
  }
  return (e - 127) + (m >= 0x3504F4u ? 1 : 0);
}
__device__ __forceinline__ uint8_t to_e8m0(float x) {
  if (!(x > 0.f)) return 0;  // zero / negative / NaN -> smallest scale (NaN cannot occur for finite weights)
  if (isinf(x)) return 0xFF;
  int n = round_log2_pos(x);
  n = n < -127 ? -127 : n;
  n = n > 127 ? 127 : n;
  return (uint8_t)(n + 127);
}
__device__ __forceinline__ float from_e8m0(uint8_t b) {
  // 2^(b-127); b = 0 -> 2^-127 (subnormal), b = 255 treated as 2^128 -> inf
  if (b == 0) return __uint_as_float(0x00400000u);
  if (b == 255) return __uint_as_float(0x7f800000u);
  return __uint_as_float((uint32_t)b << 23);
}
// fp32 -> E4M3 (fn: no inf, max 448), round-to-nearest-even, saturating; sign kept
__device__ __forceinline__ uint8_t to_e4m3(float x) {
  uint32_t u = __float_as_uint(x);
  uint8_t sign = (u >> 31) ? 0x80 : 0;
  float a = fabsf(x);
  if (a != a) return sign | 0x7F;
  if (a >= 448.f) return sign | 0x7E;  // saturate (covers inf)
  if (a < 0.015625f) {
    // subnormal range: multiples of 2^-9, RNE
    float q = rintf(__fmul_rn(a, 512.f));  // exact scaling by a power of two
    return sign | (uint8_t)q;              // q in [0,8]; 8 == smallest normal 0x08
  }
  uint32_t au = __float_as_uint(a);
  int e = (int)(au >> 23) - 127;   // [-6, 8]
  uint32_t m = au & 0x7fffff;
  uint32_t keep = m >> 20;         // 3 mantissa bits
  uint32_t rem = m & 0xfffff;
  uint32_t half = 0x80000;
  if (rem > half || (rem == half && (keep & 1))) ++keep;
  if (keep == 8) { keep = 0; ++e; }
  uint32_t code = ((uint32_t)(e + 7) << 3) | keep;
  if (code > 0x7E) code = 0x7E;
  return sign | (uint8_t)code;
}
__device__ __forceinline__ float from_e4m3(uint8_t b) {
  const float sgn = (b & 0x80) ? -1.f : 1.f;
  const int e = (b >> 3) & 0xF;
  const int m = b & 7;
  if (e == 0) return sgn * (float)m * 0.001953125f;  // m * 2^-9
  if (e == 15 && m == 7) return __uint_as_float(0x7fc00000u);
  return sgn * __uint_as_float((uint32_t)(e - 7 + 127) << 23) * (1.f + (float)m * 0.125f);
}
__device__ __forceinline__ uint8_t to_e2m1(float x) {
  const uint8_t sign = (__float_as_uint(x) >> 31) ? 0x8 : 0x0;
  const float a = fabsf(x);
  uint8_t b;
  if (a != a) b = 0x7;
  else if (a > 5.0f) b = 0x7;
  else if (a >= 3.5f) b = 0x6;
  else if (a > 2.5f) b = 0x5;
  else if (a >= 1.75f) b = 0x4;
  else if (a > 1.25f) b = 0x3;
  else if (a >= 0.75f) b = 0x2;
  else if (a > 0.25f) b = 0x1;
  else b = 0x0;
  return b | sign;
}
__device__ __forceinline__ float from_e2m1(uint8_t b) {
  const float tab[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
  const float v = tab[b & 7];
  return (b & 8) ? -v : v;
}


// ---- tcgen05 scale-factor layout (quant.cuh): byte offset of group g of `row` with ld_blocks 512 B blocks per 128-row block
__device__ __forceinline__ int64_t sf_offset(int64_t row, int64_t g, int64_t ld_blocks) {
  return ((row >> 7) * ld_blocks + (g >> 2)) * 512 + (row & 31) * 16 + ((row & 127) >> 5) * 4 + (g & 3);
}
__device__ __forceinline__ void unpack8(const uint4 raw, bool f16, float (&v)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = f16 ? __half22float2(*reinterpret_cast<const __half2*>(&w[j]))
                         : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
}

// ---- on-the-fly block-scaled activations. KIND: 1 = mxfp8, 2 = mxfp4, 3 = nvfp4.
//   mxfp8: scale = 2^ceil(log2(amax / 448)) (nothing saturates; an activation is quantised once and consumed at once, so it
//          need not follow the weight packer's round-to-nearest exponent rule), elements by cvt.rn.satfinite.e4m3x2.
//   mxfp4 / nvfp4: the weight packer's rule (amax / 6 -> E8M0 / E4M3 scale, x / scale -> E2M1 RNE, saturating), evaluated as
//          cvt.rn.satfinite.e2m1x2(f16(x * (1 / scale))). That is bit-identical to quantize_kernel / the C oracle for 16-bit
//          inputs: x (<= 11 significant bits) and threshold * scale (<= 7 bits) are dyadic rationals that either coincide or
//          differ by >= 2^-12 relative, the product carries <= 2^-22 error and the f16 rounding snaps an exact tie back onto
//          its threshold, where RNE picks the even code exactly like to_e2m1() above.
struct MxScale {
  uint32_t sb;  // the stored scale byte (E8M0 or E4M3)
  float mul;    // multiply the elements by this before the element conversion (1 / scale; 0 when the scale is 0)
};
template <int KIND>
__device__ __forceinline__ MxScale mx_scale(float amax) {
  MxScale r;
  if constexpr (KIND == 1) {
    int e = -127;  // smallest power of two s with amax / s <= 448 (exponent arithmetic only)
    if (amax > 0.f) {
      const uint32_t u = __float_as_uint(amax * (1.0f / 448.0f));
      e = (int)((u >> 23) & 0xff) - 127 + ((u & 0x7fffff) ? 1 : 0);
      e = e < -127 ? -127 : (e > 127 ? 127 : e);
    }
    const uint32_t ebits = (uint32_t)(127 - e);                         // biased exponent of 2^-e, in [0, 254]
    r.mul = __uint_as_float(ebits ? (ebits << 23) : 0x00400000u);       // 2^-e (2^-127 is subnormal)
    r.sb = (uint32_t)(e + 127);
  } else {
    // amax / 6 without the IEEE division: amax * fl(1/6) is within 2^-22 of the quotient, and rounding that to 17
    // significant bits reproduces every case the scale conversion can distinguish — a 16-bit amax (<= 11 significant bits)
    // either hits a rounding boundary 6 * m of the scale format exactly (then the snapped value is exactly m and RNE / the
    // sqrt(2) test see the same tie as the packer) or misses it by >= 2^-12 relative.
    float scale = amax * 0.16666667f;
    scale = __uint_as_float((__float_as_uint(scale) + 0x40u) & 0xFFFFFF80u);
    if constexpr (KIND == 3) {
      // E4M3, RNE, saturating: the hardware conversion equals to_e4m3() for finite non-negative input
      r.sb = (uint32_t)__nv_cvt_float_to_fp8(scale, __NV_SATFINITE, __NV_E4M3);
      const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)r.sb, __NV_E4M3);
      scale = __half2float(*reinterpret_cast<const __half*>(&h));
      float rcp;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(scale));   // scale is 0 or >= 2^-9: <= 1 ulp, no flush
      r.mul = (scale != 0.f) ? rcp : 0.f;
    } else {
      r.sb = to_e8m0(scale);
      r.mul = __uint_as_float((254u - r.sb) << 23);   // exactly 1 / 2^(sb - 127); sb = 254 -> 0 (quotients underflow to +-0)
    }
  }
  return r;
}
// four consecutive elements -> 4 bytes of E4M3 (KIND 1) or 2 bytes of E2M1 nibbles (KIND 2 / 3), first element lowest
template <int KIND>
__device__ __forceinline__ uint32_t mx_pack4(float a, float b, float c, float d, float mul) {
  if constexpr (KIND == 1) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a * mul, b * mul), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c * mul, d * mul), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
  } else {
    if (mul == 0.f) return 0u;  // zero scale: the packer stores +0 whatever the sign
    const float2 q0 = __half22float2(__floats2half2_rn(a * mul, b * mul));
    const float2 q1 = __half22float2(__floats2half2_rn(c * mul, d * mul));
    const uint32_t lo = __nv_cvt_float2_to_fp4x2(q0, __NV_E2M1, cudaRoundNearest) & 0xffu;
    const uint32_t hi = __nv_cvt_float2_to_fp4x2(q1, __NV_E2M1, cudaRoundNearest) & 0xffu;
    return lo | (hi << 8);
  }
}
__host__ __device__ inline uint8_t mx_scale_one(int kind) { return kind == 3 ? 0x38 : 127; }  // the scale byte of 1.0

}  // namespace f2b
