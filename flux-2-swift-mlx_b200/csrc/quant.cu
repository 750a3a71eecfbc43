// quant.cu — bit-exact weight quantizers / dequantizers for the five reference modes.
//
// Reference call sites: quantize(model:groupSize:bits:mode:) Flux2Pipeline.swift:567-578,
// quantized()/dequantized() WeightLoader.swift:795-815, level table QuantizationConfig.swift:51-60. The arithmetic
// itself lives in mlx-swift 0.31.6 (not in the tree); the rules restated here are the published MLX semantics:
//   groups run along the input dim of W[out,in]; element j of a uint32 word sits at bits [j*bits, (j+1)*bits).
//   affine (qint8 8/64, int4 4/64): scale/bias from the group min/max with the "edge" rule, scales/biases f16.
//   mxfp8 (8/32): E4M3 elements, E8M0 scale = 2^round(log2(amax/448)); mxfp4 (4/32): E2M1 elements, amax/6;
//   nvfp4 (4/16): E2M1 elements, scale amax/6 stored as E4M3 (no global scale). RNE, saturating.
// All fp32 steps use explicit _rn intrinsics so that no FMA contraction can make the device disagree with the
// C oracle (oracle/quant_oracle.c) by an ulp.
#include "quant.cuh"
#include <cuda_fp8.h>
#include <algorithm>
#include "ptx.cuh"

namespace f2b {

struct QSpec { int bits, group, mode; };  // mode: 0 affine, 1 mx (E8M0 scale), 2 nv (E4M3 scale)
__host__ __device__ inline QSpec qspec(int quant) {
  switch (quant) {
    case 1: return {8, 64, 0};
    case 2: return {4, 64, 0};
    case 3: return {8, 32, 1};
    case 4: return {4, 32, 1};
    case 5: return {4, 16, 2};
    default: return {16, 0, -1};
  }
}
bool quant_params(int quant, int* bits, int* group, int* has_biases, int* scale_dtype) {
  QSpec q = qspec(quant);
  if (q.mode < 0) return false;
  if (bits) *bits = q.bits;
  if (group) *group = q.group;
  if (has_biases) *has_biases = q.mode == 0;
  if (scale_dtype) *scale_dtype = q.mode == 0 ? 1 /*F16*/ : 4 /*U8*/;
  return true;
}

__device__ __forceinline__ float load_w(const void* w, int dtype, int64_t i) {
  if (dtype == 0) return reinterpret_cast<const float*>(w)[i];
  if (dtype == 1) return __half2float(reinterpret_cast<const __half*>(w)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(w)[i]);
}
__device__ __forceinline__ void store_o(void* o, int dtype, int64_t i, float v) {
  if (dtype == 0) reinterpret_cast<float*>(o)[i] = v;
  else if (dtype == 1) reinterpret_cast<__half*>(o)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(o)[i] = __float2bfloat16(v);
}

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

// ---- one thread per group
__global__ void quantize_kernel(int quant, const void* __restrict__ w, int w_dtype, int64_t rows, int64_t cols,
                                uint32_t* __restrict__ packed, void* __restrict__ scales, void* __restrict__ biases) {
  const QSpec q = qspec(quant);
  const int64_t groups_per_row = cols / q.group;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * groups_per_row) return;
  const int64_t row = gid / groups_per_row, gi = gid % groups_per_row;
  const int64_t base = row * cols + gi * q.group;
  const int per_word = 32 / q.bits;
  uint32_t* out = packed + (row * cols + gi * q.group) / per_word;

  if (q.mode == 0) {
    float wmax = -INFINITY, wmin = INFINITY;
    for (int i = 0; i < q.group; ++i) {
      const float v = load_w(w, w_dtype, base + i);
      wmax = fmaxf(wmax, v);
      wmin = fminf(wmin, v);
    }
    const float n_bins = (float)((1 << q.bits) - 1);
    float scale = fmaxf(__fdiv_rn(__fsub_rn(wmax, wmin), n_bins), 1e-7f);
    const bool side = fabsf(wmin) > fabsf(wmax);
    scale = side ? scale : -scale;
    const float edge = side ? wmin : wmax;
    const float q0 = roundf(__fdiv_rn(edge, scale));
    const bool at_zero = (q0 == 0.0f);
    scale = at_zero ? scale : __fdiv_rn(edge, q0);
    const float bias = at_zero ? 0.f : edge;
    reinterpret_cast<__half*>(scales)[gid] = __float2half_rn(scale);
    reinterpret_cast<__half*>(biases)[gid] = __float2half_rn(bias);
    for (int wd = 0; wd < q.group / per_word; ++wd) {
      uint32_t word = 0;
      for (int j = 0; j < per_word; ++j) {
        const float v = load_w(w, w_dtype, base + wd * per_word + j);
        float r = roundf(__fdiv_rn(__fsub_rn(v, bias), scale));
        r = fminf(r, n_bins);
        r = fmaxf(r, 0.f);
        word |= ((uint32_t)r) << (j * q.bits);
      }
      out[wd] = word;
    }
    return;
  }
  float amax = 0.f;
  for (int i = 0; i < q.group; ++i) amax = fmaxf(amax, fabsf(load_w(w, w_dtype, base + i)));
  float scale = __fdiv_rn(amax, q.bits == 4 ? 6.0f : 448.0f);
  uint8_t sb;
  if (q.mode == 1) { sb = to_e8m0(scale); scale = from_e8m0(sb); }
  else { sb = to_e4m3(scale); scale = from_e4m3(sb); }
  reinterpret_cast<uint8_t*>(scales)[gid] = sb;
  for (int wd = 0; wd < q.group / per_word; ++wd) {
    uint32_t word = 0;
    for (int j = 0; j < per_word; ++j) {
      const float v = load_w(w, w_dtype, base + wd * per_word + j);
      const float x = (scale == 0.f) ? 0.f : __fdiv_rn(v, scale);
      const uint32_t code = (q.bits == 4) ? to_e2m1(x) : to_e4m3(x);
      word |= code << (j * q.bits);
    }
    out[wd] = word;
  }
}

__global__ void dequantize_kernel(int quant, const uint32_t* __restrict__ packed, const void* __restrict__ scales,
                                  const void* __restrict__ biases, int64_t rows, int64_t cols, void* __restrict__ out,
                                  int out_dtype) {
  const QSpec q = qspec(quant);
  const int per_word = 32 / q.bits;
  const int64_t wid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per packed word
  const int64_t nwords = rows * cols / per_word;
  if (wid >= nwords) return;
  const int64_t e0 = wid * per_word;
  const int64_t row = e0 / cols, col = e0 % cols;
  const int64_t gid = row * (cols / q.group) + col / q.group;
  const uint32_t word = packed[wid];
  const uint32_t mask = (1u << q.bits) - 1u;
  if (q.mode == 0) {
    const float s = __half2float(reinterpret_cast<const __half*>(scales)[gid]);
    const float b = __half2float(reinterpret_cast<const __half*>(biases)[gid]);
    for (int j = 0; j < per_word; ++j) {
      const float qv = (float)((word >> (j * q.bits)) & mask);
      store_o(out, out_dtype, e0 + j, __fadd_rn(__fmul_rn(qv, s), b));
    }
  } else {
    const uint8_t sb = reinterpret_cast<const uint8_t*>(scales)[gid];
    const float s = (q.mode == 1) ? from_e8m0(sb) : from_e4m3(sb);
    for (int j = 0; j < per_word; ++j) {
      const uint8_t code = (uint8_t)((word >> (j * q.bits)) & mask);
      const float ev = (q.bits == 4) ? from_e2m1(code) : from_e4m3(code);
      store_o(out, out_dtype, e0 + j, __fmul_rn(ev, s));
    }
  }
}

cudaError_t quantize_matrix(int quant, const void* w, int w_dtype, int64_t rows, int64_t cols, uint32_t* packed,
                            void* scales, void* biases, cudaStream_t s) {
  const QSpec q = qspec(quant);
  if (q.mode < 0 || cols % q.group) return cudaErrorInvalidValue;
  const int64_t n = rows * (cols / q.group);
  if (n <= 0) return cudaSuccess;
  quantize_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(quant, w, w_dtype, rows, cols, packed, scales, biases);
  return cudaGetLastError();
}
cudaError_t dequantize_matrix(int quant, const uint32_t* packed, const void* scales, const void* biases, int64_t rows,
                              int64_t cols, void* out, int out_dtype, cudaStream_t s) {
  const QSpec q = qspec(quant);
  if (q.mode < 0 || cols % q.group) return cudaErrorInvalidValue;
  const int64_t n = rows * cols / (32 / q.bits);
  if (n <= 0) return cudaSuccess;
  dequantize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(quant, packed, scales, biases, rows, cols, out, out_dtype);
  return cudaGetLastError();
}

// W[out,in] (16-bit or f32) += scale * B[out,r] · A[r,in], computed in fp32 and rounded once to the weight dtype
// (WeightLoader.swift:825-838). Load-time only: a plain CUDA-core kernel (rank is ~16).
__global__ void lora_add_kernel(void* __restrict__ W, int w_dtype, const float* __restrict__ A, const float* __restrict__ B,
                                int64_t out_dim, int64_t in_dim, int rank, float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_dim * in_dim) return;
  const int64_t o = i / in_dim, c = i % in_dim;
  float acc = 0.f;
  for (int r = 0; r < rank; ++r) acc = __fadd_rn(acc, __fmul_rn(B[o * rank + r], A[(int64_t)r * in_dim + c]));
  // reference (WeightLoader.swift:806-810,832-836): A, B cast to the weight dtype; matmul, scale* and + each produce
  // an array of that dtype, i.e. every step rounds to it (A, B arrive here already rounded).
  auto rw = [&](float v) {
    if (w_dtype == 1) return __half2float(__float2half_rn(v));
    if (w_dtype == 2) return __bfloat162float(__float2bfloat16(v));
    return v;
  };
  const float ba = rw(acc);
  const float delta = rw(__fmul_rn(rw(scale), ba));
  const float wv = load_w(W, w_dtype, i);
  store_o(W, w_dtype, i, __fadd_rn(wv, delta));
}
cudaError_t lora_add(void* W, int w_dtype, const float* A, const float* B, int64_t out_dim, int64_t in_dim, int rank,
                     float scale, cudaStream_t s) {
  const int64_t n = out_dim * in_dim;
  if (n <= 0) return cudaSuccess;
  lora_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(W, w_dtype, A, B, out_dim, in_dim, rank, scale);
  return cudaGetLastError();
}


// ------------------------------------------------------------------ native block-scaled operands (mxfp8 / mxfp4 / nvfp4)
__host__ __device__ inline int mx_group(int kind) { return kind == 3 ? 16 : 32; }
__host__ __device__ inline int mx_bits(int kind) { return kind == 1 ? 8 : 4; }
int mx_kind_of_quant(int quant) { return quant == 3 ? 1 : quant == 4 ? 2 : quant == 5 ? 3 : 0; }
int64_t mx_sf_ld(int kind, int64_t K) { return K / mx_group(kind) / 4; }
size_t mx_sf_bytes(int kind, int64_t rows, int64_t K) { return (size_t)((rows + 127) / 128) * mx_sf_ld(kind, K) * 512; }
__device__ __forceinline__ int64_t sf_offset(int64_t row, int64_t g, int64_t ld_blocks) {
  return ((row >> 7) * ld_blocks + (g >> 2)) * 512 + (row & 31) * 16 + ((row & 127) >> 5) * 4 + (g & 3);
}
__global__ void mx_copy_rows_kernel(const uint8_t* __restrict__ src_w, const uint8_t* __restrict__ src_s, int64_t src_row0,
                                    uint8_t* __restrict__ dst_w, uint8_t* __restrict__ dst_sf, int64_t dst_row0,
                                    int64_t nrows, int64_t row_bytes, int64_t G, int tiled, int64_t Hm) {
  const int64_t vec_per_row = row_bytes / 16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto src_row = [&](int64_t r) {
    if (!tiled) return r;
    const int64_t tile = r / 256, j = r % 256;
    return (j < 128) ? tile * 128 + j : Hm + tile * 128 + (j - 128);
  };
  if (i < nrows * vec_per_row) {
    const int64_t r = i / vec_per_row, v = i % vec_per_row;
    *reinterpret_cast<uint4*>(dst_w + (dst_row0 + r) * row_bytes + v * 16) =
        *reinterpret_cast<const uint4*>(src_w + (src_row0 + src_row(r)) * row_bytes + v * 16);
  }
  if (i < nrows * G) {
    const int64_t r = i / G, g = i % G;
    dst_sf[sf_offset(dst_row0 + r, g, G / 4)] = src_s[(src_row0 + src_row(r)) * G + g];
  }
}
cudaError_t mx_copy_rows(int kind, const uint8_t* src_w, const uint8_t* src_s, int64_t src_row0, uint8_t* dst_w, uint8_t* dst_sf,
                         int64_t dst_row0, int64_t nrows, int64_t K, bool tiled, int64_t Hm, cudaStream_t s) {
  if (kind < 1 || kind > 3 || K % (kind == 1 ? 128 : 256)) return cudaErrorInvalidValue;
  const int64_t row_bytes = K * mx_bits(kind) / 8, G = K / mx_group(kind);
  const int64_t n = nrows * std::max(row_bytes / 16, G);
  if (n <= 0) return cudaSuccess;
  mx_copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src_w, src_s, src_row0, dst_w, dst_sf, dst_row0, nrows, row_bytes, G,
                                                                  tiled ? 1 : 0, Hm);
  return cudaGetLastError();
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
// mxfp8: one warp per (row, 128-element K block): lane owns 4 consecutive elements, 8 lanes share a 32-element group.
// The scale is 2^ceil(log2(amax / 448)) (nothing saturates) — an activation is quantised once and consumed at once, so it
// need not follow the weight packer's round-to-nearest exponent rule.
__global__ void __launch_bounds__(256) mx8_quantize_act_kernel(const void* __restrict__ x, int64_t ldx, int M, int K, bool f16,
                                                              uint8_t* __restrict__ a8, int64_t lda, uint8_t* __restrict__ sfa,
                                                              int64_t sf_ld, int g0) {
  const int kb4 = K / 128;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t Mpad = ((int64_t)M + 127) / 128 * 128;
  if (wid >= Mpad * kb4) return;
  const int64_t row = wid / kb4;
  const int kb = (int)(wid % kb4);
  if (row >= M) {  // padding rows of the last 128-row block: scale 1.0 (never multiplied with anything but zeros)
    if (lane < 4) sfa[sf_offset(row, g0 + kb * 4 + lane, sf_ld)] = 127;
    return;
  }
  const uint16_t* xr = reinterpret_cast<const uint16_t*>(x) + row * ldx + kb * 128 + lane * 4;
  const uint2 raw = *reinterpret_cast<const uint2*>(xr);
  float2 a, b;
  if (f16) { a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)); b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y)); }
  else { a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x)); b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y)); }
  float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y)));
#pragma unroll
  for (int o = 4; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));  // 8 lanes = one 32-element group
  // smallest power of two s with amax / s <= 448  (exponent arithmetic only)
  int e = -127;
  if (amax > 0.f) {
    const float q = amax * (1.0f / 448.0f);
    const uint32_t u = __float_as_uint(q);
    e = (int)((u >> 23) & 0xff) - 127 + ((u & 0x7fffff) ? 1 : 0);
    e = e < -127 ? -127 : (e > 127 ? 127 : e);
  }
  const uint32_t ebits = (uint32_t)(127 - e);                                      // biased exponent of 2^-e, in [0, 254]
  const float inv = __uint_as_float(ebits ? (ebits << 23) : 0x00400000u);          // 2^-e (2^-127 is subnormal)
  const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a.x * inv, a.y * inv), __NV_SATFINITE, __NV_E4M3);
  const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(b.x * inv, b.y * inv), __NV_SATFINITE, __NV_E4M3);
  *reinterpret_cast<uint32_t*>(a8 + row * lda + kb * 128 + lane * 4) = (uint32_t)lo | ((uint32_t)hi << 16);
  if ((lane & 7) == 0) sfa[sf_offset(row, g0 + kb * 4 + (lane >> 3), sf_ld)] = (uint8_t)(e + 127);
}
// fp4 kinds: one warp per (row, 256-element K block): lane owns 8 consecutive elements (16 B in, 4 B out); a 16-element
// nvfp4 group is 2 lanes, a 32-element mxfp4 group 4 lanes. Same arithmetic as the weight packer above (amax / 6 -> E4M3 or
// E8M0 scale, x / scale -> E2M1 RNE), so the result is bit-identical to quantize_kernel / the C oracle on the same matrix.
template <bool kNv>
__global__ void __launch_bounds__(256) mx4_quantize_act_kernel(const void* __restrict__ x, int64_t ldx, int M, int K, bool f16,
                                                              uint8_t* __restrict__ a4, int64_t lda, uint8_t* __restrict__ sfa,
                                                              int64_t sf_ld, int g0) {
  constexpr int GPB = kNv ? 16 : 8;  // groups per 256-element block
  const int kbn = K / 256;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t Mpad = ((int64_t)M + 127) / 128 * 128;
  if (wid >= Mpad * kbn) return;
  const int64_t row = wid / kbn;
  const int kb = (int)(wid % kbn);
  if (row >= M) {  // padding rows: scale 1.0
    if (lane < GPB) sfa[sf_offset(row, g0 + kb * GPB + lane, sf_ld)] = kNv ? 0x38 : 127;
    return;
  }
  const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(x) + row * ldx + kb * 256 + lane * 8);
  float v[8];
  unpack8(raw, f16, v);
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[j]));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  if (!kNv) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  float scale = __fdiv_rn(amax, 6.0f);
  uint8_t sb;
  if (kNv) { sb = to_e4m3(scale); scale = from_e4m3(sb); }
  else { sb = to_e8m0(scale); scale = from_e8m0(sb); }
  uint32_t word = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float q = (scale == 0.f) ? 0.f : __fdiv_rn(v[j], scale);
    word |= (uint32_t)to_e2m1(q) << (4 * j);
  }
  *reinterpret_cast<uint32_t*>(a4 + row * lda + kb * 128 + lane * 4) = word;
  constexpr int LPG = kNv ? 2 : 4;  // lanes per group
  if ((lane % LPG) == 0) sfa[sf_offset(row, g0 + kb * GPB + lane / LPG, sf_ld)] = sb;
}
cudaError_t mx_quantize_act(int kind, const void* x16, int64_t ldx, int M, int K, bool f16, uint8_t* aq, int64_t lda_bytes,
                            uint8_t* sfa, int64_t sf_ld, int64_t col0, cudaStream_t s) {
  const int kb_elems = kind == 1 ? 128 : 256;
  if (kind < 1 || kind > 3 || K % kb_elems || col0 % kb_elems || ldx % 8 || lda_bytes % 4) return cudaErrorInvalidValue;
  const int64_t Mpad = ((int64_t)M + 127) / 128 * 128;
  const int64_t warps = Mpad * (K / kb_elems);
  if (warps <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  const int g0 = (int)(col0 / mx_group(kind));
  if (kind == 1) mx8_quantize_act_kernel<<<blocks, 256, 0, s>>>(x16, ldx, M, K, f16, aq, lda_bytes, sfa, sf_ld, g0);
  else if (kind == 2) mx4_quantize_act_kernel<false><<<blocks, 256, 0, s>>>(x16, ldx, M, K, f16, aq, lda_bytes, sfa, sf_ld, g0);
  else mx4_quantize_act_kernel<true><<<blocks, 256, 0, s>>>(x16, ldx, M, K, f16, aq, lda_bytes, sfa, sf_ld, g0);
  return cudaGetLastError();
}
// scale factors back from the tcgen05 layout to row-major [M, G] (tests / debugging)
__global__ void sf_untile_kernel(const uint8_t* __restrict__ sf, uint8_t* __restrict__ out, int64_t M, int64_t G) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * G) return;
  out[i] = sf[sf_offset(i / G, i % G, G / 4)];
}
cudaError_t mx_sf_untile(const uint8_t* sf, uint8_t* out, int64_t M, int64_t G, cudaStream_t s) {
  const int64_t n = M * G;
  if (n <= 0) return cudaSuccess;
  sf_untile_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sf, out, M, G);
  return cudaGetLastError();
}

}  // namespace f2b
