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
#include "quant_dev.cuh"
#include <cuda_fp8.h>
#include <cuda_fp4.h>
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
                                  const void* __restrict__ biases, int sb_dtype, int64_t rows, int64_t cols, void* __restrict__ out,
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
    // scales / biases in the checkpoint's own float type (f16 from this packer; MLX-quantized bf16 models store bf16)
    const float s = load_w(scales, sb_dtype, gid);
    const float b = load_w(biases, sb_dtype, gid);
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
                              int64_t cols, void* out, int out_dtype, cudaStream_t s, int sb_dtype) {
  const QSpec q = qspec(quant);
  if (q.mode < 0 || cols % q.group || sb_dtype < 0 || sb_dtype > 2) return cudaErrorInvalidValue;
  const int64_t n = rows * cols / (32 / q.bits);
  if (n <= 0) return cudaSuccess;
  dequantize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(quant, packed, scales, biases, sb_dtype, rows, cols, out, out_dtype);
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
__global__ void mx_copy_rows_kernel(const uint8_t* __restrict__ src_w, const uint8_t* __restrict__ src_s, int64_t src_row0,
                                    uint8_t* __restrict__ dst_w, uint8_t* __restrict__ dst_sf, int64_t dst_row0,
                                    int64_t nrows, int64_t row_bytes, int64_t G, int tile, int64_t Hm) {
  const int64_t vec_per_row = row_bytes / 16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto src_row = [&](int64_t r) {  // SwiGLU interleave: every `tile` rows = [tile/2 gate rows | tile/2 value rows]
    if (!tile) return r;
    const int64_t h = tile / 2, t = r / tile, j = r % tile;
    return (j < h) ? t * h + j : Hm + t * h + (j - h);
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
                         int64_t dst_row0, int64_t nrows, int64_t K, int tile, int64_t Hm, cudaStream_t s) {
  if (kind < 1 || kind > 3 || K % (kind == 1 ? 128 : 256) || (tile && (nrows % tile || Hm % (tile / 2)))) return cudaErrorInvalidValue;
  const int64_t row_bytes = K * mx_bits(kind) / 8, G = K / mx_group(kind);
  const int64_t n = nrows * std::max(row_bytes / 16, G);
  if (n <= 0) return cudaSuccess;
  mx_copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src_w, src_s, src_row0, dst_w, dst_sf, dst_row0, nrows, row_bytes, G,
                                                                  tile, Hm);
  return cudaGetLastError();
}

// On-the-fly activation quantiser. A warp owns up to kIters x 256 consecutive elements of one row: lane = 8 consecutive
// elements per iteration (16 B in; 8 B of E4M3 or 4 B of E2M1 out), all loads of a warp issued before any arithmetic.
// A 32-element group is 4 lanes, a 16-element nvfp4 group 2 lanes; the four scale bytes of one 512 B-block row are gathered
// by shuffles and stored as one word.
//   mxfp8: scale = 2^ceil(log2(amax / 448)) (nothing saturates; an activation is quantised once and consumed at once, so it
//          need not follow the weight packer's round-to-nearest exponent rule), elements by cvt.rn.satfinite.e4m3x2.
//   mxfp4 / nvfp4: the weight packer's rule (amax / 6 -> E8M0 / E4M3 scale, x / scale -> E2M1 RNE, saturating), evaluated as
//          cvt.rn.satfinite.e2m1x2(f16(x * (1 / scale))). That is bit-identical to quantize_kernel / the C oracle for 16-bit
//          inputs: x (<= 11 significant bits) and threshold * scale (<= 7 bits) are dyadic rationals that either coincide or
//          differ by >= 2^-12 relative, the product carries <= 2^-22 error and the f16 rounding snaps an exact tie back onto
//          its threshold, where RNE picks the even code exactly like to_e2m1() above.
template <int KIND>
__global__ void __launch_bounds__(256) mx_quantize_act_kernel(const void* __restrict__ x, int64_t ldx, int M, int K, bool f16,
                                                             uint8_t* __restrict__ aq, int64_t lda, uint8_t* __restrict__ sfa,
                                                             int64_t sf_ld, int g0) {
  constexpr int kIters = 4;
  constexpr int GROUP = KIND == 3 ? 16 : 32;
  constexpr int LPG = GROUP / 8;        // lanes per group
  const int chunks = (K + kIters * 256 - 1) / (kIters * 256);
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t Mpad = ((int64_t)M + 127) / 128 * 128;
  if (wid >= Mpad * chunks) return;
  const int64_t row = wid / chunks;
  const int k_begin = (int)(wid % chunks) * kIters * 256;
  if (row >= M) {  // padding rows of the last 128-row block: scale 1.0 (only ever multiplied with TMA zero fill)
    const int g_end = min(K, k_begin + kIters * 256) / GROUP;
    for (int g = k_begin / GROUP + lane; g < g_end; g += 32) sfa[sf_offset(row, g0 + g, sf_ld)] = (KIND == 3) ? 0x38 : 127;
    return;
  }
  const uint16_t* xr = reinterpret_cast<const uint16_t*>(x) + row * ldx;
  uint4 raw[kIters];
#pragma unroll
  for (int i = 0; i < kIters; ++i) {
    const int k = k_begin + i * 256 + lane * 8;
    raw[i] = (k < K) ? *reinterpret_cast<const uint4*>(xr + k) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int i = 0; i < kIters; ++i) {
    const int k = k_begin + i * 256 + lane * 8;
    if (k_begin + i * 256 >= K) break;  // warp-uniform
    float v[8];
    unpack8(raw[i], f16, v);
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[j]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    if (LPG == 4) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    const MxScale sc = mx_scale<KIND>(amax);
    const uint32_t sb = sc.sb;
    if constexpr (KIND == 1) {
      if (k < K) *reinterpret_cast<uint2*>(aq + row * lda + k) =
          make_uint2(mx_pack4<1>(v[0], v[1], v[2], v[3], sc.mul), mx_pack4<1>(v[4], v[5], v[6], v[7], sc.mul));
    } else {
      if (k < K) *reinterpret_cast<uint32_t*>(aq + row * lda + (k >> 1)) =
          mx_pack4<KIND>(v[0], v[1], v[2], v[3], sc.mul) | (mx_pack4<KIND>(v[4], v[5], v[6], v[7], sc.mul) << 16);
    }
    // four consecutive group scales -> one 32-bit store (lane of the first group of each 4-group block)
    uint32_t w = sb | (__shfl_down_sync(0xffffffffu, sb, LPG) << 8);
    w |= __shfl_down_sync(0xffffffffu, w, 2 * LPG) << 16;
    if ((lane % (4 * LPG)) == 0 && k < K)
      *reinterpret_cast<uint32_t*>(sfa + sf_offset(row, g0 + k / GROUP, sf_ld)) = w;
  }
}
cudaError_t mx_quantize_act(int kind, const void* x16, int64_t ldx, int M, int K, bool f16, uint8_t* aq, int64_t lda_bytes,
                            uint8_t* sfa, int64_t sf_ld, int64_t col0, cudaStream_t s) {
  const int kb_elems = kind == 1 ? 128 : 256;
  if (kind < 1 || kind > 3 || K % kb_elems || col0 % kb_elems || ldx % 8 || lda_bytes % 8 ||
      (reinterpret_cast<uintptr_t>(x16) & 15) || (reinterpret_cast<uintptr_t>(aq) & 7))
    return cudaErrorInvalidValue;
  const int64_t Mpad = ((int64_t)M + 127) / 128 * 128;
  const int64_t warps = Mpad * ((K + 1023) / 1024);
  if (warps <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  const int g0 = (int)(col0 / mx_group(kind));
  if (kind == 1) mx_quantize_act_kernel<1><<<blocks, 256, 0, s>>>(x16, ldx, M, K, f16, aq, lda_bytes, sfa, sf_ld, g0);
  else if (kind == 2) mx_quantize_act_kernel<2><<<blocks, 256, 0, s>>>(x16, ldx, M, K, f16, aq, lda_bytes, sfa, sf_ld, g0);
  else mx_quantize_act_kernel<3><<<blocks, 256, 0, s>>>(x16, ldx, M, K, f16, aq, lda_bytes, sfa, sf_ld, g0);
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
