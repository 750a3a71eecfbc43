// elementwise.cu — the bandwidth-bound kernels of the denoising path (no tensor cores: these are HBM-bound,
// so the work is coalescing, 16 B vector access, keeping a row in registers between its reductions, and grids that
// oversubscribe the 148 SMs).
#include "elementwise.cuh"
#include "ptx.cuh"
#include "gemm.cuh"
#include "quant_dev.cuh"
#include <algorithm>

namespace f2b {

__device__ __forceinline__ uint32_t pack2(float a, float b, bool f16) {
  if (f16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16x2(a, b);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, bool f16) {
  if (f16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
__device__ __forceinline__ float load16(const void* p, int64_t i, bool f16) {
  return f16 ? __half2float(reinterpret_cast<const __half*>(p)[i])
             : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void store16(void* p, int64_t i, float v, bool f16) {
  if (f16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < NT / 32) ? sm[l] : 0.f;
  r = warp_sum(r);
  return r;
}

// ------------------------------------------------------------------ LayerNorm + modulate
// one CTA (256 threads) per row; the row stays in registers across the mean / variance reductions.
// KIND != 0 (native block-scaled path): the result is rounded to the 16-bit operand type and quantised in the same pass to
// mxfp8 / mxfp4 / nvfp4 (element bytes + tcgen05-layout scale factors, bit-identical to mx_quantize_act on the 16-bit
// output) — a thread owns 4 consecutive elements, a 16 / 32-element group is 4 / 8 adjacent lanes. CTAs of rows >= `rows`
// (up to the next multiple of 128) only fill their scale-factor row with 1.0.
template <int MAXV, int KIND>  // float4 vectors per thread
__global__ void __launch_bounds__(256) ln_modulate_kernel(const float* __restrict__ x, int64_t ldx, void* __restrict__ out,
                                                          int64_t ldo, int rows, int D, const float* __restrict__ shift,
                                                          const float* __restrict__ scale, int64_t mod_bs,
                                                          int rows_per_batch, float eps, bool f16, MxOut mx, int split_row,
                                                          const float* __restrict__ shift_lo, const float* __restrict__ scale_lo) {
  __shared__ float sm[8];
  constexpr int GROUP = KIND == 3 ? 16 : 32;
  constexpr int LPG = GROUP / 4;
  pdl_trigger();   // programmatic dependent launch (ptx.cuh): the GEMM behind this kernel may set itself up meanwhile
  pdl_wait();
  const int row = blockIdx.x;
  if (KIND != 0 && row >= rows) {
    for (int g = threadIdx.x; g < D / GROUP; g += 256) mx.sf[sf_offset(row, mx.g0 + g, mx.sf_ld)] = mx_scale_one(KIND);
    return;
  }
  const int b = row / rows_per_batch;
  const float* xr = x + (int64_t)row * ldx;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < D) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = block_sum<256>(s, sm) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < D) {
      float a = v[i].x - mean, bq = v[i].y - mean, cq = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bq * bq) + (cq * cq + d * d);
    }
  }
  const float rstd = rsqrtf(block_sum<256>(q, sm) / (float)D + eps);
  // rows below split_row form a second stream with its own modulation (text rows of a double-stream block, one launch for both)
  const float* sh = row < split_row ? shift_lo : shift + b * mod_bs;
  const float* sc = row < split_row ? scale_lo : scale + b * mod_bs;
  uint16_t* orow = reinterpret_cast<uint16_t*>(out) + (int64_t)row * ldo;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < D) {  // warp-uniform in the quantising variant (D % 128 == 0)
      const float4 h4 = __ldg(reinterpret_cast<const float4*>(sh + c));
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(sc + c));
      const float o0 = (v[i].x - mean) * rstd * (1.f + c4.x) + h4.x;
      const float o1 = (v[i].y - mean) * rstd * (1.f + c4.y) + h4.y;
      const float o2 = (v[i].z - mean) * rstd * (1.f + c4.z) + h4.z;
      const float o3 = (v[i].w - mean) * rstd * (1.f + c4.w) + h4.w;
      const uint32_t p01 = pack2(o0, o1, f16), p23 = pack2(o2, o3, f16);
      if constexpr (KIND == 0) {
        *reinterpret_cast<uint2*>(orow + c) = make_uint2(p01, p23);
      } else {
        const float2 a = unpack2(p01, f16), bb = unpack2(p23, f16);
        float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(bb.x), fabsf(bb.y)));
#pragma unroll
        for (int o = 1; o < LPG; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const MxScale ms = mx_scale<KIND>(amax);
        const uint32_t pk = mx_pack4<KIND>(a.x, a.y, bb.x, bb.y, ms.mul);
        uint8_t* qrow = mx.q + (int64_t)row * mx.ldq;
        if constexpr (KIND == 1) *reinterpret_cast<uint32_t*>(qrow + c) = pk;
        else *reinterpret_cast<uint16_t*>(qrow + (c >> 1)) = (uint16_t)pk;
        uint32_t w = ms.sb | (__shfl_down_sync(0xffffffffu, ms.sb, LPG) << 8);
        w |= __shfl_down_sync(0xffffffffu, w, 2 * LPG) << 16;
        if ((lane % (4 * LPG)) == 0) *reinterpret_cast<uint32_t*>(mx.sf + sf_offset(row, mx.g0 + c / GROUP, mx.sf_ld)) = w;
      }
    }
  }
}

// ------------------------------------------------------------------ text-encoder kernels (te.cu)
// RMSNorm over the hidden dimension with a learned weight (FluxTextEncoders/Model/RMSNorm.swift -> MLXFast.rmsNorm):
// out = x * rsqrt(mean(x^2) + eps) * w. One CTA per row, the row stays in registers; 16-bit GEMM operand or fp32 out.
template <int MAXV, bool OUT_F32>
__global__ void __launch_bounds__(256) rms_norm_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                            void* __restrict__ out, int64_t ldo, int D, float eps, bool f16) {
  __shared__ float sm[8];
  const int row = blockIdx.x;
  const float* xr = x + (int64_t)row * ldx;
  float4 v[MAXV];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < D) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  const float rstd = rsqrtf(block_sum<256>(q, sm) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < D) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
      const float o0 = v[i].x * rstd * w4.x, o1 = v[i].y * rstd * w4.y, o2 = v[i].z * rstd * w4.z, o3 = v[i].w * rstd * w4.w;
      if constexpr (OUT_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (int64_t)row * ldo + c) = make_float4(o0, o1, o2, o3);
      else *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + (int64_t)row * ldo + c) = make_uint2(pack2(o0, o1, f16), pack2(o2, o3, f16));
    }
  }
}
cudaError_t rms_norm_rows(const float* x, int64_t ldx, const float* w, void* out, int64_t ldo, int rows, int D, float eps,
                          bool out_f32, bool f16, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  if (D % 4 || D > 8192) return cudaErrorInvalidValue;
#define F2B_RMS(V)                                                                                        \
  do {                                                                                                    \
    if (out_f32) rms_norm_rows_kernel<V, true><<<rows, 256, 0, s>>>(x, ldx, w, out, ldo, D, eps, f16);     \
    else rms_norm_rows_kernel<V, false><<<rows, 256, 0, s>>>(x, ldx, w, out, ldo, D, eps, f16);            \
  } while (0)
  if (D <= 1024) F2B_RMS(1);
  else if (D <= 3072) F2B_RMS(3);
  else if (D <= 4096) F2B_RMS(4);
  else if (D <= 6144) F2B_RMS(6);
  else F2B_RMS(8);
#undef F2B_RMS
  return cudaGetLastError();
}

// Token embedding lookup (Qwen3Model.swift:66 embed_tokens): X[s, :] = table16[ids[s], :] widened to fp32. 8 elements per thread.
__global__ void embed_rows_kernel(const int32_t* __restrict__ ids, const uint16_t* __restrict__ table, int64_t vocab, int D,
                                  float* __restrict__ x, int64_t ldx, int rows, bool f16) {
  const int vec = D / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * vec) return;
  const int r = (int)(i / vec), c = (int)(i % vec) * 8;
  int64_t tok = ids[r];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);   // ids are validated on the host when they arrive from host memory
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(table + tok * D + c));
  const float2 a = unpack2(u.x, f16), b = unpack2(u.y, f16), cc = unpack2(u.z, f16), d = unpack2(u.w, f16);
  float4* o = reinterpret_cast<float4*>(x + (int64_t)r * ldx + c);
  o[0] = make_float4(a.x, a.y, b.x, b.y);
  o[1] = make_float4(cc.x, cc.y, d.x, d.y);
}
cudaError_t embed_rows(const int32_t* ids, const void* table16, int64_t vocab, int D, float* x, int64_t ldx, int rows, bool f16,
                       cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  if (D % 8) return cudaErrorInvalidValue;
  const int64_t n = (int64_t)rows * (D / 8);
  embed_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ids, reinterpret_cast<const uint16_t*>(table16), vocab, D, x, ldx, rows, f16);
  return cudaGetLastError();
}

// cos / sin [S, 128] fp32 for rotate-half RoPE at head dim 128 (MLXFast.RoPE traditional = false, scale 1): column j and
// j + 64 hold the angle (pos0 + s) * base^(-j / 64), j < 64.
__global__ void rope_half_table_kernel(int S, int pos0, float log2_base, float* __restrict__ cos_out, float* __restrict__ sin_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * 64) return;
  const int s = i >> 6, j = i & 63;
  const float inv_freq = exp2f(-(float)j * (1.0f / 64.0f) * log2_base);
  const float ang = (float)(pos0 + s) * inv_freq;
  float sn, cs;
  sincosf(ang, &sn, &cs);
  cos_out[(int64_t)s * 128 + j] = cs; cos_out[(int64_t)s * 128 + 64 + j] = cs;
  sin_out[(int64_t)s * 128 + j] = sn; sin_out[(int64_t)s * 128 + 64 + j] = sn;
}
cudaError_t rope_half_table(int S, int pos0, float base, float* cos_out, float* sin_out, cudaStream_t s) {
  if (S <= 0) return cudaSuccess;
  rope_half_table_kernel<<<(S * 64 + 255) / 256, 256, 0, s>>>(S, pos0, log2f(base), cos_out, sin_out);
  return cudaGetLastError();
}

// strided fp32 -> {fp32, bf16, f16} copy of a [rows, cols] block (hidden-state extraction: [S, H] slab of the [S, n*H] output)
__global__ void copy_f32_to_any_kernel(const float* __restrict__ in, int64_t ldi, void* __restrict__ out, int64_t ldo, int rows,
                                       int cols, int out_kind) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  const float v = in[(int64_t)r * ldi + c];
  if (out_kind == 0) reinterpret_cast<float*>(out)[(int64_t)r * ldo + c] = v;
  else if (out_kind == 1) reinterpret_cast<__half*>(out)[(int64_t)r * ldo + c] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(out)[(int64_t)r * ldo + c] = __float2bfloat16(v);
}
cudaError_t copy_f32_to_any(const float* in, int64_t ldi, void* out, int64_t ldo, int rows, int cols, int out_kind, cudaStream_t s) {
  const int64_t n = (int64_t)rows * cols;
  if (n <= 0) return cudaSuccess;
  copy_f32_to_any_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, ldi, out, ldo, rows, cols, out_kind);
  return cudaGetLastError();
}

template <int KIND>
static void ln_launch(const float* x, int64_t ldx, void* out16, int64_t ldo, int rows, int grid_rows, int D, const float* shift,
                      const float* scale, int64_t mod_bs, int rows_per_batch, float eps, bool f16, const MxOut& mx, cudaStream_t s,
                      int split_row = 0, const float* shift_lo = nullptr, const float* scale_lo = nullptr) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid_rows); cfg.blockDim = dim3(256); cfg.stream = s;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs; cfg.numAttrs = pdl_enabled() ? 1 : 0;
#define F2B_LN(V) cudaLaunchKernelEx(&cfg, ln_modulate_kernel<V, KIND>, x, ldx, out16, ldo, rows, D, shift, scale, mod_bs, rows_per_batch, eps, f16, mx, split_row, shift_lo, scale_lo)
  if (D <= 1024) F2B_LN(1);
  else if (D <= 3072) F2B_LN(3);
  else if (D <= 4096) F2B_LN(4);
  else if (D <= 6144) F2B_LN(6);
  else F2B_LN(8);
#undef F2B_LN
}
cudaError_t ln_modulate(const float* x, int64_t ldx, void* out16, int64_t ldo, int rows, int D, const float* shift,
                        const float* scale, int64_t mod_bs, int rows_per_batch, float eps, bool f16, cudaStream_t s,
                        const MxOut* mx, int split_row, const float* shift_lo, const float* scale_lo) {
  if (rows <= 0) return cudaSuccess;
  if (split_row && (mx && mx->kind)) return cudaErrorInvalidValue;
  if (D % 4 || D > 8192 || ldx % 4 || ldo % 4) return cudaErrorInvalidValue;
  MxOut m;
  if (mx) m = *mx;
  if (m.kind) {
    if (D % 128 || !m.q || !m.sf || m.g0 % 4 || m.ldq % 4) return cudaErrorInvalidValue;
    const int grid_rows = (rows + 127) / 128 * 128;
    if (m.kind == 1) ln_launch<1>(x, ldx, out16, ldo, rows, grid_rows, D, shift, scale, mod_bs, rows_per_batch, eps, f16, m, s);
    else if (m.kind == 2) ln_launch<2>(x, ldx, out16, ldo, rows, grid_rows, D, shift, scale, mod_bs, rows_per_batch, eps, f16, m, s);
    else ln_launch<3>(x, ldx, out16, ldo, rows, grid_rows, D, shift, scale, mod_bs, rows_per_batch, eps, f16, m, s);
  } else {
    ln_launch<0>(x, ldx, out16, ldo, rows, rows, D, shift, scale, mod_bs, rows_per_batch, eps, f16, m, s, split_row, shift_lo, scale_lo);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ GEMV (M = batch <= 8): weight-bandwidth bound
// STAGED: the CTA first puts the (SiLU'd) input vector(s) into shared memory. Unstaged, every warp (= output row) re-evaluated
// SiLU on all K inputs — 2 MUFU operations per element per row made the modulation GEMVs MUFU-bound (1.75 TB/s of weights)
// instead of HBM-bound. Same values, same FMA order: bit-identical results.
__device__ __forceinline__ void gemv_stage_x(float* gx, const float* __restrict__ x, int64_t ldx, int B, int K, bool silu_in) {
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    const int b = i / K, k = i - b * K;
    const float v = x[b * ldx + k];
    gx[i] = silu_in ? silu_f(v) : v;
  }
  __syncthreads();
}
template <int MAXB, bool STAGED>
__global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ x, int64_t ldx, const void* __restrict__ W,
                                                   int64_t ldw, float* __restrict__ y, int64_t ldy, int B, int N, int K,
                                                   bool silu_in, bool accumulate, bool f16) {
  extern __shared__ float gx[];
  if (STAGED) { gemv_stage_x(gx, x, ldx, B, K, silu_in); x = gx; ldx = K; silu_in = false; }
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  // grid-stride over output rows: a CTA stages the input once and then streams many weight rows (one row per warp at a time)
  for (int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp < N; warp += nwarps) {
  const uint16_t* wr = reinterpret_cast<const uint16_t*>(W) + (int64_t)warp * ldw;
  float acc[MAXB];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
#pragma unroll 4
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wr + k));
    float wf[8];
    {
      float2 t;
      t = unpack2(wv.x, f16); wf[0] = t.x; wf[1] = t.y;
      t = unpack2(wv.y, f16); wf[2] = t.x; wf[3] = t.y;
      t = unpack2(wv.z, f16); wf[4] = t.x; wf[5] = t.y;
      t = unpack2(wv.w, f16); wf[6] = t.x; wf[7] = t.y;
    }
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (b < B) {
        const float4 x0 = *reinterpret_cast<const float4*>(x + b * ldx + k);
        const float4 x1 = *reinterpret_cast<const float4*>(x + b * ldx + k + 4);
        float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xi = silu_in ? silu_f(xv[i]) : xv[i];
          acc[b] = fmaf(xi, wf[i], acc[b]);
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    if (b < B) {
      const float r = warp_sum(acc[b]);
      if (lane == 0) y[b * ldy + warp] = accumulate ? y[b * ldy + warp] + r : r;
    }
  }
  }
}
cudaError_t gemv(const float* x, int64_t ldx, const void* W16, int64_t ldw, float* y, int64_t ldy, int B, int N, int K,
                 bool silu_in, bool accumulate, bool f16, cudaStream_t s) {
  if (B > 8 || K % 8 || ldw % 8 || ldx % 4) return cudaErrorInvalidValue;
  const size_t sm = (size_t)B * K * 4;
  // staged: 4 CTAs per SM keep ~64 KB of weight loads in flight per SM; each CTA amortises its input staging over N / (8 * blocks) rows
  const int blocks = sm <= 48 * 1024 ? std::min((N * 32 + 255) / 256, 148 * 4) : (N * 32 + 255) / 256;
  if (sm <= 48 * 1024) {
    if (B <= 1) gemv_kernel<1, true><<<blocks, 256, sm, s>>>(x, ldx, W16, ldw, y, ldy, B, N, K, silu_in, accumulate, f16);
    else if (B <= 2) gemv_kernel<2, true><<<blocks, 256, sm, s>>>(x, ldx, W16, ldw, y, ldy, B, N, K, silu_in, accumulate, f16);
    else gemv_kernel<8, true><<<blocks, 256, sm, s>>>(x, ldx, W16, ldw, y, ldy, B, N, K, silu_in, accumulate, f16);
  } else if (B <= 1) gemv_kernel<1, false><<<blocks, 256, 0, s>>>(x, ldx, W16, ldw, y, ldy, B, N, K, silu_in, accumulate, f16);
  else if (B <= 2) gemv_kernel<2, false><<<blocks, 256, 0, s>>>(x, ldx, W16, ldw, y, ldy, B, N, K, silu_in, accumulate, f16);
  else gemv_kernel<8, false><<<blocks, 256, 0, s>>>(x, ldx, W16, ldw, y, ldy, B, N, K, silu_in, accumulate, f16);
  return cudaGetLastError();
}

// GEMV over a W-only quantized weight (MLX's packed codes + group scales / biases, modes 1..5 = flux2b_quant): the same loop as
// gemv_kernel — lane = 8 consecutive k per iteration, fp32 FMA chain in the same order — with the eight weights produced in
// registers by dequantize_kernel's arithmetic (quant.cu) and rounded once to the 16-bit operand type, i.e. exactly the numbers
// the dense working copy would hold: the result is bit-identical to gemv over the dequantized matrix.
template <int MAXB, bool STAGED>
__global__ void __launch_bounds__(256) gemv_q_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ codes,
                                                     int64_t row_bytes, const uint8_t* __restrict__ scales,
                                                     const uint8_t* __restrict__ biases, int64_t sb_ld, int mode, int sb_bf16,
                                                     float* __restrict__ y, int64_t ldy, int B, int N, int K, bool silu_in,
                                                     bool accumulate, bool f16) {
  extern __shared__ float gx[];
  if (STAGED) { gemv_stage_x(gx, x, ldx, B, K, silu_in); x = gx; ldx = K; silu_in = false; }
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool eight = mode == 1 || mode == 3;
  const int group = mode <= 2 ? 64 : mode == 5 ? 16 : 32;
  for (int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp < N; warp += nwarps) {
  const uint8_t* cr = codes + (int64_t)warp * row_bytes;
  float acc[MAXB];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
  auto sb = [&](const uint8_t* p, int64_t i) {
    const uint16_t h = __ldg(reinterpret_cast<const uint16_t*>(p) + i);
    return sb_bf16 ? __uint_as_float((uint32_t)h << 16) : __half2float(__ushort_as_half(h));
  };
  for (int k = lane * 8; k < K; k += 256) {
    uint32_t w0, w1 = 0;
    if (eight) { const uint2 q = __ldg(reinterpret_cast<const uint2*>(cr + k)); w0 = q.x; w1 = q.y; }
    else w0 = __ldg(reinterpret_cast<const uint32_t*>(cr + (k >> 1)));
    const int64_t gi = (int64_t)warp * sb_ld + k / group;
    float wf[8];
    if (mode <= 2) {
      const float s = sb(scales, gi), bi = sb(biases, gi);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t q = eight ? (((j < 4 ? w0 : w1) >> (8 * (j & 3))) & 0xffu) : ((w0 >> (4 * j)) & 0xfu);
        wf[j] = __fadd_rn(__fmul_rn((float)q, s), bi);
      }
    } else {
      const uint8_t sbyte = __ldg(scales + gi);
      const float s = mode == 5 ? from_e4m3(sbyte) : from_e8m0(sbyte);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ev = eight ? from_e4m3((uint8_t)(((j < 4 ? w0 : w1) >> (8 * (j & 3))) & 0xffu)) : from_e2m1((uint8_t)((w0 >> (4 * j)) & 0xfu));
        wf[j] = __fmul_rn(ev, s);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) wf[j] = f16 ? __half2float(__float2half_rn(wf[j])) : __bfloat162float(__float2bfloat16(wf[j]));
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (b < B) {
        const float4 x0 = *reinterpret_cast<const float4*>(x + b * ldx + k);
        const float4 x1 = *reinterpret_cast<const float4*>(x + b * ldx + k + 4);
        float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xi = silu_in ? silu_f(xv[i]) : xv[i];
          acc[b] = fmaf(xi, wf[i], acc[b]);
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    if (b < B) {
      const float r = warp_sum(acc[b]);
      if (lane == 0) y[b * ldy + warp] = accumulate ? y[b * ldy + warp] + r : r;
    }
  }
  }
}
cudaError_t gemv_q(const float* x, int64_t ldx, const void* codes, int64_t row_bytes, const void* scales, const void* biases, int64_t sb_ld,
                   int mode, int sb_bf16, float* y, int64_t ldy, int B, int N, int K, bool silu_in, bool accumulate, bool f16, cudaStream_t s) {
  if (B > 8 || K % 16 || row_bytes % 8 || ldx % 4 || mode < 1 || mode > 5 || !scales || (mode <= 2 && !biases)) return cudaErrorInvalidValue;
  const uint8_t *c8 = (const uint8_t*)codes, *s8 = (const uint8_t*)scales, *b8 = (const uint8_t*)biases;
  const size_t sm = (size_t)B * K * 4;
  const int blocks = sm <= 48 * 1024 ? std::min((N * 32 + 255) / 256, 148 * 4) : (N * 32 + 255) / 256;
  if (sm <= 48 * 1024) {
    if (B <= 1) gemv_q_kernel<1, true><<<blocks, 256, sm, s>>>(x, ldx, c8, row_bytes, s8, b8, sb_ld, mode, sb_bf16, y, ldy, B, N, K, silu_in, accumulate, f16);
    else if (B <= 2) gemv_q_kernel<2, true><<<blocks, 256, sm, s>>>(x, ldx, c8, row_bytes, s8, b8, sb_ld, mode, sb_bf16, y, ldy, B, N, K, silu_in, accumulate, f16);
    else gemv_q_kernel<8, true><<<blocks, 256, sm, s>>>(x, ldx, c8, row_bytes, s8, b8, sb_ld, mode, sb_bf16, y, ldy, B, N, K, silu_in, accumulate, f16);
  } else if (B <= 1) gemv_q_kernel<1, false><<<blocks, 256, 0, s>>>(x, ldx, c8, row_bytes, s8, b8, sb_ld, mode, sb_bf16, y, ldy, B, N, K, silu_in, accumulate, f16);
  else if (B <= 2) gemv_q_kernel<2, false><<<blocks, 256, 0, s>>>(x, ldx, c8, row_bytes, s8, b8, sb_ld, mode, sb_bf16, y, ldy, B, N, K, silu_in, accumulate, f16);
  else gemv_q_kernel<8, false><<<blocks, 256, 0, s>>>(x, ldx, c8, row_bytes, s8, b8, sb_ld, mode, sb_bf16, y, ldy, B, N, K, silu_in, accumulate, f16);
  return cudaGetLastError();
}

// Pre-summed weights of the folded Upsample2D (gemm.cuh: conv_up2): w OHWI [Cout, 3, 3, Cin] 16-bit -> [Cout, 16, Cin] with
// tap index (py * 2 + px) * 4 + ty * 2 + tx = sum of w[ky, kx] over ky in R(py, ty), kx in R(px, tx),
// R(0,0) = {0}, R(0,1) = {1,2}, R(1,0) = {0,1}, R(1,1) = {2}: after nearest-2x upsampling, those 3x3 taps land on the same source
// pixel. Summed in fp32, rounded once to the operand type.
__global__ void fold_upsample_weights_kernel(const uint16_t* __restrict__ w, uint16_t* __restrict__ out, int64_t Cout, int Cin, bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 16 * Cin) return;
  const int c = (int)(i % Cin);
  const int tap = (int)((i / Cin) % 16);
  const int64_t o = i / ((int64_t)16 * Cin);
  const int py = tap >> 3, px = (tap >> 2) & 1, ty = (tap >> 1) & 1, tx = tap & 1;
  const int ky0 = (py == 0) ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = (py == 0) ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
  const int kx0 = (px == 0) ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = (px == 0) ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
  float acc = 0.f;
  for (int ky = ky0; ky <= ky1; ++ky)
    for (int kx = kx0; kx <= kx1; ++kx) acc += load16(w, ((o * 3 + ky) * 3 + kx) * Cin + c, f16);
  store16(out, i, acc, f16);
}
cudaError_t fold_upsample_weights(const void* w_ohwi16, void* out16, int64_t Cout, int Cin, bool f16, cudaStream_t s) {
  const int64_t n = Cout * 16 * Cin;
  if (n <= 0) return cudaSuccess;
  fold_upsample_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint16_t*>(w_ohwi16),
                                                                          reinterpret_cast<uint16_t*>(out16), Cout, Cin, f16);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ sinusoid / rope table
__global__ void sinusoid_kernel(const float* __restrict__ t, float* __restrict__ out, int B, float pre_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 128) return;
  const int b = i / 128, j = i % 128;
  const float exponent = (-logf(10000.0f) * (float)j) / 128.0f;
  const float f = expf(exponent);
  const float a = (t[b] * pre_scale) * f;
  out[b * 256 + j] = cosf(a);
  out[b * 256 + 128 + j] = sinf(a);
}
cudaError_t timestep_sinusoid(const float* t, float* out, int B, float pre_scale, cudaStream_t s) {
  sinusoid_kernel<<<(B * 128 + 127) / 128, 128, 0, s>>>(t, out, B, pre_scale);
  return cudaGetLastError();
}

struct RopeAxes { int dims[4]; int off[4]; };
__global__ void rope_table_kernel(const int32_t* __restrict__ ids, int S, RopeAxes ax, float theta,
                                  float* __restrict__ cos_out, float* __restrict__ sin_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, pair)
  const int total_pairs = (ax.off[3] + ax.dims[3]) / 2;
  if (i >= S * total_pairs) return;
  const int row = i / total_pairs, pr = i % total_pairs;
  const int col = pr * 2;
  int a = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (col >= ax.off[k]) a = k;
  const int j2 = col - ax.off[a];  // = 2j
  const float inv_freq = 1.0f / powf(theta, (float)j2 / (float)ax.dims[a]);
  const float ang = (float)ids[row * 4 + a] * inv_freq;
  const float c = cosf(ang), sn = sinf(ang);
  const int W = ax.off[3] + ax.dims[3];
  cos_out[(int64_t)row * W + col] = c;
  cos_out[(int64_t)row * W + col + 1] = c;
  sin_out[(int64_t)row * W + col] = sn;
  sin_out[(int64_t)row * W + col + 1] = sn;
}
cudaError_t rope_table(const int32_t* ids, int S, const int* axes_dims, float theta, float* cos_out, float* sin_out,
                       cudaStream_t s) {
  RopeAxes ax;
  int off = 0;
  for (int i = 0; i < 4; ++i) { ax.dims[i] = axes_dims[i]; ax.off[i] = off; off += axes_dims[i]; }
  const int n = S * (off / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, S, ax, theta, cos_out, sin_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ QK RMSNorm + RoPE (unfused path)
// one warp per (row, head, q|k): 128 elements = 4 per lane (two interleaved pairs)
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(void* __restrict__ qkv, int64_t ld, int rows, int D,
                                                           const float* __restrict__ nq, const float* __restrict__ nk,
                                                           const float* __restrict__ cs, const float* __restrict__ sn,
                                                           float eps, bool f16) {
  const int H = D / 128;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (int64_t)rows * H * 2) return;
  const int which = (int)(wid % 2);
  const int h = (int)((wid / 2) % H);
  const int64_t row = wid / (2 * H);
  uint16_t* p = reinterpret_cast<uint16_t*>(qkv) + row * ld + which * D + h * 128 + lane * 4;
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack2(raw.x, f16), b = unpack2(raw.y, f16);
  float ss = a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss * (1.f / 128.f) + eps);
  const float4 w = __ldg(reinterpret_cast<const float4*>((which ? nk : nq) + lane * 4));
  const float4 c = __ldg(reinterpret_cast<const float4*>(cs + row * 128 + lane * 4));
  const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn + row * 128 + lane * 4));
  const float x0 = a.x * rstd * w.x, x1 = a.y * rstd * w.y, x2 = b.x * rstd * w.z, x3 = b.y * rstd * w.w;
  *reinterpret_cast<uint2*>(p) = make_uint2(pack2(x0 * c.x - x1 * s4.x, x1 * c.y + x0 * s4.y, f16),
                                            pack2(x2 * c.z - x3 * s4.z, x3 * c.w + x2 * s4.w, f16));
}
cudaError_t qk_norm_rope(void* qkv16, int64_t ld, int rows, int D, const float* norm_q, const float* norm_k,
                         const float* cos_t, const float* sin_t, float eps, bool f16, cudaStream_t s) {
  const int64_t warps = (int64_t)rows * (D / 128) * 2;
  const int64_t blocks = (warps * 32 + 255) / 256;
  qk_norm_rope_kernel<<<(unsigned)blocks, 256, 0, s>>>(qkv16, ld, rows, D, norm_q, norm_k, cos_t, sin_t, eps, f16);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ SwiGLU / gate+residual (unfused fallbacks)
__global__ void swiglu_kernel(const void* __restrict__ in, int64_t ldi, void* __restrict__ out, int64_t ldo, int rows,
                              int H, bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 8 outputs per thread
  const int per_row = H / 8;
  if (i >= (int64_t)rows * per_row) return;
  const int64_t r = i / per_row;
  const int c = (int)(i % per_row) * 8;
  const uint16_t* ip = reinterpret_cast<const uint16_t*>(in) + r * ldi + c;
  const uint4 g = *reinterpret_cast<const uint4*>(ip);
  const uint4 u = *reinterpret_cast<const uint4*>(ip + H);
  const uint32_t gg[4] = {g.x, g.y, g.z, g.w}, uu[4] = {u.x, u.y, u.z, u.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = unpack2(gg[k], f16), b = unpack2(uu[k], f16);
    o[k] = pack2(silu_f(a.x) * b.x, silu_f(a.y) * b.y, f16);
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + r * ldo + c) = make_uint4(o[0], o[1], o[2], o[3]);
}
cudaError_t swiglu(const void* in16, int64_t ldi, void* out16, int64_t ldo, int rows, int H, bool f16, cudaStream_t s) {
  if (H % 8 || ldi % 8 || ldo % 8) return cudaErrorInvalidValue;
  const int64_t n = (int64_t)rows * (H / 8);
  swiglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in16, ldi, out16, ldo, rows, H, f16);
  return cudaGetLastError();
}
__global__ void gate_residual_kernel(const void* __restrict__ y, int64_t ldy, const float* __restrict__ gate,
                                     int64_t gbs, int rpb, float* __restrict__ x, int64_t ldx, int rows, int D,
                                     bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 4 per thread
  const int per_row = D / 4;
  if (i >= (int64_t)rows * per_row) return;
  const int64_t r = i / per_row;
  const int c = (int)(i % per_row) * 4;
  const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(y) + r * ldy + c);
  const float2 a = unpack2(raw.x, f16), b = unpack2(raw.y, f16);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (r / rpb) * gbs + c));
  float4 xv = *reinterpret_cast<float4*>(x + r * ldx + c);
  xv.x = fmaf(g.x, a.x, xv.x); xv.y = fmaf(g.y, a.y, xv.y); xv.z = fmaf(g.z, b.x, xv.z); xv.w = fmaf(g.w, b.y, xv.w);
  *reinterpret_cast<float4*>(x + r * ldx + c) = xv;
}
cudaError_t gate_residual(const void* y16, int64_t ldy, const float* gate, int64_t gbs, int rpb, float* x, int64_t ldx,
                          int rows, int D, bool f16, cudaStream_t s) {
  if (D % 4) return cudaErrorInvalidValue;
  const int64_t n = (int64_t)rows * (D / 4);
  gate_residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y16, ldy, gate, gbs, rpb, x, ldx, rows, D, f16);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ conversions
__global__ void f32_to_16_kernel(const float* __restrict__ in, int64_t ldi, void* __restrict__ out, int64_t ldo, int rows,
                                 int cols, bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = (cols + 3) / 4;
  if (i >= (int64_t)rows * per_row) return;
  const int64_t r = i / per_row;
  const int c = (int)(i % per_row) * 4;
  if (c + 4 <= cols && (ldi % 4 == 0) && (ldo % 4 == 0)) {
    const float4 v = *reinterpret_cast<const float4*>(in + r * ldi + c);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + r * ldo + c) = make_uint2(pack2(v.x, v.y, f16), pack2(v.z, v.w, f16));
  } else {
    for (int k = c; k < cols && k < c + 4; ++k) store16(out, r * ldo + k, in[r * ldi + k], f16);
  }
}
cudaError_t f32_to_16(const float* in, int64_t ldi, void* out16, int64_t ldo, int rows, int cols, bool f16,
                      cudaStream_t s) {
  const int64_t n = (int64_t)rows * ((cols + 3) / 4);
  if (n <= 0) return cudaSuccess;
  f32_to_16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, ldi, out16, ldo, rows, cols, f16);
  return cudaGetLastError();
}
__global__ void any16_to_16_kernel(const void* __restrict__ in, int in_f16, void* __restrict__ out, bool f16, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store16(out, i, load16(in, i, in_f16 != 0), f16);
}
cudaError_t any16_to_16(const void* in, int in_is_f16, void* out16, bool f16, int64_t n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  any16_to_16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, in_is_f16, out16, f16, n);
  return cudaGetLastError();
}
__global__ void cvt16_to_f32_kernel(const void* __restrict__ in, float* __restrict__ out, int64_t n, bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = load16(in, i, f16);
}
cudaError_t cvt16_to_f32(const void* in16, float* out, int64_t n, bool f16, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  cvt16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in16, out, n, f16);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ scheduler math on latents
__global__ void euler_kernel(float* __restrict__ x, const float* __restrict__ pred, const float* __restrict__ un,
                             float cfg, float dt, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // separate multiply and add roundings, like the reference's array ops (no FMA contraction): bit-exact with the oracle
  float v = pred[i];
  if (un) { const float u = un[i]; v = __fadd_rn(u, __fmul_rn(cfg, __fsub_rn(v, u))); }
  x[i] = __fadd_rn(x[i], __fmul_rn(dt, v));
}
cudaError_t euler_step(float* x, const float* pred, const float* pred_uncond, float cfg, float dt, int64_t n,
                       cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  euler_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, pred, pred_uncond, cfg, dt, n);
  return cudaGetLastError();
}
__global__ void scale_noise_kernel(const float* __restrict__ a, const float* __restrict__ nz, float sigma,
                                   float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __fadd_rn(__fmul_rn(__fsub_rn(1.f, sigma), a[i]), __fmul_rn(sigma, nz[i]));
}
cudaError_t scale_noise(const float* sample, const float* noise, float sigma, float* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  scale_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sample, noise, sigma, out, n);
  return cudaGetLastError();
}
__global__ void repaint_kernel(float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ e,
                               const float* __restrict__ m, float sn, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float known = __fadd_rn(__fmul_rn(__fsub_rn(1.f, sn), x0[i]), __fmul_rn(sn, e[i]));
  x[i] = __fadd_rn(__fmul_rn(__fsub_rn(1.f, m[i]), known), __fmul_rn(m[i], x[i]));
}
cudaError_t repaint_blend(float* x, const float* x0, const float* eps, const float* mask, float sigma_next, int64_t n,
                          cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  repaint_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, x0, eps, mask, sigma_next, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ latent plumbing
struct PermuteArgs { int ndim; int shape[6]; int64_t istride[6]; };
__global__ void permute_kernel(const float* __restrict__ in, float* __restrict__ out, PermuteArgs a, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t rem = i, off = 0;
#pragma unroll
  for (int d = 5; d >= 0; --d) {
    if (d < a.ndim) {
      const int64_t idx = rem % a.shape[d];
      rem /= a.shape[d];
      off += idx * a.istride[d];
    }
  }
  out[i] = in[off];
}
cudaError_t permute_f32(const float* in, float* out, int ndim, const int* out_shape, const int64_t* in_strides,
                        cudaStream_t s) {
  if (ndim < 1 || ndim > 6) return cudaErrorInvalidValue;
  PermuteArgs a{};
  a.ndim = ndim;
  int64_t n = 1;
  for (int d = 0; d < ndim; ++d) { a.shape[d] = out_shape[d]; a.istride[d] = in_strides[d]; n *= out_shape[d]; }
  if (n <= 0) return cudaSuccess;
  permute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, a, n);
  return cudaGetLastError();
}
__global__ void bn_affine_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean,
                                 const float* __restrict__ var, float eps, int C, int64_t hw, bool denorm, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)((i / hw) % C);
  const float sd = __fsqrt_rn(__fadd_rn(var[c], eps));
  y[i] = denorm ? __fadd_rn(__fmul_rn(x[i], sd), mean[c]) : __fdiv_rn(__fsub_rn(x[i], mean[c]), sd);
}
cudaError_t bn_affine_nchw(const float* x, float* y, const float* mean, const float* var, float eps, int B, int C,
                           int64_t hw, bool denorm, cudaStream_t s) {
  const int64_t n = (int64_t)B * C * hw;
  if (n <= 0) return cudaSuccess;
  bn_affine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, mean, var, eps, C, hw, denorm, n);
  return cudaGetLastError();
}
// seq[b, y*w + x, c*4 + ph*2 + pw] -> out[b, 2y+ph, 2x+pw, c]; one thread per (token, 8 channels of one sub-pixel)
__global__ void seq_to_vae_kernel(const float* __restrict__ seq, const float* __restrict__ mean,
                                  const float* __restrict__ var, float eps, void* __restrict__ out, int B, int h, int w,
                                  bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)B * h * w * 128;
  if (i >= n) return;
  const int ch = (int)(i % 128);
  const int64_t tok = i / 128;
  const int x = (int)(tok % w), y = (int)((tok / w) % h), b = (int)(tok / ((int64_t)w * h));
  const int c = ch / 4, ph = (ch / 2) % 2, pw = ch % 2;
  const float v = seq[i] * sqrtf(var[ch] + eps) + mean[ch];
  const int64_t o = (((int64_t)b * (2 * h) + 2 * y + ph) * (2 * w) + 2 * x + pw) * 32 + c;
  store16(out, o, v, f16);
}
cudaError_t seq_to_vae_input(const float* seq, const float* mean, const float* var, float eps, void* out16, int B, int h,
                             int w, bool f16, cudaStream_t s) {
  const int64_t n = (int64_t)B * h * w * 128;
  if (n <= 0) return cudaSuccess;
  seq_to_vae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seq, mean, var, eps, out16, B, h, w, f16);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ GroupNorm (+SiLU), NHWC
// Two passes over the activation (HBM-bound: 2 reads + 1 write of the tensor), statistics in fp32 / fp64, and a FIXED
// reduction order everywhere (no atomics), so that a decode is bit-reproducible run to run and independent of batch size.
//  pass 1 (gn_stats_kernel): grid (chunks, B), blockDim = a multiple of C/8. Thread t owns channel slot t % (C/8)
//    (8 consecutive channels = one 16 B vector) for a strided set of pixels and keeps 8 (sum, sumsq) pairs in registers;
//    the CTA folds threads of equal slot, then channels of equal group, in index order and writes one partial per group.
//  pass 1b (tail of gn_stats_kernel, run by the last-arriving CTA of a batch item): sums the per-CTA partials in index order
//  (fp64) -> mean / rstd per (batch, group).
//  pass 2 (gn_apply_kernel): normalise, affine, optional SiLU.
__global__ void __launch_bounds__(512) gn_stats_kernel(const void* __restrict__ x, float* __restrict__ partial, int64_t HW,
                                                       int C, int G, bool f16, double* __restrict__ stats,
                                                       unsigned int* __restrict__ counter, double count, float eps) {
  extern __shared__ float gsm[];  // [blockDim][16] thread accumulators, then [2*C] channel sums
  const int b = blockIdx.y;
  const int vpp = C / 8;
  const int cg = C / G;
  const uint16_t* xb = reinterpret_cast<const uint16_t*>(x) + (int64_t)b * HW * C;
  const int64_t nvec = HW * vpp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;  // multiple of vpp: a thread never changes channel slot
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; }
  auto accumulate = [&](const uint4& v) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2(u[k], f16);
      s[2 * k] += f.x; q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
      s[2 * k + 1] += f.y; q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
    }
  };
  // four independent 16 B loads in flight per thread (one load per thread leaves the SM ~16 KB short of the ~35 KB in flight
  // that HBM latency x bandwidth needs); the accumulation order per thread is unchanged (ascending i)
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(xb + (i + j * stride) * 8));
#pragma unroll
    for (int j = 0; j < 4; ++j) accumulate(v[j]);
  }
  for (; i < nvec; i += stride) accumulate(__ldg(reinterpret_cast<const uint4*>(xb + i * 8)));
  float* acc = gsm;
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc[threadIdx.x * 16 + k] = s[k]; acc[threadIdx.x * 16 + 8 + k] = q[k]; }
  __syncthreads();
  // channel sums: channel c = slot*8 + k is held by threads slot, slot + vpp, slot + 2 vpp, ...
  float cs = 0.f, cq = 0.f;
  const int reps = blockDim.x / vpp;
  // each thread folds at most ceil(C / blockDim) channels; C <= 2 * blockDim is enforced by the launcher
  float ch_s[2] = {0.f, 0.f}, ch_q[2] = {0.f, 0.f};
  int n_mine = 0;
  for (int cidx = threadIdx.x; cidx < C; cidx += blockDim.x, ++n_mine) {
    const int slot = cidx >> 3, k = cidx & 7;
    cs = 0.f; cq = 0.f;
    for (int m = 0; m < reps; ++m) {
      const int t = slot + m * vpp;
      cs += acc[t * 16 + k];
      cq += acc[t * 16 + 8 + k];
    }
    ch_s[n_mine] = cs; ch_q[n_mine] = cq;
  }
  __syncthreads();
  n_mine = 0;
  for (int cidx = threadIdx.x; cidx < C; cidx += blockDim.x, ++n_mine) { gsm[cidx] = ch_s[n_mine]; gsm[C + cidx] = ch_q[n_mine]; }
  __syncthreads();
  if (threadIdx.x < G) {
    float gs = 0.f, gq = 0.f;
    for (int j = 0; j < cg; ++j) { gs += gsm[threadIdx.x * cg + j]; gq += gsm[C + threadIdx.x * cg + j]; }
    float* dst = partial + ((int64_t)b * gridDim.x + blockIdx.x) * 2 * G;
    dst[threadIdx.x] = gs;
    dst[G + threadIdx.x] = gq;
  }
  // ---- finalize in the LAST-ARRIVING CTA of this batch item (was a launch of its own: 9 us of a 66 us norm). The fold order is
  // fixed by partial index, not by arrival: sub-lane l of statistic j sums partials l, l + lanes, ... in index order (fp64), then
  // thread g folds the sub-lanes in index order, so the result does not depend on which CTA happens to be last.
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&counter[b], 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* sm = reinterpret_cast<double*>(gsm);   // blockDim x 16 floats = blockDim x 8 doubles >= lanes x 2G
  const int chunks = gridDim.x;
  const int nstat = 2 * G;
  const int lanes = blockDim.x / nstat;          // >= 1: the launcher keeps blockDim >= 2G
  const int j = threadIdx.x % nstat, l = threadIdx.x / nstat;
  if (l < lanes) {
    double a = 0.0;
    const float* src = partial + (int64_t)b * chunks * nstat + j;
    int c = l;
    for (; c + 7 * lanes < chunks; c += 8 * lanes) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = __ldcg(src + (int64_t)(c + k * lanes) * nstat);
#pragma unroll
      for (int k = 0; k < 8; ++k) a += (double)t[k];
    }
    for (; c < chunks; c += lanes) a += (double)__ldcg(src + (int64_t)c * nstat);
    sm[l * nstat + j] = a;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    double sg = 0.0, qg = 0.0;
    for (int k = 0; k < lanes; ++k) { sg += sm[k * nstat + threadIdx.x]; qg += sm[k * nstat + G + threadIdx.x]; }
    const double m = sg / count;
    const double var = qg / count - m * m;
    stats[((int64_t)b * G + threadIdx.x) * 2] = m;
    stats[((int64_t)b * G + threadIdx.x) * 2 + 1] = (double)rsqrtf(fmaxf((float)var, 0.f) + eps);
  }
  if (threadIdx.x == 0) counter[b] = 0;   // ready for the next launch
}
__global__ void __launch_bounds__(256) gn_apply_kernel(const void* __restrict__ x, void* __restrict__ y,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const double* __restrict__ stats, int64_t HW, int C, int G,
                                                       bool silu, bool f16) {
  const int b = blockIdx.y;
  const int vpp = C / 8;
  const int cg = C / G;
  const int64_t nvec = HW * vpp;
  const uint16_t* xb = reinterpret_cast<const uint16_t*>(x) + (int64_t)b * HW * C;
  uint16_t* yb = reinterpret_cast<uint16_t*>(y) + (int64_t)b * HW * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;  // multiple of vpp: per-thread scale / shift are loop invariant
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= nvec) return;
  const int c0 = (int)(i0 % vpp) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    const int g = c / cg;
    const float m = (float)stats[((int64_t)b * G + g) * 2];
    const float rstd = (float)stats[((int64_t)b * G + g) * 2 + 1];
    const float ga = __ldg(gamma + c);
    sc[k] = rstd * ga;
    sh[k] = __ldg(beta + c) - m * rstd * ga;
  }
  auto apply = [&](const uint4& v, int64_t i) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2(u[k], f16);
      float z0 = fmaf(f.x, sc[2 * k], sh[2 * k]);
      float z1 = fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]);
      if (silu) { z0 = silu_f(z0); z1 = silu_f(z1); }
      o[k] = pack2(z0, z1, f16);
    }
    *reinterpret_cast<uint4*>(yb + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  // Back to front: the statistics pass has just streamed the tensor in ascending order, so its tail is what the L2 still holds
  // when the tensor is larger than the cache (the 1024^2 layers are 100 - 200 MB against 126 MB of L2). Four loads in flight.
  const int64_t n_it = (nvec - i0 + stride - 1) / stride;   // iterations of this thread
  int64_t k = n_it - 1;
  for (; k >= 3; k -= 4) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(xb + (i0 + (k - j) * stride) * 8));
#pragma unroll
    for (int j = 0; j < 4; ++j) apply(v[j], i0 + (k - j) * stride);
  }
  for (; k >= 0; --k) apply(__ldg(reinterpret_cast<const uint4*>(xb + (i0 + k * stride) * 8)), i0 + k * stride);
}
static constexpr int GN_MAX_CHUNKS = 148 * 2;
size_t groupnorm_ws_bytes(int B, int G) {
  // [B, G] x (mean, rstd) doubles, the per-CTA partials [B, <= 296 chunks, 2G] floats, and one arrival counter per batch item
  // (zero when idle: the buffer must be zero-filled once after allocation, ensure_zeroed in ctx.h)
  return sizeof(double) * 2 * G * B + sizeof(float) * 2 * G * B * GN_MAX_CHUNKS + sizeof(unsigned int) * B;
}
cudaError_t groupnorm_silu(const void* x16, void* y16, const float* gamma, const float* beta, double* stats_ws, int B,
                           int64_t HW, int C, int G, float eps, bool silu, bool f16, cudaStream_t s) {
  const int vpp = C / 8;
  if (C % 8 || C % G || vpp > 256 || G < 1 || G > 128) return cudaErrorInvalidValue;
  // statistics pass: up to 512 threads (a multiple of the channel-slot count) and at most two CTAs per SM, so that the fold in the
  // last-arriving CTA is over <= 296 partials with >= 8 sub-lanes per statistic (~1 us instead of a 9 us launch)
  const int threads_s = (512 / vpp) * vpp;
  const int threads = (256 / vpp) * vpp;  // apply pass: largest multiple of the channel-slot count <= 256
  if (C > 2 * threads || threads_s < 2 * G) return cudaErrorInvalidValue;  // the CTA fold gives every thread at most two channels
  const int64_t nvec = HW * vpp;
  float* partial = reinterpret_cast<float*>(stats_ws + (size_t)2 * G * B);
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + (size_t)2 * G * B * GN_MAX_CHUNKS);
  const int chunks = (int)std::min<int64_t>((nvec + threads_s - 1) / threads_s, GN_MAX_CHUNKS);
  const size_t smem = std::max((size_t)threads_s * 16, (size_t)2 * C) * sizeof(float);
  gn_stats_kernel<<<dim3(chunks, B), threads_s, smem, s>>>(x16, partial, HW, C, G, f16, stats_ws, counter, (double)HW * (C / G), eps);
  const int chunks2 = (int)std::min<int64_t>((nvec + threads - 1) / threads, 148 * 8);
  gn_apply_kernel<<<dim3(chunks2, B), threads, 0, s>>>(x16, y16, gamma, beta, stats_ws, HW, C, G, silu, f16);
  return cudaGetLastError();
}

__global__ void upsample2x_kernel(const void* __restrict__ x, void* __restrict__ y, int B, int H, int W, int C) {
  const int vpp = C / 8;
  const int64_t n = (int64_t)B * (2 * H) * (2 * W) * vpp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    const int64_t pix = i / vpp;
    const int ox = (int)(pix % (2 * W)), oy = (int)((pix / (2 * W)) % (2 * H)), b = (int)(pix / ((int64_t)4 * W * H));
    const int64_t src = (((int64_t)b * H + oy / 2) * W + ox / 2) * vpp + v;
    reinterpret_cast<uint4*>(y)[i] = __ldg(reinterpret_cast<const uint4*>(x) + src);
  }
}
cudaError_t upsample_nearest2x(const void* x16, void* y16, int B, int H, int W, int C, cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  const int64_t n = (int64_t)B * 4 * H * W * (C / 8);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 32);
  upsample2x_kernel<<<blocks, 256, 0, s>>>(x16, y16, B, H, W, C);
  return cudaGetLastError();
}

// one CTA per row: max, sum(exp), write probabilities
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, int64_t ldx, void* __restrict__ y,
                                                           int64_t ldy, int cols, float scale, bool f16) {
  __shared__ float sm[8];
  const float* xr = x + (int64_t)blockIdx.x * ldx;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) mx = fmaxf(mx, xr[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = sm[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, sm[i]);
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) s += __expf((xr[c] - mx) * scale);
  const float tot = block_sum<256>(s, sm);
  const float inv = 1.f / tot;
  for (int c = threadIdx.x; c < cols; c += 256)
    store16(y, (int64_t)blockIdx.x * ldy + c, __expf((xr[c] - mx) * scale) * inv, f16);
}
cudaError_t softmax_rows(const float* x, int64_t ldx, void* y16, int64_t ldy, int rows, int cols, float scale, bool f16,
                         cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  softmax_rows_kernel<<<rows, 256, 0, s>>>(x, ldx, y16, ldy, cols, scale, f16);
  return cudaGetLastError();
}

__global__ void transpose16_kernel(const uint16_t* __restrict__ in, int64_t ldi, uint16_t* __restrict__ out, int64_t ldo,
                                   int rows, int cols) {
  __shared__ uint16_t tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = by + j, c = bx + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[(int64_t)r * ldi + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = bx + j, r = by + threadIdx.x;
    if (r < rows && c < cols) out[(int64_t)c * ldo + r] = tile[threadIdx.x][j];
  }
}
cudaError_t transpose16(const void* in16, int64_t ldi, void* out16, int64_t ldo, int rows, int cols, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose16_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const uint16_t*>(in16), ldi,
                                            reinterpret_cast<uint16_t*>(out16), ldo, rows, cols);
  return cudaGetLastError();
}
__global__ void add16_kernel(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ o, int64_t n,
                             bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store16(o, i, load16(a, i, f16) + load16(b, i, f16), f16);
}
cudaError_t add16(const void* a16, const void* b16, void* out16, int64_t n, bool f16, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  add16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a16, b16, out16, n, f16);
  return cudaGetLastError();
}

__global__ void postprocess_u8_kernel(const void* __restrict__ x, int64_t ldc, uint8_t* __restrict__ out, int64_t npix,
                                      bool f16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * 3) return;
  const int64_t p = i / 3;
  const int c = (int)(i % 3);
  float v = (load16(x, p * ldc + c, f16) + 1.0f) * 127.5f;
  v = fminf(fmaxf(v, 0.f), 255.f);
  out[i] = (uint8_t)v;  // truncation, as MLX asType(.uint8)
}
cudaError_t postprocess_u8(const void* x16, int64_t ldc, uint8_t* out, int64_t npix, bool f16, cudaStream_t s) {
  const int64_t n = npix * 3;
  if (n <= 0) return cudaSuccess;
  postprocess_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x16, ldc, out, npix, f16);
  return cudaGetLastError();
}
__global__ void nhwc16_to_nchw_kernel(const void* __restrict__ x, int64_t ldc, float* __restrict__ out, int64_t HW, int C,
                                      bool f16, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = i % HW;
  const int c = (int)((i / HW) % C);
  const int64_t b = i / (HW * C);
  out[i] = load16(x, (b * HW + p) * ldc + c, f16);
}
cudaError_t nhwc16_to_nchw_f32(const void* x16, int64_t ldc, float* out, int B, int64_t HW, int C, bool f16,
                               cudaStream_t s) {
  const int64_t n = (int64_t)B * HW * C;
  if (n <= 0) return cudaSuccess;
  nhwc16_to_nchw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x16, ldc, out, HW, C, f16, n);
  return cudaGetLastError();
}
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ in, void* __restrict__ out, int64_t HW, int C, bool f16,
                                      int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // output index
  if (i >= n) return;
  const int c = (int)(i % C);
  const int64_t p = (i / C) % HW;
  const int64_t b = i / (HW * C);
  store16(out, i, in[(b * C + c) * HW + p], f16);
}
cudaError_t nchw_f32_to_nhwc16(const float* in, void* out16, int B, int64_t HW, int C, bool f16, cudaStream_t s) {
  const int64_t n = (int64_t)B * HW * C;
  if (n <= 0) return cudaSuccess;
  nchw_to_nhwc16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out16, HW, C, f16, n);
  return cudaGetLastError();
}

// image NCHW fp32 [B, C, HW] -> NHWC 16-bit [B, HW, Cpad] with channels [C, Cpad) zero (VAE encoder input, VAEEncoder.swift:88)
__global__ void nchw_to_nhwc16_pad_kernel(const float* __restrict__ in, void* __restrict__ out, int64_t HW, int C, int Cpad, bool f16,
                                          int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % Cpad);
  const int64_t p = (i / Cpad) % HW, b = i / (Cpad * HW);
  store16(out, i, c < C ? in[(b * C + c) * HW + p] : 0.f, f16);
}
cudaError_t nchw_f32_to_nhwc16_pad(const float* in, void* out16, int B, int64_t HW, int C, int Cpad, bool f16, cudaStream_t s) {
  const int64_t n = (int64_t)B * HW * Cpad;
  if (n <= 0) return cudaSuccess;
  nchw_to_nhwc16_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out16, HW, C, Cpad, f16, n);
  return cudaGetLastError();
}
// posterior moments NHWC 16-bit [B, HW, 2L] = [mean | logvar] -> latent NCHW fp32 [B, L, HW]:
// mean, or mean + exp(0.5 * logvar) * noise when noise (NCHW fp32) is given (AutoencoderKL.swift:101-111)
__global__ void moments_to_latent_kernel(const void* __restrict__ m16, int64_t ldc, const float* __restrict__ noise, float* __restrict__ out,
                                         int64_t HW, int L, bool f16, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = i % HW; const int c = (int)((i / HW) % L); const int64_t b = i / (HW * L);
  const int64_t base = (b * HW + p) * ldc;
  float v = load16(m16, base + c, f16);
  if (noise) v = __fadd_rn(v, __fmul_rn(expf(__fmul_rn(0.5f, load16(m16, base + L + c, f16))), noise[i]));
  out[i] = v;
}
cudaError_t moments_to_latent(const void* m16, int64_t ldc, const float* noise, float* out, int B, int64_t HW, int L, bool f16,
                              cudaStream_t s) {
  const int64_t n = (int64_t)B * L * HW;
  if (n <= 0) return cudaSuccess;
  moments_to_latent_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(m16, ldc, noise, out, HW, L, f16, n);
  return cudaGetLastError();
}

}  // namespace f2b
