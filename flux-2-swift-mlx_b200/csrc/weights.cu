// weights.cu — turns the tensors handed over under the reference's flattened Swift keys
// (Loading/WeightLoader.swift:99-204,567-612; PrequantizedCheckpoint.swift:41-59) into the fused / re-tiled 16-bit
// working copies the kernels consume. The external layout is never changed: get_tensor returns what MLX would hold.
//   * double blocks: toQ|toK|toV (and addQ|addK|addV) are stacked into one [3D, D] operand,
//   * SwiGLU producers (ff.activation.proj, the gate|up slab of toQkvMlp) get their rows interleaved per 256-row
//     tile as [128 gate | 128 value] so the GEMM epilogue can apply silu(gate)*value in registers,
//   * quantized layers are quantized exactly like quantize(model:) (Flux2Pipeline.swift:567-578) and kept in MLX's
//     packed form; the dense working copy is their dequantization (W-only path: x · dequant(W)^T).
#include <cstring>

#include "ctx.h"
#include "ptx.cuh"

namespace f2b {

__global__ void copy_rows16_kernel(const uint16_t* __restrict__ src, int64_t src_ld, int64_t src_row0,
                                   uint16_t* __restrict__ dst, int64_t dst_ld, int64_t dst_row0, int64_t nrows,
                                   int64_t ncols, int tiled, int64_t Hm) {
  const int64_t vec_per_row = ncols / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * vec_per_row) return;
  const int64_t r = i / vec_per_row, v = i % vec_per_row;
  int64_t sr = r;
  if (tiled) {
    const int64_t tile = r / 256, j = r % 256;
    sr = (j < 128) ? tile * 128 + j : Hm + tile * 128 + (j - 128);
  }
  const uint4 val = *reinterpret_cast<const uint4*>(src + (src_row0 + sr) * src_ld + v * 8);
  *reinterpret_cast<uint4*>(dst + (dst_row0 + r) * dst_ld + v * 8) = val;
}
static int copy_rows16(flux2b_ctx* c, const void* src, int64_t src_ld, int64_t src_row0, void* dst, int64_t dst_ld,
                       int64_t dst_row0, int64_t nrows, int64_t ncols, bool tiled, int64_t Hm) {
  if (ncols % 8) return fail(FLUX2B_ERR_WEIGHT_LOADING, "weight inner dimension must be a multiple of 8");
  const int64_t n = nrows * (ncols / 8);
  copy_rows16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
      reinterpret_cast<const uint16_t*>(src), src_ld, src_row0, reinterpret_cast<uint16_t*>(dst), dst_ld, dst_row0, nrows,
      ncols, tiled ? 1 : 0, Hm);
  F2B_CUDA(cudaGetLastError());
  return 0;
}

__global__ void to16_kernel(const void* __restrict__ src, int src_dtype, void* __restrict__ dst, int f16, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v;
  if (src_dtype == FLUX2B_F32) v = reinterpret_cast<const float*>(src)[i];
  else if (src_dtype == FLUX2B_F16) v = __half2float(reinterpret_cast<const __half*>(src)[i]);
  else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
  if (f16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16(v);
}
__global__ void to_f32_kernel(const void* __restrict__ src, int src_dtype, float* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v;
  if (src_dtype == FLUX2B_F32) v = reinterpret_cast<const float*>(src)[i];
  else if (src_dtype == FLUX2B_F16) v = __half2float(reinterpret_cast<const __half*>(src)[i]);
  else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
  dst[i] = v;
}

static Tensor* find(flux2b_ctx* c, const std::string& key) {
  auto it = c->tensors.find(key);
  return it == c->tensors.end() ? nullptr : &it->second;
}
static bool is_float_dtype(int dt) { return dt == FLUX2B_F32 || dt == FLUX2B_F16 || dt == FLUX2B_BF16_T; }
static bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

int vector_f32_from_key(flux2b_ctx* c, const std::string& key, DevBuf* out, int64_t expect, bool required, float fill) {
  Tensor* t = find(c, key);
  if (!t) {
    if (required) return fail(FLUX2B_ERR_WEIGHT_LOADING, "missing tensor: " + key);
    std::vector<float> h((size_t)expect, fill);
    F2B_CUDA(out->alloc(sizeof(float) * expect));
    F2B_CUDA(cudaMemcpyAsync(out->p, h.data(), sizeof(float) * expect, cudaMemcpyHostToDevice, c->stream));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  }
  if (!is_float_dtype(t->dtype) || t->numel() != expect)
    return fail(FLUX2B_ERR_WEIGHT_LOADING, "bad dtype / size for " + key);
  F2B_CUDA(out->alloc(sizeof(float) * expect));
  to_f32_kernel<<<(unsigned)((expect + 255) / 256), 256, 0, c->stream>>>(t->buf.p, t->dtype, out->as<float>(), expect);
  F2B_CUDA(cudaGetLastError());
  return 0;
}

// A packed Linear and its scale / bias tensors, validated against the context's quantization mode. The loader accepts the
// float-category tensors in any float type (PrequantizedCheckpoint.swift:41-59; MLX-quantized bf16 models store bf16 scales),
// so the type is part of the metadata and every consumer dequantizes with it — a bf16 scale read as f16 is garbage.
int packed_meta(flux2b_ctx* c, const std::string& base, PackedMeta* m) {
  Tensor* w = find(c, base + ".weight");
  if (!w) return fail(FLUX2B_ERR_WEIGHT_LOADING, "missing tensor: " + base + ".weight");
  int sdt;
  if (!quant_params(c->quant, &m->bits, &m->group, &m->has_b, &sdt))
    return fail(FLUX2B_ERR_WEIGHT_LOADING, "packed weight for " + base + " but context quantization is bf16");
  if (w->dtype != FLUX2B_U32 || w->shape.size() != 2) return fail(FLUX2B_ERR_WEIGHT_LOADING, "packed weight must be 2-D uint32: " + base);
  m->rows = w->shape[0]; m->cols = w->shape[1] * 32 / m->bits;
  if (m->cols % m->group) return fail(FLUX2B_ERR_WEIGHT_LOADING, "input dim not divisible by group size: " + base);
  Tensor* s = find(c, base + ".scales");
  Tensor* b = find(c, base + ".biases");
  if (!s || (m->has_b && !b)) return fail(FLUX2B_ERR_WEIGHT_LOADING, "missing scales / biases for " + base);
  if (!m->has_b && b) return fail(FLUX2B_ERR_WEIGHT_LOADING, "unexpected biases for non-affine mode: " + base);
  const int64_t groups = m->rows * (m->cols / m->group);
  if (s->numel() != groups || (b && b->numel() != groups)) return fail(FLUX2B_ERR_WEIGHT_LOADING, "scales / biases shape mismatch: " + base);
  if (m->has_b) {
    if (!is_float_dtype(s->dtype) || b->dtype != s->dtype)
      return fail(FLUX2B_ERR_WEIGHT_LOADING, "affine scales / biases must share one float type (f16, bf16 or f32): " + base);
    m->sb_dtype = s->dtype;
  } else {
    if (s->dtype != FLUX2B_U8) return fail(FLUX2B_ERR_WEIGHT_LOADING, "block-scaled modes store one uint8 scale per group: " + base);
    m->sb_dtype = FLUX2B_F16;  // unused
  }
  m->w = w; m->s = s; m->b = b;
  return 0;
}

// Linear `base` held as float -> MLX's packed form in place (weight uint32, scales, biases), exactly like quantize(model:)
// (Flux2Pipeline.swift:567-578). Already packed layers are left alone (docs/knowledge/pitfalls/quantize-skips-quantized.md).
int ensure_packed(flux2b_ctx* c, const std::string& base) {
  Tensor* w = find(c, base + ".weight");
  if (!w) return fail(FLUX2B_ERR_WEIGHT_LOADING, "missing tensor: " + base + ".weight");
  if (w->dtype == FLUX2B_U32) return 0;
  if (!is_float_dtype(w->dtype)) return fail(FLUX2B_ERR_WEIGHT_LOADING, "unsupported weight dtype for " + base);
  int bits, group, has_b, sdt;
  if (!quant_params(c->quant, &bits, &group, &has_b, &sdt)) return fail(FLUX2B_ERR_WEIGHT_LOADING, "context quantization is bf16");
  const int64_t rows = w->shape.empty() ? 0 : w->shape[0];
  const int64_t cols = rows ? w->numel() / rows : 0;
  if (cols % group) return fail(FLUX2B_ERR_WEIGHT_LOADING, "input dim not divisible by group size: " + base);
  Tensor packed, scales, biases;
  packed.dtype = FLUX2B_U32; packed.shape = {rows, cols * bits / 32};
  scales.dtype = sdt; scales.shape = {rows, cols / group};
  biases.dtype = FLUX2B_F16; biases.shape = {rows, cols / group};
  F2B_CUDA(packed.buf.alloc((size_t)packed.numel() * 4));
  F2B_CUDA(scales.buf.alloc((size_t)scales.numel() * dtype_size(sdt)));
  if (has_b) F2B_CUDA(biases.buf.alloc((size_t)biases.numel() * 2));
  F2B_CUDA(quantize_matrix(c->quant, w->buf.p, w->dtype, rows, cols, packed.buf.as<uint32_t>(), scales.buf.p,
                           has_b ? biases.buf.p : nullptr, c->stream));
  F2B_CUDA(cudaStreamSynchronize(c->stream));  // the float weight is released below
  c->tensors[base + ".weight"] = std::move(packed);
  c->tensors[base + ".scales"] = std::move(scales);
  if (has_b) c->tensors[base + ".biases"] = std::move(biases);
  return 0;
}

// Dense [N, K] working copy (compute dtype) of Linear `base` — quantizing on the fly / dequantizing as configured.
// quantize_ok=false keeps the layer dense (VAE linears / convs are never quantized by the reference).
int dense16_from_key_ex(flux2b_ctx* c, const std::string& base, DevBuf* out, int* N, int* K, bool quantize_ok) {
  Tensor* w = find(c, base + ".weight");
  if (!w) return fail(FLUX2B_ERR_WEIGHT_LOADING, "missing tensor: " + base + ".weight");
  const bool f16 = c->f16();
  if (w->dtype == FLUX2B_U32) {
    PackedMeta m;
    F2B_TRY(packed_meta(c, base, &m));
    F2B_CUDA(out->alloc((size_t)m.rows * m.cols * 2));
    F2B_CUDA(dequantize_matrix(c->quant, m.w->buf.as<uint32_t>(), m.s->buf.p, m.has_b ? m.b->buf.p : nullptr, m.rows, m.cols, out->p,
                               f16 ? FLUX2B_F16 : FLUX2B_BF16_T, c->stream, m.sb_dtype));
    *N = (int)m.rows; *K = (int)m.cols;
    return 0;
  }
  if (!is_float_dtype(w->dtype)) return fail(FLUX2B_ERR_WEIGHT_LOADING, "unsupported weight dtype for " + base);
  int64_t rows = w->shape.empty() ? 0 : w->shape[0];
  int64_t cols = rows ? w->numel() / rows : 0;
  if (quantize_ok && c->quant != FLUX2B_BF16) {
    F2B_TRY(ensure_packed(c, base));
    return dense16_from_key_ex(c, base, out, N, K, quantize_ok);  // now takes the packed branch above
  }
  F2B_CUDA(out->alloc((size_t)rows * cols * 2));
  const int64_t n = rows * cols;
  to16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(w->buf.p, w->dtype, out->p, f16 ? 1 : 0, n);
  F2B_CUDA(cudaGetLastError());
  *N = (int)rows; *K = (int)cols;
  return 0;
}
int dense16_from_key(flux2b_ctx* c, const std::string& base, DevBuf* out, int* N, int* K) {
  return dense16_from_key_ex(c, base, out, N, K, true);
}

static int expect_shape(const std::string& base, int N, int K, int eN, int eK) {
  if (N != eN || K != eK)
    return fail(FLUX2B_ERR_WEIGHT_LOADING, "shape mismatch for " + base + ": got [" + std::to_string(N) + "," +
                                               std::to_string(K) + "] expected [" + std::to_string(eN) + "," +
                                               std::to_string(eK) + "]");
  return 0;
}

static bool wq_eligible(flux2b_ctx* c, const std::string& base, int eK);
static int wq_rows(flux2b_ctx* c, const std::string& base, int eN, int eK, int64_t src_row0, int64_t nrows, Lin* L, int N_total,
                   int64_t dst_row0, bool tiled, int Hm);
static void lin_dense_reset(Lin* L) { L->wq.release(); L->ws.release(); L->wb.release(); L->sfb.release(); L->wmode = 0; L->mx = 0; }
static int build_lin(flux2b_ctx* c, const std::string& base, Lin* L, int eN, int eK) {
  if (c->has_dit && wq_eligible(c, base, eK)) return wq_rows(c, base, eN, eK, 0, eN, L, eN, 0, false, 0);
  lin_dense_reset(L);
  int N, K;
  F2B_TRY(dense16_from_key(c, base, &L->w, &N, &K));
  F2B_TRY(expect_shape(base, N, K, eN, eK));
  L->N = N; L->K = K;
  return 0;
}
// stack several Linears row-wise into one operand
static int build_stacked(flux2b_ctx* c, const std::vector<std::string>& bases, Lin* L, int eN_each, int eK) {
  F2B_CUDA(L->w.alloc((size_t)bases.size() * eN_each * eK * 2));
  for (size_t i = 0; i < bases.size(); ++i) {
    DevBuf tmp; int N, K;
    F2B_TRY(dense16_from_key(c, bases[i], &tmp, &N, &K));
    F2B_TRY(expect_shape(bases[i], N, K, eN_each, eK));
    F2B_TRY(copy_rows16(c, tmp.p, K, 0, L->w.p, K, (int64_t)i * eN_each, N, K, false, 0));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
  }
  L->N = (int)bases.size() * eN_each; L->K = eK;
  return 0;
}
// SwiGLU producer [2*Hm, K] = [gate | value] rows -> per-tile interleave (when Hm % 128 == 0 and fusion is on)
static int build_swiglu(flux2b_ctx* c, const void* dense, int64_t ld, int64_t row0, Lin* L, int Hm, int K, bool* tiled) {
  *tiled = (Hm % 128 == 0) && c->option("fuse_swiglu", 1);
  F2B_CUDA(L->w.alloc((size_t)2 * Hm * K * 2));
  F2B_TRY(copy_rows16(c, dense, ld, row0, L->w.p, K, 0, 2 * (int64_t)Hm, K, *tiled, Hm));
  L->N = 2 * Hm; L->K = K;
  return 0;
}

// native block-scaled operand: rows [src_row0, src_row0 + nrows) of the packed Linear `base` (shape [eN, eK]) become rows
// [dst_row0, dst_row0 + nrows) of L (N_total rows, allocated on first use); the bytes are MLX's, only the row order (fusion /
// SwiGLU tile interleave) and the scale-factor tiling change.
static int mx_rows(flux2b_ctx* c, const std::string& base, int eN, int eK, int64_t src_row0, int64_t nrows, Lin* L, int N_total,
                   int64_t dst_row0, bool tiled, int Hm) {
  const int kind = c->mx_kind;
  // N tile of the block-scaled GEMM: 128 keeps two accumulator stages next to the scale-factor columns in TMEM
  const int bn = (c->option("mx_bn", 0) == 256 && N_total % 256 == 0) ? 256 : 128;
  F2B_TRY(ensure_packed(c, base));
  PackedMeta pm;
  F2B_TRY(packed_meta(c, base, &pm));
  Tensor* w = pm.w; Tensor* s = pm.s;
  const int bits = pm.bits;
  F2B_TRY(expect_shape(base, (int)pm.rows, (int)pm.cols, eN, eK));
  if (eK % (kind == 1 ? 128 : 256) || N_total % 128)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "native_mx needs in-features % 128 (fp8) / 256 (fp4) == 0 and out-features % 128 == 0: " + base);
  if (!L->wq.p || L->N != N_total || L->K != eK || L->mx != kind || L->bn != bn) {
    F2B_CUDA(L->wq.alloc((size_t)N_total * eK * bits / 8));
    F2B_CUDA(L->sfb.alloc(mx_sf_bytes(kind, N_total, eK)));
    L->w.release(); L->ws.release(); L->wb.release(); L->wmode = 0;
    L->N = N_total; L->K = eK; L->mx = kind; L->bn = bn;
  }
  if (tiled && (nrows % bn || Hm % (bn / 2))) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "SwiGLU tile does not divide the MLP width: " + base);
  F2B_CUDA(mx_copy_rows(kind, w->buf.as<uint8_t>(), s->buf.as<uint8_t>(), src_row0, L->wq.as<uint8_t>(), L->sfb.as<uint8_t>(), dst_row0,
                        nrows, eK, tiled ? bn : 0, Hm, c->stream));
  return 0;
}

// ---- W-only quantized working copies (dequantized inside the GEMM / GEMV kernels)
// rows of `src` (row_bytes each) -> rows [dst_row0, ...) of `dst`, optionally through the SwiGLU 256-row tile interleave
__global__ void copy_row_bytes_kernel(const uint8_t* __restrict__ src, int64_t src_row0, uint8_t* __restrict__ dst, int64_t dst_row0,
                                      int64_t nrows, int64_t row_bytes, int unit, int tiled, int64_t Hm) {
  const int64_t per_row = row_bytes / unit;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * per_row) return;
  const int64_t r = i / per_row, v = i % per_row;
  int64_t sr = r;
  if (tiled) {
    const int64_t tile = r / 256, j = r % 256;
    sr = (j < 128) ? tile * 128 + j : Hm + tile * 128 + (j - 128);
  }
  const uint8_t* s = src + (src_row0 + sr) * row_bytes + v * unit;
  uint8_t* d = dst + (dst_row0 + r) * row_bytes + v * unit;
  if (unit == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
  else if (unit == 2) *reinterpret_cast<uint16_t*>(d) = *reinterpret_cast<const uint16_t*>(s);
  else *d = *s;
}
static int copy_row_bytes(flux2b_ctx* c, const void* src, int64_t src_row0, void* dst, int64_t dst_row0, int64_t nrows, int64_t row_bytes,
                          bool tiled, int64_t Hm) {
  const int unit = (row_bytes % 16 == 0) ? 16 : (row_bytes % 2 == 0) ? 2 : 1;
  const int64_t n = nrows * (row_bytes / unit);
  if (n <= 0) return 0;
  copy_row_bytes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint8_t*>(src), src_row0,
                                                                           reinterpret_cast<uint8_t*>(dst), dst_row0, nrows, row_bytes, unit,
                                                                           tiled ? 1 : 0, Hm);
  F2B_CUDA(cudaGetLastError());
  return 0;
}
// Can Linear `base` ([eN, eK]) run W-only inside the kernels? (packed, K a multiple of the 64-element k-block, 16-bit affine scales)
static bool wq_eligible(flux2b_ctx* c, const std::string& base, int eK) {
  if (c->quant == FLUX2B_BF16 || !c->option("wq_inkernel", 2) || eK % 64) return false;
  Tensor* w = find(c, base + ".weight");
  if (!w) return false;
  if (w->dtype != FLUX2B_U32) return is_float_dtype(w->dtype);   // will be packed by ensure_packed with f16 scales
  Tensor* s = find(c, base + ".scales");
  return s && s->dtype != FLUX2B_F32;   // f32 scales (unusual) take the dense fallback
}
// rows [src_row0, src_row0 + nrows) of the packed Linear `base` (shape [eN, eK]) -> rows [dst_row0, ...) of L (N_total rows)
static int wq_rows(flux2b_ctx* c, const std::string& base, int eN, int eK, int64_t src_row0, int64_t nrows, Lin* L, int N_total,
                   int64_t dst_row0, bool tiled, int Hm) {
  F2B_TRY(ensure_packed(c, base));
  PackedMeta pm;
  F2B_TRY(packed_meta(c, base, &pm));
  F2B_TRY(expect_shape(base, (int)pm.rows, (int)pm.cols, eN, eK));
  const int64_t row_bytes = (int64_t)eK * pm.bits / 8, G = eK / pm.group;
  const int64_t sb_bytes = G * (pm.has_b ? 2 : 1);
  const int sb_bf16 = (pm.has_b && pm.sb_dtype == FLUX2B_BF16_T) ? 1 : 0;
  if (!L->wq.p || L->N != N_total || L->K != eK || L->wmode != c->quant) {
    F2B_CUDA(L->wq.alloc((size_t)N_total * row_bytes));
    F2B_CUDA(L->ws.alloc((size_t)N_total * sb_bytes));
    if (pm.has_b) F2B_CUDA(L->wb.alloc((size_t)N_total * sb_bytes)); else L->wb.release();
    L->w.release(); L->sfb.release();
    L->N = N_total; L->K = eK; L->wmode = c->quant; L->mx = 0;
  }
  if (dst_row0 != 0 && L->w_sb_bf16 != sb_bf16) return fail(FLUX2B_ERR_WEIGHT_LOADING, "fused layers must share one scale type: " + base);
  L->w_sb_bf16 = sb_bf16;
  if (tiled && (nrows % 256 || Hm % 128)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "SwiGLU tile does not divide the MLP width: " + base);
  F2B_TRY(copy_row_bytes(c, pm.w->buf.p, src_row0, L->wq.p, dst_row0, nrows, row_bytes, tiled, Hm));
  F2B_TRY(copy_row_bytes(c, pm.s->buf.p, src_row0, L->ws.p, dst_row0, nrows, sb_bytes, tiled, Hm));
  if (pm.has_b) F2B_TRY(copy_row_bytes(c, pm.b->buf.p, src_row0, L->wb.p, dst_row0, nrows, sb_bytes, tiled, Hm));
  return 0;
}

int finalize_dit(flux2b_ctx* c) {
  const flux2b_dit_config& g = c->dit;
  const int D = g.num_attention_heads * g.attention_head_dim;
  const int Hm = (int)((float)D * g.mlp_ratio);  // Int(Float(dim) * mlpRatio), Flux2TransformerBlock.swift:53
  c->D = D; c->H = g.num_attention_heads; c->Hm = Hm;
  c->wq_stage_max = 0;
  if (g.attention_head_dim != 128) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "attention_head_dim must be 128");
  if (g.axes_dims_rope[0] + g.axes_dims_rope[1] + g.axes_dims_rope[2] + g.axes_dims_rope[3] != 128)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "axes_dims_rope must sum to 128");
  if (g.in_channels % 8 || g.joint_attention_dim % 8)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "in_channels / joint_attention_dim must be multiples of 8");

  F2B_TRY(build_lin(c, "xEmbedder", &c->x_embed, D, g.in_channels));
  F2B_TRY(build_lin(c, "contextEmbedder", &c->ctx_embed, D, g.joint_attention_dim));
  F2B_TRY(build_lin(c, "timeGuidanceEmbed.timestepEmbedder.linear1", &c->t_lin1, D, 256));
  F2B_TRY(build_lin(c, "timeGuidanceEmbed.timestepEmbedder.linear2", &c->t_lin2, D, D));
  if (g.guidance_embeds) {
    F2B_TRY(build_lin(c, "timeGuidanceEmbed.guidanceEmbedder.linear1", &c->g_lin1, D, 256));
    F2B_TRY(build_lin(c, "timeGuidanceEmbed.guidanceEmbedder.linear2", &c->g_lin2, D, D));
  }
  F2B_TRY(build_lin(c, "doubleStreamModulationImg.linear", &c->mod_img, 6 * D, D));
  F2B_TRY(build_lin(c, "doubleStreamModulationTxt.linear", &c->mod_txt, 6 * D, D));
  F2B_TRY(build_lin(c, "singleStreamModulation.linear", &c->mod_single, 3 * D, D));
  F2B_TRY(build_lin(c, "normOut.linear", &c->norm_out, 2 * D, D));
  F2B_TRY(build_lin(c, "projOut", &c->proj_out, g.out_channels, D));

  // block linears: dense 16-bit operands, or (option native_mx with an mx* / nvfp4 quantization) MLX's packed bytes as they are
  c->mx_kind = c->option("native_mx", 0) ? mx_kind_of_quant(c->quant) : 0;
  if (c->option("native_mx", 0) && !c->mx_kind)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "native_mx needs quantization mxfp8, mxfp4 or nvfp4");
  const bool mx = c->mx_kind != 0;
  const bool tile_ok = (Hm % 128 == 0) && c->option("fuse_swiglu", 1);
  auto lin = [&](const std::string& base, Lin* L, int eN, int eK) -> int {
    if (mx) return mx_rows(c, base, eN, eK, 0, eN, L, eN, 0, false, 0);
    return build_lin(c, base, L, eN, eK);
  };
  auto stacked = [&](const std::vector<std::string>& bases, Lin* L, int eN_each, int eK) -> int {
    if (!mx) {
      bool wq = true;
      for (const auto& b : bases) wq = wq && wq_eligible(c, b, eK);
      if (!wq) { lin_dense_reset(L); return build_stacked(c, bases, L, eN_each, eK); }
      for (size_t i = 0; i < bases.size(); ++i)
        F2B_TRY(wq_rows(c, bases[i], eN_each, eK, 0, eN_each, L, (int)bases.size() * eN_each, (int64_t)i * eN_each, false, 0));
      return 0;
    }
    for (size_t i = 0; i < bases.size(); ++i)
      F2B_TRY(mx_rows(c, bases[i], eN_each, eK, 0, eN_each, L, (int)bases.size() * eN_each, (int64_t)i * eN_each, false, 0));
    return 0;
  };
  // SwiGLU producer = rows [row0, row0 + 2 Hm) of `base` ([eN, D])
  auto swiglu_lin = [&](const std::string& base, int eN, int64_t row0, Lin* L, bool* tiled) -> int {
    if (mx) {
      *tiled = tile_ok;
      return mx_rows(c, base, eN, D, row0, 2 * (int64_t)Hm, L, 2 * Hm, 0, tile_ok, Hm);
    }
    if (wq_eligible(c, base, D)) {
      *tiled = tile_ok;
      return wq_rows(c, base, eN, D, row0, 2 * (int64_t)Hm, L, 2 * Hm, 0, tile_ok, Hm);
    }
    lin_dense_reset(L);
    DevBuf tmp; int N, K;
    F2B_TRY(dense16_from_key(c, base, &tmp, &N, &K));
    F2B_TRY(expect_shape(base, N, K, eN, D));
    F2B_TRY(build_swiglu(c, tmp.p, K, row0, L, Hm, D, tiled));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  };

  c->dbl.clear(); c->dbl.resize(g.num_layers);
  for (int i = 0; i < g.num_layers; ++i) {
    const std::string p = "transformerBlocks." + std::to_string(i) + ".";
    DoubleBlockW& b = c->dbl[i];
    F2B_TRY(stacked({p + "attn.toQ", p + "attn.toK", p + "attn.toV"}, &b.qkv_img, D, D));
    F2B_TRY(stacked({p + "attn.addQProj", p + "attn.addKProj", p + "attn.addVProj"}, &b.qkv_txt, D, D));
    F2B_TRY(lin(p + "attn.toOut", &b.out_img, D, D));
    F2B_TRY(lin(p + "attn.toAddOut", &b.out_txt, D, D));
    F2B_TRY(swiglu_lin(p + "ff.activation.proj", 2 * Hm, 0, &b.ff_in_img, &b.ff_tiled));
    F2B_TRY(swiglu_lin(p + "ffContext.activation.proj", 2 * Hm, 0, &b.ff_in_txt, &b.ff_tiled));
    F2B_TRY(lin(p + "ff.linearOut", &b.ff_out_img, D, Hm));
    F2B_TRY(lin(p + "ffContext.linearOut", &b.ff_out_txt, D, Hm));
    F2B_TRY(vector_f32_from_key(c, p + "attn.normQ.weight", &b.nq_img, 128, false, 1.f));
    F2B_TRY(vector_f32_from_key(c, p + "attn.normK.weight", &b.nk_img, 128, false, 1.f));
    F2B_TRY(vector_f32_from_key(c, p + "attn.normAddedQ.weight", &b.nq_txt, 128, false, 1.f));
    F2B_TRY(vector_f32_from_key(c, p + "attn.normAddedK.weight", &b.nk_txt, 128, false, 1.f));
  }
  c->sgl.clear(); c->sgl.resize(g.num_single_layers);
  for (int i = 0; i < g.num_single_layers; ++i) {
    const std::string p = "singleTransformerBlocks." + std::to_string(i) + ".";
    SingleBlockW& b = c->sgl[i];
    // fused projection column order q | k | v | gate | up, widths D,D,D,Hm,Hm (Flux2ParallelAttention.swift:56,83-87)
    if (mx) {
      F2B_TRY(mx_rows(c, p + "attn.toQkvMlp", 3 * D + 2 * Hm, D, 0, 3 * (int64_t)D, &b.qkv, 3 * D, 0, false, 0));
    } else if (wq_eligible(c, p + "attn.toQkvMlp", D)) {
      F2B_TRY(wq_rows(c, p + "attn.toQkvMlp", 3 * D + 2 * Hm, D, 0, 3 * (int64_t)D, &b.qkv, 3 * D, 0, false, 0));
    } else {
      lin_dense_reset(&b.qkv);
      DevBuf tmp; int N, K;
      F2B_TRY(dense16_from_key(c, p + "attn.toQkvMlp", &tmp, &N, &K));
      F2B_TRY(expect_shape(p + "attn.toQkvMlp", N, K, 3 * D + 2 * Hm, D));
      F2B_CUDA(b.qkv.w.alloc((size_t)3 * D * D * 2));
      F2B_TRY(copy_rows16(c, tmp.p, K, 0, b.qkv.w.p, K, 0, 3 * (int64_t)D, K, false, 0));
      b.qkv.N = 3 * D; b.qkv.K = D;
      F2B_CUDA(cudaStreamSynchronize(c->stream));
    }
    F2B_TRY(swiglu_lin(p + "attn.toQkvMlp", 3 * D + 2 * Hm, 3 * (int64_t)D, &b.mlp, &b.mlp_tiled));
    F2B_TRY(lin(p + "attn.toOut", &b.out, D, D + Hm));
    F2B_TRY(vector_f32_from_key(c, p + "attn.normQ.weight", &b.nq, 128, false, 1.f));
    F2B_TRY(vector_f32_from_key(c, p + "attn.normK.weight", &b.nk, 128, false, 1.f));
  }
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  if (!c->option("keep_raw_weights", 1)) {
    // Forward-only context: what the kernels consume are the working copies, so the handed-over Linear tensors go — the dense
    // `.weight` matrices and, for quantized layers, `.weight` (uint32) together with its `.scales` / `.biases` (never one
    // without the others: a packed weight whose scales are gone cannot be exported or LoRA-merged, and must say so cleanly).
    // get_tensor / save_prequantized / merge_lora on such a context report the tensor as missing. With in-kernel dequantisation
    // (wq_inkernel) or native_mx the packed working copy is then the ONLY resident copy of a quantized layer.
    for (auto it = c->tensors.begin(); it != c->tensors.end();) {
      const std::string& k = it->first;
      const bool dit_key = k.rfind("decoder.", 0) != 0 && k.rfind("encoder.", 0) != 0 && k.rfind("postQuantConv", 0) != 0 &&
                           k.rfind("quantConv", 0) != 0 && k.rfind("latentBatchNorm", 0) != 0;
      const bool linear_part = it->second.shape.size() == 2 && (ends_with(k, ".weight") || ends_with(k, ".scales") || ends_with(k, ".biases"));
      if (dit_key && linear_part) it = c->tensors.erase(it);
      else ++it;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ text encoder
// Keys are the module paths of Qwen3ForCausalLM / MistralForCausalLM (FluxTextEncoders/Model/Qwen3/Qwen3Model.swift:33-55,
// Qwen3DecoderLayer.swift:14-18, Qwen3Attention.swift:54-62, Qwen3MLP.swift:19-21), i.e. the HF checkpoint names:
//   model.embed_tokens.weight, model.layers.N.{input_layernorm,post_attention_layernorm}.weight,
//   model.layers.N.self_attn.{q_proj,k_proj,v_proj,o_proj}.weight (+ .scales / .biases when MLX-quantized),
//   model.layers.N.self_attn.{q_norm,k_norm}.weight (Qwen3), model.layers.N.mlp.{gate_proj,up_proj,down_proj}.weight, model.norm.weight
// Only the layers up to the deepest extracted hidden state are ever run, so layers that are absent are simply not built
// (the Klein extractor needs 27 of 36, the Mistral one 30 of 40); lm_head is never used by the embedding path.
int finalize_te(flux2b_ctx* c) {
  te_destroy_graphs(c);   // captured prefills hold the old working weights' addresses
  const flux2b_te_config& t = c->te;
  const int Hd = t.hidden_size, I = t.intermediate_size, Nq = t.num_heads * 128, Nkv = t.num_kv_heads * 128;
  {
    int N, K;
    F2B_TRY(dense16_from_key(c, "model.embed_tokens", &c->te_embed, &N, &K));
    F2B_TRY(expect_shape("model.embed_tokens", N, K, t.vocab_size, Hd));
  }
  F2B_TRY(vector_f32_from_key(c, "model.norm.weight", &c->te_norm, Hd, false, 1.f));
  F2B_TRY(vector_f32_from_key(c, "__te_ones__", &c->te_ones, Hd, false, 1.f));
  c->te_layers.clear();
  c->te_layers.resize(t.num_layers);
  c->te_layers_built = 0;
  for (int i = 0; i < t.num_layers; ++i) {
    const std::string p = "model.layers." + std::to_string(i) + ".";
    if (!find(c, p + "self_attn.q_proj.weight")) break;   // deeper layers were not handed over
    TeLayerW& L = c->te_layers[i];
    F2B_TRY(vector_f32_from_key(c, p + "input_layernorm.weight", &L.ln1, Hd, true, 1.f));
    F2B_TRY(vector_f32_from_key(c, p + "post_attention_layernorm.weight", &L.ln2, Hd, true, 1.f));
    if (t.qk_norm) {
      F2B_TRY(vector_f32_from_key(c, p + "self_attn.q_norm.weight", &L.nq, 128, true, 1.f));
      F2B_TRY(vector_f32_from_key(c, p + "self_attn.k_norm.weight", &L.nk, 128, true, 1.f));
    }
    if (find(c, p + "self_attn.q_proj.bias")) return fail(FLUX2B_ERR_WEIGHT_LOADING, "attention_bias = true is not supported: " + p);
    {
      const char* names[3] = {"self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"};
      const int rows[3] = {Nq, Nkv, Nkv};
      F2B_CUDA(L.qkv.w.alloc((size_t)(Nq + 2 * Nkv) * Hd * 2));
      int64_t r0 = 0;
      for (int j = 0; j < 3; ++j) {
        DevBuf tmp; int N, K;
        F2B_TRY(dense16_from_key(c, p + names[j], &tmp, &N, &K));
        F2B_TRY(expect_shape(p + names[j], N, K, rows[j], Hd));
        F2B_TRY(copy_rows16(c, tmp.p, K, 0, L.qkv.w.p, K, r0, N, K, false, 0));
        F2B_CUDA(cudaStreamSynchronize(c->stream));
        r0 += N;
      }
      L.qkv.N = Nq + 2 * Nkv; L.qkv.K = Hd;
    }
    F2B_TRY(build_lin(c, p + "self_attn.o_proj", &L.o, Hd, Nq));
    {
      DevBuf both;
      F2B_CUDA(both.alloc((size_t)2 * I * Hd * 2));
      const char* names[2] = {"mlp.gate_proj", "mlp.up_proj"};
      for (int j = 0; j < 2; ++j) {
        DevBuf tmp; int N, K;
        F2B_TRY(dense16_from_key(c, p + names[j], &tmp, &N, &K));
        F2B_TRY(expect_shape(p + names[j], N, K, I, Hd));
        F2B_TRY(copy_rows16(c, tmp.p, K, 0, both.p, K, (int64_t)j * I, N, K, false, 0));
        F2B_CUDA(cudaStreamSynchronize(c->stream));
      }
      F2B_TRY(build_swiglu(c, both.p, Hd, 0, &L.gate_up, I, Hd, &L.mlp_tiled));
      F2B_CUDA(cudaStreamSynchronize(c->stream));
    }
    F2B_TRY(build_lin(c, p + "mlp.down_proj", &L.down, Hd, I));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
    c->te_layers_built = i + 1;
    if (!c->option("keep_raw_weights", 1)) {
      // the dense working copies are all the forward needs; packed (MLX-quantized) tensors stay for get_tensor
      for (auto it = c->tensors.begin(); it != c->tensors.end();) {
        if (it->first.rfind(p, 0) == 0 && is_float_dtype(it->second.dtype) && it->second.shape.size() == 2 && ends_with(it->first, ".weight")) it = c->tensors.erase(it);
        else ++it;
      }
    }
  }
  if (c->te_layers_built == 0) return fail(FLUX2B_ERR_WEIGHT_LOADING, "no text-encoder layers were loaded (model.layers.0.* missing)");
  if (!c->option("keep_raw_weights", 1)) {
    auto it = c->tensors.find("model.embed_tokens.weight");
    if (it != c->tensors.end() && is_float_dtype(it->second.dtype)) c->tensors.erase(it);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ VAE
static int build_conv(flux2b_ctx* c, const std::string& base, ConvW* w, int cout, int cin, int k) {
  Tensor* t = find(c, base + ".weight");
  if (!t) return fail(FLUX2B_ERR_WEIGHT_LOADING, "missing tensor: " + base + ".weight");
  // OHWI (WeightLoader.swift:496-498)
  if (t->shape.size() != 4 || t->shape[0] != cout || t->shape[1] != k || t->shape[2] != k || t->shape[3] != cin)
    return fail(FLUX2B_ERR_WEIGHT_LOADING, "conv weight must be OHWI [" + std::to_string(cout) + "," + std::to_string(k) +
                                               "," + std::to_string(k) + "," + std::to_string(cin) + "]: " + base);
  int N, K;
  F2B_TRY(dense16_from_key_ex(c, base, &w->w, &N, &K, false));
  w->cin = cin; w->cout = cout; w->taps = k * k;
  F2B_TRY(vector_f32_from_key(c, base + ".bias", &w->bias, cout, false, 0.f));
  return 0;
}
static int build_norm(flux2b_ctx* c, const std::string& base, NormW* n, int C) {
  n->C = C;
  F2B_TRY(vector_f32_from_key(c, base + ".weight", &n->gamma, C, false, 1.f));
  F2B_TRY(vector_f32_from_key(c, base + ".bias", &n->beta, C, false, 0.f));
  return 0;
}
static int build_resnet(flux2b_ctx* c, const std::string& base, ResnetW* r, int cin, int cout) {
  r->cin = cin; r->cout = cout;
  F2B_TRY(build_norm(c, base + ".norm1", &r->n1, cin));
  F2B_TRY(build_conv(c, base + ".conv1", &r->c1, cout, cin, 3));
  F2B_TRY(build_norm(c, base + ".norm2", &r->n2, cout));
  F2B_TRY(build_conv(c, base + ".conv2", &r->c2, cout, cout, 3));
  r->has_sc = cin != cout;
  if (r->has_sc) F2B_TRY(build_conv(c, base + ".convShortcut", &r->sc, cout, cin, 1));
  return 0;
}

// single-head VAE attention: fuse toQ|toK|toV into one [3C, C] operand (ResnetBlock.swift:258-313)
static int build_vae_attention(flux2b_ctx* c, const std::string& prefix, int C, Lin* qkv, DevBuf* qkv_bias, Lin* out, DevBuf* out_bias) {
  qkv->N = 3 * C; qkv->K = C;
  F2B_CUDA(qkv->w.alloc((size_t)3 * C * C * 2));
  F2B_CUDA(qkv_bias->alloc(sizeof(float) * 3 * C));
  const char* names[3] = {"toQ", "toK", "toV"};
  for (int i = 0; i < 3; ++i) {
    DevBuf tmp, b; int N, K;
    const std::string base = prefix + "." + names[i];
    F2B_TRY(dense16_from_key_ex(c, base, &tmp, &N, &K, false));
    F2B_TRY(expect_shape(base, N, K, C, C));
    F2B_TRY(copy_rows16(c, tmp.p, K, 0, qkv->w.p, K, (int64_t)i * C, C, C, false, 0));
    F2B_TRY(vector_f32_from_key(c, base + ".bias", &b, C, false, 0.f));
    F2B_CUDA(cudaMemcpyAsync(qkv_bias->as<float>() + (size_t)i * C, b.p, sizeof(float) * C, cudaMemcpyDeviceToDevice, c->stream));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
  }
  int N, K;
  F2B_TRY(dense16_from_key_ex(c, prefix + ".toOut", &out->w, &N, &K, false));
  F2B_TRY(expect_shape(prefix + ".toOut", N, K, C, C));
  out->N = C; out->K = C;
  F2B_TRY(vector_f32_from_key(c, prefix + ".toOut.bias", out_bias, C, false, 0.f));
  return 0;
}

__global__ void pad_cin_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int64_t rows, int cin, int cpad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cpad) return;
  const int64_t r = i / cpad; const int ch = (int)(i % cpad);
  dst[i] = ch < cin ? src[r * cin + ch] : (uint16_t)0;
}

// VAE encoder (VAE/VAEEncoder.swift:26-83; key layout WeightLoader.swift:500-547): built when its tensors were handed over
static int finalize_vae_encoder(flux2b_ctx* c) {
  const flux2b_vae_config& g = c->vae;
  VaeEncW& e = c->vw.enc;
  e.ready = false;
  if (!find(c, "encoder.convIn.weight")) return 0;
  int ch[4];
  for (int i = 0; i < 4; ++i) ch[i] = g.encoder_channels[i];
  if (!ch[0]) { ch[0] = 128; ch[1] = 256; ch[2] = 512; ch[3] = 512; }
  const int L2 = 2 * g.latent_channels;
  {
    // convIn: Cin = 3 is zero-padded to 8 input channels (the image arrives NHWC with 8 channels, 5 of them zero)
    ConvW raw;
    F2B_TRY(build_conv(c, "encoder.convIn", &raw, ch[0], g.in_channels, 3));
    const int cpad = (g.in_channels + 7) / 8 * 8;
    const int64_t rows = (int64_t)ch[0] * 9;
    F2B_CUDA(e.conv_in.w.alloc((size_t)rows * cpad * 2));
    pad_cin_kernel<<<(unsigned)((rows * cpad + 255) / 256), 256, 0, c->stream>>>(raw.w.as<uint16_t>(), e.conv_in.w.as<uint16_t>(), rows,
                                                                                g.in_channels, cpad);
    F2B_CUDA(cudaGetLastError());
    F2B_CUDA(cudaStreamSynchronize(c->stream));
    e.conv_in.bias = std::move(raw.bias);
    e.conv_in.cin = cpad; e.conv_in.cout = ch[0]; e.conv_in.taps = 9;
  }
  e.down.clear(); e.downconv.clear(); e.has_down.clear();
  e.down.resize(4); e.downconv.resize(4); e.has_down.assign(4, false);
  int prev = ch[0];
  for (int i = 0; i < 4; ++i) {
    e.down[i].resize(g.layers_per_block);
    for (int j = 0; j < g.layers_per_block; ++j) {
      F2B_TRY(build_resnet(c, "encoder.downBlocks." + std::to_string(i) + ".0." + std::to_string(j), &e.down[i][j], prev, ch[i]));
      prev = ch[i];
    }
    if (i < 3) {
      F2B_TRY(build_conv(c, "encoder.downBlocks." + std::to_string(i) + ".1.conv", &e.downconv[i], ch[i], ch[i], 3));
      e.has_down[i] = true;
    }
  }
  const int C3 = ch[3];
  F2B_TRY(build_resnet(c, "encoder.midBlock.0", &e.mid1, C3, C3));
  F2B_TRY(build_norm(c, "encoder.midBlock.1.groupNorm", &e.attn_norm, C3));
  F2B_TRY(build_vae_attention(c, "encoder.midBlock.1", C3, &e.attn_qkv, &e.attn_qkv_bias, &e.attn_out, &e.attn_out_bias));
  F2B_TRY(build_resnet(c, "encoder.midBlock.2", &e.mid2, C3, C3));
  F2B_TRY(build_norm(c, "encoder.convNormOut", &e.norm_out, C3));
  F2B_TRY(build_conv(c, "encoder.convOut", &e.conv_out, L2, C3, 3));
  e.has_quant = find(c, "quantConv.weight") != nullptr;   // optional (AutoencoderKL.swift:94-99)
  if (e.has_quant) F2B_TRY(build_conv(c, "quantConv", &e.quant, L2, L2, 1));
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  e.ready = true;
  return 0;
}

int finalize_vae(flux2b_ctx* c) {
  const flux2b_vae_config& g = c->vae;
  VaeW& v = c->vw;
  const bool saved_f16 = c->f16();
  // the VAE runs in f16 by default (bounded activations, 3 more mantissa bits than bf16); weights follow
  const int vae_f16 = c->option("vae_f16", 1);
  c->opt["compute_f16"] = vae_f16;
  auto restore = [&]() { c->opt["compute_f16"] = saved_f16 ? 1 : 0; };
  int rc = 0;
  do {
    const int C3 = g.decoder_channels[3];
    if ((rc = build_conv(c, "postQuantConv", &v.post_quant, g.latent_channels, g.latent_channels, 1))) break;
    if ((rc = build_conv(c, "decoder.convIn", &v.conv_in, C3, g.latent_channels, 3))) break;
    if ((rc = build_resnet(c, "decoder.midBlock.0", &v.mid1, C3, C3))) break;
    if ((rc = build_norm(c, "decoder.midBlock.1.groupNorm", &v.attn_norm, C3))) break;
    if ((rc = build_vae_attention(c, "decoder.midBlock.1", C3, &v.attn_qkv, &v.attn_qkv_bias, &v.attn_out, &v.attn_out_bias))) break;
    if ((rc = build_resnet(c, "decoder.midBlock.2", &v.mid2, C3, C3))) break;
    v.up.clear(); v.upconv.clear(); v.has_upconv.clear();
    v.up.resize(4); v.upconv.resize(4); v.has_upconv.assign(4, false);
    int prev = C3;
    for (int i = 0; i < 4 && !rc; ++i) {
      const int co = g.decoder_channels[3 - i];
      const int nres = g.layers_per_block + 1;  // VAEDecoder.swift:57
      v.up[i].resize(nres);
      for (int j = 0; j < nres && !rc; ++j)
        rc = build_resnet(c, "decoder.upBlocks." + std::to_string(i) + ".0." + std::to_string(j), &v.up[i][j],
                          j == 0 ? prev : co, co);
      prev = co;
      if (!rc && i < 3) {
        rc = build_conv(c, "decoder.upBlocks." + std::to_string(i) + ".1.conv", &v.upconv[i], co, co, 3);
        v.has_upconv[i] = true;
        if (!rc) {
          // Upsample2D = nearest 2x + this convolution (ResnetBlock.swift:240-252): its four 2x2 phase kernels, summed once here
          if (v.upconv[i].w_up.alloc((size_t)co * 16 * co * 2) != cudaSuccess) { cudaGetLastError(); rc = fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "upsample weights"); }
          else if (fold_upsample_weights(v.upconv[i].w.p, v.upconv[i].w_up.p, co, co, vae_f16 != 0, c->stream) != cudaSuccess) rc = fail(FLUX2B_ERR_CUDA, "fold_upsample_weights");
        }
      }
    }
    if (rc) break;
    if ((rc = build_norm(c, "decoder.convNormOut", &v.norm_out, g.decoder_channels[0]))) break;
    if ((rc = build_conv(c, "decoder.convOut", &v.conv_out, g.out_channels, g.decoder_channels[0], 3))) break;
    v.has_bn = find(c, "latentBatchNorm.runningMean") && find(c, "latentBatchNorm.runningVar");
    if (v.has_bn) {
      const int64_t n = find(c, "latentBatchNorm.runningMean")->numel();
      if ((rc = vector_f32_from_key(c, "latentBatchNorm.runningMean", &v.bn_mean, n, true, 0.f))) break;
      if ((rc = vector_f32_from_key(c, "latentBatchNorm.runningVar", &v.bn_var, n, true, 1.f))) break;
    }
    if ((rc = finalize_vae_encoder(c))) break;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(FLUX2B_ERR_CUDA, "sync after VAE weights"); break; }
    v.ready = true;
  } while (0);
  restore();
  return rc;
}

}  // namespace f2b
