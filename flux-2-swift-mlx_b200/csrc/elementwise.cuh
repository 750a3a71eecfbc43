// elementwise.cuh — host interface of the HBM-bound kernels (elementwise.cu). All launches are stream-ordered.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "quant.cuh"

namespace f2b {

// 16-bit activation storage: bf16 (default) or f16, selected at run time (same kernels, different pack/unpack).
// x fp32 [rows, D] -> out16 [rows, D] = LN(x) * (1 + scale[b]) + shift[b];   b = row / rows_per_batch
// LayerNorm: biased variance, eps, no affine (Flux2TransformerBlock.swift:56-61; Flux2Modulation.swift:96-112).
cudaError_t ln_modulate(const float* x, int64_t ldx, void* out16, int64_t ldo, int rows, int D, const float* shift,
                        const float* scale, int64_t mod_batch_stride, int rows_per_batch, float eps, bool f16,
                        cudaStream_t s, const MxOut* mx = nullptr, int split_row = 0, const float* shift_lo = nullptr,
                        const float* scale_lo = nullptr);
// split_row > 0: rows [0, split_row) take (shift_lo, scale_lo) instead — the text rows of a double-stream block, so that both
// streams are one launch (16-bit output only).
// With mx->kind != 0 (native block-scaled path) out16 is not written: the 16-bit result is quantised in the same pass to
// mxfp8 / mxfp4 / nvfp4 (bit-identical to mx_quantize_act applied to the 16-bit output); D % 128 == 0.

// ---- text-encoder prefill (te.cu)
// out[r, :] = x[r, :] * rsqrt(mean(x[r, :]^2) + eps) * w   (FluxTextEncoders/Model/RMSNorm.swift); 16-bit or fp32 output
cudaError_t rms_norm_rows(const float* x, int64_t ldx, const float* w, void* out, int64_t ldo, int rows, int D, float eps,
                          bool out_f32, bool f16, cudaStream_t s);
// x[r, :] = float(table16[ids[r], :])   (Qwen3Model.swift:66)
cudaError_t embed_rows(const int32_t* ids, const void* table16, int64_t vocab, int D, float* x, int64_t ldx, int rows, bool f16,
                       cudaStream_t s);
// rotate-half RoPE table at head dim 128: cos/sin [S, 128], columns j and j + 64 = angle (pos0 + s) * base^(-j/64)
cudaError_t rope_half_table(int S, int pos0, float base, float* cos_out, float* sin_out, cudaStream_t s);
// out_kind: 0 fp32, 1 f16, 2 bf16
cudaError_t copy_f32_to_any(const float* in, int64_t ldi, void* out, int64_t ldo, int rows, int cols, int out_kind, cudaStream_t s);

// y[b, n] = sum_k act(x[b, k]) * W[n, k] ; W 16-bit [N, K] row-major; act = SiLU if silu_in. B <= 8.
cudaError_t gemv(const float* x, int64_t ldx, const void* W16, int64_t ldw, float* y, int64_t ldy, int B, int N, int K,
                 bool silu_in, bool accumulate, bool f16, cudaStream_t s);
// the same over a W-only quantized weight (packed codes [N, row_bytes], scales / biases [N, sb_ld]; mode = flux2b_quant 1..5):
// bit-identical to gemv over the dequantized 16-bit matrix
cudaError_t gemv_q(const float* x, int64_t ldx, const void* codes, int64_t row_bytes, const void* scales, const void* biases, int64_t sb_ld,
                   int mode, int sb_bf16, float* y, int64_t ldy, int B, int N, int K, bool silu_in, bool accumulate, bool f16, cudaStream_t s);

// Timesteps(256): out[b, 0:128] = cos(t*1000*f_i), out[b, 128:256] = sin(...), f_i = exp(-ln(1e4) i / 128)
// (Flux2Embeddings.swift:27-44; the x1000 is Flux2Transformer.swift:145-146).
cudaError_t timestep_sinusoid(const float* t, float* out, int B, float pre_scale, cudaStream_t s);

// cos/sin [S,128] fp32 from ids int32 [S,4] (Flux2RoPE.swift:123-169).
cudaError_t rope_table(const int32_t* ids, int S, const int* axes_dims, float theta, float* cos_out, float* sin_out,
                       cudaStream_t s);

// In-place RMSNorm(128, learned weight) + interleaved-pair RoPE on the q and k slabs of a packed [rows, 3D] buffer
// (unfused reference path for the GEMM-epilogue version; Flux2Attention.swift:138-158).
cudaError_t qk_norm_rope(void* qkv16, int64_t ld, int rows, int D, const float* norm_q, const float* norm_k,
                         const float* cos_t, const float* sin_t, float eps, bool f16, cudaStream_t s);

// out16[r, j] = silu(in16[r, j]) * in16[r, H + j]   (Flux2FeedForward.swift:62-66)
cudaError_t swiglu(const void* in16, int64_t ldi, void* out16, int64_t ldo, int rows, int H, bool f16, cudaStream_t s);

// out[r, j] = res[r, j] + gate[b, j] * y16[r, j]   (Flux2Modulation.swift:115-122; unfused fallback)
cudaError_t gate_residual(const void* y16, int64_t ldy, const float* gate, int64_t gate_batch_stride, int rows_per_batch,
                          float* x, int64_t ldx, int rows, int D, bool f16, cudaStream_t s);

cudaError_t f32_to_16(const float* in, int64_t ldi, void* out16, int64_t ldo, int rows, int cols, bool f16,
                      cudaStream_t s);
cudaError_t any16_to_16(const void* in, int in_is_f16, void* out16, bool f16, int64_t n, cudaStream_t s);
cudaError_t cvt16_to_f32(const void* in16, float* out, int64_t n, bool f16, cudaStream_t s);

// Euler: x += (sigma_next - sigma) * v, v = pred or uncond + g (pred - uncond)
// (FlowMatchEulerScheduler.swift:136-156; CFG combine Flux2Pipeline.swift:1970)
cudaError_t euler_step(float* x, const float* pred, const float* pred_uncond, float cfg, float dt, int64_t n,
                       cudaStream_t s);
// (1 - sigma) * sample + sigma * noise  (FlowMatchEulerScheduler.swift:195-204)
cudaError_t scale_noise(const float* sample, const float* noise, float sigma, float* out, int64_t n, cudaStream_t s);
// RePaint blend used by the only in-tree step hook (Flux2MaskedInpaintingChain.swift:399-403):
// x = (1-m) * ((1-sn) * x0 + sn * eps) + m * x
cudaError_t repaint_blend(float* x, const float* x0, const float* eps, const float* mask, float sigma_next, int64_t n,
                          cudaStream_t s);

// generic strided permute of an fp32 tensor with up to 6 dims (LatentUtils pack / unpack / patchify)
cudaError_t permute_f32(const float* in, float* out, int ndim, const int* out_shape, const int64_t* in_strides,
                        cudaStream_t s);
// y = x * sqrt(var[c] + eps) + mean[c]  or the inverse; NCHW fp32 (LatentUtils.swift:460-496)
cudaError_t bn_affine_nchw(const float* x, float* y, const float* mean, const float* var, float eps, int B, int C,
                           int64_t hw, bool denorm, cudaStream_t s);
// fused tail of the denoise loop: seq [B, h*w, 128] fp32 -> BN denorm -> unpatchify -> NHWC 16-bit [B, 2h, 2w, 32]
cudaError_t seq_to_vae_input(const float* seq, const float* mean, const float* var, float eps, void* out16, int B, int h,
                             int w, bool f16, cudaStream_t s);

// GroupNorm (+ optional SiLU) over NHWC 16-bit, fp32 statistics (ResnetBlock.swift:24-54); deterministic (no atomics).
// stats_ws must hold groupnorm_ws_bytes(B, G) bytes.
size_t groupnorm_ws_bytes(int B, int G);
cudaError_t groupnorm_silu(const void* x16, void* y16, const float* gamma, const float* beta, double* stats_ws, int B,
                           int64_t HW, int C, int G, float eps, bool silu, bool f16, cudaStream_t s);
// OHWI [Cout, 3, 3, Cin] -> [Cout, 16, Cin]: the four 2x2 phase kernels of nearest-2x upsample + conv3x3 (gemm.cuh: conv_up2)
cudaError_t fold_upsample_weights(const void* w_ohwi16, void* out16, int64_t Cout, int Cin, bool f16, cudaStream_t s);
cudaError_t upsample_nearest2x(const void* x16, void* y16, int B, int H, int W, int C, cudaStream_t s);
// softmax over the last dim of fp32 scores -> 16-bit probabilities (VAE mid attention, ResnetBlock.swift:302-304)
cudaError_t softmax_rows(const float* x, int64_t ldx, void* y16, int64_t ldy, int rows, int cols, float scale, bool f16,
                         cudaStream_t s);
cudaError_t transpose16(const void* in16, int64_t ldi, void* out16, int64_t ldo, int rows, int cols, cudaStream_t s);
cudaError_t add16(const void* a16, const void* b16, void* out16, int64_t n, bool f16, cudaStream_t s);
// NHWC 16-bit [B,H,W,3(ldc)] -> uint8 HWC: clip((x+1)*127.5, 0, 255) truncated (Flux2Pipeline.swift:2425-2468)
cudaError_t postprocess_u8(const void* x16, int64_t ldc, uint8_t* out, int64_t npix, bool f16, cudaStream_t s);
// NHWC 16-bit -> NCHW fp32 (AutoencoderKL.decode returns NCHW)
cudaError_t nhwc16_to_nchw_f32(const void* x16, int64_t ldc, float* out, int B, int64_t HW, int C, bool f16,
                               cudaStream_t s);
cudaError_t nchw_f32_to_nhwc16(const float* in, void* out16, int B, int64_t HW, int C, bool f16, cudaStream_t s);
// the same with the channel count padded to Cpad with zeros (VAE encoder input: 3 -> 8 channels)
cudaError_t nchw_f32_to_nhwc16_pad(const float* in, void* out16, int B, int64_t HW, int C, int Cpad, bool f16, cudaStream_t s);
// posterior moments NHWC 16-bit [B, HW, ldc >= 2L] -> latent NCHW fp32: mean (noise == nullptr) or mean + exp(logvar / 2) * noise
cudaError_t moments_to_latent(const void* m16, int64_t ldc, const float* noise, float* out, int B, int64_t HW, int L, bool f16,
                              cudaStream_t s);

}  // namespace f2b
