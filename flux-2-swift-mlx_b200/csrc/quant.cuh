// quant.cuh — host interface of the weight quantizers (quant.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace f2b {
// quant: 1 qint8, 2 int4, 3 mxfp8, 4 mxfp4, 5 nvfp4 (== flux2b_quant). dtype codes == flux2b_dtype (0 f32, 1 f16, 2 bf16).
bool quant_params(int quant, int* bits, int* group, int* has_biases, int* scale_dtype);
cudaError_t quantize_matrix(int quant, const void* w, int w_dtype, int64_t rows, int64_t cols, uint32_t* packed,
                            void* scales, void* biases, cudaStream_t s);
// sb_dtype: float type of the affine modes' scales / biases (0 f32, 1 f16, 2 bf16); the mx modes' scales are always one byte
cudaError_t dequantize_matrix(int quant, const uint32_t* packed, const void* scales, const void* biases, int64_t rows,
                              int64_t cols, void* out, int out_dtype, cudaStream_t s, int sb_dtype = 1);
cudaError_t lora_add(void* W, int w_dtype, const float* A, const float* B, int64_t out_dim, int64_t in_dim, int rank,
                     float scale, cudaStream_t s);

// ---- native block-scaled operands. kind: 1 = mxfp8 (E4M3, E8M0 / 32), 2 = mxfp4 (E2M1, E8M0 / 32), 3 = nvfp4 (E2M1, E4M3 / 16).
// Element bytes stay exactly as MLX packs them (row-major, fp4 two per byte, low nibble first). Group scales are re-tiled into
// the tcgen05 scale-factor layout: one 512 B block per (128 rows x 4 groups), blocks row-block-major with `ld` blocks per row block:
//      offset(row, g) = ((row / 128) * ld + g / 4) * 512 + (row % 32) * 16 + ((row % 128) / 32) * 4 + g % 4,   g = k / group
int mx_kind_of_quant(int quant);                       // flux2b_quant -> kind (0 for bf16 / affine modes)
int64_t mx_sf_ld(int kind, int64_t K);                 // blocks per row block for a [*, K] operand
size_t mx_sf_bytes(int kind, int64_t rows, int64_t K);
// rows of an MLX-packed weight (bytes [*, K*bits/8], scales [*, K/group]) -> rows [dst_row0, dst_row0 + nrows) of the
// working copy; tile = 256 / 128 applies the SwiGLU interleave ([tile/2 gate | tile/2 value] rows per GEMM N tile), 0 = none
cudaError_t mx_copy_rows(int kind, const uint8_t* src_w, const uint8_t* src_s, int64_t src_row0, uint8_t* dst_w, uint8_t* dst_sf,
                         int64_t dst_row0, int64_t nrows, int64_t K, int tile, int64_t Hm, cudaStream_t s);
// 16-bit activations x16[M, K] (leading dim ldx elements) -> quantised bytes aq[M, K*bits/8] (leading dim lda_bytes) + group
// scales written at group offset col0 / group of a scale-factor tensor with `sf_ld` blocks per row block (so a column slice
// [col0, col0 + K) of a wider activation can be quantised on its own). fp4 kinds use the weight packer's arithmetic
// (bit-identical to quantize_matrix); mxfp8 uses scale = 2^ceil(log2(amax / 448)). Scale rows up to the next multiple of
// 128 are filled with 1.0.
cudaError_t mx_quantize_act(int kind, const void* x16, int64_t ldx, int M, int K, bool f16, uint8_t* aq, int64_t lda_bytes,
                            uint8_t* sfa, int64_t sf_ld, int64_t col0, cudaStream_t s);
// Destination of an activation quantisation fused into the producing kernel (LayerNorm + modulate, SwiGLU GEMM epilogue):
// row r, column c of the producer's output goes to q[r * ldq + c * bits / 8] and its group scale to group g0 + c / group of
// row r in the scale-factor tensor `sf` (sf_ld blocks per 128-row block). kind 0 = off.
struct MxOut {
  int kind = 0;
  uint8_t* q = nullptr;
  int64_t ldq = 0;
  uint8_t* sf = nullptr;
  int64_t sf_ld = 0;
  int g0 = 0;
};
cudaError_t mx_sf_untile(const uint8_t* sf, uint8_t* out, int64_t M, int64_t G, cudaStream_t s);

}  // namespace f2b
