// quant.cuh — host interface of the weight quantizers (quant.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace f2b {
// quant: 1 qint8, 2 int4, 3 mxfp8, 4 mxfp4, 5 nvfp4 (== flux2b_quant). dtype codes == flux2b_dtype (0 f32, 1 f16, 2 bf16).
bool quant_params(int quant, int* bits, int* group, int* has_biases, int* scale_dtype);
cudaError_t quantize_matrix(int quant, const void* w, int w_dtype, int64_t rows, int64_t cols, uint32_t* packed,
                            void* scales, void* biases, cudaStream_t s);
cudaError_t dequantize_matrix(int quant, const uint32_t* packed, const void* scales, const void* biases, int64_t rows,
                              int64_t cols, void* out, int out_dtype, cudaStream_t s);
cudaError_t lora_add(void* W, int w_dtype, const float* A, const float* B, int64_t out_dim, int64_t in_dim, int rank,
                     float scale, cudaStream_t s);

// ---- native block-scaled (mxfp8) operands: tcgen05 scale-factor layout, one 512 B block per (128 rows x 128 K):
//      offset(row, g) = ((row / 128) * (K / 128) + g / 4) * 512 + (row % 32) * 16 + ((row % 128) / 32) * 4 + g % 4,  g = k / 32
// rows of an MLX-packed mxfp8 weight (bytes [*, K], scales [*, K/32]) -> rows [dst_row0, dst_row0 + nrows) of the
// working copy; `tiled` applies the SwiGLU [128 gate | 128 value] interleave of weights.cu
cudaError_t mx8_copy_rows(const uint8_t* src_w, const uint8_t* src_s, int64_t src_row0, uint8_t* dst_w, uint8_t* dst_sf,
                          int64_t dst_row0, int64_t nrows, int64_t K, bool tiled, int64_t Hm, cudaStream_t s);
// 16-bit activations [M, K] (leading dim ldx) -> E4M3 bytes [M, K] + E8M0 scales (scale = 2^ceil(log2(amax / 448)), so
// nothing saturates); scale-factor rows up to the next multiple of 128 are filled with 1.0
cudaError_t mx8_quantize_act(const void* x16, int64_t ldx, int M, int K, bool f16, uint8_t* a8, uint8_t* sfa, cudaStream_t s);
size_t mx8_sf_bytes(int64_t rows, int64_t K);

}  // namespace f2b
