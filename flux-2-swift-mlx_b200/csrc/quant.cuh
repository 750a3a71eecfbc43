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
}  // namespace f2b
