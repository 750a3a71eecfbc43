// attention.cuh — host interface of the fused joint flash attention (attention.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace f2b {

struct KVSegment {
  const void* k = nullptr;  // [rows, ldk] ; head h at column h*128
  const void* v = nullptr;
  int64_t ldk = 0, ldv = 0;
  int64_t rows_total = 0;  // rows addressable from k / v (tensor-map extent)
  int row0 = 0;            // first key row of this segment (per batch item: row0 + b*batch_stride)
  int len = 0;             // number of keys
  int64_t batch_stride = 0;
};

struct AttnProblem {
  // softmax(Q K^T * scale) V per head, head_dim = 128, no mask (reference: Flux2Attention.swift:168-174,
  // Flux2ParallelAttention.swift:104-110). Q/K arrive already RMS-normed and rotated.
  const void* q = nullptr;  // [rows, ldq]
  int64_t ldq = 0;
  int64_t q_rows_total = 0;
  int q_row0 = 0;
  int64_t q_batch_stride = 0;
  int sq = 0;  // queries per batch item
  void* o = nullptr;  // [rows, ldo]
  int64_t ldo = 0;
  int o_row0 = 0;
  int64_t o_batch_stride = 0;
  int num_heads = 0;
  int batch = 1;
  float scale = 0.08838834764831845f;
  int num_segments = 1;
  KVSegment seg[3];
  // Ulysses peer-memory output (variant 3 only): query row r belongs to rank r / o_rows_per_peer; its output goes to
  // o_peer[that rank] + (r % o_rows_per_peer) * ldo + o_col0 + head * 128 (a buffer mapped over NVLink). 0 = off.
  int o_rows_per_peer = 0;
  int o_col0 = 0;
  void* o_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // Text-encoder prefill (variant 3, one segment with len == sq; Qwen3Model.createCausalMask, Qwen3Model.swift:196-231):
  int causal = 0;        // key j visible to query i iff j <= i
  int kv_group = 1;      // grouped-query attention: query head h uses K / V head h / kv_group (Qwen3Attention.swift:133-145)
  int key_lo = 0, key_hi = 0;  // attention_mask == 1 exactly on keys [key_lo, key_hi); the others get `pad_bias` added (0, 0 = no mask)
  float pad_bias = -1e9f;
  const int* mask_dev = nullptr;   // device int[2] = {key_lo, key_hi}, read by the kernel instead of the two fields above (graph replay)
  int f16 = 0;      // 16-bit storage type: 0 = bf16, 1 = f16
  int variant = 0;  // 0 = auto, 1 = P through shared memory (64-key tiles), 2 = P kept in TMEM (128-key tiles)
  int poly = 0;     // variant 3: 0 = default, -1 = all exponentials on the MUFU, n in {2, 3, 4} = one element in n by polynomial on the FMA pipe
};

cudaError_t attention_launch(const AttnProblem& p, cudaStream_t stream);

}  // namespace f2b
