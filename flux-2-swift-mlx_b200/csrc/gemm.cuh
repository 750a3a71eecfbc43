// gemm.cuh — host-side interface of the tcgen05 GEMM / implicit-GEMM convolution kernel (gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "quant.cuh"

namespace f2b {

enum EpiMode : int {
  EPI_BF16 = 0,      // out_bf16[m, n] = acc (+ bias[n]) (+ res16[m, n])
  EPI_F32 = 1,       // out_f32 [m, n] = acc (+ bias[n])
  EPI_GATE_RES = 2,  // out_f32 [m, n] = res_f32[m, n] + gate[n] * acc          (Flux2Modulation.applyGate + residual)
  EPI_SWIGLU = 3,    // out_bf16[m, nb*BN/2 + j] = silu(acc[j]) * acc[BN/2 + j]  (weight rows pre-interleaved per tile)
  EPI_QKV_ROPE = 4,  // out_bf16 = RoPE(RMSNorm_128(acc) * w) for q/k column ranges, plain for v (BN multiple of 128)
};

struct Epilogue {
  int mode = EPI_BF16;
  int f16 = 0;                  // 16-bit operand / output type: 0 = bf16, 1 = f16
  void* out = nullptr;
  int ldo = 0;                  // elements
  const float* bias = nullptr;  // [N]
  const float* gate = nullptr;  // [N]            (EPI_GATE_RES)
  const float* res = nullptr;   // [M, ldr] fp32   (EPI_GATE_RES)
  const void* res16 = nullptr;  // [M, ldr] 16-bit (EPI_BF16 residual add, VAE resnet shortcut)
  int ldr = 0;
  // EPI_QKV_ROPE: columns [0, qk_cols) get per-128 RMSNorm (weights normw[(col / D_model) ...]) and RoPE
  const float* cos = nullptr;  // [M, 128]
  const float* sin = nullptr;  // [M, 128]
  const float* norm_q = nullptr;  // [128]
  const float* norm_k = nullptr;  // [128]
  int dmodel = 0;                 // D: columns [0,D) = q, [D,2D) = k, [2D,3D) = v
  float eps = 1e-6f;
  // text-encoder variants of EPI_QKV_ROPE (te.cu): grouped-query column ranges (0 = dmodel / 2 * dmodel), rotate-half pairing
  // (element j with j + 64, MLXFast.RoPE traditional = false) instead of adjacent pairs; norm_q / norm_k may be null (no QK-norm)
  int k_col0 = 0, v_col0 = 0;
  int rope_half = 0;
  // two-problem launch (GemmProblem::B_lo): rows below split_row belong to the second problem and take these instead of
  // gate / norm_q / norm_k (the double-stream blocks' text rows: own modulation gate, own QK-norm weights)
  int split_row = 0;
  const float* gate_lo = nullptr;
  const float* norm_q_lo = nullptr;
  const float* norm_k_lo = nullptr;
  // EPI_QKV_ROPE under Ulysses sequence parallelism (sp_hp > 0): head h of q / k / v goes to rank h / sp_hp, i.e. the
  // epilogue stores straight into the all-to-all layout [dest rank][local token][q | k | v][sp_hp * 128]. sp_base[d] is
  // where rank d's slab for THIS rank's tokens starts (a local send buffer, or rank d's gather buffer mapped over
  // NVLink: the projection and its all-to-all are then one kernel); `out` is unused, ldo = 3 * sp_hp * 128.
  int sp_hp = 0;
  void* sp_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // EPI_SWIGLU on the native block-scaled path (mxo.kind != 0): the 16-bit result is not stored; it is quantised in the
  // epilogue (a thread owns 32 consecutive output columns = whole groups) to the format the next GEMM consumes, bit-identical
  // to mx_quantize_act on the 16-bit output. Rows of the last M tile beyond M get scale 1.0.
  MxOut mxo;
};

struct GemmProblem {
  // C[M,N] = A[M,K] * B[N,K]^T ; A, B bf16 row-major with K contiguous
  const void* A = nullptr;
  int64_t lda = 0;
  const void* B = nullptr;
  int64_t ldb = 0;
  int M = 0, N = 0, K = 0;
  // Two problems in one launch (plain 16-bit GEMM only): rows [0, M_lo) of A are multiplied by B_lo (same [N, K] shape and ldb),
  // rows [M_lo, M) by B. The double-stream blocks use it for their text (512 rows) and image streams, which share the A / output
  // buffers but not the weights: the text tiles ride in the image GEMM's tile schedule instead of a launch of their own that
  // occupies a third of the SMs. M_lo must be a multiple of 256.
  const void* B_lo = nullptr;
  int M_lo = 0;
  // implicit-GEMM 3x3 / 1x1 convolution (NHWC activations, OHWI weights): M = batch*H*W, K = taps*Cin
  int conv_taps = 0;  // 0 = plain GEMM, 1 = 1x1, 9 = 3x3 (stride 1: pad 1; stride 2: pad 0 top / left, 1 bottom / right)
  int batch = 1, H = 0, W = 0, Cin = 0;  // H, W = OUTPUT extent
  int conv_stride = 1;                   // 1 | 2
  int conv_no_halo = 0;                  // 1 = 3x3 stride-1 convolution through one TMA box per tap instead of the halo tile (cross-check)
  // Upsample2D (nearest 2x, then conv3x3 pad 1; VAE/ResnetBlock.swift:240-252) as ONE kernel over the low-resolution input:
  // output pixel (2y + py, 2x + px) only ever sees the 2x2 source window {y + py - 1, y + py} x {x + px - 1, x + px}, with the
  // 3x3 taps that fall on the same source pixel summed. H, W = SOURCE extent, the output is [batch, 2H, 2W, N]; B holds the
  // pre-summed weights [N, 16, Cin] = [phase py * 2 + px][tap ty * 2 + tx] (fold_upsample_weights). 4/9 of the FLOPs of the
  // convolution over the upsampled tensor, and the 4x intermediate is never written.
  int conv_up2 = 0;
  int Hin = 0, Win = 0;                  // input extent (0 = H * stride, W * stride)
  Epilogue epi;
  // native block-scaled operands (tcgen05.mma.kind::mxf8f6f4 / mxf4nvf4 .block_scale), single-CTA tiles:
  //   mx = 1 mxfp8: E4M3 bytes, E8M0 scale per 32 elements, K % 128 == 0
  //   mx = 2 mxfp4: E2M1 nibbles (two per byte, low nibble first), E8M0 scale per 32 elements, K % 256 == 0
  //   mx = 3 nvfp4: E2M1 nibbles, E4M3 scale per 16 elements, K % 256 == 0
  // A, B hold the element bytes exactly as MLX packs them (lda / ldb in BYTES); sfa / sfb are the group scales in the
  // tcgen05 scale-factor layout (quant.cuh) and sfa_ld / sfb_ld their 512 B blocks per 128-row block (0 = K / (4 groups),
  // larger when A / B is a K-slice of a wider matrix and sfa / sfb point at the slice's first block)
  int mx = 0;
  const uint8_t* sfa = nullptr;
  const uint8_t* sfb = nullptr;
  int sfa_ld = 0, sfb_ld = 0;
  int force_cta_group = 0;  // 0 = auto, 1, 2
  int force_bn = 0;         // 0 = auto
  // W-only quantized weights, dequantized inside the kernel on their way into the B stage (x · dequant(W)^T, the reference's
  // arithmetic; QuantizedLinear forward, Flux2Pipeline.swift:567-578). wq = flux2b_quant 1..5; plain GEMM, K % 64 == 0.
  // B (and B_lo) = MLX's packed codes [N, K * bits / 8] with ldb in BYTES; wq_scales / wq_biases (+ _lo) row-major
  // [N, wq_sb_ld groups] in the checkpoint's type: affine f16 (or bf16: wq_sb_bf16), block-scaled one byte, no biases.
  // A K-slice of a wider weight passes offset pointers and the full row's group count in wq_sb_ld.
  int wq = 0;
  const void* wq_scales = nullptr;
  const void* wq_biases = nullptr;
  const void* wq_scales_lo = nullptr;
  const void* wq_biases_lo = nullptr;
  int wq_sb_ld = 0;
  int wq_sb_bf16 = 0;
  // Staged variant: when wq_stage is set the layer is first dequantized by a streaming kernel into this 16-bit [N, K] scratch
  // (wq_stage_lo for B_lo; each N * K * 2 bytes, caller-owned, fixed address) and the plain 16-bit kernel reads it — same operand
  // bits as the in-kernel path. For many-row GEMMs (M in the thousands) the dequantisation inside the kernel is repeated for
  // every 256-row M tile and costs more issue slots than the tensor core leaves idle; staging does it once per launch.
  void* wq_stage = nullptr;
  void* wq_stage_lo = nullptr;
  int wq_stage_kb = 0;   // staged path: KB of 16-bit weights per N chunk (0 = the whole layer at once, the measured-fastest setting)
  // Column window of a wider problem (plain 16-bit GEMM only; used by the staged W-only path, which runs a layer in N chunks whose
  // stage stays L2-resident): B holds rows [n_off, N) of the weight, the epilogue addresses outputs / bias / gate by absolute column.
  int n_off = 0;
};

// Returns cudaSuccess or the launch error. Never synchronises.
cudaError_t gemm_launch(const GemmProblem& p, cudaStream_t stream);
// kernels gemm_launch issues for this problem (1, except the staged W-only path: per N chunk one staging kernel per weight set + the GEMM)
int gemm_launch_count(const GemmProblem& p);
// One-time driver entry point lookup; returns false if cuTensorMapEncodeTiled cannot be resolved.
bool gemm_init();
const char* gemm_last_error();
// programmatic dependent launch for the GEMM / attention / LayerNorm kernels (default on; FLUX2B_PDL=0 turns it off)
bool pdl_enabled();

}  // namespace f2b
