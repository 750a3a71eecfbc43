/* flux2b.h — C ABI of the B200-native Flux.2 denoising hot path.
 *
 * The reference (VincentGourbin/flux-2-swift-mlx) has no FFI of its own for this path: its Swift model classes call
 * MLX primitives directly. This header is the cut a maintainer binds from a SwiftPM C target:
 * every entry point names the Swift interface it stands in for (file:line under /root/reference/Sources/Flux2Core).
 *
 * Conventions
 *  - return 0 on success, a negative flux2b_status otherwise (mapped 1:1 onto Flux2Error cases, Flux2Core.swift:14-40);
 *    flux2b_last_error() returns a thread-local message. No exceptions / aborts cross the ABI.
 *  - every pointer argument may be a host or a device pointer (detected with cudaPointerGetAttributes); host buffers
 *    are staged through the context stream, device buffers are used in place. Calls enqueue on the context stream and
 *    synchronise only when an output lives in host memory.
 *  - a context is NOT re-entrant (the reference runs one generation at a time, Flux2Pipeline.swift:99,1158);
 *    distinct contexts are independent (image-parallel = one context per GPU).
 *  - there is no CPU fallback: every compute entry point fails with FLUX2B_ERR_NO_DEVICE without an sm_100 GPU.
 */
#ifndef FLUX2B_H
#define FLUX2B_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flux2b_ctx flux2b_ctx;

typedef enum {
  FLUX2B_OK = 0,
  FLUX2B_ERR_MODEL_NOT_LOADED = -1,     /* Flux2Error.modelNotLoaded */
  FLUX2B_ERR_INVALID_CONFIGURATION = -2,/* Flux2Error.invalidConfiguration */
  FLUX2B_ERR_INSUFFICIENT_MEMORY = -3,  /* Flux2Error.insufficientMemory */
  FLUX2B_ERR_WEIGHT_LOADING = -4,       /* Flux2Error.weightLoadingFailed */
  FLUX2B_ERR_IMAGE_PROCESSING = -5,     /* Flux2Error.imageProcessingFailed */
  FLUX2B_ERR_GENERATION_FAILED = -6,    /* Flux2Error.generationFailed */
  FLUX2B_ERR_CANCELLED = -7,            /* Flux2Error.generationCancelled */
  FLUX2B_ERR_NO_DEVICE = -8,            /* no CUDA device / not sm_100: the product has no CPU path */
  FLUX2B_ERR_CUDA = -9
} flux2b_status;

typedef enum { FLUX2B_F32 = 0, FLUX2B_F16 = 1, FLUX2B_BF16_T = 2, FLUX2B_U32 = 3, FLUX2B_U8 = 4, FLUX2B_I32 = 5 } flux2b_dtype;

/* == TransformerQuantization (Configuration/QuantizationConfig.swift:40-73): (bits, groupSize, mode) =
 *    bf16 (16,-,-) | qint8 (8,64,affine) | int4 (4,64,affine) | mxfp8 (8,32,mxfp8) | mxfp4 (4,32,mxfp4) | nvfp4 (4,16,nvfp4) */
typedef enum { FLUX2B_BF16 = 0, FLUX2B_QINT8 = 1, FLUX2B_INT4 = 2, FLUX2B_MXFP8 = 3, FLUX2B_MXFP4 = 4, FLUX2B_NVFP4 = 5 } flux2b_quant;

/* == Flux2TransformerConfig (Configuration/Flux2Config.swift:210-374) */
typedef struct {
  int patch_size, in_channels, out_channels;
  int num_layers, num_single_layers;
  int attention_head_dim, num_attention_heads;
  int joint_attention_dim;
  int guidance_embeds;
  int axes_dims_rope[4];
  float rope_theta;
  float mlp_ratio;
} flux2b_dit_config;

/* == VAEConfig (Configuration/VAEConfig.swift:7-81); decoder_channels = effectiveDecoderChannels,
 * encoder_channels = blockOutChannels (all zero = the default 128, 256, 512, 512) */
typedef struct {
  int in_channels, out_channels, latent_channels;
  int layers_per_block, norm_num_groups;
  int decoder_channels[4];
  float norm_eps;
  int encoder_channels[4];
} flux2b_vae_config;

/* ------------------------------------------------------------------ lifecycle */
const char* flux2b_version(void);
const char* flux2b_last_error(void);
/* number of visible sm_100 devices (0 on a CPU-only box; never an error) */
int flux2b_device_count(void);
/* dit / vae may be NULL when that half of the path is not used. Flux2Pipeline.init + loadTransformer/loadVAE
 * (Pipeline/Flux2Pipeline.swift:299-316,483-610,692-735) */
int flux2b_create(int device, const flux2b_dit_config* dit, const flux2b_vae_config* vae, int quant, flux2b_ctx** out);
void flux2b_destroy(flux2b_ctx* ctx);
/* run on the caller's CUDA stream (cudaStream_t) instead of the context's own. The context's own stream is a blocking
 * stream (ordered against the legacy default stream); device pointers handed to any call must be ready with respect to the
 * stream the context runs on — that ordering is the caller's responsibility when it supplies a non-blocking stream. */
int flux2b_set_stream(flux2b_ctx* ctx, void* cuda_stream);
int flux2b_synchronize(flux2b_ctx* ctx);
/* options: "compute_f16" (0 = bf16 activations [default], 1 = f16), "fuse_qk_rope" (1), "fuse_swiglu" (1),
 * "attn_variant" (0 auto | 1 | 2), "gemm_cta_group" (0 auto | 1 | 2), "keep_raw_weights" (1), "vae_f16" (1),
 * "uint8_round" (0 = truncate like MLX asType(.uint8) [default], 1 = round to nearest),
 * "native_mx" (0 = W-only: x · dequant(W)^T through the 16-bit GEMM [default, matches the reference's arithmetic];
 *  1 = with quant = mxfp8 the block linears run on tcgen05 block-scaled MMA with on-the-fly mxfp8 activations: faster,
 *  but activations carry E4M3 precision — set before flux2b_finalize_weights),
 * "wq_inkernel" (W-only quantized layers. 1 or 2 = they keep ONLY their packed form (codes + group scales / biases, re-tiled for
 *  the fused kernels): 1 = always dequantized inside the GEMM / GEMV kernels on the way into the operand stage; 2 [default] =
 *  the same for GEMMs of up to 1024 rows and the GEMVs, while many-row GEMMs dequantize the layer in flight once into a
 *  context-owned 16-bit stage (sized for the largest layer) and run the plain kernel on it. 0 = a dense 16-bit expansion of every
 *  layer at finalize. Same output bits in all three — set before flux2b_finalize_weights),
 * "keep_raw_weights" (1 [default]; 0 = forward-only context: after finalize the handed-over Linear tensors — dense weights, and
 *  packed weights together with their scales / biases — are released, so get_tensor / save_prequantized / merge_lora report them
 *  missing; with wq_inkernel or native_mx the packed working copy is then the only resident copy of a quantized layer),
 * "sp_disable" (1 = a sequence-parallel context runs the forward alone on its own GPU: the parity reference of the sharded forward),
 * "group_streams", "attn_poly", "te_graph", "dit_graph", "sp_mode", "sp_overlap", "mx_bn", "mx_fuse_quant", "vae_attn_chunk",
 * "vae_fold_upsample", "wq_stage_kb": see DESIGN.md / INTEGRATION.md */
int flux2b_set_option(flux2b_ctx* ctx, const char* name, int value);

/* ------------------------------------------------------------------ weights (Loading/WeightLoader.swift:567-623)
 * keys are the reference's flattened Swift module paths, e.g. "transformerBlocks.3.attn.toQ.weight",
 * "singleTransformerBlocks.7.attn.toQkvMlp.{weight,scales,biases}", "decoder.upBlocks.1.0.2.conv1.weight" (OHWI),
 * "latentBatchNorm.runningMean". Quantized layers use MLX's layout: weight uint32 [out, in*bits/32],
 * scales [out, in/group] (f16 affine | uint8 E8M0 mx* | uint8 E4M3 nvfp4), biases [out, in/group] f16 (affine only). */
int flux2b_set_tensor(flux2b_ctx* ctx, const char* key, const void* data, int dtype, const int64_t* shape, int ndim);
/* copies the stored tensor (as set, or as produced by flux2b_quantize / flux2b_merge_lora) to `dst`;
 * returns the byte size (also when dst == NULL), negative on error. */
int64_t flux2b_get_tensor(flux2b_ctx* ctx, const char* key, void* dst, size_t capacity, int* dtype, int64_t* shape, int* ndim);
/* build the internal fused / tiled working copies; on-the-fly quantization when quant != bf16 and the layer arrived
 * unquantized: == quantize(model:groupSize:bits:mode:) over every Linear (Pipeline/Flux2Pipeline.swift:567-578). */
int flux2b_finalize_weights(flux2b_ctx* ctx);

/* ------------------------------------------------------------------ safetensors + pre-quantized checkpoint (Loading/PrequantizedCheckpoint.swift)
 * Hand every tensor of a safetensors file to flux2b_set_tensor under its own key. The file must already use the Swift module
 * keys (a pre-quantized export; weights re-saved after WeightLoader's key mapping, WeightLoader.swift:99-204,397-547). The
 * payload size is checked against the header's data_offsets first. Returns the number of tensors, or a negative status. */
int flux2b_load_safetensors(flux2b_ctx* ctx, const char* path);
/* Flux2PrequantizedCheckpoint.save (:225-281): the finalized, quantized transformer of this context — packed uint32 `weight`,
 * `scales`, affine `biases`, and the float parameters — as "flux2-mlx-prequantized-v1" safetensors with the reference's metadata
 * (format, quantization, bits, group_size, mode, component, source, source_fingerprint, created_by[, lora_baked]); written
 * atomically (temporary file + rename). invalidConfiguration for a bf16 context (:234-237). */
int flux2b_save_prequantized(flux2b_ctx* ctx, const char* path, const char* source_name, const char* source_fingerprint, int lora_baked);
/* Flux2PrequantizedCheckpoint.load (:290-387). Everything is validated BEFORE the context is touched: payload integrity,
 * metadata (source_name / source_fingerprint may be NULL = not checked), key set in both directions against the
 * post-quantization manifest of the context's configuration, shapes and dtype categories (uint32 / uint8 exact, floats
 * interchangeable). 0 = tensors handed over, call flux2b_finalize_weights next (no quantize pass runs); 1 = not applied, the
 * context is untouched and the caller falls back to the standard load (reason in flux2b_last_error()); < 0 = misuse.
 * A LoRA-baked export loads with a warning left in flux2b_last_error() (:322-325). */
int flux2b_load_prequantized(flux2b_ctx* ctx, const char* path, const char* source_name, const char* source_fingerprint);
/* Flux2PrequantizedCheckpoint.isValid (:150-166): payload integrity + header + metadata, no device needed. 1 valid, 0 not. */
int flux2b_prequantized_is_valid(const char* path, int quant, const char* source_name, const char* source_fingerprint);
/* W += scale * B·A, in the weight dtype, or dequantize -> add -> requantize for quantized layers
 * (Loading/WeightLoader.swift:736-856). layer_path e.g. "transformerBlocks.0.attn.toQ". A [rank,in], B [out,rank]. */
int flux2b_merge_lora(flux2b_ctx* ctx, const char* layer_path, const void* A, const void* B, int rank, int dtype, float scale);

/* standalone quantizer entry points (mlx quantized()/dequantized(); call sites WeightLoader.swift:795-815).
 * w [rows, cols] -> packed uint32 [rows, cols*bits/32], scales [rows, cols/group], biases (affine only). */
int flux2b_quant_params(int quant, int* bits, int* group_size, int* has_biases, int* scale_dtype);
int flux2b_quantize_matrix(flux2b_ctx* ctx, int quant, const void* w, int w_dtype, int64_t rows, int64_t cols,
                           uint32_t* packed, void* scales, void* biases);
int flux2b_dequantize_matrix(flux2b_ctx* ctx, int quant, const uint32_t* packed, const void* scales, const void* biases,
                             int64_t rows, int64_t cols, void* out, int out_dtype);

/* ------------------------------------------------------------------ DiT forward
 * Flux2Transformer2DModel.callAsFunction(hiddenStates:encoderHiddenStates:timestep:guidance:imgIds:txtIds:)
 * (Transformer/Flux2Transformer.swift:123-327). hidden [B,S_img,in_ch] f32, enc [B,S_txt,joint] (enc_dtype),
 * timestep [B] (sigma in [0,1]; x1000 inside, :145), guidance [B] or NULL, img_ids [S_img,4], txt_ids [S_txt,4] i32,
 * out [B,S_img,out_ch] f32. */
int flux2b_dit_forward(flux2b_ctx* ctx, int B, int S_img, int S_txt, const float* hidden, const void* enc, int enc_dtype,
                       const float* timestep, const float* guidance, const int32_t* img_ids, const int32_t* txt_ids,
                       float* out);
/* klein-9b-kv (Flux2Transformer.swift:346-546): token order [txt | ref | img]; step 0 extracts per-layer post-RoPE
 * reference K/V into the context-owned cache (TransformerKVCache.swift:13-79), later steps attend to it. */
int flux2b_dit_forward_kv_extract(flux2b_ctx* ctx, int B, int S_img, int S_ref, int S_txt, const float* hidden,
                                  const float* ref_hidden, const void* enc, int enc_dtype, const float* timestep,
                                  const float* guidance, const int32_t* img_ids, const int32_t* ref_ids,
                                  const int32_t* txt_ids, float* out);
int flux2b_dit_forward_kv_cached(flux2b_ctx* ctx, int B, int S_img, int S_txt, const float* hidden, const void* enc,
                                 int enc_dtype, const float* timestep, const float* guidance, const int32_t* img_ids,
                                 const int32_t* txt_ids, float* out);
int flux2b_kv_cache_clear(flux2b_ctx* ctx);
/* debugging / parity taps: copy the fp32 residual stream after block `index` of the last forward
 * (0..num_layers-1 double, then single blocks) into dst [S, D] (txt rows first). Enabled by option "record_blocks". */
int64_t flux2b_get_block_output(flux2b_ctx* ctx, int index, float* dst, size_t capacity);

/* ------------------------------------------------------------------ scheduler (Scheduler/FlowMatchEulerScheduler.swift)
 * host fp32 logic, no device work. */
float flux2b_compute_empirical_mu(int image_seq_len, int num_steps);                                  /* :9-28  */
/* sigmas_out capacity >= num_steps + 1; returns number of sigmas written (effective steps + 1); *t_start = first index */
int flux2b_scheduler_set_timesteps(int num_steps, int image_seq_len /* <=0: default 4096 */, float strength,
                                   float* sigmas_out, int* t_start);                                /* :65-115 */
int flux2b_scheduler_set_custom_sigmas(const float* sigmas, int n, float* sigmas_out);               /* :236-260 */
/* x <- x + (sigma_next - sigma) * v ; v = pred, or uncond + cfg*(pred - uncond) when pred_uncond != NULL
 * (:136-156; CFG combine Pipeline/Flux2Pipeline.swift:1970) */
int flux2b_euler_step(flux2b_ctx* ctx, float* sample_inout, const float* pred, const float* pred_uncond, float cfg,
                      float sigma, float sigma_next, size_t n);
int flux2b_scale_noise(flux2b_ctx* ctx, const float* sample, const float* noise, float sigma, float* out, size_t n); /* :195-204 */

/* ------------------------------------------------------------------ latent plumbing (Pipeline/LatentUtils.swift) */
int flux2b_pack_patchified_to_sequence(flux2b_ctx* ctx, const float* in, float* out, int B, int C, int H, int W);     /* :76-86 */
int flux2b_unpack_sequence_to_patchified(flux2b_ctx* ctx, const float* in, float* out, int B, int C, int H, int W);   /* :95-110 */
int flux2b_unpatchify_latents(flux2b_ctx* ctx, const float* in, float* out, int B, int C, int H, int W);              /* :119-142 (C = 32) */
int flux2b_pack_latents_to_patchified(flux2b_ctx* ctx, const float* in, float* out, int B, int C, int H, int W);     /* :186-212 */
int flux2b_bn_latents(flux2b_ctx* ctx, const float* in, float* out, const float* mean, const float* var, float eps,
                      int B, int C, int H, int W, int denormalize);                                                   /* :460-496 */
int flux2b_image_position_ids(int height, int width, int32_t* out /* [(h/16)*(w/16), 4] */);                           /* :256-285 */
int flux2b_text_position_ids(int length, int32_t* out);                                                                /* :291-298 */
int flux2b_reference_position_ids(const int* lat_h, const int* lat_w, int n, int scale, int32_t* out);                 /* :324-346 */

/* ------------------------------------------------------------------ VAE decoder (VAE/AutoencoderKL.swift:129-143)
 * latents [B, latent_ch, h8, w8] f32 NCHW -> image [B, 3, 8*h8, 8*w8] f32 NCHW in [-1,1] */
int flux2b_vae_decode(flux2b_ctx* ctx, int B, int h8, int w8, const float* latents_nchw, float* image_nchw);
/* decode + postprocessVAEOutput (Pipeline/Flux2Pipeline.swift:2425-2468): uint8 [B, H, W, 3] */
int flux2b_vae_decode_u8(flux2b_ctx* ctx, int B, int h8, int w8, const float* latents_nchw, uint8_t* rgb_hwc);

/* ------------------------------------------------------------------ VAE encoder (needs the "encoder.*" / "quantConv.*" tensors)
 * AutoencoderKLFlux2.encode (VAE/AutoencoderKL.swift:90-127; VAE/VAEEncoder.swift:85-115; asymmetric-pad stride-2 downsample
 * VAE/ResnetBlock.swift:189-213): image [B, 3, H, W] f32 NCHW in [-1, 1] (H, W multiples of 8) -> latents
 * [B, latent_ch, H/8, W/8] f32 NCHW. noise = NULL: samplePosterior false (the mean — what every pipeline call site uses,
 * Flux2Pipeline.swift:2199); otherwise standard-normal noise [B, latent_ch, H/8, W/8] supplied by the caller (MLXRandom streams
 * cannot be reproduced): mean + exp(0.5 * logvar) * noise. No scaling factor, no BatchNorm here (:113-123). */
int flux2b_vae_encode(flux2b_ctx* ctx, int B, int H, int W, const float* image_nchw, const float* noise, float* latents_nchw);
/* encodeImageToPackedSequence (Pipeline/Flux2Pipeline+ChainHelpers.swift:75-101) = the per-image body of encodeReferenceImages
 * (Pipeline/Flux2Pipeline.swift:2196-2213): encode -> packLatentsToPatchified -> normalizeLatentsWithBatchNorm ->
 * packPatchifiedToSequence. H, W multiples of 16. seq: [B, (H/16) * (W/16), 4 * latent_ch] f32. */
int flux2b_encode_image_to_sequence(flux2b_ctx* ctx, int B, int H, int W, const float* image_nchw, const float* noise, float* seq);

/* ------------------------------------------------------------------ text-embedding producer
 * The step on the other side of the DiT boundary: one causal prefill of the text encoder and extraction of the hidden
 * states that become `encoderHiddenStates`. Replaces Qwen3Model.forwardWithHiddenStates
 * (FluxTextEncoders/Model/Qwen3/Qwen3Model.swift:104-191) as called by KleinEmbeddingExtractor.extractKleinEmbeddings
 * (Embeddings/KleinEmbeddingExtractor.swift:46-133) and MistralModel.callAsFunction(outputHiddenStates:attentionMask:)
 * (Model/MistralModel.swift:97-148) as called by EmbeddingExtractor.extractFluxEmbeddings (Embeddings/EmbeddingExtractor.swift:202-285).
 * Tokenisation, chat template, truncation and padding stay in Swift; this library receives the padded ids and the 0/1 mask.
 * == Qwen3TextConfig (Configuration/Qwen3Configuration.swift:16-130) / MistralTextConfig: head_dim must be 128 (true of
 * Qwen3-4B / 8B checkpoints, whose config.json carries head_dim = 128, and of Mistral Small 3.2); attention_bias = false. */
typedef struct {
  int vocab_size, hidden_size, intermediate_size;
  int num_layers, num_heads, num_kv_heads, head_dim;
  int qk_norm;                   /* 1 = Qwen3 (q_norm / k_norm per head before RoPE, Qwen3Attention.swift:108-111), 0 = Mistral */
  float rms_norm_eps, rope_theta;
  int max_position_embeddings;   /* Mistral original_max_position_embeddings: longer inputs are refused (Llama-4 query scale != 1); 0 = no check */
} flux2b_te_config;
/* A text-encoder context: set_tensor / load_safetensors / set_option / finalize_weights / prof_* / destroy work as for a DiT
 * context. Tensor keys are the HF / Swift module paths ("model.embed_tokens.weight", "model.layers.3.self_attn.q_proj.weight",
 * ".scales" / ".biases" for MLX-quantized checkpoints, "model.layers.3.self_attn.q_norm.weight", "model.layers.3.mlp.gate_proj.weight",
 * "model.norm.weight", ...). Layers beyond the deepest hidden state that will be requested need not be loaded. `quant` is
 * the mode of packed tensors in the checkpoint (mlx-community 8-bit / 4-bit = FLUX2B_QINT8 / FLUX2B_INT4) or FLUX2B_BF16. */
int flux2b_te_create(int device, const flux2b_te_config* cfg, int quant, flux2b_ctx** out);
/* input_ids [B, S] int32; attention_mask [B, S] int32 0/1 with the ones in one contiguous run (right padding for Klein,
 * left padding for Dev), or NULL = no padding. layer_indices as in the reference: 0 = embedding output, i = output of decoder
 * layer i (1-based), num_layers = after the final norm. out: [B, S, n_layers * hidden] in out_dtype (f32 / f16 / bf16), the
 * layers concatenated along the last axis in the order given (KleinEmbeddingExtractor.swift:111-121). Host or device pointers. */
int flux2b_te_hidden_states(flux2b_ctx* ctx, int B, int S, const int32_t* input_ids, const int32_t* attention_mask,
                            const int* layer_indices, int n_layers, void* out, int out_dtype);

/* ------------------------------------------------------------------ denoise loop (Pipeline/Flux2Pipeline.swift:1933-2052)
 * Flux2StepHook (:64): called after the Euler update of every step with the output latents [1, seq, 128] in HOST
 * memory; whatever the hook leaves in `latents` replaces them (:1985-2001). Flux2StepContext (:42-57). */
typedef struct {
  int step_idx, total_steps;
  float sigma, sigma_next;
  int height, width;
  int is_i2i;
} flux2b_step_context;
typedef int (*flux2b_step_hook)(const flux2b_step_context* sc, float* latents, size_t n, void* user); /* nonzero = cancel */

typedef struct {
  int height, width;             /* pixels, multiples of 16 */
  int num_sigmas;                /* = steps + 1 */
  const float* sigmas;           /* host */
  const float* guidance;         /* [1] or NULL (guidance embedding, Dev) */
  float cfg_scale;               /* classical CFG (klein *Base*): used when enc_uncond != NULL */
  const void* enc;               /* [1, S_txt, joint] */
  const void* enc_uncond;        /* or NULL */
  int enc_dtype;
  int S_txt;
  const float* ref_latents;      /* [S_ref, 128] packed reference tokens (I2I) or NULL */
  const int32_t* ref_ids;        /* [S_ref, 4] */
  int S_ref;
  flux2b_step_hook hook;         /* or NULL: no host round trip inside the loop */
  void* hook_user;
  int kv_cache;                  /* with ref_latents: 1 = the klein-9b-kv loop (Flux2Pipeline.swift:1555-1683): step 0 is
                                  * forwardKVExtract over [txt | refs | output], later steps forwardKVCached over [txt | output]
                                  * against the cached reference K / V; 0 = the standard I2I loop ([output | refs] every step) */
  int S_txt_uncond;              /* token count of enc_uncond [1, S_txt_uncond, joint]; 0 = S_txt. The negative prompt's position
                                  * ids are built from its own length (uncondTextIds, Flux2Pipeline.swift:1690,1960-1975) */
} flux2b_denoise_params;
/* latents [1, S_img, 128] f32 in/out (packed sequence). */
int flux2b_denoise(flux2b_ctx* ctx, const flux2b_denoise_params* p, float* latents_inout);
/* T2I tail: denoise, unpack -> BN denorm -> unpatchify -> VAE decode -> uint8 (Flux2Pipeline.swift:2059-2098) */
int flux2b_generate(flux2b_ctx* ctx, const flux2b_denoise_params* p, float* latents_inout, uint8_t* rgb_hwc);
/* RePaint blend of the only in-tree hook (Flux2Chains/Flux2MaskedInpaintingChain.swift:399-403), device side */
int flux2b_repaint_blend(flux2b_ctx* ctx, float* x, const float* x0, const float* eps, const float* mask, float sigma_next, size_t n);

/* ------------------------------------------------------------------ multi-GPU (new in this build)
 * Ulysses sequence parallelism: tokens of each stream are sharded over `world` ranks; Q/K/V are exchanged by NCCL
 * all-to-all before the fused attention and O after it. nccl_unique_id: 128 bytes from flux2b_sp_unique_id on rank 0. */
int flux2b_sp_unique_id(void* id128);
/* option "sp_mode" (set before sp_init): 0 = NCCL all-to-all, 1 = peer-memory stores fused into the QKV GEMM / attention
 * epilogues. After sp_init every rank calls flux2b_dit_forward / flux2b_denoise with the FULL inputs and receives the
 * FULL output; the library computes the rows of its own token shard. */
int flux2b_sp_init(flux2b_ctx* ctx, const void* id128, int rank, int world);
/* host-only: how a joint sequence is sharded over `world` ranks (the single source of truth used by the forward).
 * Rank r owns txt rows [txt_row0, txt_row0 + txt_rows) and img rows [img_row0, img_row0 + img_rows); after the first
 * exchange it holds all S tokens (rank-major: [txt_0 | img_0 | txt_1 | img_1 | ...]) for heads
 * [rank * heads_per_rank, (rank + 1) * heads_per_rank). Chunk sizes are 16-bit elements per peer message. */
typedef struct {
  int txt_row0, txt_rows, img_row0, img_rows, local_rows, heads_per_rank;
  int64_t qkv_chunk_elems, o_chunk_elems;
} flux2b_sp_layout_t;
int flux2b_sp_layout(int world, int rank, int S_txt, int S_img, int num_heads, flux2b_sp_layout_t* out);

/* ------------------------------------------------------------------ profiling (CUDA events on the context stream) */
enum { FLUX2B_PROF_GEMM = 0, FLUX2B_PROF_ATTN = 1, FLUX2B_PROF_ELEMWISE = 2, FLUX2B_PROF_CONV = 3, FLUX2B_PROF_GEMV = 4,
       FLUX2B_PROF_COMM = 5, FLUX2B_PROF_GROUPNORM = 6 /* VAE GroupNorm(+SiLU): statistics + finalize + apply */, FLUX2B_PROF_KINDS = 7 };
int flux2b_prof_enable(flux2b_ctx* ctx, int on);
int flux2b_prof_reset(flux2b_ctx* ctx);
/* resolves pending events; ms = summed device time, launches, flops and algorithmic bytes of that kernel class */
int flux2b_prof_get(flux2b_ctx* ctx, int kind, double* ms, int64_t* launches, double* flops, double* bytes);
int64_t flux2b_launch_count(flux2b_ctx* ctx);

/* ------------------------------------------------------------------ single-kernel entry points (parity tests) */
/* C[M,N] = A[M,K] · W[N,K]^T; a16/w16 are 16-bit (bf16, or f16 when option compute_f16); epilogue selects the fused tail:
 * 0 = store 16-bit (+bias), 1 = store f32 (+bias), 2 = out_f32 = res + gate*acc, 3 = SwiGLU (W rows pre-tiled), */
int flux2b_op_gemm(flux2b_ctx* ctx, const void* a16, const void* w16, int M, int N, int K, int epilogue, void* out,
                   const float* bias, const float* gate, const float* res, int cta_group, int bn);
/* native block-scaled GEMM (tcgen05.mma.kind::mxf8f6f4 / mxf4nvf4 .block_scale): C[M,N] f32 = q(A)[M,K] · W[N,K]^T where W arrives
 * exactly as MLX packs a quantized Linear of mode `quant` (QuantizedLinear weight / scales, Flux2Pipeline.swift:567-578):
 *   FLUX2B_MXFP8: weight uint32 [N, K/4] = E4M3 bytes,   scales uint8 [N, K/32] E8M0;  K % 128 == 0
 *   FLUX2B_MXFP4: weight uint32 [N, K/8] = E2M1 nibbles, scales uint8 [N, K/32] E8M0;  K % 256 == 0
 *   FLUX2B_NVFP4: weight uint32 [N, K/8] = E2M1 nibbles, scales uint8 [N, K/16] E4M3;  K % 256 == 0
 * and A (16-bit) is quantised on the fly to the same element / scale format. N % 128 == 0; bn = 0 (auto = 128) | 128 | 256;
 * cta_group = 0 (auto: CTA pairs, cta_group::2, for 128-wide tiles) | 1 | 2.
 * aq_out / sfa_out (optional, host or device): the quantised activation bytes [M, K*bits/8] and their scales [M, K/group]. */
int flux2b_op_gemm_mx(flux2b_ctx* ctx, int quant, const void* a16, const uint32_t* w_packed, const uint8_t* w_scales, int M, int N,
                      int K, float* out, uint8_t* aq_out, uint8_t* sfa_out, int bn, int cta_group);
/* = flux2b_op_gemm_mx(ctx, FLUX2B_MXFP8, ..., 0, 0) */
int flux2b_op_gemm_mxfp8(flux2b_ctx* ctx, const void* a16, const uint32_t* w_packed, const uint8_t* w_scales, int M, int N, int K,
                         float* out, uint8_t* a8_out, uint8_t* sfa_out);
/* QuantizedLinear forward as the reference computes it (W-only: x · dequant(W)^T; quantize(model:) Flux2Pipeline.swift:567-578,
 * dequantized() WeightLoader.swift:795-815): x16 [M, K] 16-bit, w_packed / w_scales / w_biases exactly as MLX holds them for mode
 * `quant` (biases NULL for the block-scaled modes; sb_dtype = FLUX2B_F16 or FLUX2B_BF16_T for the affine modes' scales / biases),
 * out f32 [M, N]. K % 64 == 0. in_kernel = 1: the packed codes are dequantized inside the tcgen05 GEMM on their way into the
 * operand stage (the product path: packed weights are the only copy read from HBM); 0: dense 16-bit expansion first (the
 * cross-check). Both produce identical bits. */
int flux2b_op_linear_quantized(flux2b_ctx* ctx, int quant, const void* x16, const uint32_t* w_packed, const void* w_scales,
                               const void* w_biases, int sb_dtype, int M, int N, int K, float* out, int in_kernel, int cta_group);
int flux2b_op_attention(flux2b_ctx* ctx, const void* qkv16 /* [B*S, 3*H*128] */, int B, int S, int H, void* out16 /* [B*S, H*128] */,
                        int variant);
/* causal grouped-query attention with the text encoders' additive padding mask (createCausalMask, Qwen3Model.swift:196-231):
 * qkv16 [S, (Hq + 2 Hkv) * 128] = q | k | v, attention_mask == 1 on keys [key_lo, key_hi) (key_hi = 0: no padding) */
int flux2b_op_attention_causal(flux2b_ctx* ctx, const void* qkv16, int S, int num_heads, int num_kv_heads, int key_lo, int key_hi,
                               void* out16 /* [S, Hq * 128] */);
int flux2b_op_ln_modulate(flux2b_ctx* ctx, const float* x, int rows, int D, const float* shift, const float* scale, void* out16);
int flux2b_op_qk_norm_rope(flux2b_ctx* ctx, void* qkv16, int rows, int D, const float* norm_q, const float* norm_k,
                           const float* cos_t, const float* sin_t);
int flux2b_op_rope_table(flux2b_ctx* ctx, const int32_t* ids, int S, float* cos_out, float* sin_out); /* Flux2RoPE.swift:123-169 */
int flux2b_op_timestep_embedding(flux2b_ctx* ctx, const float* t, int B, float* out /* [B,256] */);     /* Flux2Embeddings.swift:27-44 */
/* NHWC 16-bit conv (3x3 pad 1 or 1x1), OHWI weights, fp32 bias, optional 16-bit residual add */
int flux2b_op_conv2d(flux2b_ctx* ctx, const void* x16, const void* w16, const float* bias, const void* res16, void* out16,
                     int B, int H, int W, int Cin, int Cout, int ksize, int cta_group);
int flux2b_op_groupnorm_silu(flux2b_ctx* ctx, const void* x16, void* y16, const float* gamma, const float* beta, int B,
                             int HW, int C, int G, float eps, int silu);

#ifdef __cplusplus
}
#endif
#endif /* FLUX2B_H */
