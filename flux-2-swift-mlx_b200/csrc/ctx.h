// ctx.h — the context object behind the C ABI: owns the device, stream, weights, workspaces and profiler.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/flux2b.h"
#include "attention.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "quant.cuh"

namespace f2b {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define F2B_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return f2b::fail(FLUX2B_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +      \
                                            (f2b::gemm_last_error()[0] ? std::string(" / ") + f2b::gemm_last_error() : std::string()));    \
  } while (0)
#define F2B_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

inline size_t dtype_size(int dt) {
  switch (dt) {
    case FLUX2B_F32: case FLUX2B_U32: case FLUX2B_I32: return 4;
    case FLUX2B_F16: case FLUX2B_BF16_T: return 2;
    default: return 1;
  }
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  // bumped whenever any buffer of the process is (re)allocated or freed: a captured CUDA graph holds raw device addresses, so
  // a graph is valid only for the epoch it was captured in (steady-state calls allocate nothing, the epoch then stands still)
  static inline uint64_t g_epoch = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (p) { cudaFree(p); ++g_epoch; } p = nullptr; bytes = 0; }
  cudaError_t alloc(size_t n) {
    release();
    ++g_epoch;
    if (n == 0) n = 16;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n; else p = nullptr;
    return e;
  }
  cudaError_t ensure(size_t n) { return (n <= bytes && p) ? cudaSuccess : alloc(n); }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ensure() for buffers whose contents must start as zeros (GroupNorm's arrival counters): zero-filled whenever (re)allocated
inline cudaError_t ensure_zeroed(DevBuf& b, size_t n, cudaStream_t s) {
  void* before = b.p;
  cudaError_t e = b.ensure(n);
  if (e == cudaSuccess && b.p != before) e = cudaMemsetAsync(b.p, 0, b.bytes, s);
  return e;
}

// non-owning view of a context-owned scratch buffer
struct Buf {
  void* p = nullptr;
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Tensor {
  DevBuf buf;
  int dtype = FLUX2B_F32;
  std::vector<int64_t> shape;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

// dense 16-bit working copy of a (possibly fused / re-tiled) Linear weight [N, K]
struct Lin {
  DevBuf w;
  int N = 0, K = 0;
  // native block-scaled path (quant = mxfp8 / mxfp4 / nvfp4 with option "native_mx"): the element bytes [N, K*bits/8]
  // exactly as MLX packs them (rows fused / re-tiled like `w`) and the group scales re-tiled into the tcgen05 scale-factor
  // layout (quant.cuh). `w` is then not materialised.
  DevBuf wq, sfb;
  int mx = 0;  // 0 = dense 16-bit operand in `w`; 1 / 2 / 3 = mxfp8 / mxfp4 / nvfp4 operand in `wq` + `sfb`
  int bn = 0;  // block-scaled operands: GEMM N tile (128 | 256); SwiGLU producers are row-interleaved per tile of this size
  // W-only quantized layers dequantized inside the kernels (option wq_inkernel, default 1): `wq` = MLX's packed codes
  // [N, K * bits / 8] (rows fused / re-tiled like `w`), `ws` / `wb` = group scales / biases row-major [N, K / group] in the
  // checkpoint's type. `w` is then not materialised: the packed form is the only resident copy of the layer.
  DevBuf ws, wb;
  int wmode = 0;       // 0 = off; 1..5 = flux2b_quant of the codes
  int w_sb_bf16 = 0;   // affine scales / biases are bf16 (else f16)
};
struct DoubleBlockW {
  Lin qkv_img, qkv_txt, out_img, out_txt, ff_in_img, ff_out_img, ff_in_txt, ff_out_txt;
  DevBuf nq_img, nk_img, nq_txt, nk_txt;  // fp32 [128]
  bool ff_tiled = false;
};
struct SingleBlockW {
  Lin qkv, mlp, out;
  DevBuf nq, nk;
  bool mlp_tiled = false;
};
struct ConvW {
  DevBuf w;     // 16-bit OHWI [Cout, taps, Cin]
  DevBuf w_up;  // Upsample2D convolutions only: the pre-summed [Cout, 16, Cin] phase kernels of the folded upsample (gemm.cuh: conv_up2)
  DevBuf bias;  // fp32 [Cout]
  int cin = 0, cout = 0, taps = 0;
};
struct NormW { DevBuf gamma, beta; int C = 0; };
struct ResnetW { NormW n1, n2; ConvW c1, c2, sc; bool has_sc = false; int cin = 0, cout = 0; };
// VAE encoder (VAE/VAEEncoder.swift:26-115): optional, built when the "encoder.*" tensors are present
struct VaeEncW {
  bool ready = false;
  ConvW conv_in;                            // Cin zero-padded 3 -> 8 (TMA rows are 16 B)
  std::vector<std::vector<ResnetW>> down;   // [4][layers_per_block]
  std::vector<ConvW> downconv;              // stride-2 3x3, pad bottom / right (ResnetBlock.swift:189-213)
  std::vector<bool> has_down;
  ResnetW mid1, mid2;
  NormW attn_norm;
  Lin attn_qkv, attn_out;
  DevBuf attn_qkv_bias, attn_out_bias;
  NormW norm_out;
  ConvW conv_out, quant;
  bool has_quant = false;
};
struct VaeW {
  bool ready = false;
  VaeEncW enc;
  ConvW post_quant, conv_in, conv_out;
  ResnetW mid1, mid2;
  NormW attn_norm;
  Lin attn_qkv, attn_out;        // qkv fused [3C, C]
  DevBuf attn_qkv_bias, attn_out_bias;
  std::vector<std::vector<ResnetW>> up;
  std::vector<ConvW> upconv;
  std::vector<bool> has_upconv;
  NormW norm_out;
  DevBuf bn_mean, bn_var;  // fp32 [128]
  bool has_bn = false;
};

// Text encoder (FluxTextEncoders/Model/Qwen3/*, Model/Mistral*): one decoder layer's working weights
struct TeLayerW {
  Lin qkv;       // q_proj | k_proj | v_proj stacked: [(Hq + 2 Hkv) * 128, hidden]
  Lin o;         // o_proj [hidden, Hq * 128]
  Lin gate_up;   // gate_proj | up_proj, rows interleaved per 256-row tile as [128 gate | 128 up] (SwiGLU epilogue)
  Lin down;      // down_proj [hidden, intermediate]
  DevBuf ln1, ln2;   // input_layernorm / post_attention_layernorm weights, fp32 [hidden]
  DevBuf nq, nk;     // q_norm / k_norm weights, fp32 [128] (Qwen3 only)
  bool mlp_tiled = false;
};

struct TeGraph {
  int S = 0;
  std::vector<int> layers;
  void* exec = nullptr;        // cudaGraphExec_t
  DevBuf ids, mask, out32;     // stable device addresses baked into the graph: token ids [S], {key_lo, key_hi}, fp32 [S, n * hidden]
  int64_t launches = 0;        // kernels per replay
  uint64_t ws_gen = 0;         // flux2b_ctx::te_ws_gen at capture time: the graph holds the workspaces' addresses of that generation
};

// Ulysses sequence-parallel state (sp.cu). world == 1: off.
struct SpState {
  int world = 1, rank = 0;
  void* comm = nullptr;          // ncclComm_t
  int mode = 0;                  // 0 = NCCL all-to-all, 1 = peer-memory stores fused into the producing kernels
  // peer mappings (mode 1): [rank] -> base of that rank's gather / CAT buffer and barrier flags in this process
  void* gather_peer[8] = {};
  void* cat_peer[8] = {};
  uint32_t* flag_peer[8] = {};
  // side stream + events: the NCCL exchanges of a single-stream block run beside the MLP-in / MLP-part-of-out GEMMs
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
  void* gather_exported = nullptr;  // local buffers whose handles the peers currently hold
  void* cat_exported = nullptr;
  void* flag_exported = nullptr;
  uint32_t epoch = 0;
};

// state that survives between flux2b_denoise calls: position ids per (height, width, S_txt) and the schedule + guidance scalar live on
// the device and are re-uploaded only when they change (host copies are context-owned: an asynchronous upload never reads a
// caller's or a local buffer after the call returned)
struct DenoiseCache {
  int height = 0, width = 0, S_txt = 0, S_txt_u = 0;
  std::vector<int32_t> ids_img, ids_txt, ids_txt_u;
  const void *ids_img_dev = nullptr, *ids_txt_dev = nullptr, *ids_txt_u_dev = nullptr, *sigmas_dev = nullptr;
  std::vector<float> sigmas;   // [sigmas ..., guidance]
};

// A captured launch sequence (the denoise loop of one shape / schedule, the VAE decode of one resolution): replayed with one
// cudaGraphLaunch instead of ~740 kernel launches per image, each of which re-encodes 2 - 4 tensor maps on the host.
struct CachedGraph {
  std::string key;          // every scalar / pointer the sequence depends on
  void* exec = nullptr;     // cudaGraphExec_t; nullptr = this key could not be captured (plain launches are used)
  int64_t launches = 0;     // kernels per replay (for flux2b_launch_count)
  uint64_t epoch = 0, opt_gen = 0;
};

struct ProfKind {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t used = 0;
  double flops = 0, bytes = 0;
  int64_t launches = 0;
};

}  // namespace f2b

struct flux2b_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool has_dit = false, has_vae = false;
  flux2b_dit_config dit{};
  flux2b_vae_config vae{};
  int quant = 0;
  std::map<std::string, int> opt;
  std::unordered_map<std::string, f2b::Tensor> tensors;
  bool finalized = false;
  bool dit_dirty = false;   // tensors changed under a finalized context (flux2b_merge_lora): working copies are rebuilt before the next forward

  // ---- DiT working weights
  int D = 0, H = 0, Hm = 0;
  f2b::Lin x_embed, ctx_embed, t_lin1, t_lin2, g_lin1, g_lin2, mod_img, mod_txt, mod_single, norm_out, proj_out;
  std::vector<f2b::DoubleBlockW> dbl;
  std::vector<f2b::SingleBlockW> sgl;
  f2b::VaeW vw;

  // ---- text encoder (context made by flux2b_te_create)
  bool has_te = false;
  flux2b_te_config te{};
  std::vector<f2b::TeLayerW> te_layers;
  f2b::DevBuf te_embed;   // 16-bit [vocab, hidden] (dequantized when the checkpoint holds a QuantizedEmbedding)
  f2b::DevBuf te_norm;    // final norm weight fp32 [hidden]
  f2b::DevBuf te_ones;    // fp32 ones [hidden]: "gate" of the residual-add GEMM epilogue
  int te_layers_built = 0;
  // the prefill captured as a CUDA graph per (tokens, layer set): ~190 launches of 5 - 40 us each are host-bound otherwise
  std::vector<f2b::TeGraph> te_graphs;
  uint64_t te_ws_gen = 0;   // bumped whenever a text-encoder workspace buffer is reallocated (captured graphs of older generations are stale)

  // ---- workspaces (grown on demand)
  f2b::DevBuf ws_x, ws_xn, ws_qkv, ws_cat, ws_cos, ws_sin, ws_ids, ws_small, ws_hid16, ws_enc16, ws_out;
  // W-only quantized layers, staged variant (option wq_inkernel = 2): 16-bit scratch the layer in flight is dequantized into
  // (one per weight set of a two-stream launch), sized once for the largest quantized Linear of the model
  f2b::DevBuf wq_stage[2];
  size_t wq_stage_max = 0;
  // on-the-fly block-scaled activations (native path): quantised XN and CAT, and their scale factors per row range
  // (0 = text rows of XN, 1 = image rows / whole sequence of XN, 2 and 3 likewise for CAT)
  f2b::DevBuf ws_aq_xn, ws_aq_cat, ws_sfa[4];
  int mx_kind = 0;  // block-scaled kind the DiT block linears run in (0 = 16-bit operands); fixed at finalize
  f2b::DevBuf ws_rec;  // recorded block outputs
  int rec_S = 0, rec_count = 0;
  std::vector<f2b::DevBuf> vae_ws;
  f2b::DevBuf gn_stats;
  // KV cache (klein-9b-kv): per layer K and V of the reference tokens, 16-bit [S_ref, D]
  std::vector<f2b::DevBuf> kv_k, kv_v;
  int kv_S_ref = 0;

  // ---- staging for host <-> device marshalling: a pool that is recycled (not freed) at the end of every API call,
  // so that a steady-state call performs no cudaMalloc / cudaFree (both synchronise the device)
  std::vector<f2b::DevBuf> staging;
  size_t staging_used = 0;
  f2b::DevBuf* stage(size_t bytes) {
    if (staging_used == staging.size()) staging.emplace_back();
    f2b::DevBuf& b = staging[staging_used];
    if (b.ensure(bytes) != cudaSuccess) return nullptr;
    ++staging_used;
    return &b;
  }
  // ---- named, persistent scratch buffers (denoise-loop state, VAE mid-attention scores, ...)
  std::map<std::string, f2b::DevBuf> scratch;
  void* scratch_buf(const char* name, size_t bytes) {
    f2b::DevBuf& b = scratch[name];
    return b.ensure(bytes) == cudaSuccess ? b.p : nullptr;
  }

  f2b::DenoiseCache dn_cache;
  std::vector<f2b::CachedGraph> graphs;   // option dit_graph (default 1)
  uint64_t opt_gen = 0;                   // bumped by set_option / set_stream / sp_init: cached graphs of older generations are dropped

  // ---- sequence parallelism
  f2b::SpState sp;
  f2b::DevBuf ws_sp_gather, ws_sp_o, ws_sp_orecv, ws_sp_flags;

  // ---- profiler
  bool prof_on = false;
  f2b::ProfKind prof[FLUX2B_PROF_KINDS];
  int64_t launches = 0;

  int option(const char* k, int dflt) const { auto it = opt.find(k); return it == opt.end() ? dflt : it->second; }
  bool f16() const { return option("compute_f16", 0) != 0; }
};

namespace f2b {

// RAII profiler bracket: records events around one kernel launch when profiling is on.
struct ProfScope {
  flux2b_ctx* c; int kind; bool on;
  cudaEvent_t stop = nullptr;
  ProfScope(flux2b_ctx* ctx, int k, double flops, double bytes, int n_launches = 1) : c(ctx), kind(k), on(ctx->prof_on) {
    c->launches += n_launches;
    ProfKind& pk = c->prof[kind];
    pk.launches += n_launches; pk.flops += flops; pk.bytes += bytes;
    if (!on) return;
    if (pk.used == pk.ev.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      pk.ev.emplace_back(a, b);
    }
    cudaEventRecord(pk.ev[pk.used].first, c->stream);
    stop = pk.ev[pk.used].second;
    pk.used++;
  }
  ~ProfScope() { if (on) cudaEventRecord(stop, c->stream); }
};

// marshalling helpers (api.cu)
int dev_in(flux2b_ctx* c, const void* src, size_t bytes, const void** out);     // host -> staged device copy, device -> as is
int dev_out(flux2b_ctx* c, void* dst, size_t bytes, void** dev, bool* is_host);  // device scratch for a host destination
int finish_out(flux2b_ctx* c, void* dst, const void* dev, size_t bytes, bool is_host);
int end_call(flux2b_ctx* c, bool sync);
bool is_device_ptr(const void* p);

// weights (weights.cu)
struct PackedMeta { int bits = 0, group = 0, has_b = 0, sb_dtype = FLUX2B_F16; int64_t rows = 0, cols = 0; Tensor* w = nullptr; Tensor* s = nullptr; Tensor* b = nullptr; };
int packed_meta(flux2b_ctx* c, const std::string& base, PackedMeta* m);   // validates a packed Linear's scales / biases (shape, dtype)
int dense16_from_key(flux2b_ctx* c, const std::string& base, DevBuf* out, int* N, int* K);
int finalize_dit(flux2b_ctx* c);
int finalize_vae(flux2b_ctx* c);
int finalize_te(flux2b_ctx* c);
// text-encoder prefill (te.cu): ids [S] int32 on the device; attention_mask == 1 exactly on [key_lo, key_hi) (key_hi == 0: no mask);
// hidden states after layers `layers[i]` (0 = embeddings, num_layers = after the final norm) -> out_f32[:, i * hidden ...], row stride ldo
int te_forward_device(flux2b_ctx* c, int S, const int32_t* ids, int key_lo, int key_hi, const int* layers, int n_layers,
                      float* out_f32, int64_t ldo, const int* mask_dev = nullptr);
void te_destroy_graphs(flux2b_ctx* c);

// Run `body` (pure device work enqueued on c->stream, no allocation after its first execution) through the graph cache: the
// first call for a key executes it directly and then captures it; later calls replay the graph. Falls back to plain launches
// when graphs are off, the profiler is on, or capture fails.
int run_graphed(flux2b_ctx* c, const std::string& key, const std::function<int()>& body);
void destroy_graphs(flux2b_ctx* c);

// forward passes
struct DitIO {
  int B, S_img, S_txt;
  const float* hidden; const void* enc; int enc_dtype;
  const float* timestep; const float* guidance;
  const int32_t* img_ids; const int32_t* txt_ids;
  float* out;
  // kv variants
  int kv_mode = 0;  // 0 none, 1 extract, 2 cached
  int S_ref = 0; const float* ref_hidden = nullptr; const int32_t* ref_ids = nullptr;
};
int dit_forward_device(flux2b_ctx* c, const DitIO& io);  // all pointers already on device

// sequence parallelism (sp.cu)
int sp_all_to_all(flux2b_ctx* c, const void* send, void* recv, size_t chunk_elems16, cudaStream_t stream = nullptr);  // 16-bit elements per peer chunk
int sp_fork(flux2b_ctx* c);   // comm stream waits for everything enqueued on the context stream so far
int sp_join(flux2b_ctx* c);   // context stream waits for everything enqueued on the comm stream so far
int sp_all_gather_f32(flux2b_ctx* c, float* buf, size_t elems_per_rank);                // in place: rank r's slice at r * elems
int sp_map_peers(flux2b_ctx* c);   // (re-)exchange cudaIpc handles of ws_sp_gather / ws_cat / flags when they changed (collective)
int sp_barrier(flux2b_ctx* c);     // all ranks: everything enqueued before it on every rank is visible after it
void sp_destroy(flux2b_ctx* c);
int vae_decode_device(flux2b_ctx* c, int B, int h8, int w8, const void* latents_nhwc16, void** out_nhwc16, int* out_ld);
// image NHWC 16-bit [B, H, W, 8] (3 channels + zero padding) -> moments NHWC 16-bit [B, H/8, W/8, 2 * latent_ch] (after quantConv)
int vae_encode_device(flux2b_ctx* c, int B, int H, int W, const void* image_nhwc16, void** out_nhwc16, int* out_ld);

}  // namespace f2b
