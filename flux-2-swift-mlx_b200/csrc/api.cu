// api.cu — the extern "C" surface declared in include/flux2b.h: lifecycle, tensor hand-over, marshalling of
// host / device pointers, and the thin wrappers that enqueue the kernels. No compute happens on the host.
#include <cstring>
#include <mutex>

#include "ctx.h"

namespace f2b {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int dev_in(flux2b_ctx* c, const void* src, size_t bytes, const void** out) {
  if (!src) { *out = nullptr; return 0; }
  if (is_device_ptr(src)) { *out = src; return 0; }
  DevBuf* b = c->stage(bytes);
  if (!b) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "staging buffer allocation failed"); }
  F2B_CUDA(cudaMemcpyAsync(b->p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  *out = b->p;
  return 0;
}
int dev_out(flux2b_ctx* c, void* dst, size_t bytes, void** dev, bool* is_host) {
  if (is_device_ptr(dst)) { *dev = dst; *is_host = false; return 0; }
  DevBuf* b = c->stage(bytes);
  if (!b) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "staging buffer allocation failed"); }
  *dev = b->p; *is_host = true;
  return 0;
}
int finish_out(flux2b_ctx* c, void* dst, const void* dev, size_t bytes, bool is_host) {
  if (is_host) F2B_CUDA(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
int end_call(flux2b_ctx* c, bool sync) {
  // host buffers were staged: the copies must land (and the staging slots must be idle) before the call returns
  if (sync || c->staging_used) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    c->staging_used = 0;
    if (e != cudaSuccess) return fail(FLUX2B_ERR_CUDA, std::string("stream sync: ") + cudaGetErrorString(e));
  }
  return 0;
}
// in/out buffer: staged copy for host memory, in place for device memory
struct InOut {
  void* dev = nullptr; bool host = false; void* user = nullptr; size_t bytes = 0;
};
static int dev_inout(flux2b_ctx* c, void* p, size_t bytes, InOut* io) {
  io->user = p; io->bytes = bytes;
  if (is_device_ptr(p)) { io->dev = p; io->host = false; return 0; }
  DevBuf* b = c->stage(bytes);
  if (!b) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "staging buffer allocation failed"); }
  F2B_CUDA(cudaMemcpyAsync(b->p, p, bytes, cudaMemcpyHostToDevice, c->stream));
  io->dev = b->p; io->host = true;
  return 0;
}

static int check_ctx(flux2b_ctx* c) {
  if (!c) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null context");
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return fail(FLUX2B_ERR_NO_DEVICE, cudaGetErrorString(e));
  return 0;
}

}  // namespace f2b

using namespace f2b;

extern "C" {

const char* flux2b_version(void) { return "flux2b 0.1.0 (sm_100a)"; }
const char* flux2b_last_error(void) { return g_last_error.c_str(); }

int flux2b_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

static int create_common(int device, int quant, std::unique_ptr<flux2b_ctx>* holder) {
  if (quant < FLUX2B_BF16 || quant > FLUX2B_NVFP4) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "unknown quantization");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(FLUX2B_ERR_NO_DEVICE, "no CUDA device: flux2b has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(FLUX2B_ERR_NO_DEVICE, "device index out of range");
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) return fail(FLUX2B_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200): kernels are sm_100a only");
  if (cudaSetDevice(device) != cudaSuccess) return fail(FLUX2B_ERR_NO_DEVICE, "cudaSetDevice failed");
  if (!gemm_init()) return fail(FLUX2B_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available from the driver");
  std::unique_ptr<flux2b_ctx> c(new flux2b_ctx());
  c->device = device;
  c->quant = quant;
  // a blocking stream: ordered against the legacy default stream, so device buffers a caller produced there (e.g. a pageable
  // cudaMemcpy whose DMA is still in flight when it returns) are complete before this context reads them, and results are
  // visible to it afterwards. Callers that want full concurrency hand in their own stream (flux2b_set_stream).
  if (cudaStreamCreate(&c->stream) != cudaSuccess)
    return fail(FLUX2B_ERR_CUDA, "cudaStreamCreate failed");
  c->own_stream = true;
  *holder = std::move(c);
  return 0;
}

int flux2b_create(int device, const flux2b_dit_config* dit, const flux2b_vae_config* vae, int quant, flux2b_ctx** out) {
  if (!out) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "out is null");
  *out = nullptr;
  std::unique_ptr<flux2b_ctx> c;
  F2B_TRY(create_common(device, quant, &c));
  if (dit) { c->dit = *dit; c->has_dit = true; }
  if (vae) { c->vae = *vae; c->has_vae = true; }
  *out = c.release();
  return 0;
}

int flux2b_te_create(int device, const flux2b_te_config* cfg, int quant, flux2b_ctx** out) {
  if (!out) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "out is null");
  *out = nullptr;
  if (!cfg) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "text-encoder config is null");
  // configuration errors are reported before the device is touched (CPU-testable)
  if (cfg->head_dim != 128) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "text encoder head_dim must be 128");
  if (cfg->vocab_size < 1 || cfg->num_layers < 1 || cfg->num_heads < 1 || cfg->num_kv_heads < 1 ||
      cfg->num_heads % cfg->num_kv_heads != 0 || cfg->hidden_size < 8 || cfg->hidden_size % 8 != 0 || cfg->hidden_size > 8192 ||
      cfg->intermediate_size < 8 || cfg->intermediate_size % 8 != 0 || !(cfg->rope_theta > 1.0f) || !(cfg->rms_norm_eps > 0.0f))
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "inconsistent text-encoder configuration");
  std::unique_ptr<flux2b_ctx> c;
  F2B_TRY(create_common(device, quant, &c));
  c->te = *cfg;
  c->has_te = true;
  *out = c.release();
  return 0;
}

void flux2b_destroy(flux2b_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  sp_destroy(c);
  te_destroy_graphs(c);
  destroy_graphs(c);
  for (auto& pk : c->prof)
    for (auto& e : pk.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int flux2b_set_stream(flux2b_ctx* c, void* s) {
  F2B_TRY(check_ctx(c));
  cudaStreamSynchronize(c->stream);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = reinterpret_cast<cudaStream_t>(s);
  c->own_stream = false;
  ++c->opt_gen;
  return 0;
}
int flux2b_synchronize(flux2b_ctx* c) {
  F2B_TRY(check_ctx(c));
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int flux2b_set_option(flux2b_ctx* c, const char* name, int value) {
  if (!c || !name) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null argument");
  static const char* known[] = {"compute_f16", "fuse_qk_rope", "fuse_swiglu", "attn_variant", "gemm_cta_group",
                                "keep_raw_weights", "vae_f16", "uint8_round", "record_blocks", "vae_conv_cta_group", "sp_mode", "sp_overlap", "native_mx", "mx_bn", "mx_fuse_quant", "attn_poly", "group_streams", "te_graph", "sp_disable", "vae_attn_chunk", "wq_inkernel", "dit_graph", "vae_fold_upsample", "wq_stage_kb"};
  bool ok = false;
  for (const char* k : known) ok = ok || !strcmp(k, name);
  if (!ok) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, std::string("unknown option: ") + name);
  if (c->finalized && (!strcmp(name, "compute_f16") || !strcmp(name, "fuse_swiglu") || !strcmp(name, "vae_f16") || !strcmp(name, "native_mx") || !strcmp(name, "mx_bn") || !strcmp(name, "wq_inkernel")))
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, std::string(name) + " must be set before flux2b_finalize_weights");
  c->opt[name] = value;
  ++c->opt_gen;   // captured launch sequences may depend on any option
  return 0;
}

int flux2b_set_tensor(flux2b_ctx* c, const char* key, const void* data, int dtype, const int64_t* shape, int ndim) {
  F2B_TRY(check_ctx(c));
  if (!key || !data || !shape || ndim < 1 || ndim > 6) return fail(FLUX2B_ERR_WEIGHT_LOADING, "bad set_tensor arguments");
  if (dtype < FLUX2B_F32 || dtype > FLUX2B_I32) return fail(FLUX2B_ERR_WEIGHT_LOADING, "bad dtype");
  Tensor t;
  t.dtype = dtype;
  t.shape.assign(shape, shape + ndim);
  const size_t bytes = (size_t)t.numel() * dtype_size(dtype);
  if (t.buf.alloc(bytes) != cudaSuccess) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, std::string("cudaMalloc failed for ") + key); }
  F2B_CUDA(cudaMemcpyAsync(t.buf.p, data, bytes, is_device_ptr(data) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  c->tensors[key] = std::move(t);
  return 0;
}

int64_t flux2b_get_tensor(flux2b_ctx* c, const char* key, void* dst, size_t capacity, int* dtype, int64_t* shape, int* ndim) {
  F2B_TRY(check_ctx(c));
  auto it = c->tensors.find(key ? key : "");
  if (it == c->tensors.end()) return fail(FLUX2B_ERR_WEIGHT_LOADING, std::string("no such tensor: ") + (key ? key : "(null)"));
  const Tensor& t = it->second;
  const size_t bytes = (size_t)t.numel() * dtype_size(t.dtype);
  if (dtype) *dtype = t.dtype;
  if (ndim) *ndim = (int)t.shape.size();
  if (shape) for (size_t i = 0; i < t.shape.size(); ++i) shape[i] = t.shape[i];
  if (dst) {
    if (capacity < bytes) return fail(FLUX2B_ERR_WEIGHT_LOADING, "destination too small");
    F2B_CUDA(cudaMemcpyAsync(dst, t.buf.p, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
  }
  return (int64_t)bytes;
}

int flux2b_finalize_weights(flux2b_ctx* c) {
  F2B_TRY(check_ctx(c));
  if (c->has_te) F2B_TRY(finalize_te(c));
  if (c->has_dit) { F2B_TRY(finalize_dit(c)); c->dit_dirty = false; }
  if (c->has_vae && c->tensors.count("decoder.convIn.weight")) F2B_TRY(finalize_vae(c));
  c->finalized = true;
  return 0;
}

int flux2b_quant_params(int quant, int* bits, int* group_size, int* has_biases, int* scale_dtype) {
  if (!quant_params(quant, bits, group_size, has_biases, scale_dtype))
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "not a quantized mode");
  return 0;
}

int flux2b_quantize_matrix(flux2b_ctx* c, int quant, const void* w, int w_dtype, int64_t rows, int64_t cols,
                           uint32_t* packed, void* scales, void* biases) {
  F2B_TRY(check_ctx(c));
  int bits, group, has_b, sdt;
  if (!quant_params(quant, &bits, &group, &has_b, &sdt)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "not a quantized mode");
  if (cols % group) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "cols must be a multiple of the group size");
  const void* dw; void *dp, *ds, *db = nullptr; bool hp, hs, hb = false;
  const size_t pbytes = (size_t)rows * cols * bits / 8, sbytes = (size_t)rows * (cols / group) * dtype_size(sdt);
  F2B_TRY(dev_in(c, w, (size_t)rows * cols * dtype_size(w_dtype), &dw));
  F2B_TRY(dev_out(c, packed, pbytes, &dp, &hp));
  F2B_TRY(dev_out(c, scales, sbytes, &ds, &hs));
  if (has_b) {
    if (!biases) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "affine modes need a biases output");
    F2B_TRY(dev_out(c, biases, (size_t)rows * (cols / group) * 2, &db, &hb));
  }
  F2B_CUDA(quantize_matrix(quant, dw, w_dtype, rows, cols, (uint32_t*)dp, ds, db, c->stream));
  F2B_TRY(finish_out(c, packed, dp, pbytes, hp));
  F2B_TRY(finish_out(c, scales, ds, sbytes, hs));
  if (has_b) F2B_TRY(finish_out(c, biases, db, (size_t)rows * (cols / group) * 2, hb));
  return end_call(c, true);
}

int flux2b_dequantize_matrix(flux2b_ctx* c, int quant, const uint32_t* packed, const void* scales, const void* biases,
                             int64_t rows, int64_t cols, void* out, int out_dtype) {
  F2B_TRY(check_ctx(c));
  int bits, group, has_b, sdt;
  if (!quant_params(quant, &bits, &group, &has_b, &sdt)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "not a quantized mode");
  if (cols % group) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "cols must be a multiple of the group size");
  if (out_dtype > FLUX2B_BF16_T) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "output must be a float type");
  const void *dp, *ds, *db = nullptr; void* dout; bool ho;
  F2B_TRY(dev_in(c, packed, (size_t)rows * cols * bits / 8, &dp));
  F2B_TRY(dev_in(c, scales, (size_t)rows * (cols / group) * dtype_size(sdt), &ds));
  if (has_b) F2B_TRY(dev_in(c, biases, (size_t)rows * (cols / group) * 2, &db));
  const size_t obytes = (size_t)rows * cols * dtype_size(out_dtype);
  F2B_TRY(dev_out(c, out, obytes, &dout, &ho));
  F2B_CUDA(dequantize_matrix(quant, (const uint32_t*)dp, ds, db, rows, cols, dout, out_dtype, c->stream));
  F2B_TRY(finish_out(c, out, dout, obytes, ho));
  return end_call(c, true);
}

// ------------------------------------------------------------------------------------------------ DiT forward
static int dit_forward_common(flux2b_ctx* c, int B, int S_img, int S_ref, int S_txt, const float* hidden,
                              const float* ref_hidden, const void* enc, int enc_dtype, const float* timestep,
                              const float* guidance, const int32_t* img_ids, const int32_t* ref_ids,
                              const int32_t* txt_ids, float* out, int kv_mode) {
  F2B_TRY(check_ctx(c));
  if (!c->has_dit) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "context was created without a transformer config");
  if (!hidden || !enc || !timestep || !img_ids || !txt_ids || !out) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null argument");
  if (enc_dtype > FLUX2B_BF16_T) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "enc_dtype must be f32 / f16 / bf16");
  const flux2b_dit_config& g = c->dit;
  DitIO io{};
  io.B = B; io.S_img = S_img; io.S_txt = S_txt; io.enc_dtype = enc_dtype; io.kv_mode = kv_mode; io.S_ref = S_ref;
  const void* p;
  F2B_TRY(dev_in(c, hidden, (size_t)B * S_img * g.in_channels * 4, &p)); io.hidden = (const float*)p;
  F2B_TRY(dev_in(c, enc, (size_t)B * S_txt * g.joint_attention_dim * dtype_size(enc_dtype), &p)); io.enc = p;
  F2B_TRY(dev_in(c, timestep, (size_t)B * 4, &p)); io.timestep = (const float*)p;
  F2B_TRY(dev_in(c, guidance, (size_t)B * 4, &p)); io.guidance = (const float*)p;
  F2B_TRY(dev_in(c, img_ids, (size_t)S_img * 16, &p)); io.img_ids = (const int32_t*)p;
  F2B_TRY(dev_in(c, txt_ids, (size_t)S_txt * 16, &p)); io.txt_ids = (const int32_t*)p;
  if (kv_mode == 1) {
    if (!ref_hidden || !ref_ids || S_ref < 1) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "kv_extract needs reference tokens");
    F2B_TRY(dev_in(c, ref_hidden, (size_t)S_ref * g.in_channels * 4, &p)); io.ref_hidden = (const float*)p;
    F2B_TRY(dev_in(c, ref_ids, (size_t)S_ref * 16, &p)); io.ref_ids = (const int32_t*)p;
  }
  void* dout; bool ho;
  const size_t obytes = (size_t)B * S_img * g.out_channels * 4;
  F2B_TRY(dev_out(c, out, obytes, &dout, &ho));
  io.out = (float*)dout;
  F2B_TRY(dit_forward_device(c, io));
  F2B_TRY(finish_out(c, out, dout, obytes, ho));
  return end_call(c, false);
}

int flux2b_dit_forward(flux2b_ctx* c, int B, int S_img, int S_txt, const float* hidden, const void* enc, int enc_dtype,
                       const float* timestep, const float* guidance, const int32_t* img_ids, const int32_t* txt_ids,
                       float* out) {
  return dit_forward_common(c, B, S_img, 0, S_txt, hidden, nullptr, enc, enc_dtype, timestep, guidance, img_ids, nullptr,
                            txt_ids, out, 0);
}
int flux2b_dit_forward_kv_extract(flux2b_ctx* c, int B, int S_img, int S_ref, int S_txt, const float* hidden,
                                  const float* ref_hidden, const void* enc, int enc_dtype, const float* timestep,
                                  const float* guidance, const int32_t* img_ids, const int32_t* ref_ids,
                                  const int32_t* txt_ids, float* out) {
  return dit_forward_common(c, B, S_img, S_ref, S_txt, hidden, ref_hidden, enc, enc_dtype, timestep, guidance, img_ids,
                            ref_ids, txt_ids, out, 1);
}
int flux2b_dit_forward_kv_cached(flux2b_ctx* c, int B, int S_img, int S_txt, const float* hidden, const void* enc,
                                 int enc_dtype, const float* timestep, const float* guidance, const int32_t* img_ids,
                                 const int32_t* txt_ids, float* out) {
  return dit_forward_common(c, B, S_img, 0, S_txt, hidden, nullptr, enc, enc_dtype, timestep, guidance, img_ids, nullptr,
                            txt_ids, out, 2);
}
int flux2b_kv_cache_clear(flux2b_ctx* c) {
  if (!c) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null context");
  cudaStreamSynchronize(c->stream);
  c->kv_k.clear(); c->kv_v.clear(); c->kv_S_ref = 0;
  return 0;
}
int64_t flux2b_get_block_output(flux2b_ctx* c, int index, float* dst, size_t capacity) {
  F2B_TRY(check_ctx(c));
  if (index < 0 || index >= c->rec_count || !c->ws_rec.p) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "no recorded block output (set option record_blocks=1)");
  const size_t bytes = (size_t)c->rec_S * c->D * 4;
  if (dst) {
    if (capacity < bytes) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "destination too small");
    F2B_CUDA(cudaMemcpyAsync(dst, c->ws_rec.as<float>() + (size_t)index * c->rec_S * c->D, bytes,
                             is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
  }
  return (int64_t)bytes;
}

// ------------------------------------------------------------------------------------------------ profiler
int flux2b_prof_enable(flux2b_ctx* c, int on) { if (!c) return -2; c->prof_on = on != 0; return 0; }
int flux2b_prof_reset(flux2b_ctx* c) {
  if (!c) return -2;
  cudaStreamSynchronize(c->stream);
  for (auto& pk : c->prof) { pk.used = 0; pk.flops = pk.bytes = 0; pk.launches = 0; }
  c->launches = 0;
  return 0;
}
int flux2b_prof_get(flux2b_ctx* c, int kind, double* ms, int64_t* launches, double* flops, double* bytes) {
  F2B_TRY(check_ctx(c));
  if (kind < 0 || kind >= FLUX2B_PROF_KINDS) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad kind");
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  ProfKind& pk = c->prof[kind];
  double total = 0;
  for (size_t i = 0; i < pk.used; ++i) {
    float t = 0;
    if (cudaEventElapsedTime(&t, pk.ev[i].first, pk.ev[i].second) == cudaSuccess) total += t; else cudaGetLastError();
  }
  if (ms) *ms = total;
  if (launches) *launches = pk.launches;
  if (flops) *flops = pk.flops;
  if (bytes) *bytes = pk.bytes;
  return 0;
}
int64_t flux2b_launch_count(flux2b_ctx* c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------------------------------------ single-kernel ops
int flux2b_op_gemm(flux2b_ctx* c, const void* a16, const void* w16, int M, int N, int K, int epilogue, void* out,
                   const float* bias, const float* gate, const float* res, int cta_group, int bn) {
  F2B_TRY(check_ctx(c));
  const void *da, *dw, *db, *dg, *dr;
  F2B_TRY(dev_in(c, a16, (size_t)M * K * 2, &da));
  F2B_TRY(dev_in(c, w16, (size_t)N * K * 2, &dw));
  F2B_TRY(dev_in(c, bias, (size_t)N * 4, &db));
  F2B_TRY(dev_in(c, gate, (size_t)N * 4, &dg));
  F2B_TRY(dev_in(c, res, (size_t)M * N * 4, &dr));
  const int No = (epilogue == EPI_SWIGLU) ? N / 2 : N;
  const size_t obytes = (size_t)M * No * ((epilogue == EPI_F32 || epilogue == EPI_GATE_RES) ? 4 : 2);
  void* dout; bool ho;
  F2B_TRY(dev_out(c, out, obytes, &dout, &ho));
  GemmProblem g;
  g.A = da; g.lda = K; g.B = dw; g.ldb = K; g.M = M; g.N = N; g.K = K;
  g.epi.mode = epilogue; g.epi.f16 = c->f16(); g.epi.out = dout; g.epi.ldo = No;
  g.epi.bias = (const float*)db; g.epi.gate = (const float*)dg; g.epi.res = (const float*)dr; g.epi.ldr = N;
  g.force_cta_group = cta_group; g.force_bn = bn;
  if (epilogue == EPI_GATE_RES && (!gate || !res)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "gate/res required");
  if (epilogue == EPI_SWIGLU && (N % 256)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "SwiGLU epilogue needs N % 256 == 0");
  {
    ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * M * N * K, 2.0 * ((double)M * K + (double)N * K + (double)M * N));
    F2B_CUDA(gemm_launch(g, c->stream));
  }
  F2B_TRY(finish_out(c, out, dout, obytes, ho));
  return end_call(c, false);
}

int flux2b_op_gemm_mx(flux2b_ctx* c, int quant, const void* a16, const uint32_t* w_packed, const uint8_t* w_scales, int M, int N,
                      int K, float* out, uint8_t* aq_out, uint8_t* sfa_out, int bn, int cta_group) {
  F2B_TRY(check_ctx(c));
  const int kind = mx_kind_of_quant(quant);
  if (!kind) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "block-scaled GEMM: quant must be mxfp8, mxfp4 or nvfp4");
  const int kbe = kind == 1 ? 128 : 256, group = kind == 3 ? 16 : 32, bits = kind == 1 ? 8 : 4;
  if (M < 1 || N % 128 || K % kbe || N < 128 || K < kbe)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "block-scaled GEMM needs N % 128 == 0 and K % 128 (fp8) / 256 (fp4) == 0");
  const size_t wrow = (size_t)K * bits / 8, G = (size_t)K / group;
  const void *da, *dw, *ds;
  F2B_TRY(dev_in(c, a16, (size_t)M * K * 2, &da));
  F2B_TRY(dev_in(c, w_packed, (size_t)N * wrow, &dw));
  F2B_TRY(dev_in(c, w_scales, (size_t)N * G, &ds));
  void* dout; bool ho;
  F2B_TRY(dev_out(c, out, (size_t)M * N * 4, &dout, &ho));
  DevBuf wq, sfb, aq, sfa, sfa_plain;
  F2B_CUDA(wq.alloc((size_t)N * wrow));
  F2B_CUDA(sfb.alloc(mx_sf_bytes(kind, N, K)));
  F2B_CUDA(aq.alloc((size_t)M * wrow));
  F2B_CUDA(sfa.alloc(mx_sf_bytes(kind, M, K)));
  F2B_CUDA(mx_copy_rows(kind, (const uint8_t*)dw, (const uint8_t*)ds, 0, wq.as<uint8_t>(), sfb.as<uint8_t>(), 0, N, K, 0, 0, c->stream));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (2.0 + bits / 8.0) * M * K);
    F2B_CUDA(mx_quantize_act(kind, da, K, M, K, c->f16(), aq.as<uint8_t>(), (int64_t)wrow, sfa.as<uint8_t>(), mx_sf_ld(kind, K), 0, c->stream));
  }
  GemmProblem g;
  g.A = aq.p; g.lda = (int64_t)wrow; g.B = wq.p; g.ldb = (int64_t)wrow; g.M = M; g.N = N; g.K = K;
  g.mx = kind; g.sfa = sfa.as<uint8_t>(); g.sfb = sfb.as<uint8_t>();
  g.epi.mode = EPI_F32; g.epi.out = dout; g.epi.ldo = N;
  g.force_bn = bn; g.force_cta_group = cta_group;
  {
    ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * M * N * (double)K, ((double)M + N) * wrow + 4.0 * M * N);
    F2B_CUDA(gemm_launch(g, c->stream));
  }
  F2B_TRY(finish_out(c, out, dout, (size_t)M * N * 4, ho));
  if (aq_out) F2B_CUDA(cudaMemcpyAsync(aq_out, aq.p, (size_t)M * wrow, is_device_ptr(aq_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  if (sfa_out) {
    F2B_CUDA(sfa_plain.alloc((size_t)M * G));
    F2B_CUDA(mx_sf_untile(sfa.as<uint8_t>(), sfa_plain.as<uint8_t>(), M, (int64_t)G, c->stream));
    F2B_CUDA(cudaMemcpyAsync(sfa_out, sfa_plain.p, (size_t)M * G, is_device_ptr(sfa_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  }
  return end_call(c, true);
}
// QuantizedLinear forward, W-only (x · dequant(W)^T): the packed codes are dequantized inside the GEMM kernel (in_kernel = 1) or
// expanded to a dense 16-bit matrix first (in_kernel = 0, the cross-check); both give the same bits.
int flux2b_op_linear_quantized(flux2b_ctx* c, int quant, const void* x16, const uint32_t* w_packed, const void* w_scales,
                               const void* w_biases, int sb_dtype, int M, int N, int K, float* out, int in_kernel, int cta_group) {
  F2B_TRY(check_ctx(c));
  int bits, group, has_b, sdt;
  if (!quant_params(quant, &bits, &group, &has_b, &sdt)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "not a quantized mode");
  if (M < 1 || N < 1 || K < 64 || K % 64) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "quantized linear: K must be a positive multiple of 64");
  if (has_b && (!w_biases || (sb_dtype != FLUX2B_F16 && sb_dtype != FLUX2B_BF16_T)))
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "affine modes need biases and 16-bit scales / biases (f16 or bf16)");
  const size_t wrow = (size_t)K * bits / 8, G = (size_t)K / group, sb_bytes = G * (has_b ? 2 : 1);
  const void *dx, *dw, *ds, *db;
  F2B_TRY(dev_in(c, x16, (size_t)M * K * 2, &dx));
  F2B_TRY(dev_in(c, w_packed, (size_t)N * wrow, &dw));
  F2B_TRY(dev_in(c, w_scales, (size_t)N * sb_bytes, &ds));
  F2B_TRY(dev_in(c, has_b ? w_biases : nullptr, (size_t)N * sb_bytes, &db));
  void* dout; bool ho;
  F2B_TRY(dev_out(c, out, (size_t)M * N * 4, &dout, &ho));
  GemmProblem g;
  g.A = dx; g.lda = K; g.M = M; g.N = N; g.K = K;
  g.epi.mode = EPI_F32; g.epi.f16 = c->f16(); g.epi.out = dout; g.epi.ldo = N;
  g.force_cta_group = cta_group;
  DevBuf dense;
  if (in_kernel) {
    g.wq = quant; g.B = dw; g.ldb = (int64_t)wrow; g.wq_scales = ds; g.wq_biases = db; g.wq_sb_ld = (int)G;
    g.wq_sb_bf16 = (has_b && sb_dtype == FLUX2B_BF16_T) ? 1 : 0;
  } else {
    F2B_CUDA(dense.alloc((size_t)N * K * 2));
    F2B_CUDA(dequantize_matrix(quant, (const uint32_t*)dw, ds, db, N, K, dense.p, c->f16() ? FLUX2B_F16 : FLUX2B_BF16_T, c->stream,
                               has_b ? sb_dtype : FLUX2B_F16));
    g.B = dense.p; g.ldb = K; g.force_bn = 256;
  }
  {
    ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * M * N * (double)K, 2.0 * M * K + (double)N * wrow + 4.0 * M * N);
    F2B_CUDA(gemm_launch(g, c->stream));
  }
  F2B_TRY(finish_out(c, out, dout, (size_t)M * N * 4, ho));
  return end_call(c, true);
}

int flux2b_op_gemm_mxfp8(flux2b_ctx* c, const void* a16, const uint32_t* w_packed, const uint8_t* w_scales, int M, int N, int K,
                         float* out, uint8_t* a8_out, uint8_t* sfa_out) {
  return flux2b_op_gemm_mx(c, FLUX2B_MXFP8, a16, w_packed, w_scales, M, N, K, out, a8_out, sfa_out, 0, 0);
}

int flux2b_op_attention(flux2b_ctx* c, const void* qkv16, int B, int S, int H, void* out16, int variant) {
  F2B_TRY(check_ctx(c));
  const int D = H * 128;
  const void* dq;
  F2B_TRY(dev_in(c, qkv16, (size_t)B * S * 3 * D * 2, &dq));
  void* dout; bool ho;
  const size_t obytes = (size_t)B * S * D * 2;
  F2B_TRY(dev_out(c, out16, obytes, &dout, &ho));
  AttnProblem a;
  a.q = dq; a.ldq = 3 * D; a.q_rows_total = (int64_t)B * S; a.q_row0 = 0; a.q_batch_stride = S; a.sq = S;
  a.o = dout; a.ldo = D; a.o_row0 = 0; a.o_batch_stride = S;
  a.num_heads = H; a.batch = B; a.scale = 1.0f / sqrtf(128.f);
  a.num_segments = 1;
  a.seg[0].k = (const uint16_t*)dq + D; a.seg[0].v = (const uint16_t*)dq + 2 * D;
  a.seg[0].ldk = a.seg[0].ldv = 3 * D; a.seg[0].rows_total = (int64_t)B * S; a.seg[0].row0 = 0; a.seg[0].len = S;
  a.seg[0].batch_stride = S;
  a.f16 = c->f16(); a.variant = variant; a.poly = c->option("attn_poly", 0);
  {
    ProfScope ps(c, FLUX2B_PROF_ATTN, 4.0 * B * (double)S * S * D, 8.0 * B * (double)S * D);
    F2B_CUDA(attention_launch(a, c->stream));
  }
  F2B_TRY(finish_out(c, out16, dout, obytes, ho));
  return end_call(c, false);
}

int flux2b_op_ln_modulate(flux2b_ctx* c, const float* x, int rows, int D, const float* shift, const float* scale, void* out16) {
  F2B_TRY(check_ctx(c));
  const void *dx, *dsh, *dsc;
  F2B_TRY(dev_in(c, x, (size_t)rows * D * 4, &dx));
  F2B_TRY(dev_in(c, shift, (size_t)D * 4, &dsh));
  F2B_TRY(dev_in(c, scale, (size_t)D * 4, &dsc));
  void* dout; bool ho;
  F2B_TRY(dev_out(c, out16, (size_t)rows * D * 2, &dout, &ho));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)rows * D * 6);
    F2B_CUDA(ln_modulate((const float*)dx, D, dout, D, rows, D, (const float*)dsh, (const float*)dsc, 0, rows, 1e-6f, c->f16(), c->stream));
  }
  F2B_TRY(finish_out(c, out16, dout, (size_t)rows * D * 2, ho));
  return end_call(c, false);
}

int flux2b_op_qk_norm_rope(flux2b_ctx* c, void* qkv16, int rows, int D, const float* norm_q, const float* norm_k,
                           const float* cos_t, const float* sin_t) {
  F2B_TRY(check_ctx(c));
  InOut io;
  F2B_TRY(dev_inout(c, qkv16, (size_t)rows * 3 * D * 2, &io));
  const void *nq, *nk, *cs, *sn;
  F2B_TRY(dev_in(c, norm_q, 512, &nq));
  F2B_TRY(dev_in(c, norm_k, 512, &nk));
  F2B_TRY(dev_in(c, cos_t, (size_t)rows * 512, &cs));
  F2B_TRY(dev_in(c, sin_t, (size_t)rows * 512, &sn));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)rows * D * 8);
    F2B_CUDA(qk_norm_rope(io.dev, 3 * D, rows, D, (const float*)nq, (const float*)nk, (const float*)cs, (const float*)sn, 1e-6f, c->f16(), c->stream));
  }
  F2B_TRY(finish_out(c, qkv16, io.dev, io.bytes, io.host));
  return end_call(c, false);
}

int flux2b_op_rope_table(flux2b_ctx* c, const int32_t* ids, int S, float* cos_out, float* sin_out) {
  F2B_TRY(check_ctx(c));
  const void* di;
  F2B_TRY(dev_in(c, ids, (size_t)S * 16, &di));
  void *dc, *ds; bool hc, hs;
  F2B_TRY(dev_out(c, cos_out, (size_t)S * 512, &dc, &hc));
  F2B_TRY(dev_out(c, sin_out, (size_t)S * 512, &ds, &hs));
  int axes[4] = {32, 32, 32, 32};
  float theta = 2000.f;
  if (c->has_dit) { for (int i = 0; i < 4; ++i) axes[i] = c->dit.axes_dims_rope[i]; theta = c->dit.rope_theta; }
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * 1024);
    F2B_CUDA(rope_table((const int32_t*)di, S, axes, theta, (float*)dc, (float*)ds, c->stream));
  }
  F2B_TRY(finish_out(c, cos_out, dc, (size_t)S * 512, hc));
  F2B_TRY(finish_out(c, sin_out, ds, (size_t)S * 512, hs));
  return end_call(c, false);
}

int flux2b_op_timestep_embedding(flux2b_ctx* c, const float* t, int B, float* out) {
  F2B_TRY(check_ctx(c));
  const void* dt;
  F2B_TRY(dev_in(c, t, (size_t)B * 4, &dt));
  void* dout; bool ho;
  F2B_TRY(dev_out(c, out, (size_t)B * 1024, &dout, &ho));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 0);
    F2B_CUDA(timestep_sinusoid((const float*)dt, (float*)dout, B, 1000.0f, c->stream));
  }
  F2B_TRY(finish_out(c, out, dout, (size_t)B * 1024, ho));
  return end_call(c, false);
}

int flux2b_op_conv2d(flux2b_ctx* c, const void* x16, const void* w16, const float* bias, const void* res16, void* out16,
                     int B, int H, int W, int Cin, int Cout, int ksize, int cta_group) {
  F2B_TRY(check_ctx(c));
  if (ksize != 1 && ksize != 3) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "ksize must be 1 or 3");
  const void *dx, *dw, *db, *dr;
  const size_t npix = (size_t)B * H * W;
  F2B_TRY(dev_in(c, x16, npix * Cin * 2, &dx));
  F2B_TRY(dev_in(c, w16, (size_t)Cout * ksize * ksize * Cin * 2, &dw));
  F2B_TRY(dev_in(c, bias, (size_t)Cout * 4, &db));
  F2B_TRY(dev_in(c, res16, npix * Cout * 2, &dr));
  void* dout; bool ho;
  F2B_TRY(dev_out(c, out16, npix * Cout * 2, &dout, &ho));
  GemmProblem g;
  g.A = dx; g.lda = Cin; g.B = dw; g.ldb = (int64_t)ksize * ksize * Cin;
  g.M = (int)npix; g.N = Cout; g.K = ksize * ksize * Cin;
  g.conv_taps = ksize * ksize; g.batch = B; g.H = H; g.W = W; g.Cin = Cin;
  g.epi.mode = EPI_BF16; g.epi.f16 = c->f16(); g.epi.out = dout; g.epi.ldo = Cout; g.epi.bias = (const float*)db;
  g.epi.res16 = dr; g.epi.ldr = Cout;
  g.force_cta_group = cta_group;
  {
    ProfScope ps(c, FLUX2B_PROF_CONV, 2.0 * npix * Cout * (double)g.K, 2.0 * (npix * Cin + npix * Cout + (double)Cout * g.K));
    F2B_CUDA(gemm_launch(g, c->stream));
  }
  F2B_TRY(finish_out(c, out16, dout, npix * Cout * 2, ho));
  return end_call(c, false);
}

int flux2b_op_groupnorm_silu(flux2b_ctx* c, const void* x16, void* y16, const float* gamma, const float* beta, int B,
                             int HW, int C, int G, float eps, int silu) {
  F2B_TRY(check_ctx(c));
  const void *dx, *dg, *db;
  const size_t bytes = (size_t)B * HW * C * 2;
  F2B_TRY(dev_in(c, x16, bytes, &dx));
  F2B_TRY(dev_in(c, gamma, (size_t)C * 4, &dg));
  F2B_TRY(dev_in(c, beta, (size_t)C * 4, &db));
  void* dout; bool ho;
  F2B_TRY(dev_out(c, y16, bytes, &dout, &ho));
  F2B_CUDA(ensure_zeroed(c->gn_stats, groupnorm_ws_bytes(B, G), c->stream));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)bytes * 3);
    F2B_CUDA(groupnorm_silu(dx, dout, (const float*)dg, (const float*)db, c->gn_stats.as<double>(), B, HW, C, G, eps, silu != 0, c->f16(), c->stream));
  }
  F2B_TRY(finish_out(c, y16, dout, bytes, ho));
  return end_call(c, false);
}

}  // extern "C"
