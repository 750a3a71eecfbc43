// pipeline.cu — scheduler (host fp32), latent plumbing, Euler step, LoRA merge and the denoise loop body.
#include <cmath>
#include <cstring>

#include "ctx.h"

using namespace f2b;

namespace f2b {
int vector_f32_from_key(flux2b_ctx* c, const std::string& key, DevBuf* out, int64_t expect, bool required, float fill);
}

namespace f2b {

void destroy_graphs(flux2b_ctx* c) {
  for (CachedGraph& g : c->graphs)
    if (g.exec) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(g.exec));
  c->graphs.clear();
}

int run_graphed(flux2b_ctx* c, const std::string& key, const std::function<int()>& body) {
  if (!c->option("dit_graph", 1) || c->prof_on) return body();
  // graphs of an older allocation epoch hold stale addresses; graphs of an older option generation a different launch sequence
  for (size_t i = 0; i < c->graphs.size();) {
    if (c->graphs[i].epoch != DevBuf::g_epoch || c->graphs[i].opt_gen != c->opt_gen) {
      if (c->graphs[i].exec) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(c->graphs[i].exec));
      c->graphs.erase(c->graphs.begin() + (long)i);
    } else ++i;
  }
  for (CachedGraph& g : c->graphs)
    if (g.key == key) {
      if (!g.exec) return body();
      F2B_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(g.exec), c->stream));
      c->launches += g.launches;
      return 0;
    }
  // first call for this key: the real execution doubles as the warm-up that grows every workspace
  const uint64_t epoch0 = DevBuf::g_epoch;
  const int64_t l0 = c->launches;
  F2B_TRY(body());
  if (DevBuf::g_epoch != epoch0) return 0;   // buffers moved during this call: capture on the next one
  if (c->graphs.size() >= 16) destroy_graphs(c);
  CachedGraph g;
  g.key = key; g.launches = c->launches - l0; g.epoch = DevBuf::g_epoch; g.opt_gen = c->opt_gen;
  int64_t prof_launches[FLUX2B_PROF_KINDS]; double prof_flops[FLUX2B_PROF_KINDS], prof_bytes[FLUX2B_PROF_KINDS];
  for (int i = 0; i < FLUX2B_PROF_KINDS; ++i) { prof_launches[i] = c->prof[i].launches; prof_flops[i] = c->prof[i].flops; prof_bytes[i] = c->prof[i].bytes; }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  if (ok) {
    const int rc = body();
    ok = cudaStreamEndCapture(c->stream, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
  }
  // the capture enqueued nothing: undo its bookkeeping
  c->launches = l0 + g.launches;
  for (int i = 0; i < FLUX2B_PROF_KINDS; ++i) { c->prof[i].launches = prof_launches[i]; c->prof[i].flops = prof_flops[i]; c->prof[i].bytes = prof_bytes[i]; }
  if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok || DevBuf::g_epoch != g.epoch) {   // not capturable (or it allocated): remember, so that later calls do not retry
    cudaGetLastError();
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
    g.epoch = DevBuf::g_epoch;
  }
  g.exec = exec;
  c->graphs.push_back(std::move(g));
  return 0;
}

}  // namespace f2b

extern "C" {

// ------------------------------------------------------------------ scheduler (host; FlowMatchEulerScheduler.swift)
float flux2b_compute_empirical_mu(int image_seq_len, int num_steps) {
  const float a1 = 8.73809524e-05f, b1 = 1.89833333f, a2 = 0.00016927f, b2 = 0.45666666f;
  if (image_seq_len > 4300) return a2 * (float)image_seq_len + b2;
  const float m_200 = a2 * (float)image_seq_len + b2;
  const float m_10 = a1 * (float)image_seq_len + b1;
  const float a = (m_200 - m_10) / 190.0f;
  const float b = m_200 - 200.0f * a;
  return a * (float)num_steps + b;
}

int flux2b_scheduler_set_timesteps(int num_steps, int image_seq_len, float strength, float* sigmas_out, int* t_start) {
  if (num_steps < 1 || !sigmas_out) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad scheduler arguments");
  const float mu = flux2b_compute_empirical_mu(image_seq_len > 0 ? image_seq_len : 4096, num_steps);
  std::vector<float> all((size_t)num_steps + 1);
  const float e_mu = expf(mu);
  for (int i = 0; i < num_steps; ++i) {
    const float sigma = 1.0f - (float)i / (float)num_steps;
    all[i] = e_mu / (e_mu + powf(1.0f / sigma - 1.0f, 1.0f));  // timeShift(mu, sigma: 1.0, t) :123-128
  }
  all[num_steps] = 0.0f;
  const float cs = fmaxf(0.01f, fminf(1.0f, strength));
  int start = num_steps - (int)((float)num_steps * cs);
  if (start < 0) start = 0;
  if (t_start) *t_start = start;
  int n = 0;
  for (int i = start; i <= num_steps; ++i) sigmas_out[n++] = all[i];
  return n;
}

int flux2b_scheduler_set_custom_sigmas(const float* sigmas, int n, float* sigmas_out) {
  if (!sigmas || n < 1 || !sigmas_out) return 0;  // empty input is ignored by the reference (:237-240)
  memcpy(sigmas_out, sigmas, sizeof(float) * n);
  if (sigmas[n - 1] != 0.0f) { sigmas_out[n] = 0.0f; return n + 1; }
  return n;
}

static int check(flux2b_ctx* c) {
  if (!c) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null context");
  if (cudaSetDevice(c->device) != cudaSuccess) return fail(FLUX2B_ERR_NO_DEVICE, "cudaSetDevice failed");
  return 0;
}

int flux2b_euler_step(flux2b_ctx* c, float* sample, const float* pred, const float* pred_uncond, float cfg, float sigma,
                      float sigma_next, size_t n) {
  F2B_TRY(check(c));
  const bool host = !is_device_ptr(sample);
  const void *dp, *du;
  void* dx = sample;
  if (host) {
    const void* tmp;
    F2B_TRY(dev_in(c, sample, n * 4, &tmp));
    dx = const_cast<void*>(tmp);
  }
  F2B_TRY(dev_in(c, pred, n * 4, &dp));
  F2B_TRY(dev_in(c, pred_uncond, n * 4, &du));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 12.0 * n);
    F2B_CUDA(euler_step((float*)dx, (const float*)dp, (const float*)du, cfg, sigma_next - sigma, (int64_t)n, c->stream));
  }
  if (host) F2B_CUDA(cudaMemcpyAsync(sample, dx, n * 4, cudaMemcpyDeviceToHost, c->stream));
  return end_call(c, false);
}

int flux2b_scale_noise(flux2b_ctx* c, const float* sample, const float* noise, float sigma, float* out, size_t n) {
  F2B_TRY(check(c));
  const void *da, *db; void* dout; bool ho;
  F2B_TRY(dev_in(c, sample, n * 4, &da));
  F2B_TRY(dev_in(c, noise, n * 4, &db));
  F2B_TRY(dev_out(c, out, n * 4, &dout, &ho));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 12.0 * n);
    F2B_CUDA(scale_noise((const float*)da, (const float*)db, sigma, (float*)dout, (int64_t)n, c->stream));
  }
  F2B_TRY(finish_out(c, out, dout, n * 4, ho));
  return end_call(c, false);
}

int flux2b_repaint_blend(flux2b_ctx* c, float* x, const float* x0, const float* eps, const float* mask, float sigma_next, size_t n) {
  F2B_TRY(check(c));
  const bool host = !is_device_ptr(x);
  void* dx = x;
  const void *d0, *de, *dm;
  if (host) { const void* tmp; F2B_TRY(dev_in(c, x, n * 4, &tmp)); dx = const_cast<void*>(tmp); }
  F2B_TRY(dev_in(c, x0, n * 4, &d0));
  F2B_TRY(dev_in(c, eps, n * 4, &de));
  F2B_TRY(dev_in(c, mask, n * 4, &dm));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 20.0 * n);
    F2B_CUDA(repaint_blend((float*)dx, (const float*)d0, (const float*)de, (const float*)dm, sigma_next, (int64_t)n, c->stream));
  }
  if (host) F2B_CUDA(cudaMemcpyAsync(x, dx, n * 4, cudaMemcpyDeviceToHost, c->stream));
  return end_call(c, false);
}

// ------------------------------------------------------------------ latent plumbing (LatentUtils.swift)
static int permute_call(flux2b_ctx* c, const float* in, float* out, size_t n, int ndim, const int* oshape, const int64_t* istr) {
  F2B_TRY(check(c));
  const void* di; void* dout; bool ho;
  F2B_TRY(dev_in(c, in, n * 4, &di));
  F2B_TRY(dev_out(c, out, n * 4, &dout, &ho));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 8.0 * n);
    F2B_CUDA(permute_f32((const float*)di, (float*)dout, ndim, oshape, istr, c->stream));
  }
  F2B_TRY(finish_out(c, out, dout, n * 4, ho));
  return end_call(c, false);
}
// [B,C,H,W] -> [B,H*W,C]
int flux2b_pack_patchified_to_sequence(flux2b_ctx* c, const float* in, float* out, int B, int C, int H, int W) {
  const int os[4] = {B, H, W, C};
  const int64_t is[4] = {(int64_t)C * H * W, W, 1, (int64_t)H * W};
  return permute_call(c, in, out, (size_t)B * C * H * W, 4, os, is);
}
// [B,H*W,C] -> [B,C,H,W]
int flux2b_unpack_sequence_to_patchified(flux2b_ctx* c, const float* in, float* out, int B, int C, int H, int W) {
  const int os[4] = {B, C, H, W};
  const int64_t is[4] = {(int64_t)H * W * C, 1, (int64_t)W * C, C};
  return permute_call(c, in, out, (size_t)B * C * H * W, 4, os, is);
}
// [B, C*2*2, H, W] -> [B, C, 2H, 2W]: reshape [B,C,p,p,H,W] -> transpose (0,1,4,2,5,3)
int flux2b_unpatchify_latents(flux2b_ctx* c, const float* in, float* out, int B, int C, int H, int W) {
  const int p = 2;
  const int os[6] = {B, C, H, p, W, p};
  const int64_t hw = (int64_t)H * W;
  // input dims [B, C, ph, pw, H, W] strides
  const int64_t sB = (int64_t)C * p * p * hw, sC = (int64_t)p * p * hw, sPh = (int64_t)p * hw, sPw = hw, sH = W, sW = 1;
  const int64_t is[6] = {sB, sC, sH, sPh, sW, sPw};
  return permute_call(c, in, out, (size_t)B * C * p * p * hw, 6, os, is);
}
// [B, C, H, W] -> [B, C*4, H/2, W/2]: reshape [B,C,pH,p,pW,p] -> transpose (0,2,4,1,3,5) -> [B,pH,pW,C*4] -> NCHW
int flux2b_pack_latents_to_patchified(flux2b_ctx* c, const float* in, float* out, int B, int C, int H, int W) {
  const int p = 2, pH = H / p, pW = W / p;
  // output [B, (C, ph, pw), pH, pW]
  const int os[6] = {B, C, p, p, pH, pW};
  const int64_t is[6] = {(int64_t)C * H * W, (int64_t)H * W, W, 1, (int64_t)p * W, p};
  return permute_call(c, in, out, (size_t)B * C * H * W, 6, os, is);
}
int flux2b_bn_latents(flux2b_ctx* c, const float* in, float* out, const float* mean, const float* var, float eps, int B,
                      int C, int H, int W, int denormalize) {
  F2B_TRY(check(c));
  const size_t n = (size_t)B * C * H * W;
  const void *di, *dm, *dv; void* dout; bool ho;
  F2B_TRY(dev_in(c, in, n * 4, &di));
  F2B_TRY(dev_in(c, mean, (size_t)C * 4, &dm));
  F2B_TRY(dev_in(c, var, (size_t)C * 4, &dv));
  F2B_TRY(dev_out(c, out, n * 4, &dout, &ho));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 8.0 * n);
    F2B_CUDA(bn_affine_nchw((const float*)di, (float*)dout, (const float*)dm, (const float*)dv, eps, B, C, (int64_t)H * W, denormalize != 0, c->stream));
  }
  F2B_TRY(finish_out(c, out, dout, n * 4, ho));
  return end_call(c, false);
}
int flux2b_image_position_ids(int height, int width, int32_t* out) {
  const int h = height / 8 / 2, w = width / 8 / 2;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int32_t* p = out + ((size_t)y * w + x) * 4;
      p[0] = 0; p[1] = y; p[2] = x; p[3] = 0;
    }
  return h * w;
}
int flux2b_text_position_ids(int length, int32_t* out) {
  for (int l = 0; l < length; ++l) { out[4 * l] = 0; out[4 * l + 1] = 0; out[4 * l + 2] = 0; out[4 * l + 3] = l; }
  return length;
}
int flux2b_reference_position_ids(const int* lat_h, const int* lat_w, int n, int scale, int32_t* out) {
  size_t k = 0;
  for (int i = 0; i < n; ++i) {
    const int32_t t = scale + scale * i;
    for (int y = 0; y < lat_h[i]; ++y)
      for (int x = 0; x < lat_w[i]; ++x) { out[k++] = t; out[k++] = y; out[k++] = x; out[k++] = 0; }
  }
  return (int)(k / 4);
}

// ------------------------------------------------------------------ LoRA merge (WeightLoader.swift:736-856)
int flux2b_merge_lora(flux2b_ctx* c, const char* layer_path, const void* A, const void* B, int rank, int dtype, float scale) {
  F2B_TRY(check(c));
  if (!layer_path || !A || !B || rank < 1 || dtype > FLUX2B_BF16_T) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad merge_lora arguments");
  const std::string base(layer_path);
  auto it = c->tensors.find(base + ".weight");
  if (it == c->tensors.end()) return fail(FLUX2B_ERR_WEIGHT_LOADING, "no weight found for layer: " + base);
  Tensor& w = it->second;
  const bool quantized = w.dtype == FLUX2B_U32;
  int bits = 16, group = 0, has_b = 0, sdt = 0;
  int64_t out_dim = w.shape[0], in_dim = 0;
  if (quantized) {
    quant_params(c->quant, &bits, &group, &has_b, &sdt);
    in_dim = w.shape[1] * 32 / bits;
  } else {
    in_dim = w.numel() / out_dim;
  }
  // merged weight dtype: f16 for the quantized path (dequantized(...).asType(.float16)), else the weight's own dtype
  const int wd = quantized ? FLUX2B_F16 : w.dtype;
  // A, B -> weight dtype -> fp32 staging (the casts of :806-807, :832-833)
  auto stage_rounded = [&](const void* src, size_t n, DevBuf* out32) -> int {
    const void* d;
    F2B_TRY(dev_in(c, src, n * dtype_size(dtype), &d));
    DevBuf t16;
    F2B_CUDA(out32->alloc(n * 4));
    if (wd == FLUX2B_F32) {
      if (dtype == FLUX2B_F32) F2B_CUDA(cudaMemcpyAsync(out32->p, d, n * 4, cudaMemcpyDeviceToDevice, c->stream));
      else F2B_CUDA(cvt16_to_f32(d, out32->as<float>(), (int64_t)n, dtype == FLUX2B_F16, c->stream));
      return 0;
    }
    F2B_CUDA(t16.alloc(n * 2));
    const bool w_f16 = wd == FLUX2B_F16;
    if (dtype == FLUX2B_F32) F2B_CUDA(f32_to_16((const float*)d, (int64_t)n, t16.p, (int64_t)n, 1, (int)n, w_f16, c->stream));
    else F2B_CUDA(any16_to_16(d, dtype == FLUX2B_F16, t16.p, w_f16, (int64_t)n, c->stream));
    F2B_CUDA(cvt16_to_f32(t16.p, out32->as<float>(), (int64_t)n, w_f16, c->stream));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  };
  DevBuf A32, B32;
  F2B_TRY(stage_rounded(A, (size_t)rank * in_dim, &A32));
  F2B_TRY(stage_rounded(B, (size_t)out_dim * rank, &B32));
  if (quantized) {
    // find(), not operator[]: a missing scales / biases tensor is an error, not an empty tensor to dereference on the device
    PackedMeta m;
    F2B_TRY(packed_meta(c, base, &m));
    DevBuf dense;
    F2B_CUDA(dense.alloc((size_t)out_dim * in_dim * 2));
    F2B_CUDA(dequantize_matrix(c->quant, w.buf.as<uint32_t>(), m.s->buf.p, m.b ? m.b->buf.p : nullptr, out_dim, in_dim, dense.p, FLUX2B_F16, c->stream, m.sb_dtype));
    F2B_CUDA(lora_add(dense.p, FLUX2B_F16, A32.as<float>(), B32.as<float>(), out_dim, in_dim, rank, scale, c->stream));
    if (m.has_b && m.sb_dtype != FLUX2B_F16) {
      // the re-quantization of the f16 merged weight produces f16 scales / biases (quantized(...) of a .float16 array)
      const size_t groups = (size_t)m.s->numel();
      if (m.sb_dtype == FLUX2B_F32) {
        F2B_CUDA(cudaStreamSynchronize(c->stream));
        F2B_CUDA(m.s->buf.alloc(groups * 2));
        F2B_CUDA(m.b->buf.alloc(groups * 2));
      }
      m.s->dtype = FLUX2B_F16; m.b->dtype = FLUX2B_F16;
    }
    F2B_CUDA(quantize_matrix(c->quant, dense.p, FLUX2B_F16, out_dim, in_dim, w.buf.as<uint32_t>(), m.s->buf.p, m.b ? m.b->buf.p : nullptr, c->stream));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
  } else {
    F2B_CUDA(lora_add(w.buf.p, w.dtype, A32.as<float>(), B32.as<float>(), out_dim, in_dim, rank, scale, c->stream));
    F2B_CUDA(cudaStreamSynchronize(c->stream));
  }
  c->staging_used = 0;
  // The fused / re-tiled working copies are stale now. They are rebuilt ONCE, lazily, before the next forward (or by an explicit
  // flux2b_finalize_weights): the reference merges a LoRA as a loop over ~200 layers (WeightLoader.swift:736-856), and a
  // rebuild of the whole model per merged layer would make that loop quadratic.
  if (c->finalized && c->has_dit) c->dit_dirty = true;
  return 0;
}

// ------------------------------------------------------------------ denoise loop (Flux2Pipeline.swift:1933-2052,1696-1767)
int flux2b_denoise(flux2b_ctx* c, const flux2b_denoise_params* p, float* latents) {
  F2B_TRY(check(c));
  if (!p || !latents || !p->sigmas || p->num_sigmas < 2 || !p->enc) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad denoise parameters");
  if (!c->has_dit || !c->finalized) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "transformer not loaded");
  const flux2b_dit_config& g = c->dit;
  if (p->height < 16 || p->width < 16 || p->height % 16 || p->width % 16 || p->S_txt < 1)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "height / width must be positive multiples of 16, S_txt >= 1");
  const int h = p->height / 16, w = p->width / 16;
  const int S_img = h * w, S_ref = p->ref_latents ? p->S_ref : 0, S_all = S_img + S_ref;
  // the negative prompt has its own length and position ids (uncondTextIds, Flux2Pipeline.swift:1690,1960-1975); 0 = same as S_txt
  const int S_txt_u = (p->enc_uncond && p->S_txt_uncond > 0) ? p->S_txt_uncond : p->S_txt;
  const size_t n_lat = (size_t)S_img * g.in_channels;
  const int n_sig = p->num_sigmas;
  // device-resident state for the whole loop; persistent (context-owned) buffers: a steady-state call allocates nothing
  Buf x{c->scratch_buf("dn.x", n_lat * 4)}, hid{c->scratch_buf("dn.hid", (size_t)S_all * g.in_channels * 4)},
      pred{c->scratch_buf("dn.pred", (size_t)S_all * g.out_channels * 4)},
      pred_u{p->enc_uncond ? c->scratch_buf("dn.pred_u", (size_t)S_all * g.out_channels * 4) : nullptr},
      ids_img{c->scratch_buf("dn.ids_img", (size_t)S_all * 16)}, ids_txt{c->scratch_buf("dn.ids_txt", (size_t)p->S_txt * 16)},
      ids_txt_u{c->scratch_buf("dn.ids_txt_u", (size_t)S_txt_u * 16)},
      tbuf{c->scratch_buf("dn.sigmas", (size_t)(n_sig + 1) * 4)};
  if (!x.p || !hid.p || !pred.p || (p->enc_uncond && !pred_u.p) || !ids_img.p || !ids_txt.p || !ids_txt_u.p || !tbuf.p) {
    cudaGetLastError();
    return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "denoise state allocation failed");
  }
  const bool lat_host = !is_device_ptr(latents);
  F2B_CUDA(cudaMemcpyAsync(x.p, latents, n_lat * 4, lat_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
  // Position ids depend on (height, width, S_txt) only and the schedule + guidance scalar are a few floats: both live on the device
  // across calls and are uploaded only when they change, from context-owned host storage, so a call with device pointers
  // enqueues and returns without a single stream synchronisation (include/flux2b.h: "synchronise only when the destination is host
  // memory"; the reference forces completion once per step only because its hook / progress consumers need it, Flux2Pipeline.swift:1983).
  {
    DenoiseCache& dc = c->dn_cache;
    if (dc.height != p->height || dc.width != p->width || dc.S_txt != p->S_txt || dc.S_txt_u != S_txt_u ||
        dc.ids_img_dev != ids_img.p || dc.ids_txt_dev != ids_txt.p || dc.ids_txt_u_dev != ids_txt_u.p) {
      F2B_CUDA(cudaStreamSynchronize(c->stream));   // an earlier upload may still be reading the host vectors (shape change: rare)
      dc.ids_img.resize((size_t)S_img * 4); dc.ids_txt.resize((size_t)p->S_txt * 4); dc.ids_txt_u.resize((size_t)S_txt_u * 4);
      flux2b_image_position_ids(p->height, p->width, dc.ids_img.data());
      flux2b_text_position_ids(p->S_txt, dc.ids_txt.data());
      flux2b_text_position_ids(S_txt_u, dc.ids_txt_u.data());
      F2B_CUDA(cudaMemcpyAsync(ids_img.p, dc.ids_img.data(), dc.ids_img.size() * 4, cudaMemcpyHostToDevice, c->stream));
      F2B_CUDA(cudaMemcpyAsync(ids_txt.p, dc.ids_txt.data(), dc.ids_txt.size() * 4, cudaMemcpyHostToDevice, c->stream));
      F2B_CUDA(cudaMemcpyAsync(ids_txt_u.p, dc.ids_txt_u.data(), dc.ids_txt_u.size() * 4, cudaMemcpyHostToDevice, c->stream));
      dc.height = p->height; dc.width = p->width; dc.S_txt = p->S_txt; dc.S_txt_u = S_txt_u;
      dc.ids_img_dev = ids_img.p; dc.ids_txt_dev = ids_txt.p; dc.ids_txt_u_dev = ids_txt_u.p;
    }
    // [sigmas ..., guidance]: the guidance scalar is read on the host when it is host memory (one float, as the reference's
    // MLXArray([guidance]), Flux2Pipeline.swift:1916), so it never occupies a staging slot
    std::vector<float> sg(p->sigmas, p->sigmas + n_sig);
    const bool guid_dev = p->guidance && is_device_ptr(p->guidance);
    sg.push_back(p->guidance && !guid_dev ? *p->guidance : 0.f);
    if (dc.sigmas_dev != tbuf.p || dc.sigmas != sg) {
      F2B_CUDA(cudaStreamSynchronize(c->stream));
      dc.sigmas = sg;
      F2B_CUDA(cudaMemcpyAsync(tbuf.p, dc.sigmas.data(), dc.sigmas.size() * 4, cudaMemcpyHostToDevice, c->stream));
      dc.sigmas_dev = tbuf.p;
    }
  }
  if (S_ref) {
    if (!p->ref_ids) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "ref_ids required with ref_latents");
    const void *rid, *rlat;
    F2B_TRY(dev_in(c, p->ref_ids, (size_t)S_ref * 16, &rid));
    F2B_TRY(dev_in(c, p->ref_latents, (size_t)S_ref * g.in_channels * 4, &rlat));
    F2B_CUDA(cudaMemcpyAsync(ids_img.as<int32_t>() + (size_t)S_img * 4, rid, (size_t)S_ref * 16, cudaMemcpyDeviceToDevice, c->stream));
    // [output | refs] (Flux2Pipeline.swift:1703); reference tokens do not change across steps
    F2B_CUDA(cudaMemcpyAsync(hid.as<float>() + n_lat, rlat, (size_t)S_ref * g.in_channels * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  const void *enc_d, *encu_d;
  const size_t enc_bytes = (size_t)p->S_txt * g.joint_attention_dim * dtype_size(p->enc_dtype);
  const size_t encu_bytes = (size_t)S_txt_u * g.joint_attention_dim * dtype_size(p->enc_dtype);
  F2B_TRY(dev_in(c, p->enc, enc_bytes, &enc_d));
  F2B_TRY(dev_in(c, p->enc_uncond, encu_bytes, &encu_d));
  const float* guid_d = nullptr;
  if (p->guidance) guid_d = is_device_ptr(p->guidance) ? p->guidance : tbuf.as<float>() + n_sig;
  std::vector<float> host_lat;
  const int steps = n_sig - 1;
  const bool kv = p->kv_cache != 0 && S_ref > 0;
  if (kv && p->enc_uncond) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "the KV-cached loop has no classical-CFG branch (as the reference)");
  // Without a hook the whole loop is one fixed launch sequence over context-owned buffers: it is captured once per (shape,
  // schedule) as a CUDA graph and replayed (option dit_graph, default 1; not under sequence parallelism — NCCL calls and the
  // peer-mapping handshake stay outside graphs). The text embeddings are then read from a context-owned copy, so the graph does
  // not depend on the caller's pointers.
  const bool graphed = !p->hook && c->sp.world == 1 && c->option("dit_graph", 1) && !c->prof_on;
  if (graphed) {
    Buf enc_c{c->scratch_buf("dn.enc", enc_bytes)}, encu_c{p->enc_uncond ? c->scratch_buf("dn.enc_u", encu_bytes) : nullptr};
    if (!enc_c.p || (p->enc_uncond && !encu_c.p)) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "denoise state allocation failed"); }
    F2B_CUDA(cudaMemcpyAsync(enc_c.p, enc_d, enc_bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (p->enc_uncond) F2B_CUDA(cudaMemcpyAsync(encu_c.p, encu_d, encu_bytes, cudaMemcpyDeviceToDevice, c->stream));
    enc_d = enc_c.p; encu_d = encu_c.p;
    if (guid_d && guid_d != tbuf.as<float>() + n_sig) {   // a device-resident guidance scalar: into the schedule buffer's tail
      F2B_CUDA(cudaMemcpyAsync(tbuf.as<float>() + n_sig, guid_d, 4, cudaMemcpyDeviceToDevice, c->stream));
      guid_d = tbuf.as<float>() + n_sig;
    }
  }
  auto loop_body = [&]() -> int {
  for (int i = 0; i < steps; ++i) {
    const float sigma = p->sigmas[i], sigma_next = p->sigmas[i + 1];
    F2B_CUDA(cudaMemcpyAsync(hid.p, x.p, n_lat * 4, cudaMemcpyDeviceToDevice, c->stream));
    DitIO io{};
    io.B = 1; io.S_img = S_all; io.S_txt = p->S_txt; io.hidden = hid.as<float>(); io.enc = enc_d; io.enc_dtype = p->enc_dtype;
    io.timestep = tbuf.as<float>() + i; io.guidance = guid_d;
    io.img_ids = ids_img.as<int32_t>(); io.txt_ids = ids_txt.as<int32_t>(); io.out = pred.as<float>();
    if (kv) {
      // klein-9b-kv (Flux2Pipeline.swift:1565-1644): the reference tokens enter the transformer once, at step 0
      io.S_img = S_img;
      io.kv_mode = i == 0 ? 1 : 2;
      if (i == 0) {
        io.S_ref = S_ref;
        io.ref_hidden = hid.as<float>() + n_lat;
        io.ref_ids = ids_img.as<int32_t>() + (size_t)S_img * 4;
      }
    }
    F2B_TRY(dit_forward_device(c, io));
    if (p->enc_uncond) {
      io.enc = encu_d; io.S_txt = S_txt_u; io.txt_ids = ids_txt_u.as<int32_t>(); io.out = pred_u.as<float>();
      F2B_TRY(dit_forward_device(c, io));
    }
    {
      // only the first S_img predictions feed the Euler step (:1743)
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (p->enc_uncond ? 16.0 : 12.0) * n_lat);
      F2B_CUDA(euler_step(x.as<float>(), pred.as<float>(), p->enc_uncond ? pred_u.as<float>() : nullptr, p->cfg_scale,
                          sigma_next - sigma, (int64_t)n_lat, c->stream));
    }
    if (p->hook) {
      host_lat.resize(n_lat);
      F2B_CUDA(cudaMemcpyAsync(host_lat.data(), x.p, n_lat * 4, cudaMemcpyDeviceToHost, c->stream));
      F2B_CUDA(cudaStreamSynchronize(c->stream));
      flux2b_step_context sc{i, steps, sigma, sigma_next, p->height, p->width, S_ref > 0 ? 1 : 0};
      if (p->hook(&sc, host_lat.data(), n_lat, p->hook_user) != 0) {
        c->staging_used = 0;
        return fail(FLUX2B_ERR_CANCELLED, "generation cancelled by step hook");
      }
      F2B_CUDA(cudaMemcpyAsync(x.p, host_lat.data(), n_lat * 4, cudaMemcpyHostToDevice, c->stream));
      F2B_CUDA(cudaStreamSynchronize(c->stream));
    }
  }
  return 0;
  };
  if (graphed) {
    std::string key = "denoise";
    auto add = [&](const void* v, size_t n) { key.append(reinterpret_cast<const char*>(v), n); };
    const int scalars[] = {p->height, p->width, p->S_txt, S_txt_u, S_ref, (int)kv, p->enc_uncond ? 1 : 0, guid_d ? 1 : 0, p->enc_dtype, n_sig};
    const void* ptrs[] = {x.p, hid.p, pred.p, pred_u.p, ids_img.p, ids_txt.p, ids_txt_u.p, tbuf.p, enc_d, encu_d};
    add(scalars, sizeof(scalars)); add(ptrs, sizeof(ptrs)); add(&p->cfg_scale, 4); add(p->sigmas, (size_t)n_sig * 4);
    F2B_TRY(run_graphed(c, key, loop_body));
  } else {
    F2B_TRY(loop_body());
  }
  F2B_CUDA(cudaMemcpyAsync(latents, x.p, n_lat * 4, lat_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
  // host latents: the copy must have landed when the call returns; device pointers: stream-ordered, no synchronisation
  return end_call(c, lat_host);
}

int flux2b_generate(flux2b_ctx* c, const flux2b_denoise_params* p, float* latents, uint8_t* rgb) {
  F2B_TRY(check(c));
  if (!c->vw.ready) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "VAE not loaded");
  if (!c->vw.has_bn) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "latentBatchNorm running stats not set");
  if (!p || !latents || !rgb) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null argument");
  const int h = p->height / 16, w = p->width / 16;
  const size_t n_lat = (size_t)h * w * 128;
  // keep latents on the device between the loop and the decoder
  Buf xdev{c->scratch_buf("gen.x", n_lat * 4)}, z{c->scratch_buf("gen.z", (size_t)4 * h * w * 32 * 2)};
  if (!xdev.p || !z.p) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "generate state allocation failed"); }
  const bool lat_host = !is_device_ptr(latents);
  F2B_CUDA(cudaMemcpyAsync(xdev.p, latents, n_lat * 4, lat_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
  F2B_TRY(flux2b_denoise(c, p, xdev.as<float>()));   // device latents: synchronises only if p->enc / refs were staged from host memory
  F2B_CUDA(cudaMemcpyAsync(latents, xdev.p, n_lat * 4, lat_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
  // unpack -> BN denorm (eps 1e-4) -> unpatchify -> NHWC 16-bit, one fused gather (Flux2Pipeline.swift:2059-2079)
  const bool vf16 = c->option("vae_f16", 1) != 0;
  const int64_t npix = (int64_t)p->height * p->width;
  void* dout; bool ho;
  F2B_TRY(dev_out(c, rgb, (size_t)npix * 3, &dout, &ho));
  // latents -> uint8 image: ~200 launches over context-owned buffers, captured per resolution like the denoise loop
  auto decode_body = [&]() -> int {
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 6.0 * n_lat);
      F2B_CUDA(seq_to_vae_input(xdev.as<float>(), c->vw.bn_mean.as<float>(), c->vw.bn_var.as<float>(), 1e-4f, z.p, 1, h, w, vf16, c->stream));
    }
    void* img16; int ld;
    F2B_TRY(vae_decode_device(c, 1, 2 * h, 2 * w, z.p, &img16, &ld));
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 9.0 * npix);
      F2B_CUDA(postprocess_u8(img16, ld, (uint8_t*)dout, npix, vf16, c->stream));
    }
    return 0;
  };
  {
    std::string key = "decode";
    const int scalars[] = {p->height, p->width, (int)vf16};
    const void* ptrs[] = {xdev.p, z.p, dout};
    key.append(reinterpret_cast<const char*>(scalars), sizeof(scalars));
    key.append(reinterpret_cast<const char*>(ptrs), sizeof(ptrs));
    F2B_TRY(run_graphed(c, key, decode_body));
  }
  F2B_TRY(finish_out(c, rgb, dout, (size_t)npix * 3, ho));
  return end_call(c, lat_host || ho);   // all-device call: stream-ordered, returns without synchronising
}

// ------------------------------------------------------------------ VAE entry points
static int vae_decode_common(flux2b_ctx* c, int B, int h8, int w8, const float* lat, float* img_f32, uint8_t* img_u8) {
  F2B_TRY(check(c));
  if (!c->vw.ready) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "VAE weights not finalized");
  const int L = c->vae.latent_channels;
  const bool vf16 = c->option("vae_f16", 1) != 0;
  const void* dl;
  F2B_TRY(dev_in(c, lat, (size_t)B * L * h8 * w8 * 4, &dl));
  void* zp = c->scratch_buf("vae.z", (size_t)B * h8 * w8 * L * 2);
  if (!zp) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "VAE input allocation failed"); }
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 6.0 * B * L * h8 * w8);
    F2B_CUDA(nchw_f32_to_nhwc16((const float*)dl, zp, B, (int64_t)h8 * w8, L, vf16, c->stream));
  }
  void* img16; int ld;
  F2B_TRY(vae_decode_device(c, B, h8, w8, zp, &img16, &ld));
  const int64_t npix = (int64_t)B * 64 * h8 * w8;
  const int Co = c->vae.out_channels;
  if (img_f32) {
    void* dout; bool ho;
    F2B_TRY(dev_out(c, img_f32, (size_t)npix * Co * 4, &dout, &ho));
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 6.0 * npix * Co);
      F2B_CUDA(nhwc16_to_nchw_f32(img16, ld, (float*)dout, B, (int64_t)64 * h8 * w8, Co, vf16, c->stream));
    }
    F2B_TRY(finish_out(c, img_f32, dout, (size_t)npix * Co * 4, ho));
  } else {
    void* dout; bool ho;
    F2B_TRY(dev_out(c, img_u8, (size_t)npix * 3, &dout, &ho));
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 9.0 * npix);
      F2B_CUDA(postprocess_u8(img16, ld, (uint8_t*)dout, npix, vf16, c->stream));
    }
    F2B_TRY(finish_out(c, img_u8, dout, (size_t)npix * 3, ho));
  }
  return end_call(c, true);
}
int flux2b_vae_decode(flux2b_ctx* c, int B, int h8, int w8, const float* lat, float* img) {
  return vae_decode_common(c, B, h8, w8, lat, img, nullptr);
}
int flux2b_vae_decode_u8(flux2b_ctx* c, int B, int h8, int w8, const float* lat, uint8_t* rgb) {
  return vae_decode_common(c, B, h8, w8, lat, nullptr, rgb);
}

// ------------------------------------------------------------------ VAE encoder entry points
// image (host or device) -> latent NCHW fp32 in the scratch buffer "vae.enc.lat"
static int vae_encode_to_latent(flux2b_ctx* c, int B, int H, int W, const float* image, const float* noise, float** lat_out) {
  F2B_TRY(check(c));
  if (!c->vw.ready || !c->vw.enc.ready) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "VAE encoder weights not loaded (encoder.* tensors)");
  if (B < 1 || H % 8 || W % 8 || H < 8 || W < 8) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "VAE encode: batch >= 1, height / width multiples of 8");
  const int L = c->vae.latent_channels, Cin = c->vae.in_channels, Cpad = c->vw.enc.conv_in.cin;
  const bool vf16 = c->option("vae_f16", 1) != 0;
  const int h8 = H / 8, w8 = W / 8;
  const void *di, *dn;
  F2B_TRY(dev_in(c, image, (size_t)B * Cin * H * W * 4, &di));
  F2B_TRY(dev_in(c, noise, (size_t)B * L * h8 * w8 * 4, &dn));
  void* xp = c->scratch_buf("vae.enc.x", (size_t)B * H * W * Cpad * 2);
  float* lat = (float*)c->scratch_buf("vae.enc.lat", (size_t)B * L * h8 * w8 * 4);
  if (!xp || !lat) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "VAE encoder input allocation failed"); }
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)B * H * W * (4.0 * Cin + 2.0 * Cpad));
    F2B_CUDA(nchw_f32_to_nhwc16_pad((const float*)di, xp, B, (int64_t)H * W, Cin, Cpad, vf16, c->stream));
  }
  void* m16; int ld;
  F2B_TRY(vae_encode_device(c, B, H, W, xp, &m16, &ld));
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)B * h8 * w8 * L * 8.0);
    F2B_CUDA(moments_to_latent(m16, ld, (const float*)dn, lat, B, (int64_t)h8 * w8, L, vf16, c->stream));
  }
  *lat_out = lat;
  return 0;
}
int flux2b_vae_encode(flux2b_ctx* c, int B, int H, int W, const float* image, const float* noise, float* latents) {
  float* lat;
  F2B_TRY(vae_encode_to_latent(c, B, H, W, image, noise, &lat));
  const size_t n = (size_t)B * c->vae.latent_channels * (H / 8) * (W / 8);
  F2B_CUDA(cudaMemcpyAsync(latents, lat, n * 4, is_device_ptr(latents) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  return end_call(c, true);
}
int flux2b_encode_image_to_sequence(flux2b_ctx* c, int B, int H, int W, const float* image, const float* noise, float* seq) {
  if (H % 16 || W % 16) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "encode to sequence: height / width must be multiples of 16");
  float* lat;
  F2B_TRY(vae_encode_to_latent(c, B, H, W, image, noise, &lat));
  if (!c->vw.has_bn) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "latentBatchNorm.runningMean / runningVar not loaded");
  const int L = c->vae.latent_channels, C4 = 4 * L, h8 = H / 8, w8 = W / 8, pH = h8 / 2, pW = w8 / 2;
  const size_t n = (size_t)B * L * h8 * w8;
  float* t1 = (float*)c->scratch_buf("vae.enc.t1", n * 4);
  float* t2 = (float*)c->scratch_buf("vae.enc.t2", n * 4);
  if (!t1 || !t2) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "VAE encoder scratch"); }
  void* dout; bool ho;
  F2B_TRY(dev_out(c, seq, n * 4, &dout, &ho));
  ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 24.0 * n);
  {  // packLatentsToPatchified (LatentUtils.swift:186-212): [B, L, h8, w8] -> [B, 4L, pH, pW], channel = c * 4 + ph * 2 + pw
    const int p = 2;
    const int os[6] = {B, L, p, p, pH, pW};
    const int64_t is[6] = {(int64_t)L * h8 * w8, (int64_t)h8 * w8, w8, 1, (int64_t)p * w8, p};
    F2B_CUDA(permute_f32(lat, t1, 6, os, is, c->stream));
  }
  // normalizeLatentsWithBatchNorm (:460-473): (x - mean) / sqrt(var + 1e-4)
  F2B_CUDA(bn_affine_nchw(t1, t2, c->vw.bn_mean.as<float>(), c->vw.bn_var.as<float>(), 1e-4f, B, C4, (int64_t)pH * pW, false, c->stream));
  {  // packPatchifiedToSequence (:76-86): [B, 4L, pH, pW] -> [B, pH * pW, 4L]
    const int os[4] = {B, pH, pW, C4};
    const int64_t is[4] = {(int64_t)C4 * pH * pW, pW, 1, (int64_t)pH * pW};
    F2B_CUDA(permute_f32(t2, (float*)dout, 4, os, is, c->stream));
  }
  F2B_TRY(finish_out(c, seq, dout, n * 4, ho));
  return end_call(c, true);
}

}  // extern "C"
