// vae.cu — Flux.2 VAE decoder on the device (reference: VAE/AutoencoderKL.swift:129-143, VAE/VAEDecoder.swift:91-121,
// VAE/ResnetBlock.swift:24-54,168-186,229-253,281-313). Activations stay NHWC 16-bit (f16 by default) end to end;
// every convolution is the implicit-GEMM mode of the tcgen05 kernel in gemm.cu with bias / shortcut add in its
// epilogue; GroupNorm(+SiLU) is a two-pass HBM-bound kernel with fp32 statistics.
#include "ctx.h"

namespace f2b {

struct VaeRun {
  flux2b_ctx* c;
  bool f16;
  int B;
  std::vector<DevBuf>& bufs;  // pool of activation buffers (5, at most 4 live at a time)
  // a buffer that is none of the live ones
  void* pick(const void* a = nullptr, const void* b = nullptr, const void* c2 = nullptr, const void* d = nullptr) {
    for (auto& bf : bufs)
      if (bf.p != a && bf.p != b && bf.p != c2 && bf.p != d) return bf.p;
    return nullptr;
  }
};

// H, W: OUTPUT extent; stride 2 = the encoder's downsample (input 2H x 2W, zero pad bottom / right only)
static int conv(flux2b_ctx* c, bool f16, const ConvW& w, const void* x, void* y, const void* res, int B, int H, int W, int stride = 1) {
  GemmProblem g;
  g.A = x; g.lda = w.cin; g.B = w.w.p; g.ldb = (int64_t)w.taps * w.cin;
  g.M = B * H * W; g.N = w.cout; g.K = w.taps * w.cin;
  g.conv_taps = w.taps; g.batch = B; g.H = H; g.W = W; g.Cin = w.cin; g.conv_stride = stride;
  g.epi.mode = EPI_BF16; g.epi.f16 = f16; g.epi.out = y; g.epi.ldo = w.cout; g.epi.bias = w.bias.as<float>();
  g.epi.res16 = res; g.epi.ldr = w.cout;
  g.force_cta_group = c->option("vae_conv_cta_group", 0);
  const double npix = (double)B * H * W;
  ProfScope ps(c, FLUX2B_PROF_CONV, 2.0 * npix * w.cout * g.K, 2.0 * (npix * w.cin + npix * w.cout + (double)w.cout * g.K));
  F2B_CUDA(gemm_launch(g, c->stream));
  return 0;
}
static int gn(flux2b_ctx* c, bool f16, const NormW& n, const void* x, void* y, int B, int64_t HW, bool silu) {
  const int G = c->vae.norm_num_groups;
  F2B_CUDA(ensure_zeroed(c->gn_stats, groupnorm_ws_bytes(B, G), c->stream));
  ProfScope ps(c, FLUX2B_PROF_GROUPNORM, 0, (double)B * HW * n.C * 6, 2);   // statistics (+ finalize in its last CTA), apply
  F2B_CUDA(groupnorm_silu(x, y, n.gamma.as<float>(), n.beta.as<float>(), c->gn_stats.as<double>(), B, HW, n.C, G,
                          c->vae.norm_eps, silu, f16, c->stream));
  return 0;
}
// x -> resnet(x); returns the buffer holding the result
static int resnet(VaeRun& r, const ResnetW& w, void*& x, int H, int W) {
  flux2b_ctx* c = r.c;
  void* t1 = r.pick(x);
  F2B_TRY(gn(c, r.f16, w.n1, x, t1, r.B, (int64_t)H * W, true));
  void* t2 = r.pick(x, t1);
  F2B_TRY(conv(c, r.f16, w.c1, t1, t2, nullptr, r.B, H, W));
  F2B_TRY(gn(c, r.f16, w.n2, t2, t1, r.B, (int64_t)H * W, true));
  const void* shortcut = x;
  if (w.has_sc) {
    F2B_TRY(conv(c, r.f16, w.sc, x, t2, nullptr, r.B, H, W));  // t2 is free again after norm2
    shortcut = t2;
  }
  void* y = r.pick(x, t1, t2);
  F2B_TRY(conv(c, r.f16, w.c2, t1, y, shortcut, r.B, H, W));
  x = y;
  return 0;
}

struct VaeAttnRef {
  const NormW& attn_norm; const Lin& attn_qkv; const DevBuf& attn_qkv_bias; const Lin& attn_out; const DevBuf& attn_out_bias;
};
static int mid_attention(VaeRun& r, const VaeAttnRef& v, void*& x, int H, int W) {
  flux2b_ctx* c = r.c;
  const int C = v.attn_norm.C;
  const int N = H * W;
  const bool f16 = r.f16;
  void* hn = r.pick(x);
  F2B_TRY(gn(c, f16, v.attn_norm, x, hn, r.B, N, false));
  void* y = r.pick(x, hn);
  // scratch sized for one batch item
  int chunk = std::max(128, std::min(N, (int)(((size_t)1 << 28) / (size_t)N) / 128 * 128));  // <= 1 GiB of fp32 scores
  if (c->option("vae_attn_chunk", 0) > 0) chunk = std::min(N, std::max(128, c->option("vae_attn_chunk", 0) / 128 * 128));  // tests: chunk < N at small N
  const int ldn = (N + 7) & ~7;  // TMA needs 16 B row strides; columns [N, ldn) are never read (tensor-map extent is N)
  // context-owned scratch: allocated once per resolution, so a steady-state decode performs no cudaMalloc / cudaFree
  Buf qkv{c->scratch_buf("vae.attn.qkv", (size_t)N * 3 * C * 2)}, vt{c->scratch_buf("vae.attn.vt", (size_t)C * ldn * 2)},
      scores{c->scratch_buf("vae.attn.scores", (size_t)chunk * ldn * 4)}, probs{c->scratch_buf("vae.attn.probs", (size_t)chunk * ldn * 2)},
      o{c->scratch_buf("vae.attn.o", (size_t)N * C * 2)};
  if (!qkv.p || !vt.p || !scores.p || !probs.p || !o.p) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "VAE attention scratch"); }
  const float scale = 1.0f / sqrtf((float)C);
  for (int b = 0; b < r.B; ++b) {
    const uint16_t* hb = reinterpret_cast<const uint16_t*>(hn) + (size_t)b * N * C;
    GemmProblem g;
    g.A = hb; g.lda = C; g.B = v.attn_qkv.w.p; g.ldb = C; g.M = N; g.N = 3 * C; g.K = C;
    g.epi.mode = EPI_BF16; g.epi.f16 = f16; g.epi.out = qkv.p; g.epi.ldo = 3 * C; g.epi.bias = v.attn_qkv_bias.as<float>();
    { ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * N * 3.0 * C * C, 2.0 * (N * 4.0 * C + 3.0 * C * C)); F2B_CUDA(gemm_launch(g, c->stream)); }
    { ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 4.0 * N * C);
      F2B_CUDA(transpose16(qkv.as<uint16_t>() + 2 * C, 3 * C, vt.p, ldn, N, C, c->stream)); }
    for (int q0 = 0; q0 < N; q0 += chunk) {
      const int rows = std::min(chunk, N - q0);
      GemmProblem s;
      s.A = qkv.as<uint16_t>() + (size_t)q0 * 3 * C; s.lda = 3 * C; s.B = qkv.as<uint16_t>() + C; s.ldb = 3 * C;
      s.M = rows; s.N = N; s.K = C;
      s.epi.mode = EPI_F32; s.epi.f16 = f16; s.epi.out = scores.p; s.epi.ldo = ldn;
      { ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * rows * (double)N * C, 4.0 * rows * (double)N); F2B_CUDA(gemm_launch(s, c->stream)); }
      { ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 6.0 * rows * (double)N);
        F2B_CUDA(softmax_rows(scores.as<float>(), ldn, probs.p, ldn, rows, N, scale, f16, c->stream)); }
      GemmProblem pv;
      pv.A = probs.p; pv.lda = ldn; pv.B = vt.p; pv.ldb = ldn; pv.M = rows; pv.N = C; pv.K = N;
      pv.epi.mode = EPI_BF16; pv.epi.f16 = f16; pv.epi.out = o.as<uint16_t>() + (size_t)q0 * C; pv.epi.ldo = C;
      { ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * rows * (double)N * C, 2.0 * rows * (double)N); F2B_CUDA(gemm_launch(pv, c->stream)); }
    }
    GemmProblem og;
    og.A = o.p; og.lda = C; og.B = v.attn_out.w.p; og.ldb = C; og.M = N; og.N = C; og.K = C;
    og.epi.mode = EPI_BF16; og.epi.f16 = f16; og.epi.out = reinterpret_cast<uint16_t*>(y) + (size_t)b * N * C; og.epi.ldo = C;
    og.epi.bias = v.attn_out_bias.as<float>();
    og.epi.res16 = reinterpret_cast<const uint16_t*>(x) + (size_t)b * N * C; og.epi.ldr = C;
    { ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * N * (double)C * C, 6.0 * N * C); F2B_CUDA(gemm_launch(og, c->stream)); }
  }
  x = y;
  return 0;
}

int vae_decode_device(flux2b_ctx* c, int B, int h8, int w8, const void* z, void** out, int* out_ld) {
  const VaeW& v = c->vw;
  if (!v.ready) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "VAE weights not finalized");
  const flux2b_vae_config& g = c->vae;
  const bool f16 = c->option("vae_f16", 1) != 0;
  // largest activation: full resolution x max(decoder_channels[0] (and [1] at half res x4 after upsample))
  size_t max_elems = 0;
  {
    int H = h8, W = w8;
    max_elems = (size_t)H * W * g.decoder_channels[3];
    for (int i = 0; i < 4; ++i) {
      const int co = g.decoder_channels[3 - i];
      const int ci = i == 0 ? g.decoder_channels[3] : g.decoder_channels[4 - i];
      max_elems = std::max(max_elems, (size_t)H * W * std::max(co, ci));
      if (i < 3) { H *= 2; W *= 2; max_elems = std::max(max_elems, (size_t)H * W * co); }
    }
    max_elems *= B;
  }
  if (c->vae_ws.size() != 5) { c->vae_ws.clear(); c->vae_ws.resize(5); }
  for (auto& b : c->vae_ws) F2B_CUDA(b.ensure(max_elems * 2));
  VaeRun r{c, f16, B, c->vae_ws};
  int H = h8, W = w8;
  void* x = r.pick();
  F2B_TRY(conv(c, f16, v.post_quant, z, x, nullptr, B, H, W));   // AutoencoderKL.swift:135-139
  void* t = r.pick(x);
  F2B_TRY(conv(c, f16, v.conv_in, x, t, nullptr, B, H, W));      // VAEDecoder.swift:97
  x = t;
  F2B_TRY(resnet(r, v.mid1, x, H, W));
  F2B_TRY(mid_attention(r, VaeAttnRef{v.attn_norm, v.attn_qkv, v.attn_qkv_bias, v.attn_out, v.attn_out_bias}, x, H, W));
  F2B_TRY(resnet(r, v.mid2, x, H, W));
  for (int i = 0; i < 4; ++i) {
    for (const ResnetW& rw : v.up[i]) F2B_TRY(resnet(r, rw, x, H, W));
    if (v.has_upconv[i] && v.upconv[i].w_up.p && c->option("vae_fold_upsample", 1)) {
      // nearest-2x upsample folded into its convolution: four 2x2 phase kernels over the low-resolution tensor, 4/9 of the
      // FLOPs and no 4x intermediate (option vae_fold_upsample = 0: upsample kernel + 3x3 convolution, the cross-check)
      const ConvW& w = v.upconv[i];
      void* y = r.pick(x);
      GemmProblem g;
      g.A = x; g.lda = w.cin; g.B = w.w_up.p; g.ldb = (int64_t)16 * w.cin;
      g.M = B * H * W; g.N = w.cout; g.K = 9 * w.cin;
      g.conv_taps = 9; g.conv_up2 = 1; g.batch = B; g.H = H; g.W = W; g.Cin = w.cin;
      g.epi.mode = EPI_BF16; g.epi.f16 = f16; g.epi.out = y; g.epi.ldo = w.cout; g.epi.bias = w.bias.as<float>();
      g.force_cta_group = c->option("vae_conv_cta_group", 0);
      const double npix = (double)B * H * W;
      {
        ProfScope ps(c, FLUX2B_PROF_CONV, 2.0 * npix * 16.0 * w.cout * w.cin, 2.0 * (npix * w.cin + 4.0 * npix * w.cout + 16.0 * w.cout * w.cin));
        F2B_CUDA(gemm_launch(g, c->stream));
      }
      H *= 2; W *= 2;
      x = y;
    } else if (v.has_upconv[i]) {
      void* up = r.pick(x);
      {
        ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 10.0 * B * H * W * v.upconv[i].cin);
        F2B_CUDA(upsample_nearest2x(x, up, B, H, W, v.upconv[i].cin, c->stream));  // ResnetBlock.swift:242-250
      }
      H *= 2; W *= 2;
      void* y = r.pick(up);
      F2B_TRY(conv(c, f16, v.upconv[i], up, y, nullptr, B, H, W));
      x = y;
    }
  }
  void* n = r.pick(x);
  F2B_TRY(gn(c, f16, v.norm_out, x, n, B, (int64_t)H * W, true));
  void* y = r.pick(n);
  F2B_TRY(conv(c, f16, v.conv_out, n, y, nullptr, B, H, W));
  *out = y;
  *out_ld = g.out_channels;
  return 0;
}

// VAEEncoder.callAsFunction (VAE/VAEEncoder.swift:85-115) + quantConv (VAE/AutoencoderKL.swift:94-99).
// image NHWC 16-bit [B, H, W, 8] (channels 3..7 zero) -> moments NHWC 16-bit [B, H/8, W/8, 2 * latent_ch]
int vae_encode_device(flux2b_ctx* c, int B, int H, int W, const void* image, void** out, int* out_ld) {
  const VaeEncW& e = c->vw.enc;
  if (!c->vw.ready || !e.ready) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "VAE encoder weights not loaded (encoder.* tensors)");
  if (H % 8 || W % 8 || H < 8 || W < 8) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "VAE encode: height / width must be multiples of 8");
  const bool f16 = c->option("vae_f16", 1) != 0;
  // largest activation: full resolution x the widest of the first block's channels
  size_t max_elems = 0;
  {
    int h = H, w = W, prev = e.conv_in.cout;
    for (int i = 0; i < 4; ++i) {
      const int co = e.down[i].empty() ? prev : e.down[i][0].cout;
      max_elems = std::max(max_elems, (size_t)h * w * std::max(prev, co));
      prev = co;
      if (e.has_down[i]) { h /= 2; w /= 2; }
    }
    max_elems *= B;
  }
  if (c->vae_ws.size() != 5) { c->vae_ws.clear(); c->vae_ws.resize(5); }
  for (auto& b : c->vae_ws) F2B_CUDA(b.ensure(max_elems * 2));
  VaeRun r{c, f16, B, c->vae_ws};
  int h = H, w = W;
  void* x = r.pick();
  F2B_TRY(conv(c, f16, e.conv_in, image, x, nullptr, B, h, w));
  for (int i = 0; i < 4; ++i) {
    for (const ResnetW& rw : e.down[i]) F2B_TRY(resnet(r, rw, x, h, w));
    if (e.has_down[i]) {
      h /= 2; w /= 2;
      void* y = r.pick(x);
      F2B_TRY(conv(c, f16, e.downconv[i], x, y, nullptr, B, h, w, 2));
      x = y;
    }
  }
  F2B_TRY(resnet(r, e.mid1, x, h, w));
  F2B_TRY(mid_attention(r, VaeAttnRef{e.attn_norm, e.attn_qkv, e.attn_qkv_bias, e.attn_out, e.attn_out_bias}, x, h, w));
  F2B_TRY(resnet(r, e.mid2, x, h, w));
  void* n = r.pick(x);
  F2B_TRY(gn(c, f16, e.norm_out, x, n, B, (int64_t)h * w, true));
  void* y = r.pick(n);
  F2B_TRY(conv(c, f16, e.conv_out, n, y, nullptr, B, h, w));
  if (e.has_quant) {
    void* q = r.pick(y);
    F2B_TRY(conv(c, f16, e.quant, y, q, nullptr, B, h, w));
    y = q;
  }
  *out = y;
  *out_ld = e.conv_out.cout;
  return 0;
}

}  // namespace f2b
