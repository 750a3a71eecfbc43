// dit.cu — one Flux.2 DiT forward on the device (reference: Transformer/Flux2Transformer.swift:123-327).
//
// Data layout in HBM for one batch item with S = S_txt + S_img tokens (txt rows first — the reference's concat
// order, Flux2Attention.swift:161-163, Flux2Transformer.swift:252):
//   X    fp32  [S, D]      residual stream; double blocks update the txt / img row ranges in place, single blocks
//                          the whole matrix, so the [txt|img] concatenation of :252 is free
//   XN   16bit [S, D]      LayerNorm+modulate output (A operand of the next GEMM)
//   QKV  16bit [S, 3D]     q|k|v, already RMS-normed and rotated by the QKV GEMM epilogue
//   CAT  16bit [S, D+Hm]   attention output in columns [0,D), SwiGLU output in [D, D+Hm): the single-stream
//                          concat([attn, mlp]) of Flux2ParallelAttention.swift:119 is again free
// Kernel sequence per double block: 2x ln_modulate, 2x GEMM(QKV+norm+RoPE), attention, 2x GEMM(out, gate+residual),
// 2x ln_modulate, 2x GEMM(FF-in, SwiGLU), 2x GEMM(FF-out, gate+residual); per single block: ln_modulate,
// GEMM(QKV+norm+RoPE), GEMM(MLP-in, SwiGLU), attention, GEMM(out, gate+residual).
#include "ctx.h"

namespace f2b {

// quantised view of an activation operand (native block-scaled path): element bytes + scale factors (quant.cuh layout)
struct QView {
  const uint8_t* q = nullptr;   // [rows, ldq bytes]
  int64_t ldq = 0;
  const uint8_t* sf = nullptr;  // first 512 B block of the operand's K range
  int sf_ld = 0;                // blocks per 128-row block
};

// W-only quantized GEMMs with many rows dequantize the layer once into a context-owned 16-bit stage and run the plain kernel
// (gemm.cuh: wq_stage); few-row GEMMs (HBM-bound on the weights) dequantize inside the kernel. Option wq_inkernel: 1 = always
// inside the kernel, 2 (default) = staged above WQ_STAGE_MIN_ROWS rows.
static constexpr int WQ_STAGE_MIN_ROWS = 1024;
static int wq_stage_buffers(flux2b_ctx* c, int M, const Lin& W, bool two, void** s0, void** s1) {
  *s0 = *s1 = nullptr;
  if (!W.wmode || c->option("wq_inkernel", 2) != 2 || M <= WQ_STAGE_MIN_ROWS) return 0;
  if (!c->wq_stage_max) {
    auto upd = [&](const Lin& L) { if (L.wmode) c->wq_stage_max = std::max(c->wq_stage_max, (size_t)L.N * L.K * 2); };
    for (const DoubleBlockW& b : c->dbl) { upd(b.qkv_img); upd(b.qkv_txt); upd(b.out_img); upd(b.out_txt); upd(b.ff_in_img); upd(b.ff_out_img); upd(b.ff_in_txt); upd(b.ff_out_txt); }
    for (const SingleBlockW& b : c->sgl) { upd(b.qkv); upd(b.mlp); upd(b.out); }
    upd(c->x_embed); upd(c->ctx_embed); upd(c->proj_out);
  }
  const size_t need = std::max(c->wq_stage_max, (size_t)W.N * W.K * 2);
  F2B_CUDA(c->wq_stage[0].ensure(need));
  *s0 = c->wq_stage[0].p;
  if (two) { F2B_CUDA(c->wq_stage[1].ensure(need)); *s1 = c->wq_stage[1].p; }
  return 0;
}

// C = A · W^T (+ epilogue). k_off / K_override select a K-slice of W (the single-stream out projection split under SP).
// Weights in block-scaled form (W.mx) need the activation's QView; the 16-bit A pointer is then unused.
static int run_gemm(flux2b_ctx* c, const void* A, int64_t lda, const Lin& W, int M, Epilogue epi, const QView* qa = nullptr,
                    int64_t k_off = 0, int K_override = 0) {
  GemmProblem g;
  g.M = M; g.N = W.N; g.K = K_override ? K_override : W.K;
  epi.f16 = c->f16() ? 1 : 0;
  g.epi = epi;
  double bytes;
  if (W.wmode) {
    // W-only: x · dequant(W)^T with the weight dequantized inside the kernel (packed codes are all that is read from HBM)
    const int bits = (W.wmode == 1 || W.wmode == 3) ? 8 : 4, group = W.wmode <= 2 ? 64 : W.wmode == 5 ? 16 : 32;
    const int esz = W.wmode <= 2 ? 2 : 1;
    g.A = A; g.lda = lda;
    g.wq = W.wmode; g.wq_sb_bf16 = W.w_sb_bf16; g.wq_sb_ld = W.K / group;
    g.B = W.wq.as<uint8_t>() + k_off * bits / 8; g.ldb = (int64_t)W.K * bits / 8;
    g.wq_scales = W.ws.as<uint8_t>() + (k_off / group) * esz;
    g.wq_biases = W.wb.p ? W.wb.as<uint8_t>() + (k_off / group) * esz : nullptr;
    g.force_cta_group = c->option("gemm_cta_group", 0);
    F2B_TRY(wq_stage_buffers(c, M, W, false, &g.wq_stage, &g.wq_stage_lo));
    g.wq_stage_kb = c->option("wq_stage_kb", 0);
    bytes = 2.0 * ((double)M * g.K + (double)M * g.N) + (double)g.N * g.K * bits / 8 + (double)g.N * (g.K / group) * esz * (W.wb.p ? 2 : 1);
    if (g.wq_stage) bytes += 2.0 * 2.0 * (double)g.N * g.K;   // the 16-bit stage is written once and read once
  } else if (W.mx) {
    if (!qa || !qa->q) return fail(FLUX2B_ERR_GENERATION_FAILED, "internal: block-scaled weight without a quantised activation");
    const int bits = W.mx == 1 ? 8 : 4, group = W.mx == 3 ? 16 : 32;
    g.mx = W.mx; g.force_bn = W.bn;
    g.force_cta_group = c->option("gemm_cta_group", 0);
    g.A = qa->q; g.lda = qa->ldq; g.sfa = qa->sf; g.sfa_ld = qa->sf_ld;
    g.B = W.wq.as<uint8_t>() + k_off * bits / 8; g.ldb = (int64_t)W.K * bits / 8;
    g.sfb = W.sfb.as<uint8_t>() + (k_off / group / 4) * 512; g.sfb_ld = W.K / group / 4;
    bytes = ((double)M + g.N) * g.K * bits / 8 + 2.0 * M * g.N;
  } else {
    g.A = A; g.lda = lda;
    g.B = W.w.as<uint16_t>() + k_off; g.ldb = W.K;
    g.force_cta_group = c->option("gemm_cta_group", 0);
    bytes = 2.0 * ((double)M * g.K + (double)g.N * g.K + (double)M * g.N);
  }
  const double flops = 2.0 * M * (double)g.N * g.K;
  ProfScope ps(c, FLUX2B_PROF_GEMM, flops, bytes, gemm_launch_count(g));
  F2B_CUDA(gemm_launch(g, c->stream));
  return 0;
}

static int ensure_ws(flux2b_ctx* c, int S, int S_img, int S_txt) {
  const int D = c->D, Hm = c->Hm;
  F2B_CUDA(c->ws_x.ensure((size_t)S * D * 4));
  F2B_CUDA(c->ws_xn.ensure((size_t)S * D * 2));
  F2B_CUDA(c->ws_qkv.ensure((size_t)S * 3 * D * 2));
  F2B_CUDA(c->ws_cat.ensure((size_t)S * (D + 3 * Hm) * 2));  // CAT [S, D+Hm] + room for the unfused [gate|value] fallback
  F2B_CUDA(c->ws_cos.ensure((size_t)S * 128 * 4));
  F2B_CUDA(c->ws_sin.ensure((size_t)S * 128 * 4));
  F2B_CUDA(c->ws_ids.ensure((size_t)S * 4 * 4));
  F2B_CUDA(c->ws_small.ensure((size_t)(32 * D + 1024) * 4));
  F2B_CUDA(c->ws_hid16.ensure((size_t)S_img * c->dit.in_channels * 2));
  F2B_CUDA(c->ws_enc16.ensure((size_t)S_txt * c->dit.joint_attention_dim * 2));
  if (c->mx_kind) {
    const int bits = c->mx_kind == 1 ? 8 : 4;
    F2B_CUDA(c->ws_aq_xn.ensure((size_t)S * D * bits / 8));
    F2B_CUDA(c->ws_aq_cat.ensure((size_t)S * (D + Hm) * bits / 8));
    for (int i = 0; i < 4; ++i) F2B_CUDA(c->ws_sfa[i].ensure(mx_sf_bytes(c->mx_kind, S, D + Hm)));
  }
  return 0;
}

static int attention(flux2b_ctx* c, int S_q, int q_row0, const KVSegment* segs, int nseg, void* out, int64_t ldo,
                     int o_row0, int64_t rows_total) {
  AttnProblem a;
  a.q = c->ws_qkv.p; a.ldq = 3 * c->D; a.q_rows_total = rows_total; a.q_row0 = q_row0; a.sq = S_q;
  a.o = out; a.ldo = ldo; a.o_row0 = o_row0;
  a.num_heads = c->H; a.batch = 1;
  a.scale = 1.0f / sqrtf(128.0f);
  a.num_segments = nseg;
  double keys = 0;
  for (int i = 0; i < nseg; ++i) { a.seg[i] = segs[i]; keys += segs[i].len; }
  a.f16 = c->f16() ? 1 : 0;
  a.variant = c->option("attn_variant", 0);
  a.poly = c->option("attn_poly", 0);
  ProfScope ps(c, FLUX2B_PROF_ATTN, 4.0 * S_q * keys * c->D, 2.0 * (2.0 * S_q * c->D + 2.0 * keys * c->D));
  F2B_CUDA(attention_launch(a, c->stream));
  return 0;
}

static int record_block(flux2b_ctx* c, int idx, int S) {
  if (!c->option("record_blocks", 0)) return 0;
  const int total = c->dit.num_layers + c->dit.num_single_layers;
  F2B_CUDA(c->ws_rec.ensure((size_t)total * S * c->D * 4));
  c->rec_S = S; c->rec_count = total;
  F2B_CUDA(cudaMemcpyAsync(c->ws_rec.as<float>() + (size_t)idx * S * c->D, c->ws_x.p, (size_t)S * c->D * 4,
                           cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

// Ulysses attention for the local token shard (see sp.cu), in three phases so that a single-stream block can run its
// MLP GEMMs beside the two exchanges:
//   sp_exchange_qkv : QKV of the local rows ([dest rank][token][q|k|v][Hp*128], written by the GEMM epilogue) -> gather
//   sp_attend       : attention over all S tokens for this rank's heads
//   sp_exchange_o   : O back to the token owners, heads of rank j into columns [j*Dp, (j+1)*Dp) of `o`
// `async` = run the NCCL exchange on the side stream (caller joins before consuming). Peer-memory mode (sp.mode 1) has
// no exchange kernels at all: the producers already stored remotely, only flag barriers remain.
static int sp_exchange_qkv(flux2b_ctx* c, int Sl, bool async) {
  const int P = c->sp.world, Dp = (c->H / P) * 128;
  if (c->sp.mode == 1) return sp_barrier(c);  // every rank's QKV epilogue stores have landed in my gather buffer
  if (async) F2B_TRY(sp_fork(c));
  return sp_all_to_all(c, c->ws_qkv.p, c->ws_sp_gather.p, (size_t)Sl * 3 * Dp, async ? c->sp.comm_stream : c->stream);
}
static int sp_attend(flux2b_ctx* c, int Sl, void* o, int64_t ldo) {
  const int P = c->sp.world, Hp = c->H / P, Dp = Hp * 128;
  const int S = Sl * P;
  uint16_t* G = c->ws_sp_gather.as<uint16_t>();
  AttnProblem a;
  a.q = G; a.ldq = 3 * Dp; a.q_rows_total = S; a.q_row0 = 0; a.sq = S;
  a.o = c->ws_sp_o.p; a.ldo = Dp; a.o_row0 = 0;
  a.num_heads = Hp; a.batch = 1;
  a.scale = 1.0f / sqrtf(128.0f);
  a.num_segments = 1;
  a.seg[0].k = G + Dp; a.seg[0].v = G + 2 * Dp; a.seg[0].ldk = a.seg[0].ldv = 3 * Dp;
  a.seg[0].rows_total = S; a.seg[0].row0 = 0; a.seg[0].len = S;
  a.f16 = c->f16() ? 1 : 0;
  a.variant = c->option("attn_variant", 0);
  a.poly = c->option("attn_poly", 0);
  if (c->sp.mode == 1) {
    // the attention epilogue stores each query row's heads straight into the owning rank's attention-output buffer
    a.variant = 3;
    a.ldo = ldo; a.o_rows_per_peer = Sl; a.o_col0 = c->sp.rank * Dp;
    const size_t off = reinterpret_cast<uint8_t*>(o) - reinterpret_cast<uint8_t*>(c->ws_cat.p);
    for (int d = 0; d < P; ++d) a.o_peer[d] = reinterpret_cast<uint8_t*>(c->sp.cat_peer[d]) + off;
  }
  ProfScope ps(c, FLUX2B_PROF_ATTN, 4.0 * S * (double)S * Dp, 2.0 * (4.0 * S * Dp));
  F2B_CUDA(attention_launch(a, c->stream));
  return 0;
}
static int sp_exchange_o(flux2b_ctx* c, int Sl, void* o, int64_t ldo, bool async) {
  const int P = c->sp.world, Dp = (c->H / P) * 128;
  if (c->sp.mode == 1) return sp_barrier(c);  // all heads of my rows have arrived; peers are done with their gather buffers
  const size_t o_chunk = (size_t)Sl * Dp;
  if (async) F2B_TRY(sp_fork(c));
  cudaStream_t st = async ? c->sp.comm_stream : c->stream;
  F2B_TRY(sp_all_to_all(c, c->ws_sp_o.p, c->ws_sp_orecv.p, o_chunk, st));
  for (int j = 0; j < P; ++j)
    F2B_CUDA(cudaMemcpy2DAsync(reinterpret_cast<uint16_t*>(o) + (size_t)j * Dp, (size_t)ldo * 2,
                               c->ws_sp_orecv.as<uint16_t>() + (size_t)j * o_chunk, (size_t)Dp * 2, (size_t)Dp * 2, Sl,
                               cudaMemcpyDeviceToDevice, st));
  return 0;
}
static int sp_attention(flux2b_ctx* c, int Sl, void* o, int64_t ldo) {
  F2B_TRY(sp_exchange_qkv(c, Sl, false));
  F2B_TRY(sp_attend(c, Sl, o, ldo));
  return sp_exchange_o(c, Sl, o, ldo, false);
}

// One batch item. hidden [S_img, in_ch] f32, enc [S_txt, joint] (dtype), t/g scalars on device, ids on device.
// Under sequence parallelism every rank receives the FULL inputs and returns the FULL output; it computes the rows of
// its own token shard and the [S_img, out_ch] prediction is all-gathered at the end.
static int forward_one(flux2b_ctx* c, int S_img, int S_txt, const float* hidden, const void* enc, int enc_dtype,
                       const float* t, const float* g, const int32_t* img_ids, const int32_t* txt_ids, float* out,
                       int kv_mode, int S_ref, const float* ref_hidden, const int32_t* ref_ids) {
  const flux2b_dit_config& cfg = c->dit;
  const int D = c->D, Hm = c->Hm;
  const bool f16 = c->f16();
  cudaStream_t st = c->stream;
  // option sp_disable: a sequence-parallel context runs this forward on its own GPU alone (the single-GPU reference the
  // parity of the sharded forward is measured against, same weights, same process)
  const int P = c->option("sp_disable", 0) ? 1 : c->sp.world;
  float* out_full = out;
  const int S_img_full = S_img;
  if (P > 1) {
    if (kv_mode) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "the KV-cached forward is not sequence-parallel");
    if (!c->option("fuse_qk_rope", 1)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "sequence parallelism needs fuse_qk_rope = 1");
    flux2b_sp_layout_t L;
    F2B_TRY(flux2b_sp_layout(P, c->sp.rank, S_txt, S_img, c->H, &L));
    S_txt = L.txt_rows; S_img = L.img_rows;
    hidden += (size_t)L.img_row0 * cfg.in_channels;
    enc = reinterpret_cast<const uint8_t*>(enc) + (size_t)L.txt_row0 * cfg.joint_attention_dim * dtype_size(enc_dtype);
    img_ids += (size_t)L.img_row0 * 4;
    txt_ids += (size_t)L.txt_row0 * 4;
    out += (size_t)L.img_row0 * cfg.out_channels;
  }
  // token layout of this pass: [txt | (ref) | img]
  const int S_mid = (kv_mode == 1) ? S_ref : 0;
  const int S_im_all = S_mid + S_img;   // rows of the "image" stream
  const int S = S_txt + S_im_all;
  F2B_TRY(ensure_ws(c, S, S_im_all, S_txt));
  if (P > 1) {
    F2B_CUDA(c->ws_sp_gather.ensure((size_t)S * 3 * D * 2));
    F2B_CUDA(c->ws_sp_o.ensure((size_t)S * D * 2));
    F2B_CUDA(c->ws_sp_orecv.ensure((size_t)S * D * 2));
    if (c->sp.mode == 1) F2B_TRY(sp_map_peers(c));
  }
  float* X = c->ws_x.as<float>();
  float* Ximg = X + (size_t)S_txt * D;
  uint16_t* XN = c->ws_xn.as<uint16_t>();
  uint16_t* QKV = c->ws_qkv.as<uint16_t>();
  uint16_t* CAT = c->ws_cat.as<uint16_t>();
  float* cosT = c->ws_cos.as<float>();
  float* sinT = c->ws_sin.as<float>();
  float* sm = c->ws_small.as<float>();
  float* sinus = sm;                 // [256] (+256 guidance)
  float* h1 = sm + 512;              // [D]
  float* temb = h1 + D;              // [D]
  float* h2 = temb + D;              // [D] guidance hidden
  float* mod_img = h2 + D;           // [6D]
  float* mod_txt = mod_img + 6 * D;  // [6D]
  float* mod_sgl = mod_txt + 6 * D;  // [3D]
  float* mod_out = mod_sgl + 3 * D;  // [2D]

  // ---- timestep (+guidance) embedding: Flux2Transformer.swift:145-149, Flux2Embeddings.swift:124-141
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 0);
    F2B_CUDA(timestep_sinusoid(t, sinus, 1, 1000.0f, st));
  }
  auto gemv_p = [&](const float* x, const Lin& W, float* y, bool silu_in, bool accumulate) -> int {
    if (W.wmode) {
      const int bits = (W.wmode == 1 || W.wmode == 3) ? 8 : 4, group = W.wmode <= 2 ? 64 : W.wmode == 5 ? 16 : 32;
      ProfScope ps(c, FLUX2B_PROF_GEMV, 2.0 * W.N * W.K, (double)W.N * W.K * bits / 8 + (double)W.N * (W.K / group) * (W.wmode <= 2 ? 4 : 1));
      F2B_CUDA(gemv_q(x, W.K, W.wq.p, (int64_t)W.K * bits / 8, W.ws.p, W.wb.p, W.K / group, W.wmode, W.w_sb_bf16, y, W.N, 1, W.N, W.K,
                      silu_in, accumulate, f16, st));
      return 0;
    }
    ProfScope ps(c, FLUX2B_PROF_GEMV, 2.0 * W.N * W.K, 2.0 * W.N * W.K);
    F2B_CUDA(gemv(x, W.K, W.w.p, W.K, y, W.N, 1, W.N, W.K, silu_in, accumulate, f16, st));
    return 0;
  };
  F2B_TRY(gemv_p(sinus, c->t_lin1, h1, false, false));
  F2B_TRY(gemv_p(h1, c->t_lin2, temb, true, false));
  if (cfg.guidance_embeds && g) {
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, 0);
      F2B_CUDA(timestep_sinusoid(g, sinus + 256, 1, 1000.0f, st));
    }
    F2B_TRY(gemv_p(sinus + 256, c->g_lin1, h2, false, false));
    F2B_TRY(gemv_p(h2, c->g_lin2, temb, true, true));  // temb += guidance embedding (Flux2Embeddings.swift:134-138)
  }
  // ---- modulation, once per forward (Flux2Transformer.swift:160-161,256; Flux2Modulation.swift:49-75)
  F2B_TRY(gemv_p(temb, c->mod_img, mod_img, true, false));
  F2B_TRY(gemv_p(temb, c->mod_txt, mod_txt, true, false));
  F2B_TRY(gemv_p(temb, c->mod_single, mod_sgl, true, false));
  F2B_TRY(gemv_p(temb, c->norm_out, mod_out, true, false));

  // ---- RoPE table over [txt | (ref) | img] ids (Flux2Transformer.swift:153-154)
  {
    int32_t* ids = c->ws_ids.as<int32_t>();
    F2B_CUDA(cudaMemcpyAsync(ids, txt_ids, (size_t)S_txt * 16, cudaMemcpyDeviceToDevice, st));
    if (S_mid) F2B_CUDA(cudaMemcpyAsync(ids + (size_t)S_txt * 4, ref_ids, (size_t)S_mid * 16, cudaMemcpyDeviceToDevice, st));
    F2B_CUDA(cudaMemcpyAsync(ids + (size_t)(S_txt + S_mid) * 4, img_ids, (size_t)S_img * 16, cudaMemcpyDeviceToDevice, st));
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * 128 * 8);
    F2B_CUDA(rope_table(ids, S, cfg.axes_dims_rope, cfg.rope_theta, cosT, sinT, st));
  }

  // ---- input embedders (Flux2Transformer.swift:137-138)
  {
    uint16_t* hid16 = c->ws_hid16.as<uint16_t>();
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S_im_all * cfg.in_channels * 6);
      if (S_mid) F2B_CUDA(f32_to_16(ref_hidden, cfg.in_channels, hid16, cfg.in_channels, S_mid, cfg.in_channels, f16, st));
      F2B_CUDA(f32_to_16(hidden, cfg.in_channels, hid16 + (size_t)S_mid * cfg.in_channels, cfg.in_channels, S_img,
                         cfg.in_channels, f16, st));
    }
    Epilogue e; e.mode = EPI_F32; e.out = Ximg; e.ldo = D;
    F2B_TRY(run_gemm(c, hid16, cfg.in_channels, c->x_embed, S_im_all, e));
    uint16_t* enc16 = c->ws_enc16.as<uint16_t>();
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S_txt * cfg.joint_attention_dim * 6);
      const int64_t n = (int64_t)S_txt * cfg.joint_attention_dim;
      if (enc_dtype == FLUX2B_F32) F2B_CUDA(f32_to_16((const float*)enc, cfg.joint_attention_dim, enc16, cfg.joint_attention_dim, S_txt, cfg.joint_attention_dim, f16, st));
      else F2B_CUDA(any16_to_16(enc, enc_dtype == FLUX2B_F16, enc16, f16, n, st));
    }
    Epilogue e2; e2.mode = EPI_F32; e2.out = X; e2.ldo = D;
    F2B_TRY(run_gemm(c, enc16, cfg.joint_attention_dim, c->ctx_embed, S_txt, e2));
  }

  const bool fuse_qk = c->option("fuse_qk_rope", 1) != 0;
  // ---- native block-scaled path: every block linear consumes its activation quantised on the fly to the weight's format.
  // slot 0 / 1 = text rows / image rows (or the whole sequence) of XN, 2 / 3 likewise of CAT; `src` rows are quantised into
  // the same row range of the slot's byte buffer, scale factors into the slot's own tile space (row 0 = first row of `src`).
  const int mxk = c->mx_kind;
  const int mx_bits = mxk == 1 ? 8 : 4, mx_grp = mxk == 3 ? 16 : 32;
  // where the quantised version of rows [row0, ...) x columns [col0, col0 + K) of a [*, Ktot] activation lives
  auto qdest = [&](int slot, int64_t row0, int64_t Ktot, int64_t col0, QView* qv, MxOut* mo) {
    uint8_t* base = (slot < 2 ? c->ws_aq_xn : c->ws_aq_cat).as<uint8_t>();
    const int64_t ldq = Ktot * mx_bits / 8;
    uint8_t* q = base + row0 * ldq + col0 * mx_bits / 8;
    uint8_t* sf = c->ws_sfa[slot].as<uint8_t>();
    const int sf_ld = (int)(Ktot / mx_grp / 4);
    qv->q = q; qv->ldq = ldq; qv->sf = sf + (col0 / mx_grp / 4) * 512; qv->sf_ld = sf_ld;
    mo->kind = mxk; mo->q = q; mo->ldq = ldq; mo->sf = sf; mo->sf_ld = sf_ld; mo->g0 = (int)(col0 / mx_grp);
  };
  auto quantize = [&](const uint16_t* src, int64_t ld, int rows, int K, int slot, int64_t row0, int64_t Ktot, int64_t col0,
                      QView* qv) -> int {
    if (!mxk) return 0;
    MxOut mo;
    qdest(slot, row0, Ktot, col0, qv, &mo);
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)rows * K * (2.0 + mx_bits / 8.0));
    F2B_CUDA(mx_quantize_act(mxk, src, ld, rows, K, f16, mo.q, mo.ldq, mo.sf, mo.sf_ld, col0, st));
    return 0;
  };
  // option mx_fuse_quant (default 1): LayerNorm + modulate and the SwiGLU epilogue emit the block-scaled operand themselves
  // (no 16-bit XN / MLP activation is ever written); 0 = separate quantisation pass (same bits, kept as the cross-check)
  const bool fuse_q = mxk && c->option("mx_fuse_quant", 1) != 0;
  // LN + modulate of `rows` rows into XN rows starting at `o` (slot 0 = text range, 1 = image range / everything)
  auto ln_mod = [&](const float* x, int rows, const float* shift, const float* scale, uint16_t* o, int slot, QView* qv) -> int {
    if (fuse_q) {
      MxOut mo;
      qdest(slot, (o - XN) / D, D, 0, qv, &mo);
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)rows * D * (4.0 + mx_bits / 8.0));
      F2B_CUDA(ln_modulate(x, D, o, D, rows, D, shift, scale, 0, rows, 1e-6f, f16, st, &mo));
      return 0;
    }
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)rows * D * 6);
      F2B_CUDA(ln_modulate(x, D, o, D, rows, D, shift, scale, 0, rows, 1e-6f, f16, st));
    }
    return quantize(o, D, rows, D, slot, (o - XN) / D, D, 0, qv);
  };
  auto qkv_gemm = [&](const uint16_t* a, const QView* qa, const Lin& W, int rows, int row0, const DevBuf& nq, const DevBuf& nk) -> int {
    Epilogue e;
    e.out = QKV + (size_t)row0 * 3 * D; e.ldo = 3 * D;
    if (fuse_qk) {
      e.mode = EPI_QKV_ROPE; e.cos = cosT + (size_t)row0 * 128; e.sin = sinT + (size_t)row0 * 128;
      e.norm_q = nq.as<float>(); e.norm_k = nk.as<float>(); e.dmodel = D; e.eps = 1e-6f;
      if (P > 1) {
        // all-to-all layout [dest][local token][q|k|v][Dp]; S is the local row count here
        const int Hp = c->H / P, Dp = Hp * 128;
        e.sp_hp = Hp; e.ldo = 3 * Dp;
        if (c->sp.mode == 1) {
          // fused projection + all-to-all: store into rank d's gather buffer, slab of THIS rank's tokens
          for (int d = 0; d < P; ++d)
            e.sp_base[d] = reinterpret_cast<uint16_t*>(c->sp.gather_peer[d]) + ((size_t)c->sp.rank * S + row0) * 3 * Dp;
        } else {
          for (int d = 0; d < P; ++d) e.sp_base[d] = QKV + ((size_t)d * S + row0) * 3 * Dp;
        }
      }
      F2B_TRY(run_gemm(c, a, D, W, rows, e, qa));
    } else {
      e.mode = EPI_BF16;
      F2B_TRY(run_gemm(c, a, D, W, rows, e, qa));
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)rows * D * 8);
      F2B_CUDA(qk_norm_rope(QKV + (size_t)row0 * 3 * D, 3 * D, rows, D, nq.as<float>(), nk.as<float>(),
                            cosT + (size_t)row0 * 128, sinT + (size_t)row0 * 128, 1e-6f, f16, st));
    }
    return 0;
  };
  auto gate_res_gemm = [&](const uint16_t* a, int64_t lda, const QView* qa, const Lin& W, int rows, float* x, const float* gate) -> int {
    Epilogue e; e.mode = EPI_GATE_RES; e.out = x; e.ldo = D; e.res = x; e.ldr = D; e.gate = gate;
    return run_gemm(c, a, lda, W, rows, e, qa);
  };
  // SwiGLU producer: out[rows, Hm] (leading dim ldo) = silu(gate) * value
  // With `mo` (native path, fused quantisation) the 16-bit result is not stored; the epilogue emits the block-scaled operand.
  auto swiglu_gemm = [&](const uint16_t* a, const QView* qa, const Lin& W, bool tiled, int rows, uint16_t* o, int64_t ldo, uint16_t* scratch,
                         const MxOut* mo = nullptr) -> int {
    if (tiled) {
      Epilogue e; e.mode = EPI_SWIGLU; e.out = o; e.ldo = (int)ldo;
      if (mo) e.mxo = *mo;
      return run_gemm(c, a, D, W, rows, e, qa);
    }
    Epilogue e; e.mode = EPI_BF16; e.out = scratch; e.ldo = 2 * Hm;
    F2B_TRY(run_gemm(c, a, D, W, rows, e, qa));
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)rows * Hm * 6);
    F2B_CUDA(swiglu(scratch, 2 * Hm, o, ldo, rows, Hm, f16, st));
    return 0;
  };

  // attention key segments for this pass
  auto full_attention = [&](int layer, void* o, int64_t ldo) -> int {
    if (P > 1) return sp_attention(c, S, o, ldo);
    const int64_t rows_total = S;
    uint16_t* Kp = QKV + D;
    uint16_t* Vp = QKV + 2 * D;
    if (kv_mode == 0) {
      KVSegment sg; sg.k = Kp; sg.v = Vp; sg.ldk = sg.ldv = 3 * D; sg.rows_total = rows_total; sg.row0 = 0; sg.len = S;
      return attention(c, S, 0, &sg, 1, o, ldo, 0, rows_total);
    }
    if (kv_mode == 1) {
      // extraction pass [txt | ref | img] (Flux2Attention.swift:245-330): the additive mask of :422-437 blocks
      // reference queries from the output keys; text and output queries see everything. The mask is block-aligned, so
      // it is expressed as query ranges x key segments instead of an S x S float matrix.
      // (a) txt + img queries over all keys
      KVSegment all; all.k = Kp; all.v = Vp; all.ldk = all.ldv = 3 * D; all.rows_total = rows_total; all.row0 = 0; all.len = S;
      F2B_TRY(attention(c, S_txt, 0, &all, 1, o, ldo, 0, rows_total));
      F2B_TRY(attention(c, S_img, S_txt + S_ref, &all, 1, o, ldo, S_txt + S_ref, rows_total));
      // (b) ref queries over [txt | ref] keys
      KVSegment rf = all; rf.row0 = 0; rf.len = S_txt + S_ref;
      F2B_TRY(attention(c, S_ref, S_txt, &rf, 1, o, ldo, S_txt, rows_total));
      // cache post-norm, post-RoPE reference K / V of this layer (TransformerKVCache.swift:13-79)
      F2B_CUDA(c->kv_k[layer].ensure((size_t)S_ref * D * 2));
      F2B_CUDA(c->kv_v[layer].ensure((size_t)S_ref * D * 2));
      F2B_CUDA(cudaMemcpy2DAsync(c->kv_k[layer].p, (size_t)D * 2, Kp + (size_t)S_txt * 3 * D, (size_t)3 * D * 2,
                                 (size_t)D * 2, S_ref, cudaMemcpyDeviceToDevice, st));
      F2B_CUDA(cudaMemcpy2DAsync(c->kv_v[layer].p, (size_t)D * 2, Vp + (size_t)S_txt * 3 * D, (size_t)3 * D * 2,
                                 (size_t)D * 2, S_ref, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
    // cached pass: queries [txt | img], keys [txt | cachedRef | img] (Flux2Attention.swift:393-395)
    KVSegment sg[3];
    sg[0].k = Kp; sg[0].v = Vp; sg[0].ldk = sg[0].ldv = 3 * D; sg[0].rows_total = rows_total; sg[0].row0 = 0; sg[0].len = S_txt;
    sg[1].k = c->kv_k[layer].p; sg[1].v = c->kv_v[layer].p; sg[1].ldk = sg[1].ldv = D; sg[1].rows_total = c->kv_S_ref;
    sg[1].row0 = 0; sg[1].len = c->kv_S_ref;
    sg[2] = sg[0]; sg[2].row0 = S_txt; sg[2].len = S_img;
    return attention(c, S, 0, sg, 3, o, ldo, 0, rows_total);
  };

  // ---- double-stream blocks (Flux2TransformerBlock.swift:80-168)
  QView q_img, q_txt, q_all, q_mlp;
  // option group_streams (default 1): one launch per operation for both streams (16-bit operands, fused epilogues, single GPU)
  bool group = P == 1 && !mxk && fuse_qk && S_txt > 0 && S_txt % 256 == 0 && S_im_all > 0 && c->option("group_streams", 1) != 0;
  for (int i = 0; group && i < cfg.num_layers; ++i) {
    const DoubleBlockW& b = c->dbl[i];
    group = b.ff_tiled && !b.qkv_img.mx && b.qkv_img.wmode == b.qkv_txt.wmode && b.out_img.wmode == b.out_txt.wmode &&
            b.ff_in_img.wmode == b.ff_in_txt.wmode && b.ff_out_img.wmode == b.ff_out_txt.wmode &&
            b.qkv_img.w_sb_bf16 == b.qkv_txt.w_sb_bf16 && b.out_img.w_sb_bf16 == b.out_txt.w_sb_bf16 &&
            b.ff_in_img.w_sb_bf16 == b.ff_in_txt.w_sb_bf16 && b.ff_out_img.w_sb_bf16 == b.ff_out_txt.w_sb_bf16;
  }
  for (int i = 0; i < cfg.num_layers; ++i) {
    DoubleBlockW& b = c->dbl[i];
    uint16_t* XNi = XN + (size_t)S_txt * D;
    if (group) {
      // Both streams in one launch per operation (7 launches instead of 13): text rows [0, S_txt) and image rows behind them
      // share X / XN / QKV / CAT and differ only in weights, modulation and QK-norm weights. A text-stream GEMM of its own
      // (M = 512) fills a third of the SMs for ~35 us; as two extra M units of the image GEMM's tile schedule it costs ~1/9 more.
      auto ln2 = [&](int set) -> int {
        ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * D * 6);
        F2B_CUDA(ln_modulate(X, D, XN, D, S, D, mod_img + (3 * set) * D, mod_img + (3 * set + 1) * D, 0, S, 1e-6f, f16, st, nullptr,
                             S_txt, mod_txt + (3 * set) * D, mod_txt + (3 * set + 1) * D));
        return 0;
      };
      auto gemm2 = [&](const void* A, int64_t lda, const Lin& Wimg, const Lin& Wtxt, Epilogue e) -> int {
        GemmProblem g;
        g.M = S; g.N = Wimg.N; g.K = Wimg.K;
        g.A = A; g.lda = lda; g.B = Wimg.w.p; g.ldb = Wimg.K; g.B_lo = Wtxt.w.p; g.M_lo = S_txt;
        double wbytes = 2.0 * 2.0 * (double)g.N * g.K;
        if (Wimg.wmode) {
          const int bits = (Wimg.wmode == 1 || Wimg.wmode == 3) ? 8 : 4, group = Wimg.wmode <= 2 ? 64 : Wimg.wmode == 5 ? 16 : 32;
          g.wq = Wimg.wmode; g.wq_sb_bf16 = Wimg.w_sb_bf16; g.wq_sb_ld = Wimg.K / group;
          g.B = Wimg.wq.p; g.B_lo = Wtxt.wq.p; g.ldb = (int64_t)Wimg.K * bits / 8;
          g.wq_scales = Wimg.ws.p; g.wq_biases = Wimg.wb.p; g.wq_scales_lo = Wtxt.ws.p; g.wq_biases_lo = Wtxt.wb.p;
          wbytes = 2.0 * ((double)g.N * g.K * bits / 8 + (double)g.N * (g.K / group) * (Wimg.wmode <= 2 ? 4 : 1));
          F2B_TRY(wq_stage_buffers(c, S, Wimg, true, &g.wq_stage, &g.wq_stage_lo));
          g.wq_stage_kb = c->option("wq_stage_kb", 0);
          if (g.wq_stage) wbytes += 2.0 * 2.0 * 2.0 * (double)g.N * g.K;
        }
        e.f16 = f16 ? 1 : 0; e.split_row = S_txt;
        g.epi = e;
        g.force_cta_group = c->option("gemm_cta_group", 0);
        ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * S * (double)g.N * g.K, 2.0 * ((double)S * g.K + (double)S * g.N) + wbytes, gemm_launch_count(g));
        F2B_CUDA(gemm_launch(g, st));
        return 0;
      };
      F2B_TRY(ln2(0));
      {
        Epilogue e; e.mode = EPI_QKV_ROPE; e.out = QKV; e.ldo = 3 * D; e.cos = cosT; e.sin = sinT; e.dmodel = D; e.eps = 1e-6f;
        e.norm_q = b.nq_img.as<float>(); e.norm_k = b.nk_img.as<float>();
        e.norm_q_lo = b.nq_txt.as<float>(); e.norm_k_lo = b.nk_txt.as<float>();
        F2B_TRY(gemm2(XN, D, b.qkv_img, b.qkv_txt, e));
      }
      F2B_TRY(full_attention(i, CAT, D));
      {
        Epilogue e; e.mode = EPI_GATE_RES; e.out = X; e.ldo = D; e.res = X; e.ldr = D;
        e.gate = mod_img + 2 * D; e.gate_lo = mod_txt + 2 * D;
        F2B_TRY(gemm2(CAT, D, b.out_img, b.out_txt, e));
      }
      F2B_TRY(ln2(1));
      {
        Epilogue e; e.mode = EPI_SWIGLU; e.out = CAT; e.ldo = Hm;
        F2B_TRY(gemm2(XN, D, b.ff_in_img, b.ff_in_txt, e));
      }
      {
        Epilogue e; e.mode = EPI_GATE_RES; e.out = X; e.ldo = D; e.res = X; e.ldr = D;
        e.gate = mod_img + 5 * D; e.gate_lo = mod_txt + 5 * D;
        F2B_TRY(gemm2(CAT, Hm, b.ff_out_img, b.ff_out_txt, e));
      }
      F2B_TRY(record_block(c, i, S));
      continue;
    }
    F2B_TRY(ln_mod(Ximg, S_im_all, mod_img + 0, mod_img + D, XNi, 1, &q_img));
    F2B_TRY(ln_mod(X, S_txt, mod_txt + 0, mod_txt + D, XN, 0, &q_txt));
    F2B_TRY(qkv_gemm(XNi, &q_img, b.qkv_img, S_im_all, S_txt, b.nq_img, b.nk_img));
    F2B_TRY(qkv_gemm(XN, &q_txt, b.qkv_txt, S_txt, 0, b.nq_txt, b.nk_txt));
    F2B_TRY(full_attention(i, CAT, D));
    F2B_TRY(quantize(CAT + (size_t)S_txt * D, D, S_im_all, D, 3, S_txt, D, 0, &q_img));
    F2B_TRY(quantize(CAT, D, S_txt, D, 2, 0, D, 0, &q_txt));
    F2B_TRY(gate_res_gemm(CAT + (size_t)S_txt * D, D, &q_img, b.out_img, S_im_all, Ximg, mod_img + 2 * D));
    F2B_TRY(gate_res_gemm(CAT, D, &q_txt, b.out_txt, S_txt, X, mod_txt + 2 * D));
    F2B_TRY(ln_mod(Ximg, S_im_all, mod_img + 3 * D, mod_img + 4 * D, XNi, 1, &q_img));
    F2B_TRY(ln_mod(X, S_txt, mod_txt + 3 * D, mod_txt + 4 * D, XN, 0, &q_txt));
    uint16_t* Hbuf = CAT;                       // [S, Hm]
    uint16_t* scratch = CAT + (size_t)S * Hm;   // [S, 2Hm] unfused fallback
    if (fuse_q && b.ff_tiled) {
      MxOut mo; QView q_h;
      qdest(3, S_txt, Hm, 0, &q_h, &mo);
      F2B_TRY(swiglu_gemm(XNi, &q_img, b.ff_in_img, true, S_im_all, nullptr, Hm, nullptr, &mo));
      F2B_TRY(gate_res_gemm(nullptr, Hm, &q_h, b.ff_out_img, S_im_all, Ximg, mod_img + 5 * D));
      qdest(2, 0, Hm, 0, &q_h, &mo);
      F2B_TRY(swiglu_gemm(XN, &q_txt, b.ff_in_txt, true, S_txt, nullptr, Hm, nullptr, &mo));
      F2B_TRY(gate_res_gemm(nullptr, Hm, &q_h, b.ff_out_txt, S_txt, X, mod_txt + 5 * D));
    } else {
      F2B_TRY(swiglu_gemm(XNi, &q_img, b.ff_in_img, b.ff_tiled, S_im_all, Hbuf + (size_t)S_txt * Hm, Hm, scratch));
      F2B_TRY(quantize(Hbuf + (size_t)S_txt * Hm, Hm, S_im_all, Hm, 3, S_txt, Hm, 0, &q_img));
      F2B_TRY(gate_res_gemm(Hbuf + (size_t)S_txt * Hm, Hm, &q_img, b.ff_out_img, S_im_all, Ximg, mod_img + 5 * D));
      F2B_TRY(swiglu_gemm(XN, &q_txt, b.ff_in_txt, b.ff_tiled, S_txt, Hbuf, Hm, scratch));
      F2B_TRY(quantize(Hbuf, Hm, S_txt, Hm, 2, 0, Hm, 0, &q_txt));
      F2B_TRY(gate_res_gemm(Hbuf, Hm, &q_txt, b.ff_out_txt, S_txt, X, mod_txt + 5 * D));
    }
    F2B_TRY(record_block(c, i, S));
  }
  // ---- single-stream blocks (Flux2SingleBlock.swift:59-98, Flux2ParallelAttention.swift:72-123)
  const int ldc = D + Hm;
  for (int i = 0; i < cfg.num_single_layers; ++i) {
    SingleBlockW& b = c->sgl[i];
    F2B_TRY(ln_mod(X, S, mod_sgl + 0, mod_sgl + D, XN, 1, &q_all));
    F2B_TRY(qkv_gemm(XN, &q_all, b.qkv, S, 0, b.nq, b.nk));
    uint16_t* scratch = CAT + (size_t)S * ldc;  // [S, 2Hm] unfused fallback lives behind CAT (see ensure_ws)
    if (P > 1 && c->sp.mode == 0 && c->option("sp_overlap", 1)) {
      // NCCL transport: hide both exchanges behind the MLP GEMMs of the block.
      //   QKV all-to-all  ||  MLP-in GEMM (SwiGLU)
      //   attention
      //   O all-to-all    ||  out GEMM over the MLP columns (K = Hm, no dependency on the attention)
      //   out GEMM over the attention columns (K = D)
      F2B_TRY(sp_exchange_qkv(c, S, true));
      if (fuse_q && b.mlp_tiled) {
        MxOut mo;
        qdest(3, 0, ldc, D, &q_mlp, &mo);
        F2B_TRY(swiglu_gemm(XN, &q_all, b.mlp, true, S, nullptr, ldc, nullptr, &mo));
      } else {
        F2B_TRY(swiglu_gemm(XN, &q_all, b.mlp, b.mlp_tiled, S, CAT + D, ldc, b.mlp_tiled ? nullptr : scratch));
        F2B_TRY(quantize(CAT + D, ldc, S, Hm, 3, 0, ldc, D, &q_mlp));
      }
      F2B_TRY(sp_join(c));
      F2B_TRY(sp_attend(c, S, CAT, ldc));
      F2B_TRY(sp_exchange_o(c, S, CAT, ldc, true));
      {
        Epilogue e; e.mode = EPI_GATE_RES; e.out = X; e.ldo = D; e.res = X; e.ldr = D; e.gate = mod_sgl + 2 * D;
        F2B_TRY(run_gemm(c, CAT + D, ldc, b.out, S, e, &q_mlp, D, Hm));
        F2B_TRY(sp_join(c));
        F2B_TRY(quantize(CAT, ldc, S, D, 3, 0, ldc, 0, &q_all));
        F2B_TRY(run_gemm(c, CAT, ldc, b.out, S, e, &q_all, 0, D));
      }
      F2B_TRY(record_block(c, cfg.num_layers + i, S));
      continue;
    }
    if (fuse_q && b.mlp_tiled) {
      // MLP columns [D, D + Hm) of the out-projection operand come quantised from the SwiGLU epilogue; only the attention
      // columns [0, D) need the separate pass
      MxOut mo;
      qdest(3, 0, ldc, D, &q_mlp, &mo);
      F2B_TRY(swiglu_gemm(XN, &q_all, b.mlp, true, S, nullptr, ldc, nullptr, &mo));
      F2B_TRY(full_attention(cfg.num_layers + i, CAT, ldc));
      F2B_TRY(quantize(CAT, ldc, S, D, 3, 0, ldc, 0, &q_all));
    } else {
      F2B_TRY(swiglu_gemm(XN, &q_all, b.mlp, b.mlp_tiled, S, CAT + D, ldc, b.mlp_tiled ? nullptr : scratch));
      F2B_TRY(full_attention(cfg.num_layers + i, CAT, ldc));
      F2B_TRY(quantize(CAT, ldc, S, ldc, 3, 0, ldc, 0, &q_all));
    }
    F2B_TRY(gate_res_gemm(CAT, ldc, &q_all, b.out, S, X, mod_sgl + 2 * D));
    F2B_TRY(record_block(c, cfg.num_layers + i, S));
  }
  // ---- output: AdaLayerNormContinuous (scale first, Flux2Modulation.swift:146-148) + projOut (:321-324)
  float* Xout = X + (size_t)(S_txt + S_mid) * D;
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S_img * D * 6);
    F2B_CUDA(ln_modulate(Xout, D, XN, D, S_img, D, mod_out + D, mod_out + 0, 0, S_img, 1e-6f, f16, st));
  }
  Epilogue e; e.mode = EPI_F32; e.out = out; e.ldo = cfg.out_channels;
  F2B_TRY(run_gemm(c, XN, D, c->proj_out, S_img, e));
  if (P > 1) F2B_TRY(sp_all_gather_f32(c, out_full, (size_t)S_img * cfg.out_channels));
  (void)S_img_full;
  return 0;
}

int dit_forward_device(flux2b_ctx* c, const DitIO& io) {
  if (!c->has_dit || !c->finalized) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "transformer weights not finalized");
  if (c->dit_dirty) {   // LoRA merges since the last forward: one rebuild of the working copies for all of them
    F2B_TRY(finalize_dit(c));
    c->dit_dirty = false;
  }
  const flux2b_dit_config& cfg = c->dit;
  if (io.B < 1 || io.S_img < 1 || io.S_txt < 1) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "empty batch / sequence");
  if (io.kv_mode) {
    const int layers = cfg.num_layers + cfg.num_single_layers;
    if ((int)c->kv_k.size() != layers) { c->kv_k.clear(); c->kv_v.clear(); c->kv_k.resize(layers); c->kv_v.resize(layers); }
    if (io.kv_mode == 1) c->kv_S_ref = io.S_ref;
    if (io.kv_mode == 2 && c->kv_S_ref <= 0) return fail(FLUX2B_ERR_GENERATION_FAILED, "kv cache is empty: run kv_extract first");
    if (io.B != 1) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "kv-cached forward supports batch 1 (as the reference pipeline)");
  }
  for (int b = 0; b < io.B; ++b) {
    const size_t enc_elems = (size_t)io.S_txt * cfg.joint_attention_dim;
    const void* enc_b = reinterpret_cast<const uint8_t*>(io.enc) + (size_t)b * enc_elems * dtype_size(io.enc_dtype);
    F2B_TRY(forward_one(c, io.S_img, io.S_txt, io.hidden + (size_t)b * io.S_img * cfg.in_channels, enc_b, io.enc_dtype,
                        io.timestep + b, io.guidance ? io.guidance + b : nullptr, io.img_ids, io.txt_ids,
                        io.out + (size_t)b * io.S_img * cfg.out_channels, io.kv_mode, io.S_ref, io.ref_hidden, io.ref_ids));
  }
  return 0;
}

}  // namespace f2b
