// te.cu — the text-embedding producer on the other side of the DiT boundary: one causal prefill of the
// text encoder and extraction of the hidden states the Flux.2 pipelines concatenate into `encoderHiddenStates`.
//
// Reference: FluxTextEncoders/Embeddings/KleinEmbeddingExtractor.swift:46-133 (Qwen3, right padding to 512, layers 9/18/27),
// Embeddings/EmbeddingExtractor.swift:202-285 (Mistral Small 3.2, left padding, layers 10/20/30),
// Model/Qwen3/Qwen3Model.swift:104-231 (forwardWithHiddenStates, createCausalMask), Qwen3DecoderLayer.swift:32-48,
// Qwen3Attention.swift:92-163 (q/k RMSNorm before rotate-half RoPE, grouped-query attention), Qwen3MLP.swift:42-47,
// Model/MistralAttention.swift:393-474 (same without the QK-norm; the Llama-4 query scaling is exactly 1 below
// original_max_position_embeddings and therefore checked, not computed).
// Tokenisation and chat templating stay on the host side of the boundary (KleinEmbeddingExtractor.swift:56-90).
//
// Data layout (one prompt, S tokens, hidden size Hd, Hq query / Hkv key-value heads of 128):
//   X    fp32  [S, Hd]                      residual stream
//   XN   16bit [S, Hd]                      RMSNorm output = A operand of the next GEMM
//   QKV  16bit [S, (Hq + 2 Hkv) * 128]      q | k | v; q / k leave the GEMM epilogue normed and rotated
//   ATT  16bit [S, Hq * 128]                attention output
//   ACT  16bit [S, I]                       silu(gate) * up from the SwiGLU epilogue of the fused gate|up GEMM
// Kernel sequence per layer (7 launches): rms_norm, GEMM(QKV + q/k-norm + RoPE), causal GQA flash attention,
// GEMM(o_proj, + residual), rms_norm, GEMM(gate|up, SwiGLU), GEMM(down_proj, + residual).
#include "ctx.h"

namespace f2b {

static int te_gemm(flux2b_ctx* c, const void* A, int64_t lda, const Lin& W, int M, Epilogue epi) {
  GemmProblem g;
  g.M = M; g.N = W.N; g.K = W.K;
  epi.f16 = c->f16() ? 1 : 0;
  g.epi = epi;
  g.A = A; g.lda = lda;
  g.B = W.w.p; g.ldb = W.K;
  g.force_cta_group = c->option("gemm_cta_group", 0);
  const double bytes = 2.0 * ((double)M * g.K + (double)g.N * g.K + (double)M * g.N);
  ProfScope ps(c, FLUX2B_PROF_GEMM, 2.0 * M * (double)g.N * g.K, bytes);
  F2B_CUDA(gemm_launch(g, c->stream));
  return 0;
}

int te_forward_device(flux2b_ctx* c, int S, const int32_t* ids, int key_lo, int key_hi, const int* layers, int n_layers,
                      float* out_f32, int64_t ldo, const int* mask_dev) {
  const flux2b_te_config& t = c->te;
  const int Hd = t.hidden_size, I = t.intermediate_size, Hq = t.num_heads, Hkv = t.num_kv_heads;
  const int Nq = Hq * 128, Nkv = Hkv * 128, Nqkv = Nq + 2 * Nkv;
  const bool f16 = c->f16();
  cudaStream_t st = c->stream;
  int deepest = 0;
  for (int i = 0; i < n_layers; ++i) {
    if (layers[i] < 0 || layers[i] > t.num_layers)
      return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "hidden-state layer index out of range: " + std::to_string(layers[i]));
    deepest = std::max(deepest, layers[i]);
  }
  if (deepest > c->te_layers_built)
    return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "hidden state of layer " + std::to_string(deepest) + " requested but only " +
                                                 std::to_string(c->te_layers_built) + " layers are loaded");
  {
    // A captured prefill graph bakes in these buffers' addresses. DevBuf::ensure frees and reallocates when a longer sequence
    // needs more room, so any growth starts a new workspace generation and graphs of older generations are never replayed
    // (te_graph_for drops them): S = 128, then 512, then 128 again must not run the first graph against freed memory.
    const void* before[6] = {c->ws_x.p, c->ws_xn.p, c->ws_qkv.p, c->ws_cat.p, c->ws_cos.p, c->ws_sin.p};
    F2B_CUDA(c->ws_x.ensure((size_t)S * Hd * 4));
    F2B_CUDA(c->ws_xn.ensure((size_t)S * Hd * 2));
    F2B_CUDA(c->ws_qkv.ensure((size_t)S * Nqkv * 2));
    F2B_CUDA(c->ws_cat.ensure((size_t)S * ((size_t)Nq + 3 * (size_t)I) * 2));   // ATT | ACT (+ the unfused [gate | up] fallback)
    F2B_CUDA(c->ws_cos.ensure((size_t)S * 128 * 4));
    F2B_CUDA(c->ws_sin.ensure((size_t)S * 128 * 4));
    const void* after[6] = {c->ws_x.p, c->ws_xn.p, c->ws_qkv.p, c->ws_cat.p, c->ws_cos.p, c->ws_sin.p};
    for (int i = 0; i < 6; ++i)
      if (before[i] != after[i]) { ++c->te_ws_gen; break; }
  }
  float* X = c->ws_x.as<float>();
  uint16_t* XN = c->ws_xn.as<uint16_t>();
  uint16_t* QKV = c->ws_qkv.as<uint16_t>();
  uint16_t* ATT = c->ws_cat.as<uint16_t>();
  uint16_t* ACT = ATT + (size_t)S * Nq;
  uint16_t* GU = ACT + (size_t)S * I;
  float* cosT = c->ws_cos.as<float>();
  float* sinT = c->ws_sin.as<float>();

  auto emit = [&](int layer_idx, const float* src) -> int {
    for (int i = 0; i < n_layers; ++i)
      if (layers[i] == layer_idx) {
        ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * Hd * 8);
        F2B_CUDA(copy_f32_to_any(src, Hd, out_f32 + (size_t)i * Hd, ldo, S, Hd, 0, st));
      }
    return 0;
  };

  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * Hd * 6);
    F2B_CUDA(embed_rows(ids, c->te_embed.p, t.vocab_size, Hd, X, Hd, S, f16, st));
  }
  {
    ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * 128 * 8);
    F2B_CUDA(rope_half_table(S, 0, t.rope_theta, cosT, sinT, st));
  }
  F2B_TRY(emit(0, X));

  for (int l = 0; l < deepest; ++l) {
    const TeLayerW& L = c->te_layers[l];
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * Hd * 6);
      F2B_CUDA(rms_norm_rows(X, Hd, L.ln1.as<float>(), XN, Hd, S, Hd, t.rms_norm_eps, false, f16, st));
    }
    {
      Epilogue e;
      e.mode = EPI_QKV_ROPE; e.out = QKV; e.ldo = Nqkv;
      e.cos = cosT; e.sin = sinT;
      e.norm_q = t.qk_norm ? L.nq.as<float>() : nullptr;
      e.norm_k = t.qk_norm ? L.nk.as<float>() : nullptr;
      e.dmodel = Nq; e.k_col0 = Nq; e.v_col0 = Nq + Nkv; e.rope_half = 1; e.eps = t.rms_norm_eps;
      F2B_TRY(te_gemm(c, XN, Hd, L.qkv, S, e));
    }
    {
      AttnProblem a;
      a.q = QKV; a.ldq = Nqkv; a.q_rows_total = S; a.q_row0 = 0; a.sq = S;
      a.o = ATT; a.ldo = Nq; a.o_row0 = 0;
      a.num_heads = Hq; a.batch = 1;
      a.scale = 1.0f / sqrtf(128.0f);
      a.num_segments = 1;
      a.seg[0].k = QKV + Nq; a.seg[0].v = QKV + Nq + Nkv; a.seg[0].ldk = a.seg[0].ldv = Nqkv;
      a.seg[0].rows_total = S; a.seg[0].row0 = 0; a.seg[0].len = S;
      a.causal = 1; a.kv_group = Hq / Hkv; a.key_lo = key_lo; a.key_hi = key_hi; a.pad_bias = -1e9f; a.mask_dev = mask_dev;
      a.f16 = f16 ? 1 : 0; a.variant = 3; a.poly = c->option("attn_poly", 0);
      ProfScope ps(c, FLUX2B_PROF_ATTN, 2.0 * S * (double)S * Nq, 2.0 * S * (2.0 * Nq + 2.0 * Nkv));   // causal: half of 4 S^2 D
      F2B_CUDA(attention_launch(a, st));
    }
    {
      Epilogue e; e.mode = EPI_GATE_RES; e.out = X; e.ldo = Hd; e.res = X; e.ldr = Hd; e.gate = c->te_ones.as<float>();
      F2B_TRY(te_gemm(c, ATT, Nq, L.o, S, e));
    }
    {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * Hd * 6);
      F2B_CUDA(rms_norm_rows(X, Hd, L.ln2.as<float>(), XN, Hd, S, Hd, t.rms_norm_eps, false, f16, st));
    }
    if (L.mlp_tiled) {
      Epilogue e; e.mode = EPI_SWIGLU; e.out = ACT; e.ldo = I;
      F2B_TRY(te_gemm(c, XN, Hd, L.gate_up, S, e));
    } else {
      Epilogue e; e.mode = EPI_BF16; e.out = GU; e.ldo = 2 * I;
      F2B_TRY(te_gemm(c, XN, Hd, L.gate_up, S, e));
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * I * 6);
      F2B_CUDA(swiglu(GU, 2 * I, ACT, I, S, I, f16, st));
    }
    {
      Epilogue e; e.mode = EPI_GATE_RES; e.out = X; e.ldo = Hd; e.res = X; e.ldr = Hd; e.gate = c->te_ones.as<float>();
      F2B_TRY(te_gemm(c, ACT, I, L.down, S, e));
    }
    if (l + 1 < t.num_layers) F2B_TRY(emit(l + 1, X));
  }
  if (deepest == t.num_layers) {
    // the last hidden state is taken after the final norm (Qwen3Model.swift:183-189)
    for (int i = 0; i < n_layers; ++i)
      if (layers[i] == t.num_layers) {
        ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * Hd * 8);
        F2B_CUDA(rms_norm_rows(X, Hd, c->te_norm.as<float>(), out_f32 + (size_t)i * Hd, ldo, S, Hd, t.rms_norm_eps, true, f16, st));
      }
  }
  return 0;
}

void te_destroy_graphs(flux2b_ctx* c) {
  for (TeGraph& g : c->te_graphs)
    if (g.exec) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(g.exec));
  c->te_graphs.clear();
}

// The prefill as a CUDA graph (option te_graph, default 1): the ~190 kernels of a 27-layer prefill run 5 - 40 us each, less than
// the host needs to encode their tensor maps and launch them, so the un-captured forward is host-bound. One graph per
// (token count, layer set); token ids, the padding bounds and the fp32 result live at fixed device addresses owned by the graph
// entry, the bounds are read by the attention kernel from device memory, so one graph serves every prompt length.
static int te_graph_for(flux2b_ctx* c, int S, const int* layers, int n_layers, TeGraph** out) {
  // graphs captured against an older workspace generation point at freed buffers: drop them before looking anything up
  for (size_t i = 0; i < c->te_graphs.size();) {
    TeGraph& g = c->te_graphs[i];
    if (g.ws_gen != c->te_ws_gen) {
      if (g.exec) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(g.exec));
      c->te_graphs.erase(c->te_graphs.begin() + (long)i);
    } else ++i;
  }
  for (TeGraph& g : c->te_graphs)
    if (g.S == S && (int)g.layers.size() == n_layers && std::equal(g.layers.begin(), g.layers.end(), layers)) { *out = &g; return 0; }
  if (c->te_graphs.size() >= 8) te_destroy_graphs(c);   // bounded cache (a pipeline uses one or two shapes)
  c->te_graphs.emplace_back();
  TeGraph& g = c->te_graphs.back();
  g.S = S; g.layers.assign(layers, layers + n_layers);
  const int64_t ldo = (int64_t)n_layers * c->te.hidden_size;
  auto drop = [&](int rc) { c->te_graphs.pop_back(); return rc; };
  if (g.ids.alloc((size_t)S * 4) != cudaSuccess || g.mask.alloc(8) != cudaSuccess || g.out32.alloc((size_t)S * ldo * 4) != cudaSuccess) {
    cudaGetLastError();
    return drop(fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "text-encoder graph buffers"));
  }
  cudaError_t e = cudaMemsetAsync(g.ids.p, 0, (size_t)S * 4, c->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g.mask.p, 0, 8, c->stream);
  if (e != cudaSuccess) return drop(fail(FLUX2B_ERR_CUDA, cudaGetErrorString(e)));
  // un-captured warm-up: grows the workspaces and sets the kernels' attributes, so that the capture allocates nothing
  const int64_t l0 = c->launches;
  int rc = te_forward_device(c, S, g.ids.as<int32_t>(), 0, 0, layers, n_layers, g.out32.as<float>(), ldo, g.mask.as<int>());
  if (rc) return drop(rc);
  g.launches = c->launches - l0;
  e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return drop(fail(FLUX2B_ERR_CUDA, cudaGetErrorString(e)));
  // Capture problems are not errors of the call: the plain launch sequence is always available (option te_graph is switched off).
  auto no_graph = [&](const std::string& why) {
    cudaGetLastError();
    set_error("text-encoder CUDA graph disabled: " + why);
    c->opt["te_graph"] = 0;
    c->te_graphs.pop_back();
    *out = nullptr;
    return 0;
  };
  e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) return no_graph(std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
  rc = te_forward_device(c, S, g.ids.as<int32_t>(), 0, 0, layers, n_layers, g.out32.as<float>(), ldo, g.mask.as<int>());
  c->launches -= g.launches;   // the capture enqueued nothing
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(c->stream, &graph);
  if (rc || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    return no_graph(rc ? std::string("launch failed during capture") : std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return no_graph(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
  g.exec = exec;
  g.ws_gen = c->te_ws_gen;   // the warm-up above grew the workspaces if needed; the capture saw their final addresses
  *out = &g;
  return 0;
}

}  // namespace f2b

using namespace f2b;

extern "C" {

int flux2b_te_hidden_states(flux2b_ctx* c, int B, int S, const int32_t* input_ids, const int32_t* attention_mask,
                            const int* layer_indices, int n_layers, void* out, int out_dtype) {
  if (!c) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null context");
  if (cudaSetDevice(c->device) != cudaSuccess) return fail(FLUX2B_ERR_NO_DEVICE, "cudaSetDevice failed");
  if (!c->has_te) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "not a text-encoder context (flux2b_te_create)");
  if (!c->finalized) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "text-encoder weights are not finalized");
  if (B < 1 || S < 1 || !input_ids || !layer_indices || n_layers < 1 || n_layers > 64 || !out)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad te_hidden_states arguments");
  if (out_dtype != FLUX2B_F32 && out_dtype != FLUX2B_F16 && out_dtype != FLUX2B_BF16_T)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "out_dtype must be f32, f16 or bf16");
  if (S > c->te.max_position_embeddings && c->te.max_position_embeddings > 0)
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "sequence longer than original_max_position_embeddings: the Llama-4 query scaling "
                                                  "(MistralAttention.swift:15-32) would no longer be 1");
  // ids and mask are tiny: validated on the host (the reference would trap on an out-of-range embedding row)
  std::vector<int32_t> h_ids((size_t)B * S), h_mask;
  if (is_device_ptr(input_ids)) F2B_CUDA(cudaMemcpy(h_ids.data(), input_ids, h_ids.size() * 4, cudaMemcpyDeviceToHost));
  else memcpy(h_ids.data(), input_ids, h_ids.size() * 4);
  for (int32_t v : h_ids)
    if (v < 0 || v >= c->te.vocab_size) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "token id out of range: " + std::to_string(v));
  std::vector<int> lo(B, 0), hi(B, 0);
  if (attention_mask) {
    h_mask.resize((size_t)B * S);
    if (is_device_ptr(attention_mask)) F2B_CUDA(cudaMemcpy(h_mask.data(), attention_mask, h_mask.size() * 4, cudaMemcpyDeviceToHost));
    else memcpy(h_mask.data(), attention_mask, h_mask.size() * 4);
    for (int b = 0; b < B; ++b) {
      // the extractors build left- or right-padded masks: the ones must form one run [lo, hi)
      const int32_t* m = h_mask.data() + (size_t)b * S;
      int first = -1, last = -1, count = 0;
      for (int i = 0; i < S; ++i)
        if (m[i] == 1) { if (first < 0) first = i; last = i; ++count; }
      if (count == 0 || last - first + 1 != count)
        return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "attention_mask must be 1 on one contiguous, non-empty run of tokens (left or right padding)");
      lo[b] = first; hi[b] = last + 1;
      if (lo[b] == 0 && hi[b] == S) hi[b] = 0;   // nothing is padded: no mask term at all
    }
  }
  const int Hd = c->te.hidden_size;
  const int64_t ldo = (int64_t)n_layers * Hd;
  const size_t esz = dtype_size(out_dtype);
  const size_t obytes = (size_t)B * S * ldo * esz;
  const void* d_ids;
  F2B_TRY(dev_in(c, h_ids.data(), h_ids.size() * 4, &d_ids));
  void* dout; bool ho;
  F2B_TRY(dev_out(c, out, obytes, &dout, &ho));
  float* acc = reinterpret_cast<float*>(dout);
  if (out_dtype != FLUX2B_F32) {
    acc = reinterpret_cast<float*>(c->scratch_buf("te_out_f32", (size_t)S * ldo * 4));
    if (!acc) { cudaGetLastError(); return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "te output scratch allocation failed"); }
  }
  const bool use_graph = c->option("te_graph", 1) != 0 && !c->prof_on;
  TeGraph* tg = nullptr;
  if (use_graph) F2B_TRY(te_graph_for(c, S, layer_indices, n_layers, &tg));
  for (int b = 0; b < B; ++b) {
    float* dst32 = out_dtype == FLUX2B_F32 ? acc + (size_t)b * S * ldo : acc;
    const int32_t* ids_b = reinterpret_cast<const int32_t*>(d_ids) + (size_t)b * S;
    if (tg) {
      const int bounds[2] = {lo[b], hi[b]};
      F2B_CUDA(cudaMemcpyAsync(tg->ids.p, ids_b, (size_t)S * 4, cudaMemcpyDeviceToDevice, c->stream));
      F2B_CUDA(cudaMemcpyAsync(tg->mask.p, bounds, 8, cudaMemcpyHostToDevice, c->stream));   // pageable source: staged before return
      F2B_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(tg->exec), c->stream));
      c->launches += tg->launches;
      c->prof[FLUX2B_PROF_ELEMWISE].launches++; c->launches++;
      F2B_CUDA(copy_f32_to_any(tg->out32.as<float>(), ldo, reinterpret_cast<uint8_t*>(dout) + (size_t)b * S * ldo * esz, ldo, S, (int)ldo,
                               out_dtype == FLUX2B_F32 ? 0 : out_dtype == FLUX2B_F16 ? 1 : 2, c->stream));
      continue;
    }
    F2B_TRY(te_forward_device(c, S, ids_b, lo[b], hi[b], layer_indices, n_layers, dst32, ldo));
    if (out_dtype != FLUX2B_F32) {
      ProfScope ps(c, FLUX2B_PROF_ELEMWISE, 0, (double)S * ldo * 6);
      F2B_CUDA(copy_f32_to_any(acc, ldo, reinterpret_cast<uint8_t*>(dout) + (size_t)b * S * ldo * esz, ldo, S, (int)ldo,
                               out_dtype == FLUX2B_F16 ? 1 : 2, c->stream));
    }
  }
  F2B_TRY(finish_out(c, out, dout, obytes, ho));
  return end_call(c, false);
}

// kernel-level entry for the parity tests: causal grouped-query attention with the reference's padding mask over a packed
// [S, (Hq + 2 Hkv) * 128] q | k | v buffer (q / k already normed and rotated) -> out [S, Hq * 128]
int flux2b_op_attention_causal(flux2b_ctx* c, const void* qkv16, int S, int num_heads, int num_kv_heads, int key_lo, int key_hi,
                               void* out16) {
  if (!c) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null context");
  if (cudaSetDevice(c->device) != cudaSuccess) return fail(FLUX2B_ERR_NO_DEVICE, "cudaSetDevice failed");
  if (S < 1 || num_heads < 1 || num_kv_heads < 1 || num_heads % num_kv_heads || key_lo < 0 || key_hi > S || (key_hi > 0 && key_lo >= key_hi))
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad causal attention arguments");
  const int Nq = num_heads * 128, Nkv = num_kv_heads * 128, Nqkv = Nq + 2 * Nkv;
  const void* dq;
  F2B_TRY(dev_in(c, qkv16, (size_t)S * Nqkv * 2, &dq));
  void* dout; bool ho;
  const size_t obytes = (size_t)S * Nq * 2;
  F2B_TRY(dev_out(c, out16, obytes, &dout, &ho));
  AttnProblem a;
  a.q = dq; a.ldq = Nqkv; a.q_rows_total = S; a.sq = S;
  a.o = dout; a.ldo = Nq;
  a.num_heads = num_heads; a.batch = 1; a.scale = 1.0f / sqrtf(128.f);
  a.num_segments = 1;
  a.seg[0].k = (const uint16_t*)dq + Nq; a.seg[0].v = (const uint16_t*)dq + Nq + Nkv;
  a.seg[0].ldk = a.seg[0].ldv = Nqkv; a.seg[0].rows_total = S; a.seg[0].len = S;
  a.causal = 1; a.kv_group = num_heads / num_kv_heads; a.key_lo = key_lo; a.key_hi = key_hi;
  a.f16 = c->f16(); a.variant = 3; a.poly = c->option("attn_poly", 0);
  {
    ProfScope ps(c, FLUX2B_PROF_ATTN, 2.0 * S * (double)S * Nq, 2.0 * S * (2.0 * Nq + 2.0 * Nkv));
    F2B_CUDA(attention_launch(a, c->stream));
  }
  F2B_TRY(finish_out(c, out16, dout, obytes, ho));
  return end_call(c, false);
}

}  // extern "C"
