"""Pins the text-encoder oracle against an INDEPENDENT published implementation of the same architectures:
Hugging Face transformers' Qwen3Model / MistralModel (the checkpoints the reference's Swift encoders load are HF checkpoints, and
Sources/FluxTextEncoders/Model/Qwen3/*.swift / MistralModel.swift are ports of these modules). Same random weights, same token ids:
every hidden state the extractors use (KleinEmbeddingExtractor.swift:98-121, EmbeddingExtractor.swift:252-285) must agree in fp32.

This is not MLX (the arithmetic of mlx-swift stays unpinned) — it pins the STRUCTURE the oracle restates: layer
indexing of hidden_states, q/k RMSNorm before RoPE, rotate-half pair layout and theta, grouped-query head mapping, SwiGLU, the final
norm on the last index, causal + right-padding masking of the valid rows."""
import pytest
import torch

from oracle import flux2_oracle as O

transformers = pytest.importorskip("transformers")


def _load(model, W, cfg):
    sd = {}
    for k, v in W.items():
        assert k.startswith("model.")
        sd[k[len("model."):]] = v.clone()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    return model.eval()


def _ids(S, n_valid, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (1, S), generator=g)
    mask = torch.zeros(1, S, dtype=torch.int64)
    mask[:, :n_valid] = 1   # right padding, as the Klein extractor pads
    return ids, mask


@pytest.mark.parametrize("n_valid", [48, 29])
def test_qwen3_oracle_matches_hf_transformers(n_valid):
    cfg = O.TEConfig(vocab_size=257, hidden_size=256, intermediate_size=512, num_layers=4, num_heads=4, num_kv_heads=2, head_dim=128,
                     qk_norm=True, rms_norm_eps=1e-6, rope_theta=1_000_000.0)
    W = O.random_te_weights(cfg, seed=5, round_to=None)
    hf_cfg = transformers.Qwen3Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                                      num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_kv_heads,
                                      head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, max_position_embeddings=512,
                                      attention_bias=False, tie_word_embeddings=False, use_sliding_window=False, attn_implementation="eager")
    model = _load(transformers.Qwen3Model(hf_cfg).to(torch.float32), W, cfg)
    S = 48
    ids, mask = _ids(S, n_valid, cfg.vocab_size, 11)
    with torch.no_grad():
        hs = model(input_ids=ids, attention_mask=mask, output_hidden_states=True).hidden_states
        layers = (1, 2, 3, 4)   # 4 = num_layers: after the final norm in both
        ref = O.te_hidden_states(W, cfg, ids.to(torch.int32), mask.to(torch.int32), layers)
    Hd = cfg.hidden_size
    for n, li in enumerate(layers):
        a, b = ref[0, :n_valid, n * Hd:(n + 1) * Hd].double(), hs[li][0, :n_valid].double()
        assert float((a - b).norm() / b.norm()) < 2e-5, (li, float((a - b).norm() / b.norm()))


def test_mistral_oracle_matches_hf_transformers():
    cfg = O.TEConfig(vocab_size=301, hidden_size=256, intermediate_size=384, num_layers=3, num_heads=4, num_kv_heads=1, head_dim=128,
                     qk_norm=False, rms_norm_eps=1e-5, rope_theta=1_000_000_000.0)
    W = O.random_te_weights(cfg, seed=6, round_to=None)
    hf_cfg = transformers.MistralConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                                        num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_kv_heads,
                                        head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, max_position_embeddings=512,
                                        sliding_window=None, tie_word_embeddings=False, attn_implementation="eager")
    model = _load(transformers.MistralModel(hf_cfg).to(torch.float32), W, cfg)
    S, n_valid = 40, 40
    ids, mask = _ids(S, n_valid, cfg.vocab_size, 12)
    with torch.no_grad():
        hs = model(input_ids=ids, attention_mask=mask, output_hidden_states=True).hidden_states
        layers = (1, 2)
        ref = O.te_hidden_states(W, cfg, ids.to(torch.int32), mask.to(torch.int32), layers)
    Hd = cfg.hidden_size
    for n, li in enumerate(layers):
        a, b = ref[0, :, n * Hd:(n + 1) * Hd].double(), hs[li][0].double()
        assert float((a - b).norm() / b.norm()) < 2e-5, (li, float((a - b).norm() / b.norm()))
