"""CPU: pin the oracle to everything the reference's own XCTests pin on this path.
Each test names the reference test it re-expresses (Tests/Flux2CoreTests/...)."""
import math

import numpy as np
import pytest
import torch

from oracle import flux2_oracle as O
from oracle import quant_oracle as Q


def test_scheduler_step_counts_and_monotonic():
    # Flux2CoreTests.swift:177-204 — N steps -> N+1 sigmas, strictly decreasing, first 1.0, last 0
    for n in (4, 20, 28, 50):
        s = O.FlowMatchEulerScheduler()
        s.set_timesteps(n, 4096)
        assert len(s.sigmas) == n + 1 and s.sigmas[0] == 1.0 and s.sigmas[-1] == 0.0
        assert all(a > b for a, b in zip(s.sigmas, s.sigmas[1:]))
        assert s.timesteps[0] == 1000.0


def test_strength_half_of_50_gives_25_steps():
    # Flux2CoreTests.swift:348-361
    s = O.FlowMatchEulerScheduler()
    t0 = s.set_timesteps(50, 4096, strength=0.5)
    assert t0 == 25 and len(s.sigmas) - 1 == 25


def test_custom_sigmas_append_terminal_zero():
    # Flux2CoreTests.swift:336-346
    s = O.FlowMatchEulerScheduler()
    s.set_custom_sigmas([1.0, 0.5, 0.25])
    assert s.sigmas == [1.0, 0.5, 0.25, 0.0]
    s.set_custom_sigmas([1.0, 0.0])
    assert s.sigmas == [1.0, 0.0]


def test_scale_noise_sigma_zero_is_identity():
    # Flux2CoreTests.swift:405-415
    x, n = torch.randn(1, 8, 128), torch.randn(1, 8, 128)
    assert torch.equal(O.FlowMatchEulerScheduler.scale_noise(x, 0.0, n), x)
    assert torch.allclose(O.FlowMatchEulerScheduler.scale_noise(x, 1.0, n), n)


def test_empirical_mu_inequalities():
    # Flux2CoreTests.swift:1211-1233 — mu grows with sequence length; branch switch at 4300
    assert O.compute_empirical_mu(4096, 4) > O.compute_empirical_mu(256, 4)
    assert O.compute_empirical_mu(16384, 28) > O.compute_empirical_mu(4096, 28)
    assert abs(O.compute_empirical_mu(4301, 10) - (0.00016927 * 4301 + 0.45666666)) < 1e-5


def test_pack_unpack_roundtrip_shapes():
    # Flux2CoreTests.swift:139-162, 802-819
    x = torch.randn(1, 128, 64, 48)
    seq = O.pack_patchified_to_sequence(x)
    assert seq.shape == (1, 64 * 48, 128)
    assert torch.equal(O.unpack_sequence_to_patchified(seq, 1024, 768), x)
    lat = O.unpatchify_latents(x)
    assert lat.shape == (1, 32, 128, 96)
    assert torch.equal(O.pack_latents_to_patchified(lat), x)
    # channel index = c*4 + ph*2 + pw (LatentUtils.swift:135-141)
    assert lat[0, 5, 2 * 7 + 1, 2 * 9 + 0] == x[0, 5 * 4 + 1 * 2 + 0, 7, 9]


def test_position_ids_layout_and_reference_t_coordinate():
    # ImageToImageTrainingTests.swift:286-319 — shapes, (T,H,W,L), reference T = 10, 20, ...
    ids = O.image_position_ids(1024, 1024)
    assert ids.shape == (4096, 4) and ids[65].tolist() == [0, 1, 1, 0]
    t = O.text_position_ids(512)
    assert t.shape == (512, 4) and t[7].tolist() == [0, 0, 0, 7]
    r = O.reference_position_ids([2, 2], [3, 3])
    assert r.shape == (12, 4) and r[0, 0] == 10 and r[6, 0] == 20 and r[7].tolist() == [20, 0, 1, 0]


def test_kv_extraction_mask_pattern():
    # Flux2CoreTests.swift:2577-2681 — exactly 0 / -inf, reference queries blocked from output keys only
    m = O.kv_extraction_mask(4, 3, 5)[0, 0]
    assert m.shape == (12, 12)
    assert torch.isinf(m[4:7, 7:]).all() and (m[4:7, 7:] < 0).all()
    assert (m[:4] == 0).all() and (m[7:] == 0).all() and (m[4:7, :7] == 0).all()


def test_timestep_embedding_and_rope_shapes():
    # Flux2CoreTests.swift:227-257
    e = O.timesteps_proj(torch.tensor([500.0, 1.0]))
    assert e.shape == (2, 256) and torch.allclose(e[:, 0], torch.cos(torch.tensor([500.0, 1.0])))
    cos, sin = O.rope_embeddings(O.image_position_ids(64, 64))
    assert cos.shape == (16, 128) and torch.equal(cos[:, 0::2], cos[:, 1::2])
    assert torch.allclose(cos ** 2 + sin ** 2, torch.ones_like(cos), atol=1e-6)


def test_modulation_and_feedforward_shapes():
    # Flux2CoreTests.swift:264-288
    cfg = O.DiTConfig(num_layers=1, num_single_layers=1, num_attention_heads=1, joint_attention_dim=64, guidance_embeds=False)
    W = O.random_dit_weights(cfg)
    mods = O.modulation(W, "doubleStreamModulationImg", torch.randn(2, 128), 2, 128)
    assert len(mods) == 2 and all(t.shape == (2, 128) for m in mods for t in m)
    assert O.feed_forward(W, "transformerBlocks.0.ff", torch.randn(1, 5, 128)).shape == (1, 5, 128)


def test_quantization_modes_runtime_contract():
    # Flux2CoreTests.swift:64-85,1100-1142 — (bits, group, mode), biases iff affine, shapes, requant shape stability
    w = (torch.randn(32, 64) * 0.1).half().numpy()
    for name, (bits, group, has_b) in {"qint8": (8, 64, True), "int4": (4, 64, True), "mxfp8": (8, 32, False),
                                        "mxfp4": (4, 32, False), "nvfp4": (4, 16, False)}.items():
        q = Q.QUANT[name]
        assert Q.params(q) == (bits, group, has_b)
        p, s, b = Q.quantize(q, w)
        assert p.shape == (32, 64 * bits // 32) and s.shape == (32, 64 // group)
        assert (b is not None) == has_b
        d = Q.dequantize(q, p, s, b, 64)
        p2, s2, b2 = Q.quantize(q, d)
        assert p2.shape == p.shape and s2.shape == s.shape


def test_bf16_to_f16_direct_equals_via_f32():
    # Flux2CoreTests.swift:1770-1849 — bit-identical on 3072^2 normal weights (here 512^2 to stay fast)
    w = torch.randn(512, 512).to(torch.bfloat16)
    assert torch.equal(w.to(torch.float16), w.to(torch.float32).to(torch.float16))


def test_mask_blend_semantics():
    # Flux2ChainsTests.swift:48-110 — all-white mask keeps x, all-black replaces with the re-noised known latents
    x, x0, e = torch.randn(4, 128), torch.randn(4, 128), torch.randn(4, 128)
    assert torch.equal(O.repaint_blend(x, x0, e, torch.ones(4, 128), 0.3), x)
    assert torch.allclose(O.repaint_blend(x, x0, e, torch.zeros(4, 128), 0.3), 0.7 * x0 + 0.3 * e)


def test_postprocess_uint8_truncates():
    img = torch.tensor([[[[-1.0, 1.0], [0.0, 0.999]]]]).repeat(1, 3, 1, 1)
    out = O.postprocess_vae_output(img)
    assert out.shape == (2, 2, 3) and out[0, 0, 0] == 0 and out[0, 1, 0] == 255 and out[1, 0, 0] == 127 and out[1, 1, 0] == 254


# ---- format tables of the quantizer oracle (exhaustive where the code space is small)
def test_e4m3_roundtrip_all_codes():
    L = Q.lib()
    for c in range(256):
        if (c & 0x7F) == 0x7F:
            continue
        v = L.oracle_from_e4m3(c)
        assert L.oracle_to_e4m3(v) == c or v == 0.0
    assert L.oracle_to_e4m3(1e9) == 0x7E and L.oracle_from_e4m3(0x7E) == 448.0
    # ties to even: halfway between 1.0 (0x38) and 1.125 (0x39) -> 1.0 ; between 1.125 and 1.25 -> 1.25
    assert L.oracle_to_e4m3(1.0625) == 0x38 and L.oracle_to_e4m3(1.1875) == 0x3A


def test_e2m1_grid_and_ties():
    L = Q.lib()
    grid = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]
    for i, v in enumerate(grid):
        assert L.oracle_to_e2m1(v) == i and L.oracle_from_e2m1(i) == v
        assert L.oracle_from_e2m1(L.oracle_to_e2m1(-v)) == -v
    for x, want in [(0.25, 0.0), (0.75, 1.0), (1.25, 1.0), (1.75, 2.0), (2.5, 2.0), (3.5, 4.0), (5.0, 4.0), (100.0, 6.0)]:
        assert L.oracle_from_e2m1(L.oracle_to_e2m1(x)) == want


def test_e8m0_is_round_log2():
    L = Q.lib()
    rng = np.random.default_rng(0)
    for x in np.exp2(rng.uniform(-20, 20, 2000)).astype(np.float32):
        want = int(np.clip(np.round(np.log2(np.float64(x))), -127, 127)) + 127
        assert L.oracle_to_e8m0(float(x)) == want
    assert L.oracle_to_e8m0(0.0) == 0 and L.oracle_from_e8m0(127) == 1.0


def test_affine_reconstruction_error_bound():
    w = (torch.randn(64, 256) * 0.05).half().numpy()
    for q, bits in ((1, 8), (2, 4)):
        p, s, b = Q.quantize(q, w)
        d = Q.dequantize(q, p, s, b, 256)
        step = np.abs(s.astype(np.float32)).repeat(64, axis=1)
        # MLX's edge rule re-fits the scale so that 0 is exactly representable, which can clip the far end of the
        # range by up to one step; everything else is within half a step (+ f16 rounding of scale / bias)
        assert np.all(np.abs(d - w.astype(np.float32)) <= 1.0 * step + (2 ** bits) * step * 2.0 ** -11 + 2e-4)
        assert np.mean(np.abs(d - w.astype(np.float32)) <= 0.5 * step + (2 ** bits) * step * 2.0 ** -11 + 2e-4) > 0.97
        assert p.max() < 2 ** 32 and (bits == 8 or True)


def test_fp4_dequantized_values_lie_on_the_scaled_grid():
    # every dequantized element equals (E2M1 grid value) x (decoded group scale), exactly
    L = Q.lib()
    w = (torch.randn(16, 128) * 0.05).half().numpy()
    grid = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=np.float32)
    for q, group, dec in ((4, 32, L.oracle_from_e8m0), (5, 16, L.oracle_from_e4m3)):
        p, s, _ = Q.quantize(q, w)
        d = Q.dequantize(q, p, s, None, 128)
        sc = np.vectorize(lambda b: dec(int(b)))(s).astype(np.float32).repeat(group, axis=1)
        ratio = np.where(sc > 0, np.abs(d) / np.where(sc > 0, sc, 1), 0)
        assert np.isin(ratio, grid).all()
        # the largest element of every group survives to within the E2M1 spacing at the top of the range
        amax = np.abs(w.astype(np.float32)).reshape(16, -1, group).max(-1)
        dmax = np.abs(d).reshape(16, -1, group).max(-1)
        assert np.all(dmax <= 6.0 * sc.reshape(16, -1, group)[:, :, 0] + 1e-12) and np.all(dmax >= 0.5 * amax)


# ------------------------------------------------------------------ text encoder: what the reference's source fixes
def test_te_causal_mask_values_and_padding():
    """createCausalMask (FluxTextEncoders/Model/Qwen3/Qwen3Model.swift:196-231): 0 on/below the diagonal, -inf above, -1e9 added on
    padded keys; right padding for Klein (KleinEmbeddingExtractor.swift:78-95), left padding for Dev (EmbeddingExtractor.swift:227-248)."""
    ids, m = O.te_pad_tokens([11, 12, 13], 6, 151643, "right")
    assert ids.tolist() == [[11, 12, 13, 151643, 151643, 151643]] and m.tolist() == [[1, 1, 1, 0, 0, 0]]
    ids, m = O.te_pad_tokens([11, 12, 13], 6, 11, "left")
    assert ids.tolist() == [[11, 11, 11, 11, 12, 13]] and m.tolist() == [[0, 0, 0, 1, 1, 1]]
    ids, m = O.te_pad_tokens(list(range(10)), 6, 0, "right")          # truncation keeps the prefix
    assert ids.tolist() == [list(range(6))] and m.sum() == 6
    mask = O.te_causal_mask(4, torch.tensor([[1, 1, 0, 0]]))[0, 0]
    assert mask[0, 0] == 0 and mask[1, 0] == 0 and mask[1, 1] == 0
    assert torch.isinf(mask[0, 1]) and mask[0, 1] < 0 and torch.isinf(mask[2, 3])
    assert mask[2, 2] == -1e9 and mask[3, 2] == -1e9 and mask[3, 3] == -1e9 and mask[3, 1] == 0
    assert O.te_causal_mask(3, None).shape == (1, 1, 3, 3)


def test_te_left_padded_rows_attend_uniformly():
    """-1e9 absorbs every realistic score in fp32 (ulp(1e9) = 64), so a padded query row that sees only padded keys gets a uniform
    softmax — the behaviour the device kernel reproduces instead of treating the bias as -inf."""
    s = torch.tensor([3.5, -7.25, 0.125]) + torch.tensor(-1e9)
    assert torch.all(s == -1e9)
    p = torch.softmax(s, dim=-1)
    assert torch.allclose(p, torch.full((3,), 1 / 3))


def test_te_rope_is_rotate_half_and_position_zero_is_identity():
    """MLXFast.RoPE(traditional: false): pairs (j, j + hd/2), angle pos * base^(-j/(hd/2)) (Qwen3Attention.swift:29-35)."""
    x = torch.randn(1, 2, 5, 128, generator=torch.Generator().manual_seed(0))
    y = O.te_rope_half(x, 1e6)
    assert torch.equal(y[:, :, 0], x[:, :, 0])                             # position 0: angle 0
    # the norm of each (j, j + 64) pair is preserved
    n0 = x[..., :64] ** 2 + x[..., 64:] ** 2
    n1 = y[..., :64] ** 2 + y[..., 64:] ** 2
    assert torch.allclose(n0, n1, rtol=1e-4, atol=1e-5)
    j = 1
    ang = 3 * 1e6 ** (-j / 64)
    assert torch.allclose(y[0, 0, 3, j], x[0, 0, 3, j] * math.cos(ang) - x[0, 0, 3, j + 64] * math.sin(ang), atol=1e-5)
    assert torch.allclose(y[0, 0, 3, j + 64], x[0, 0, 3, j + 64] * math.cos(ang) + x[0, 0, 3, j] * math.sin(ang), atol=1e-5)


def test_te_hidden_state_indexing_and_gqa():
    """forwardWithHiddenStates (Qwen3Model.swift:104-191): index 0 = embeddings, i = after layer i, num_layers = after the final norm;
    Klein extracts 9/18/27, Dev 10/20/30; concatenation order = request order. GQA: query head h uses kv head h // rep."""
    assert O.KLEIN_HIDDEN_STATE_LAYERS == (9, 18, 27) and O.FLUX_HIDDEN_STATE_LAYERS == (10, 20, 30)
    assert 3 * O.qwen3_4b().hidden_size == 7680 and 3 * O.qwen3_8b().hidden_size == 12288     # == joint_attention_dim of Klein 4B / 9B
    cfg = O.TEConfig(vocab_size=64, hidden_size=128, intermediate_size=128, num_layers=2, num_heads=2, num_kv_heads=1)
    W = O.random_te_weights(cfg, seed=0)
    ids = torch.tensor([[5, 9, 33, 2]], dtype=torch.int32)
    h = O.te_hidden_states(W, cfg, ids, None, (0, 1, 2))
    assert h.shape == (1, 4, 3 * 128)
    assert torch.equal(h[..., :128], W["model.embed_tokens.weight"][ids.long()])
    mask = O.te_causal_mask(4, None)
    l1 = O.te_decoder_layer(W, 0, cfg, h[..., :128], mask)
    assert torch.allclose(h[..., 128:256], l1, atol=1e-6)
    l2 = O.rms_norm(O.te_decoder_layer(W, 1, cfg, l1, mask), W["model.norm.weight"], cfg.rms_norm_eps)
    assert torch.allclose(h[..., 256:], l2, atol=1e-6)
    swapped = O.te_hidden_states(W, cfg, ids, None, (2, 0))
    assert torch.equal(swapped[..., :128], h[..., 256:]) and torch.equal(swapped[..., 128:], h[..., :128])
    # causality: changing the last token leaves every earlier row untouched
    ids2 = ids.clone(); ids2[0, 3] = 7
    h2 = O.te_hidden_states(W, cfg, ids2, None, (1,))
    assert torch.equal(h2[:, :3], h[:, :3, 128:256]) and not torch.equal(h2[:, 3], h[:, 3, 128:256])
    # right padding never changes the real tokens' rows
    idp, mp = O.te_pad_tokens([5, 9, 33, 2], 8, 63, "right")
    hp = O.te_hidden_states(W, cfg, idp, mp, (1,))
    assert torch.allclose(hp[:, :4], h[..., 128:256], atol=1e-6)


def test_llama4_query_scale_is_one_below_the_original_context():
    """MistralAttention.swift:15-32,422-432: the query scale only departs from 1 at positions >= original_max_position_embeddings."""
    s = O.llama4_attention_scale(0, 512, beta=0.1, max_position_embeddings=8192)
    assert torch.all(s == 1.0)
    s = O.llama4_attention_scale(8190, 8194, beta=0.1, max_position_embeddings=8192)
    assert s[0] == 1.0 and s[1] == 1.0 and torch.allclose(s[2:], torch.tensor(1.0 + 0.1 * math.log(2.0)))
