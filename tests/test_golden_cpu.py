"""CPU: the oracle still reproduces the committed fixtures (tests/golden/golden.npz, made by tools/make_golden.py).
Integer / packed results must match bit for bit; fp32 results within accumulation-order noise of the BLAS in use."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import flux2_oracle as O
from oracle import quant_oracle as Q
import make_golden as MG


def test_quantizer_fixtures_bit_exact(golden):
    w = golden["quant_w_f16"]
    for name, q in Q.QUANT.items():
        if q == 0:
            continue
        p, s, b = Q.quantize(q, w)
        assert np.array_equal(p, golden[f"quant_{name}_packed"])
        assert np.array_equal(s.view(np.uint8), golden[f"quant_{name}_scales"].view(np.uint8))
        if b is not None:
            assert np.array_equal(b.view(np.uint16), golden[f"quant_{name}_biases"].view(np.uint16))
        d = Q.dequantize(q, p, s, b, 256)
        assert np.array_equal(d.view(np.uint32), golden[f"quant_{name}_dequant"].view(np.uint32))


def test_scheduler_fixtures(golden):
    for steps, seq, strength in ((4, 4096, 1.0), (4, 256, 1.0), (28, 16384, 1.0), (50, 4096, 0.5)):
        s = O.FlowMatchEulerScheduler()
        s.set_timesteps(steps, seq, strength)
        np.testing.assert_allclose(np.array(s.sigmas, np.float32), golden[f"sigmas_{steps}_{seq}_{int(strength * 100)}"], rtol=1e-6)


def test_dit_and_vae_fixtures(golden):
    cfg, W, hidden, enc = MG.tiny_inputs()
    assert np.array_equal(hidden.numpy(), golden["dit_hidden"]) and np.array_equal(enc.numpy(), golden["dit_enc"])
    rec = []
    y = O.dit_forward(W, cfg, hidden, enc, torch.tensor([0.7]), torch.tensor([4.0]), O.image_position_ids(MG.HW, MG.HW),
                      O.text_position_ids(MG.S_TXT), record=rec)
    np.testing.assert_allclose(y.numpy(), golden["dit_out"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(torch.stack(rec).numpy(), golden["dit_blocks"], rtol=1e-4, atol=1e-4)
    vcfg = O.vae_small_decoder()
    VW = O.random_vae_weights(vcfg, seed=1)
    img = O.vae_decode(VW, vcfg, torch.from_numpy(golden["vae_z"]))
    np.testing.assert_allclose(img.numpy(), golden["vae_out"], rtol=1e-4, atol=1e-4)


def test_vae_encoder_fixture():
    g = dict(np.load(MG.GOLDEN_ENC))
    vcfg = O.vae_small_decoder()
    VW = O.random_vae_weights(vcfg, seed=3, encoder=True)
    # adding the encoder tensors after the decoder ones leaves every decoder tensor as it was
    VW0 = O.random_vae_weights(vcfg, seed=3)
    assert all(torch.equal(VW[k], v) for k, v in VW0.items())
    img = torch.from_numpy(g["img"])
    np.testing.assert_allclose(O.vae_encode_moments(VW, vcfg, img).numpy(), g["moments"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(O.encode_image_to_packed_sequence(VW, vcfg, img).numpy(), g["seq"], rtol=1e-4, atol=1e-4)
    # Downsample2D pads bottom / right only (ResnetBlock.swift:203-213): equals an explicit padded stride-2 correlation
    x = torch.randn(1, 6, 8, 4)
    W = {"d.conv.weight": torch.randn(5, 3, 3, 4), "d.conv.bias": torch.randn(5)}
    y = O.downsample2d(W, "d", x)
    xp = torch.zeros(1, 7, 9, 4); xp[:, :6, :8] = x
    ref = torch.zeros(1, 3, 4, 5)
    for oy in range(3):
        for ox in range(4):
            patch = xp[0, 2 * oy:2 * oy + 3, 2 * ox:2 * ox + 3]          # [3, 3, 4]
            ref[0, oy, ox] = (W["d.conv.weight"] * patch[None]).sum(dim=(1, 2, 3)) + W["d.conv.bias"]
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)


def test_text_encoder_fixture():
    g = dict(np.load(MG.GOLDEN_TE))
    for name, (cfg, seed, side, layers, toks) in MG.te_configs().items():
        W = O.random_te_weights(cfg, seed=seed)
        ids, mask = O.te_pad_tokens(toks, 32, 3, side)
        assert np.array_equal(ids.numpy(), g[f"{name}_ids"]) and np.array_equal(mask.numpy(), g[f"{name}_mask"])
        np.testing.assert_allclose(O.te_hidden_states(W, cfg, ids, mask, layers).numpy(), g[f"{name}_hidden"], rtol=1e-4, atol=1e-5)
