#!/usr/bin/env python
"""Generates tests/golden/golden.npz from the ORACLE (oracle/flux2_oracle.py + oracle/quant_oracle.c).

The reference itself (Swift on mlx-swift 0.31.6) cannot run in this container and holds no numeric golden vectors for
this path, so these fixtures pin the *restatement*: the CPU suite checks the oracle still
reproduces them (guards against torch / compiler drift), the GPU suite checks the CUDA path against them without
needing anything but the committed file.

usage: python tools/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flux2_oracle as O  # noqa: E402
from oracle import quant_oracle as Q  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "golden.npz")
GOLDEN_ENC = os.path.join(ROOT, "tests", "golden", "golden_vae_encoder.npz")   # added later: kept apart so golden.npz never changes
GOLDEN_TE = os.path.join(ROOT, "tests", "golden", "golden_text_encoder.npz")    # likewise

TINY = dict(num_layers=1, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256, guidance_embeds=True)
S_IMG, S_TXT, HW = 64, 64, 128  # 128x128 pixels -> 8x8 tokens


def tiny_inputs():
    cfg = O.DiTConfig(**TINY)
    W = O.random_dit_weights(cfg, seed=0, round_to=torch.bfloat16)
    hidden = torch.randn(1, S_IMG, 128, generator=torch.Generator().manual_seed(42))
    enc = torch.randn(1, S_TXT, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43))
    return cfg, W, hidden, enc


def main():
    torch.set_num_threads(1)  # fixed reduction order
    out = {}
    # ---- quantizers (C oracle)
    g = torch.Generator().manual_seed(5)
    w = torch.randn(64, 256, generator=g) * 0.05
    w[0, :64] = 0
    w[1, :64] = 0.03125
    w[2, 3] = 1000.0
    w[3, :16] = -w[3, :16].abs()
    wn = w.half().numpy()
    out["quant_w_f16"] = wn
    for name, q in Q.QUANT.items():
        if q == 0:
            continue
        p, s, b = Q.quantize(q, wn)
        out[f"quant_{name}_packed"], out[f"quant_{name}_scales"] = p, s
        if b is not None:
            out[f"quant_{name}_biases"] = b
        out[f"quant_{name}_dequant"] = Q.dequantize(q, p, s, b, 256)
    # ---- scheduler
    for steps, seq, strength in ((4, 4096, 1.0), (4, 256, 1.0), (28, 16384, 1.0), (50, 4096, 0.5)):
        s = O.FlowMatchEulerScheduler()
        s.set_timesteps(steps, seq, strength)
        out[f"sigmas_{steps}_{seq}_{int(strength * 100)}"] = np.array(s.sigmas, dtype=np.float32)
    # ---- tiny DiT forward + 3-step denoise
    cfg, W, hidden, enc = tiny_inputs()
    rec = []
    t, gd = torch.tensor([0.7]), torch.tensor([4.0])
    y = O.dit_forward(W, cfg, hidden, enc, t, gd, O.image_position_ids(HW, HW), O.text_position_ids(S_TXT), record=rec)
    out["dit_hidden"], out["dit_enc"], out["dit_out"] = hidden.numpy(), enc.numpy(), y.numpy()
    out["dit_blocks"] = torch.stack(rec).numpy()
    sched = O.FlowMatchEulerScheduler()
    sched.set_timesteps(3, S_IMG)
    out["denoise_sigmas"] = np.array(sched.sigmas, dtype=np.float32)
    out["denoise_out"] = O.denoise(W, cfg, hidden, enc, sched.sigmas, HW, HW, guidance=4.0).numpy()
    # ---- small-decoder VAE, 4x4 latent -> 32x32 image
    vcfg = O.vae_small_decoder()
    VW = O.random_vae_weights(vcfg, seed=1)
    z = torch.randn(1, 32, 4, 4, generator=torch.Generator().manual_seed(7))
    out["vae_z"], out["vae_out"] = z.numpy(), O.vae_decode(VW, vcfg, z).numpy()
    os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
    np.savez_compressed(GOLDEN, **out)
    print(f"wrote {GOLDEN}: {os.path.getsize(GOLDEN) / 1024:.0f} KiB, {len(out)} arrays")
    encoder_fixture()


def encoder_fixture():
    """VAE encoder: 32x48 image -> 4x6 latent -> 6 packed tokens; standard encoder widths."""
    torch.set_num_threads(1)
    vcfg = O.vae_small_decoder()
    VW = O.random_vae_weights(vcfg, seed=3, encoder=True)
    img = torch.rand(1, 3, 32, 48, generator=torch.Generator().manual_seed(11)) * 2 - 1
    enc = {"img": img.numpy(), "moments": O.vae_encode_moments(VW, vcfg, img).numpy(),
           "seq": O.encode_image_to_packed_sequence(VW, vcfg, img).numpy()}
    np.savez_compressed(GOLDEN_ENC, **enc)
    print(f"wrote {GOLDEN_ENC}: {os.path.getsize(GOLDEN_ENC) / 1024:.0f} KiB")


def te_configs():
    """Qwen3-style (QK-norm, right padding, final-norm layer included) and Mistral-style (no QK-norm, left padding) tiny encoders."""
    q = O.TEConfig(vocab_size=300, hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, num_kv_heads=1)
    m = O.TEConfig(vocab_size=300, hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, num_kv_heads=2, qk_norm=False,
                   rope_theta=1e9)
    toks = [7, 250, 31, 4, 199, 42, 42, 8, 120, 77, 5]
    return {"qwen3": (q, 4, "right", (1, 2, 3), toks), "mistral": (m, 5, "left", (0, 2), toks)}


def text_encoder_fixture():
    """Text-embedding producer: token ids -> concatenated hidden states, 32 positions."""
    torch.set_num_threads(1)
    out = {}
    for name, (cfg, seed, side, layers, toks) in te_configs().items():
        W = O.random_te_weights(cfg, seed=seed)
        ids, mask = O.te_pad_tokens(toks, 32, 3, side)
        out[f"{name}_ids"], out[f"{name}_mask"] = ids.numpy(), mask.numpy()
        out[f"{name}_hidden"] = O.te_hidden_states(W, cfg, ids, mask, layers).numpy()
    np.savez_compressed(GOLDEN_TE, **out)
    print(f"wrote {GOLDEN_TE}: {os.path.getsize(GOLDEN_TE) / 1024:.0f} KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        encoder_fixture()      # golden.npz untouched
    elif len(sys.argv) > 1 and sys.argv[1] == "text_encoder":
        text_encoder_fixture()
    else:
        main()
