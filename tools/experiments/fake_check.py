#!/usr/bin/env python
"""Timing-experiment sanity check (FLUX2B_GEMM_FAKE_HALF_B): are the activations of the faked forward still finite and of the usual
magnitude? (If they degenerate to NaN / Inf / zeros the power draw of the faked run says nothing about the real one.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "flux-2-swift-mlx_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, flux2b
from flux2b import configs
cfg = configs.klein_4b()
cfg.num_layers, cfg.num_single_layers = 2, 4
ctx = flux2b.Context(dit=cfg, options={"record_blocks": 1})
g = torch.Generator().manual_seed(0)
for k, shp in configs.dit_weight_manifest(cfg).items():
    ctx.set_tensor(k, ((torch.rand(shp, generator=g) * 2 - 1) / shp[1] ** 0.5).bfloat16())
ctx.finalize()
S_img, S_txt = 4096, 512
hidden = torch.randn(1, S_img, 128, generator=g).numpy()
enc = torch.randn(1, S_txt, cfg.joint_attention_dim, generator=g).numpy()
import oracle.flux2_oracle as O
img_ids, txt_ids = O.image_position_ids(1024, 1024).numpy(), O.text_position_ids(S_txt).numpy()
out = ctx.dit_forward(hidden, enc, np.array([0.7], np.float32), None, img_ids, txt_ids)
D = cfg.num_attention_heads * cfg.attention_head_dim
bl = [ctx.block_output(i, S_img + S_txt, D) for i in range(6)]
print("FAKE" if os.environ.get("FLUX2B_GEMM_FAKE_HALF_B") else "REAL", "out finite", bool(np.isfinite(out).all()), "std", float(out.std()),
      "block stds", [round(float(b.std()), 3) for b in bl], "block finite", [bool(np.isfinite(b).all()) for b in bl])
