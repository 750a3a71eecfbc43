#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): one DiT forward whose double block runs grouped
(both streams in one launch per operation: the two-problem GEMM, the CTA-pair tiles whose cross-CTA barriers use CTA-scope
release / acquire, ptx.cuh:52-59), one single block, the flash attention, Euler, and a tiny VAE decode. Prints SANITIZE_OK."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-2-swift-mlx_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import flux2b  # noqa: E402
from flux2b import configs  # noqa: E402


def main():
    cfg = configs.Flux2TransformerConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, guidance_embeds=False)
    vcfg = configs.vae_small_decoder()
    ctx = flux2b.Context(dit=cfg, vae=vcfg)
    g = torch.Generator().manual_seed(0)
    for k, shp in configs.dit_weight_manifest(cfg).items():
        ctx.set_tensor(k, ((torch.rand(shp, generator=g) * 2 - 1) / shp[1] ** 0.5).bfloat16())
    for k, shp in configs.vae_weight_manifest(vcfg).items():
        if k.endswith("runningVar"):
            t = 1.0 + 0.1 * torch.rand(shp, generator=g)
        elif len(shp) == 1:
            t = (1.0 if (k.endswith(".weight") and ("orm" in k)) else 0.0) + 0.05 * torch.randn(shp, generator=g)
        else:
            fan = int(np.prod(shp[1:]))
            t = ((torch.rand(shp, generator=g) * 2 - 1) / fan ** 0.5).half().float()
        ctx.set_tensor(k, t)
    ctx.finalize()
    H = W = 256   # 16 x 16 = 256 image tokens, 256 text tokens: the grouped (two-problem) launches are taken
    S_img, S_txt = 256, 256
    lat = torch.randn(1, S_img, 128, generator=g).numpy()
    enc = torch.randn(1, S_txt, 256, generator=g).numpy()
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(1, S_img)
    rgb = ctx.generate(lat, enc, sched.sigmas, H, W)
    assert np.isfinite(lat).all() and rgb.shape == (H, W, 3)
    n1 = ctx.launch_count()
    ctx.close()
    # W-only int4 weights dequantized inside the GEMM kernels: generic-proxy stores of the dequant warps into the B stage, read
    # by the tensor core's async proxy (fence.proxy.async + mbarrier hand-off), and the staged variant's streaming kernel
    n2 = 0
    for ink in (1, 2):
        qctx = flux2b.Context(dit=cfg, quant=flux2b.QUANT["int4"], options={"wq_inkernel": ink})
        g2 = torch.Generator().manual_seed(1)
        for k, shp in configs.dit_weight_manifest(cfg).items():
            qctx.set_tensor(k, ((torch.rand(shp, generator=g2) * 2 - 1) / shp[1] ** 0.5).half())
        qctx.finalize()
        S_img2 = 1024 if ink == 2 else 256
        lat2 = torch.randn(1, S_img2, 128, generator=g2).numpy()
        sched.set_timesteps(1, S_img2)
        qctx.denoise(lat2, enc, sched.sigmas, 16 * int(S_img2 ** 0.5), 16 * int(S_img2 ** 0.5))
        assert np.isfinite(lat2).all()
        n2 += qctx.launch_count()
        qctx.close()
    print("SANITIZE_OK launches", n1, n2)


if __name__ == "__main__":
    main()
