#!/usr/bin/env python
"""Runs a handful of isolated kernels (for `ncu --set full`): python tools/prof_kernels.py gemm_cg1 gemm_cg2 attn3 ..."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-2-swift-mlx_b200"))
import torch  # noqa: E402
import flux2b  # noqa: E402


def main():
    what = sys.argv[1:] or ["gemm_cg1", "attn3"]
    ctx = flux2b.Context()
    g = torch.Generator().manual_seed(0)
    M, N, K = 4608, 3072, 3072
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
    w2 = (torch.randn(18432, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
    S, H = 4608, 24
    qkv = torch.randn(S, 3 * H * 128, generator=g).to(torch.bfloat16).cuda()
    for name in what:
        for _ in range(3):
            if name == "gemm_cg1":
                ctx.op_gemm(a, w, epilogue=0, cta_group=1)
            elif name == "gemm_cg2":
                ctx.op_gemm(a, w, epilogue=0, cta_group=2)
            elif name == "gemm_ff_cg1":
                ctx.op_gemm(a, w2, epilogue=3, cta_group=1)
            elif name == "gemm_ff_cg2":
                ctx.op_gemm(a, w2, epilogue=3, cta_group=2)
            elif name.startswith("wq_"):
                # W-only quantized GEMM with in-kernel dequantisation (random codes: timing only), Klein-9B width
                q = name[3:]
                Kq = Nq = 4096
                bits = 8 if q in ("qint8", "mxfp8") else 4
                group = 64 if q in ("qint8", "int4") else 16 if q == "nvfp4" else 32
                xq = torch.randn(M, Kq, generator=g).to(torch.bfloat16).cuda()
                wp = torch.randint(0, 2 ** 31 - 1, (Nq, Kq * bits // 32), generator=g, dtype=torch.int32).cuda()
                if q in ("qint8", "int4"):
                    sc = torch.full((Nq, Kq // group), 0.01, dtype=torch.float16).cuda()
                    bi = torch.full((Nq, Kq // group), -0.08, dtype=torch.float16).cuda()
                else:
                    sc = torch.full((Nq, Kq // group), 120 if q != "nvfp4" else 0x38, dtype=torch.uint8).cuda()
                    bi = None
                ctx.op_linear_quantized(q, xq, wp, sc, bi, in_kernel=True)
            elif name.startswith("attn"):
                ctx.op_attention(qkv, 1, S, H, variant=int(name[4:]))
        ctx.synchronize()


if __name__ == "__main__":
    main()
