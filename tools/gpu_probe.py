#!/usr/bin/env python
"""GPU bring-up probe: runs each kernel check in its OWN subprocess (a trapped kernel poisons the CUDA context) with a
timeout, prints one PASS/FAIL line per probe and writes gpurun_out/probe.json. Not a benchmark, not a test-suite
replacement — it exists so that one gpurun call can evaluate many hypotheses.

usage: python tools/gpu_probe.py [probe-name-substring ...]
"""
from __future__ import annotations

import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-2-swift-mlx_b200"))
sys.path.insert(0, ROOT)


def rel_l2(a, b):
    import torch
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def probe_gemm(M, N, K, epi, cg, bn=0, dtype="bf16"):
    import torch
    import flux2b
    dt = torch.bfloat16 if dtype == "bf16" else torch.float16
    ctx = flux2b.Context(options={"compute_f16": int(dtype == "f16")})
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(M, K, generator=g)).to(dt).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dt).cuda()
    ref = a.float() @ w.float().t()
    bias = gate = res = None
    if epi == 2:
        gate = torch.randn(N, generator=g).cuda()
        res = torch.randn(M, N, generator=g).cuda()
        ref = res + gate[None] * ref
    if epi == 3:
        # rows interleaved per 256-tile: [128 gate | 128 value]
        r = ref.reshape(M, N // 256, 2, 128)
        ref = torch.nn.functional.silu(r[:, :, 0]) * r[:, :, 1]
        ref = ref.reshape(M, N // 2)
    out = ctx.op_gemm(a, w, epilogue=epi, bias=bias, gate=gate, res=res, cta_group=cg, bn=bn)
    ctx.synchronize()
    err = rel_l2(out.float(), ref)
    tol = 1e-2 if epi in (0, 3) else 2e-3
    # timing
    ctx.prof_enable(True); ctx.prof_reset()
    for _ in range(5):
        ctx.op_gemm(a, w, epilogue=epi, bias=bias, gate=gate, res=res, cta_group=cg, bn=bn, out=out)
    p = ctx.prof_get(flux2b.PROF_GEMM)
    tf = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0
    return err < tol, {"rel_l2": err, "tflops": round(tf, 1), "ms": round(p["ms"] / 5, 4)}


def probe_attention(B, S, H, variant, dtype="bf16", poly=0):
    import torch
    import flux2b
    dt = torch.bfloat16 if dtype == "bf16" else torch.float16
    ctx = flux2b.Context(options={"compute_f16": int(dtype == "f16"), "attn_poly": poly})
    g = torch.Generator().manual_seed(2)
    D = H * 128
    qkv = torch.randn(B * S, 3 * D, generator=g).to(dt).cuda()
    out = ctx.op_attention(qkv, B, S, H, variant=variant)
    ctx.synchronize()
    q, k, v = (qkv.float().reshape(B, S, 3, H, 128)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * S, D)
    err = rel_l2(out.float(), ref)
    ctx.prof_enable(True); ctx.prof_reset()
    for _ in range(5):
        ctx.op_attention(qkv, B, S, H, variant=variant)
    p = ctx.prof_get(flux2b.PROF_ATTN)
    tf = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0
    return err < 1e-2, {"rel_l2": err, "tflops": round(tf, 1), "ms": round(p["ms"] / 5, 4)}


def probe_conv(B, H, W, Cin, Cout, k, cg, residual=False):
    import torch
    import flux2b
    ctx = flux2b.Context(options={"compute_f16": 1})
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, W, Cin, generator=g).half().cuda()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / math.sqrt(Cin * k * k)).half().cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).half().cuda() if residual else None
    out = ctx.op_conv2d(x, w, bias, res, cta_group=cg)
    ctx.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=k // 2)
    ref = ref.permute(0, 2, 3, 1)
    if residual:
        ref = ref + res.float()
    err = rel_l2(out.float(), ref)
    ctx.prof_enable(True); ctx.prof_reset()
    for _ in range(3):
        ctx.op_conv2d(x, w, bias, res, cta_group=cg)
    p = ctx.prof_get(flux2b.PROF_CONV)
    tf = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0
    return err < 3e-3, {"rel_l2": err, "tflops": round(tf, 1), "ms": round(p["ms"] / 3, 4)}


def probe_dit(name, fuse_qk=1, fuse_swiglu=1, attn_variant=0, cg=0, f16=0, S_img=256, S_txt=512):
    import torch
    import flux2b
    from oracle import flux2_oracle as O
    if name == "tiny":
        cfg = O.DiTConfig(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256, guidance_embeds=True)
    elif name == "tiny_nofuse_dims":  # Hm % 128 != 0 is impossible with heads*128*3; kept for symmetry
        cfg = O.DiTConfig(num_layers=1, num_single_layers=1, num_attention_heads=1, joint_attention_dim=64, guidance_embeds=False)
    else:
        cfg = O.klein_4b()
    rt = torch.float16 if f16 else torch.bfloat16
    W = O.random_dit_weights(cfg, seed=0, round_to=rt)
    ctx = flux2b.Context(dit=cfg, options={"fuse_qk_rope": fuse_qk, "fuse_swiglu": fuse_swiglu, "attn_variant": attn_variant,
                                           "gemm_cta_group": cg, "compute_f16": f16, "record_blocks": 1})
    ctx.load_weights(W, dtype=rt)
    ctx.finalize()
    g = torch.Generator().manual_seed(42)
    side = int(math.isqrt(S_img))
    hidden = torch.randn(1, S_img, 128, generator=g)
    enc = torch.randn(1, S_txt, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43))
    t = torch.tensor([0.7])
    gd = torch.tensor([4.0]) if cfg.guidance_embeds else None
    img_ids = O.image_position_ids(side * 16, side * 16)
    txt_ids = O.text_position_ids(S_txt)
    t0 = time.time()
    out = ctx.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), gd.numpy() if gd is not None else None,
                          img_ids.numpy(), txt_ids.numpy())
    t_gpu = time.time() - t0
    rec = []
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    ref = O.dit_forward(W, cfg, hidden, enc, t, gd, img_ids, txt_ids, record=rec)
    t_cpu = time.time() - t0
    errs = []
    S = S_txt + S_img
    for i, r in enumerate(rec):
        b = torch.from_numpy(ctx.block_output(i, S, cfg.inner_dim))
        errs.append(round(rel_l2(b, r), 5))
    err = rel_l2(torch.from_numpy(out), ref)
    cos = float(torch.nn.functional.cosine_similarity(torch.from_numpy(out).flatten().double(), ref.flatten().double(), dim=0))
    ok = err < 2e-2 and max(errs) < 1e-2
    return ok, {"rel_l2_out": err, "cos": cos, "block_rel_l2_max": max(errs), "block_rel_l2": errs[:4] + errs[-2:],
                "gpu_s_first_call": round(t_gpu, 3), "cpu_oracle_s": round(t_cpu, 2)}


def probe_gemm_mx(quant, M, N, K, iters=5, bn=0, cg=0):
    """native block-scaled GEMM (mxfp8 / mxfp4 / nvfp4): exact check against dequant(aq, sfa) @ dequant(W)^T in fp64; for the
    fp4 kinds the quantised activations must also be bit-identical to the oracle's weight packer run on the same matrix."""
    import numpy as np
    import torch
    import flux2b
    from oracle import quant_oracle as Q
    qi = Q.QUANT[quant]
    bits, group, _ = Q.params(qi)
    ctx = flux2b.Context()
    g = torch.Generator().manual_seed(11)
    a = torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, K // 32, generator=g)).repeat_interleave(32, dim=1)
    a = a.to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).half().numpy()
    packed, scales, _ = Q.quantize(qi, w)
    out, aq, sfa = ctx.op_gemm_mx(quant, a.cuda(), packed, scales, return_quantized=True, bn=bn, cta_group=cg)
    ctx.synchronize()
    A_deq = Q.dequantize(qi, np.ascontiguousarray(aq).view(np.uint32), sfa, None, K).astype(np.float64)
    W_deq = Q.dequantize(qi, packed, scales, None, K).astype(np.float64)
    ref = A_deq @ W_deq.T
    err = rel_l2(out.cpu(), torch.from_numpy(ref))
    qerr = rel_l2(torch.from_numpy(A_deq), a.double())
    info = {"rel_l2_vs_exact_emulation": err, "activation_quant_rel_err": qerr}
    ok = err < 1e-5 and qerr < (0.05 if bits == 8 else 0.2)
    if bits == 4:
        a_bits = a.view(torch.int16).numpy().view(np.uint16)   # raw bf16 bits
        p_ref, s_ref, _ = Q.quantize(qi, a_bits)
        info["act_packed_bit_exact"] = bool(np.array_equal(p_ref.view(np.uint8).reshape(M, -1), aq))
        info["act_scales_bit_exact"] = bool(np.array_equal(s_ref, sfa))
        ok = ok and info["act_packed_bit_exact"] and info["act_scales_bit_exact"]
    ctx.prof_enable(True); ctx.prof_reset()
    ad = a.cuda()
    for _ in range(iters):
        ctx.op_gemm_mx(quant, ad, packed, scales, bn=bn, cta_group=cg)
    p = ctx.prof_get(flux2b.PROF_GEMM)
    e = ctx.prof_get(flux2b.PROF_ELEMWISE)
    tf = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0
    info.update({"tflops": round(tf, 1), "ms": round(p["ms"] / iters, 4), "act_quant_ms": round(e["ms"] / iters, 4),
                 "act_quant_gbs": round(e["bytes"] / (e["ms"] * 1e-3) / 1e9, 1) if e["ms"] > 0 else None})
    return ok, info


def probe_quant(quant):
    import numpy as np
    import torch
    import flux2b
    from oracle import quant_oracle as Q
    ctx = flux2b.Context()
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(256, 512, generator=g) * 0.05).to(torch.bfloat16).to(torch.float16).numpy()
    w[0, :64] = 0  # an all-zero group
    p0, s0, b0 = Q.quantize(quant, w)
    p1, s1, b1 = ctx.quantize_matrix(quant, w)
    ok = np.array_equal(p0, p1) and np.array_equal(s0.view(np.uint8), s1.view(np.uint8))
    if b0 is not None:
        ok = ok and np.array_equal(b0.view(np.uint16), b1.view(np.uint16))
    d0 = Q.dequantize(quant, p0, s0, b0, 512)
    d1 = ctx.dequantize_matrix(quant, p1, s1, b1, 512)
    ok = ok and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    return ok, {"packed_mismatch": int((p0 != p1).sum()), "dequant_mismatch": int((d0 != d1).sum()),
                "quant_rel_err": float(np.linalg.norm(d0 - w.astype(np.float32)) / np.linalg.norm(w.astype(np.float32)))}


def probe_vae(small=True, hw=8):
    import torch
    import flux2b
    from oracle import flux2_oracle as O
    vcfg = O.vae_small_decoder() if small else O.VAEConfig()
    W = O.random_vae_weights(vcfg, seed=1)
    ctx = flux2b.Context(vae=vcfg)
    ctx.load_weights(W)
    ctx.finalize()
    z = torch.randn(1, 32, hw, hw, generator=torch.Generator().manual_seed(7))
    out = ctx.vae_decode(z.numpy())
    ref = O.vae_decode(W, vcfg, z)
    err = rel_l2(torch.from_numpy(out), ref)
    return err < 2e-2, {"rel_l2": err}


PROBES = {
    "gemm_small_cg1": lambda: probe_gemm(256, 256, 128, 1, 1),
    "gemm_tail_cg1": lambda: probe_gemm(300, 200, 192, 1, 1),
    "gemm_bf16out_cg1": lambda: probe_gemm(512, 384, 256, 0, 1),
    "gemm_gate_res_cg1": lambda: probe_gemm(640, 512, 512, 2, 1),
    "gemm_swiglu_cg1": lambda: probe_gemm(384, 1024, 256, 3, 1),
    "gemm_f16_cg1": lambda: probe_gemm(512, 384, 256, 0, 1, dtype="f16"),
    "gemm_bn128_cg1": lambda: probe_gemm(512, 128, 3072, 1, 1),
    "gemm_small_cg2": lambda: probe_gemm(256, 256, 128, 1, 2),
    "gemm_tail_cg2": lambda: probe_gemm(300, 200, 192, 1, 2),
    "gemm_m512_out": lambda: probe_gemm(512, 3072, 3072, 2, 0),
    "gemm_m512_out_bn256": lambda: probe_gemm(512, 3072, 3072, 2, 2, bn=256),
    "gemm_m512_out_bn64": lambda: probe_gemm(512, 3072, 3072, 2, 2, bn=64),
    "gemm_m512_out_cg1_bn128": lambda: probe_gemm(512, 3072, 3072, 2, 1, bn=128),
    "gemm_m512_down": lambda: probe_gemm(512, 2560, 9216, 2, 0),
    "gemm_m512_down_bn64": lambda: probe_gemm(512, 2560, 9216, 2, 2, bn=64),
    "gemm_m512_down_cg1_bn64": lambda: probe_gemm(512, 2560, 9216, 2, 1, bn=64),
    "gemm_m512_qkv": lambda: probe_gemm(512, 6144, 2560, 0, 0),
    "gemm_m512_qkv_bn128": lambda: probe_gemm(512, 6144, 2560, 0, 2, bn=128),
    "gemm_big_cg1": lambda: probe_gemm(4608, 3072, 3072, 0, 1),
    "gemm_big_cg2": lambda: probe_gemm(4608, 3072, 3072, 0, 2),
    "gemm_ffin_cg1": lambda: probe_gemm(4608, 18432, 3072, 3, 1),
    "gemm_ffin_cg2": lambda: probe_gemm(4608, 18432, 3072, 3, 2),
    "gemm_out_cg1": lambda: probe_gemm(4608, 3072, 12288, 2, 1),
    "gemm_out_cg2": lambda: probe_gemm(4608, 3072, 12288, 2, 2),
    "mx8_small": lambda: probe_gemm_mx("mxfp8", 128, 128, 128),
    "mx8_k512": lambda: probe_gemm_mx("mxfp8", 256, 256, 512),
    "mx8_tail": lambda: probe_gemm_mx("mxfp8", 300, 384, 256),
    "mx8_big": lambda: probe_gemm_mx("mxfp8", 4608, 3072, 3072, iters=3),
    "mx8_ffin": lambda: probe_gemm_mx("mxfp8", 4608, 18432, 3072, iters=2),
    "mx8_ffin_bn128": lambda: probe_gemm_mx("mxfp8", 4608, 18432, 3072, iters=2, bn=128, cg=1),
    "nv4_small": lambda: probe_gemm_mx("nvfp4", 128, 128, 256),
    "nv4_k1024": lambda: probe_gemm_mx("nvfp4", 256, 256, 1024),
    "nv4_tail": lambda: probe_gemm_mx("nvfp4", 300, 384, 512),
    "nv4_big": lambda: probe_gemm_mx("nvfp4", 4608, 4096, 4096, iters=3),
    "nv4_ffin": lambda: probe_gemm_mx("nvfp4", 4608, 24576, 4096, iters=2),
    "nv4_ffin_bn128": lambda: probe_gemm_mx("nvfp4", 4608, 24576, 4096, iters=2, bn=128, cg=1),
    "nv4_ffin_cg2": lambda: probe_gemm_mx("nvfp4", 4608, 24576, 4096, iters=2, bn=128, cg=2),
    "nv4_small_cg2": lambda: probe_gemm_mx("nvfp4", 256, 128, 256, cg=2),
    "nv4_tail_cg2": lambda: probe_gemm_mx("nvfp4", 300, 384, 512, cg=2),
    "mx8_ffin_cg2": lambda: probe_gemm_mx("mxfp8", 4608, 18432, 3072, iters=2, bn=128, cg=2),
    "mx4_ffin_cg2": lambda: probe_gemm_mx("mxfp4", 4608, 24576, 4096, iters=2, bn=128, cg=2),
    "mx4_small": lambda: probe_gemm_mx("mxfp4", 128, 128, 256),
    "mx4_k1024": lambda: probe_gemm_mx("mxfp4", 256, 256, 1024),
    "mx4_tail": lambda: probe_gemm_mx("mxfp4", 300, 384, 512),
    "mx4_ffin": lambda: probe_gemm_mx("mxfp4", 4608, 24576, 4096, iters=2),
    "attn_poly0_big": lambda: probe_attention(1, 4608, 24, 3, poly=-1),
    "attn_poly4_big": lambda: probe_attention(1, 4608, 24, 3, poly=4),
    "attn_poly3_big": lambda: probe_attention(1, 4608, 24, 3, poly=3),
    "attn_poly2_big": lambda: probe_attention(1, 4608, 24, 3, poly=2),
    "attn_poly3_tail": lambda: probe_attention(2, 328, 2, 3, poly=3),
    "attn_poly2_f16": lambda: probe_attention(1, 1024, 4, 3, dtype="f16", poly=2),
    "attn_v1_small": lambda: probe_attention(1, 256, 2, 1),
    "attn_v2_small": lambda: probe_attention(1, 256, 2, 2),
    "attn_v1_tail": lambda: probe_attention(2, 328, 2, 1),
    "attn_v2_tail": lambda: probe_attention(2, 328, 2, 2),
    "attn_v1_big": lambda: probe_attention(1, 4608, 24, 1),
    "attn_v2_big": lambda: probe_attention(1, 4608, 24, 2),
    "attn_v3_big": lambda: probe_attention(1, 4608, 24, 3),
    "attn_v3_small": lambda: probe_attention(1, 256, 2, 3),
    "attn_v3_tail": lambda: probe_attention(2, 328, 2, 3),
    "attn_v4_big": lambda: probe_attention(1, 4608, 24, 4),
    "attn_v4_big_p4": lambda: probe_attention(1, 4608, 24, 4, poly=4),
    "attn_v4_big_p2": lambda: probe_attention(1, 4608, 24, 4, poly=2),
    "attn_v4_big_p0": lambda: probe_attention(1, 4608, 24, 4, poly=-1),
    "attn_v4_small": lambda: probe_attention(1, 256, 2, 4),
    "attn_v4_tail": lambda: probe_attention(2, 328, 2, 4),
    "attn_v4_dev16k": lambda: probe_attention(1, 16896, 6, 4),
    "attn_v3_k9": lambda: probe_attention(1, 4608, 32, 3),
    "attn_v3_dev16k": lambda: probe_attention(1, 16896, 6, 3),
    "conv3_cg1": lambda: probe_conv(1, 32, 32, 64, 64, 3, 1),
    "conv3_c96_cg1": lambda: probe_conv(1, 40, 24, 96, 96, 3, 1, residual=True),
    "conv1_cg1": lambda: probe_conv(2, 16, 16, 32, 32, 1, 1),
    "conv3_out3_cg1": lambda: probe_conv(1, 32, 32, 96, 3, 3, 1),
    "conv3_cg2": lambda: probe_conv(1, 32, 32, 64, 64, 3, 2),
    "conv3_big_cg1": lambda: probe_conv(1, 512, 512, 192, 192, 3, 1),
    "conv3_big_cg2": lambda: probe_conv(1, 512, 512, 192, 192, 3, 2),
    # the small decoder's layers at 1024^2 output (FLUX2B_CONV_HALO=0 in the environment: the per-tap kernel)
    "vaeconv_1024_96": lambda: probe_conv(1, 1024, 1024, 96, 96, 3, 0),
    "vaeconv_1024_192_96": lambda: probe_conv(1, 1024, 1024, 192, 96, 3, 0),
    "vaeconv_1024_96_3": lambda: probe_conv(1, 1024, 1024, 96, 3, 3, 0),
    "vaeconv_512_192": lambda: probe_conv(1, 512, 512, 192, 192, 3, 0),
    "vaeconv_512_384_192": lambda: probe_conv(1, 512, 512, 384, 192, 3, 0),
    "vaeconv_256_384": lambda: probe_conv(1, 256, 256, 384, 384, 3, 0),
    "vaeconv_128_384": lambda: probe_conv(1, 128, 128, 384, 384, 3, 0),
    "quant_qint8": lambda: probe_quant(1),
    "quant_int4": lambda: probe_quant(2),
    "quant_mxfp8": lambda: probe_quant(3),
    "quant_mxfp4": lambda: probe_quant(4),
    "quant_nvfp4": lambda: probe_quant(5),
    "dit_tiny_unfused": lambda: probe_dit("tiny", fuse_qk=0, fuse_swiglu=0, attn_variant=1, cg=1, S_img=64, S_txt=128),
    "dit_tiny_fused": lambda: probe_dit("tiny", fuse_qk=1, fuse_swiglu=1, attn_variant=1, cg=1, S_img=64, S_txt=128),
    "dit_tiny_fused_v2": lambda: probe_dit("tiny", fuse_qk=1, fuse_swiglu=1, attn_variant=2, cg=1, S_img=64, S_txt=128),
    "dit_tiny_f16": lambda: probe_dit("tiny", f16=1, attn_variant=1, cg=1, S_img=64, S_txt=128),
    "vae_small_8": lambda: probe_vae(True, 8),
    "vae_std_8": lambda: probe_vae(False, 8),
    "dit_klein4b_256": lambda: probe_dit("klein4b", attn_variant=1, cg=1),
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        name = sys.argv[2]
        try:
            ok, info = PROBES[name]()
        except Exception as e:  # noqa
            ok, info = False, {"exception": repr(e)[:500]}
        print("PROBE_RESULT " + json.dumps({"name": name, "ok": bool(ok), **info}))
        return
    sel = sys.argv[1:]
    names = [n for n in PROBES if not sel or any(s in n for s in sel)]
    results = []
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--run", n], capture_output=True, text=True, timeout=420)
            line = [l for l in r.stdout.splitlines() if l.startswith("PROBE_RESULT ")]
            if line:
                res = json.loads(line[-1][len("PROBE_RESULT "):])
            else:
                res = {"name": n, "ok": False, "rc": r.returncode, "stderr": r.stderr[-600:], "stdout": r.stdout[-300:]}
        except subprocess.TimeoutExpired:
            res = {"name": n, "ok": False, "timeout": True}
        res["wall_s"] = round(time.time() - t0, 1)
        results.append(res)
        print(("PASS " if res.get("ok") else "FAIL ") + json.dumps(res), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
            json.dump(results, f, indent=1)
    print(f"{sum(1 for r in results if r.get('ok'))}/{len(results)} probes passed")


if __name__ == "__main__":
    main()
