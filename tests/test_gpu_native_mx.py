"""GPU parity tests of the OPTIONAL native block-scaled mode (option "native_mx": tcgen05.mma.kind::mxf8f6f4 / mxf4nvf4
.block_scale consuming MLX's packed mxfp8 / mxfp4 / nvfp4 weights and group scales directly, activations quantised on the fly).

The reference computes x_fp32 · dequant(W)^T, so this mode is NOT held to the 2e-3 activation tolerance of the W-only path; it
is checked against a checker-side emulation instead (oracle.block_activation_quant + quant_oracle.fake_quant_activation):
  * the GEMM itself is exact: C == dequant(Aq, SFA) · dequant(W)^T to fp32 accumulation error,
  * the on-the-fly fp4 activation quantiser is bit-identical to the oracle's weight packer on the same matrix,
  * a whole DiT forward tracks the emulation within TOL_EMU (residual differences: 16-bit rounding of the operand that is
    quantised can flip a 4-bit / 8-bit code that sits on a rounding boundary).
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_l2
from test_gpu_model import cosine, dit_inputs, make_ctx, tiny_cfg

pytestmark = pytest.mark.gpu

KINDS = ["mxfp8", "mxfp4", "nvfp4"]
TOL_GEMM = 1e-5                                            # vs exact fp64 product of the dequantised operands
TOL_EMU = {"mxfp8": 4e-3, "mxfp4": 5e-3, "nvfp4": 5e-3}     # per-block residual stream vs the emulation oracle, rel-L2 (measured 1.7e-3 .. 2.3e-3)
COS_WONLY = {"mxfp8": 0.9995, "mxfp4": 0.999, "nvfp4": 0.999}  # model output vs the reference's W-only arithmetic (reported mode)


@pytest.mark.parametrize("name", KINDS)
@pytest.mark.parametrize("shape", [(128, 128, 256), (300, 384, 512), (1024, 768, 1024), (64, 256, 3072)])
def test_block_scaled_gemm_exact(flux2b, name, shape):
    from oracle import quant_oracle as Q
    M, N, K = shape
    q = Q.QUANT[name]
    bits, group, _ = Q.params(q)
    ctx = flux2b.Context()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, K // 32, generator=g)).repeat_interleave(32, dim=1)
    a[0, :64] = 0.0          # an all-zero group: scale 0 / smallest, elements 0
    a[1, 5] = 3.0e4          # an outlier
    a = a.to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).half().numpy()
    packed, scales, _ = Q.quantize(q, w)
    for bn, cg in ((0, 0), (128, 1), (256, 1), (128, 2)):   # auto = 128-wide tiles on CTA pairs
        out, aq, sfa = ctx.op_gemm_mx(name, a.cuda(), packed, scales, return_quantized=True, bn=bn, cta_group=cg)
        ctx.synchronize()
        A_deq = Q.dequantize(q, np.ascontiguousarray(aq).view(np.uint32), sfa, None, K).astype(np.float64)
        W_deq = Q.dequantize(q, packed, scales, None, K).astype(np.float64)
        ref = torch.from_numpy(A_deq @ W_deq.T)
        assert torch.isfinite(out).all()
        assert rel_l2(out.cpu(), ref) < TOL_GEMM, (name, shape, bn, cg)
        # the activation quantiser against the checker
        want = Q.fake_quant_activation(q, a.float())
        assert np.array_equal(A_deq.astype(np.float32), want.numpy()), (name, shape)
        if bits == 4:
            p_ref, s_ref, _ = Q.quantize(q, a.view(torch.int16).numpy().view(np.uint16))
            assert np.array_equal(p_ref.view(np.uint8).reshape(M, -1), aq)
            assert np.array_equal(s_ref, sfa)
    ctx.close()


def _dequantized_weights(Q, q, W):
    Wd = {}
    for k, w in W.items():
        if w.dim() != 2:
            Wd[k] = w
            continue
        p0, s0, b0 = Q.quantize(q, w.half().numpy())
        Wd[k] = torch.from_numpy(Q.dequantize(q, p0, s0, b0, w.shape[1]))
    return Wd


@pytest.mark.parametrize("name", KINDS)
def test_dit_forward_native_block_scaled(flux2b, name):
    from oracle import flux2_oracle as O
    from oracle import quant_oracle as Q
    q = flux2b.QUANT[name]
    cfg = tiny_cfg(O, guidance=True, layers=(2, 2))
    W = O.random_dit_weights(cfg, seed=4, round_to=torch.float16)
    Wd = _dequantized_weights(Q, q, W)
    hidden, enc, t, gd, img_ids, txt_ids = dit_inputs(O, cfg, 64, 128)
    args = (hidden.numpy(), enc.numpy(), t.numpy(), gd.numpy(), img_ids.numpy(), txt_ids.numpy())

    ctx = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16, opts={"native_mx": 1})
    out = ctx.dit_forward(*args)
    # the exported tensors are still MLX's packed layout, bit for bit
    base = "singleTransformerBlocks.0.attn.toQkvMlp"
    p0, s0, _ = Q.quantize(q, W[base + ".weight"].half().numpy())
    assert np.array_equal(ctx.get_tensor(base + ".weight"), p0)
    assert np.array_equal(ctx.get_tensor(base + ".scales").view(np.uint8), s0)

    rec = []
    with O.block_activation_quant(lambda x: Q.fake_quant_activation(q, x)):
        ref = O.dit_forward(Wd, cfg, hidden, enc, t, gd, img_ids, txt_ids, record=rec)
    errs = [rel_l2(ctx.block_output(i, 192, cfg.inner_dim), r) for i, r in enumerate(rec)]
    ref_wonly = O.dit_forward(Wd, cfg, hidden, enc, t, gd, img_ids, txt_ids)
    print(f"{name} native: block rel-L2 vs emulation max {max(errs):.2e}, out {rel_l2(out, ref):.2e}; "
          f"vs W-only arithmetic: out rel-L2 {rel_l2(out, ref_wonly):.2e} cos {cosine(out, ref_wonly):.5f}")
    assert max(errs) < TOL_EMU[name], errs
    assert rel_l2(out, ref) < 2 * TOL_EMU[name]
    assert cosine(out, ref_wonly) >= COS_WONLY[name]

    # W-only context on the same weights: the two modes agree to activation-quantisation noise
    ctx_w = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16)
    out_w = ctx_w.dit_forward(*args)
    assert cosine(out, out_w) >= COS_WONLY[name]

    # pre-quantized hand-over (PrequantizedCheckpoint.swift:290-387) into a native context: same output bits
    ctx2 = flux2b.Context(dit=cfg, quant=q, options={"native_mx": 1})
    for k, w in W.items():
        if w.dim() != 2:
            ctx2.set_tensor(k, w)
            continue
        b = k[:-len(".weight")]
        for suffix in (".weight", ".scales"):
            ctx2.set_tensor(b + suffix, ctx.get_tensor(b + suffix))
    ctx2.finalize()
    out2 = ctx2.dit_forward(*args)
    assert np.array_equal(out, out2)
    # deterministic
    assert np.array_equal(out, ctx.dit_forward(*args))
    # the quantisation fused into LayerNorm + modulate and the SwiGLU epilogue produces the same bits as the separate pass
    ctx3 = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16, opts={"native_mx": 1, "mx_fuse_quant": 0})
    assert np.array_equal(out, ctx3.dit_forward(*args))
    # 256-wide N tiles (one accumulator stage, SwiGLU rows interleaved per 256) / single-CTA tiles: same arithmetic
    for o in ({"mx_bn": 256}, {"gemm_cta_group": 1}):
        ctx4 = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16, opts={"native_mx": 1, **o})
        assert rel_l2(ctx4.dit_forward(*args), out) < 1e-5, o
        ctx4.close()
    ctx.close(); ctx_w.close(); ctx2.close(); ctx3.close()


def test_native_mx_ragged_and_batch(flux2b):
    """sequence lengths that are not multiples of the 128-row scale-factor block, batch of 2"""
    from oracle import flux2_oracle as O
    from oracle import quant_oracle as Q
    q = flux2b.QUANT["nvfp4"]
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    W = O.random_dit_weights(cfg, seed=9, round_to=torch.float16)
    Wd = _dequantized_weights(Q, q, W)
    ctx = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16, opts={"native_mx": 1})
    S_img, S_txt = 36, 77
    hidden = torch.randn(2, S_img, 128, generator=torch.Generator().manual_seed(1))
    enc = torch.randn(2, S_txt, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(2))
    t = torch.tensor([0.7, 0.3])
    img_ids, txt_ids = O.image_position_ids(96, 96), O.text_position_ids(S_txt)
    out = ctx.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), None, img_ids.numpy(), txt_ids.numpy())
    with O.block_activation_quant(lambda x: Q.fake_quant_activation(q, x)):
        ref = O.dit_forward(Wd, cfg, hidden, enc, t, None, img_ids, txt_ids)
    print(f"ragged native nvfp4: out rel-L2 {rel_l2(out, ref):.2e}")
    assert rel_l2(out, ref) < 2 * TOL_EMU["nvfp4"]
    ctx.close()


def test_native_mx_needs_a_block_scaled_quantization(flux2b):
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    W = O.random_dit_weights(cfg, seed=4, round_to=torch.float16)
    for q in (0, flux2b.QUANT["qint8"]):
        ctx = flux2b.Context(dit=cfg, quant=q, options={"native_mx": 1})
        ctx.load_weights(W, dtype=torch.float16)
        with pytest.raises(flux2b.Flux2Error) as ei:
            ctx.finalize()
        assert ei.value.case == "invalidConfiguration"
        ctx.close()
