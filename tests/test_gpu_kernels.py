"""GPU parity tests of the single kernels, through the C ABI (flux2b_op_*), against PyTorch fp32 / the oracle.

Tolerances (rel-L2 unless stated): 16-bit outputs carry one rounding of the result (bf16: 2^-9 per element ->
~1.5e-3 rel-L2 at most; f16: 2^-12), fp32 outputs only the tensor-core accumulation order. BASELINE.json north_star:
per-block activations rel-L2 <= 2e-3 for bf16; integer / packed-weight work bit-exact.
"""
import math

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL_BF16_OUT = 2e-3
TOL_F16_OUT = 3e-4
TOL_F32_OUT = 2e-5


@pytest.fixture(scope="module")
def ctx(flux2b):
    c = flux2b.Context()
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_f16(flux2b):
    c = flux2b.Context(options={"compute_f16": 1})
    yield c
    c.close()


def _gemm_case(ctx, M, N, K, epi, cg, bn=0, dt=torch.bfloat16, with_bias=False):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + epi)
    a = torch.randn(M, K, generator=g).to(dt).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dt).cuda()
    ref = a.double() @ w.double().t()
    bias = gate = res = None
    if with_bias:
        bias = torch.randn(N, generator=g).cuda()
        ref = ref + bias.double()[None]
    if epi == 2:
        gate = torch.randn(N, generator=g).cuda()
        res = torch.randn(M, N, generator=g).cuda()
        ref = res.double() + gate.double()[None] * ref
    if epi == 3:
        r = ref.reshape(M, N // 256, 2, 128)
        ref = (torch.nn.functional.silu(r[:, :, 0]) * r[:, :, 1]).reshape(M, N // 2)
    out = ctx.op_gemm(a, w, epilogue=epi, bias=bias, gate=gate, res=res, cta_group=cg, bn=bn)
    ctx.synchronize()
    return rel_l2(out, ref)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("M,N,K,epi", [
    (256, 256, 128, 1),       # exact tiles, fp32 out
    (300, 200, 192, 1),       # ragged M and N tails
    (1, 128, 64, 1),          # single row
    (129, 40, 8, 1),          # K smaller than one k-block (TMA zero fill), N < 64
    (512, 384, 256, 0),       # 16-bit out
    (640, 512, 512, 2),       # gate * acc + residual
    (384, 1024, 256, 3),      # SwiGLU epilogue
    (4608, 3072, 128, 1),     # xEmbedder shape of Klein-4B @1024^2 (K = 128)
    (4096, 128, 3072, 1),     # projOut shape (N = 128)
])
def test_gemm_epilogues(ctx, M, N, K, epi, cg):
    err = _gemm_case(ctx, M, N, K, epi, cg)
    assert err < (TOL_BF16_OUT if epi in (0, 3) else TOL_F32_OUT), err


@pytest.mark.parametrize("bn", [32, 64, 128, 256])
def test_gemm_every_tile_width(ctx, bn):
    assert _gemm_case(ctx, 520, 512, 320, 1, 1, bn=bn) < TOL_F32_OUT


def test_gemm_bias_and_f16(ctx, ctx_f16):
    assert _gemm_case(ctx, 260, 200, 96, 1, 1, with_bias=True) < TOL_F32_OUT
    assert _gemm_case(ctx_f16, 512, 384, 256, 0, 1, dt=torch.float16) < TOL_F16_OUT
    assert _gemm_case(ctx_f16, 512, 384, 256, 0, 2, dt=torch.float16) < TOL_F16_OUT


@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_klein4b_shapes(ctx, cg):
    # QKV / out / MLP shapes of Klein-4B at 1024^2 (S = 4608): SURVEY §7 step 2
    assert _gemm_case(ctx, 4608, 3072, 3072, 0, cg) < TOL_BF16_OUT
    assert _gemm_case(ctx, 4608, 3072, 12288, 2, cg) < TOL_F32_OUT
    assert _gemm_case(ctx, 4608, 18432, 3072, 3, cg) < TOL_BF16_OUT


def _attn_ref(qkv, B, S, H):
    D = H * 128
    q, k, v = (qkv.float().reshape(B, S, 3, H, 128)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    return torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * S, D)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("B,S,H", [(1, 256, 2), (2, 328, 2), (1, 77, 1), (1, 768, 3), (1, 4608, 4), (1, 64, 1), (1, 130, 2), (2, 1000, 1)])
def test_attention(ctx, B, S, H, variant):
    qkv = torch.randn(B * S, 3 * H * 128, generator=torch.Generator().manual_seed(S + H)).to(torch.bfloat16).cuda()
    out = ctx.op_attention(qkv, B, S, H, variant=variant)
    ctx.synchronize()
    ref = _attn_ref(qkv, B, S, H)
    e = rel_l2(out, ref)
    # the kernel's own arithmetic against fp32 SDPA rounded once to the bf16 output type: what is left is the bf16 rounding of P
    e_round = rel_l2(out, ref.to(torch.bfloat16).float())
    print(f"attention v{variant} B={B} S={S} H={H}: rel-L2 {e:.2e} (vs bf16-rounded reference {e_round:.2e})")
    # P is rounded to bf16 before the PV product and the result once more: two roundings of 2^-9 relative each
    assert e < 3.2e-3   # measured 2.3 - 2.6e-3 (profiles/r02_parity_margins.md)


def test_attention_f16_and_large_logits(ctx_f16):
    # scaled-up Q/K: a peaked softmax exercises the online-max correction path
    B, S, H = 1, 640, 2
    qkv = torch.randn(B * S, 3 * H * 128, generator=torch.Generator().manual_seed(9))
    qkv[:, :2 * H * 128] *= 4.0
    qkv = qkv.to(torch.float16).cuda()
    for variant in (1, 2, 3, 4):
        out = ctx_f16.op_attention(qkv, B, S, H, variant=variant)
        ctx_f16.synchronize()
        assert rel_l2(out, _attn_ref(qkv, B, S, H)) < 2e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,cg,res", [
    (1, 32, 32, 64, 64, 3, 1, False), (1, 40, 24, 96, 96, 3, 1, True), (2, 16, 16, 32, 32, 1, 1, False),
    (1, 32, 32, 96, 3, 3, 1, False), (1, 32, 32, 64, 64, 3, 2, False), (1, 9, 21, 32, 384, 3, 1, False),
    (2, 24, 24, 384, 192, 1, 2, False), (1, 128, 128, 192, 192, 3, 2, True),
    # halo-tile kernel (3x3, stride 1): 16 x 8-pixel tiles with ragged borders, Cin below / not a multiple of the 64-channel block
    # (tail MMAs), CTA pairs with an odd tile count, batch > 1, the small decoder's widths
    (1, 17, 9, 32, 96, 3, 0, False), (2, 33, 23, 96, 192, 3, 2, True), (1, 64, 64, 384, 384, 3, 2, False),
    (1, 48, 40, 192, 96, 3, 0, True), (3, 16, 8, 64, 64, 3, 2, False), (1, 100, 60, 96, 3, 3, 0, False),
])
def test_conv2d(ctx_f16, B, H, W, Cin, Cout, k, cg, res):
    g = torch.Generator().manual_seed(H * W + Cin)
    x = torch.randn(B, H, W, Cin, generator=g).half().cuda()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / math.sqrt(Cin * k * k)).half().cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    r = torch.randn(B, H, W, Cout, generator=g).half().cuda() if res else None
    out = ctx_f16.op_conv2d(x, w, bias, r, cta_group=cg)
    ctx_f16.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=k // 2)
    ref = ref.permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    assert rel_l2(out, ref) < TOL_F16_OUT


@pytest.mark.parametrize("rows,D", [(5, 256), (512, 3072), (300, 4096), (64, 6144)])
def test_ln_modulate(ctx, rows, D):
    from oracle import flux2_oracle as O
    g = torch.Generator().manual_seed(D)
    x = torch.randn(rows, D, generator=g) * 3 + 0.5
    shift, scale = torch.randn(D, generator=g), torch.randn(D, generator=g) * 0.2
    out = ctx.op_ln_modulate(x.cuda(), shift.cuda(), scale.cuda(), torch.bfloat16)
    ref = O.apply_modulation(O.layer_norm(x[None]), shift[None], scale[None])[0]
    assert rel_l2(out, ref) < TOL_BF16_OUT


def test_rope_table_and_timestep_embedding(ctx):
    from oracle import flux2_oracle as O
    ids = torch.cat([O.text_position_ids(512), O.image_position_ids(1024, 768), O.reference_position_ids([8], [8])])
    cos, sin = ctx.op_rope_table(ids.numpy())
    rc, rs = O.rope_embeddings(ids)
    # fp32 sin/cos of angles up to ~511: a few ulp of the argument
    assert np.abs(cos - rc.numpy()).max() < 2e-4 and np.abs(sin - rs.numpy()).max() < 2e-4
    t = np.array([0.0, 1.0, 0.5, 0.0313], dtype=np.float32)
    e = ctx.op_timestep_embedding(t)
    assert np.abs(e - O.timesteps_proj(torch.from_numpy(t) * 1000.0).numpy()).max() < 2e-4


def test_qk_norm_rope_unfused_kernel(ctx):
    from oracle import flux2_oracle as O
    rows, H = 200, 3
    D = H * 128
    g = torch.Generator().manual_seed(4)
    qkv = torch.randn(rows, 3 * D, generator=g).to(torch.bfloat16)
    nq, nk = 1 + 0.1 * torch.randn(128, generator=g), 1 + 0.1 * torch.randn(128, generator=g)
    cos, sin = O.rope_embeddings(O.image_position_ids(160, 320)[:rows])
    out = ctx.op_qk_norm_rope(qkv.clone().cuda(), D, nq.cuda(), nk.cuda(), cos.cuda(), sin.cuda()).float().cpu()
    x = qkv.float().reshape(1, rows, 3, H, 128)
    q = O.apply_rope(O.rms_norm(x[:, :, 0].permute(0, 2, 1, 3), nq), cos, sin).permute(0, 2, 1, 3).reshape(rows, D)
    k = O.apply_rope(O.rms_norm(x[:, :, 1].permute(0, 2, 1, 3), nk), cos, sin).permute(0, 2, 1, 3).reshape(rows, D)
    assert rel_l2(out[:, :D], q) < TOL_BF16_OUT and rel_l2(out[:, D:2 * D], k) < TOL_BF16_OUT
    assert torch.equal(out[:, 2 * D:], qkv.float()[:, 2 * D:])  # V untouched


@pytest.mark.parametrize("B,HW,C,G,silu", [(1, 64, 384, 32, True), (2, 1000, 96, 32, True), (1, 4096, 192, 32, False)])
def test_groupnorm_silu(ctx_f16, B, HW, C, G, silu):
    from oracle import flux2_oracle as O
    g = torch.Generator().manual_seed(C)
    side = int(math.isqrt(HW)) if int(math.isqrt(HW)) ** 2 == HW else None
    H, W = (side, side) if side else (HW // 8, 8)
    x = (torch.randn(B, H, W, C, generator=g) * 2 + 0.3).half()
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    out = ctx_f16.op_groupnorm_silu(x.cuda(), gamma.cuda(), beta.cuda(), G, 1e-6, silu)
    ref = O.group_norm_nhwc(x, gamma, beta, G, 1e-6)
    if silu:
        ref = torch.nn.functional.silu(ref)
    assert rel_l2(out, ref) < TOL_F16_OUT


# ------------------------------------------------------------------ quantizers: bit-exact against the C oracle
@pytest.mark.parametrize("name", ["qint8", "int4", "mxfp8", "mxfp4", "nvfp4"])
@pytest.mark.parametrize("src", ["f16", "f32", "bf16"])
def test_quantize_bit_exact(ctx, flux2b, name, src):
    from oracle import quant_oracle as Q
    q = flux2b.QUANT[name]
    g = torch.Generator().manual_seed(5)
    w = torch.randn(256, 512, generator=g) * 0.05
    w[0, :64] = 0                      # all-zero groups
    w[1, :64] = 0.03125                # constant group (wmax == wmin)
    w[2, 3] = 1000.0                   # outlier: saturating element conversion in neighbouring modes
    w[3, :16] = -w[3, :16].abs()       # all-negative group
    w[4, :64] *= 1e-6                  # tiny magnitudes (scale floor 1e-7 / E8M0 small exponents)
    if src == "f16":
        wn = w.half().numpy()
    elif src == "f32":
        wn = w.numpy()
    else:
        wn = w.to(torch.bfloat16).view(torch.uint16).numpy()
    wt = {"f16": w.half(), "f32": w, "bf16": w.to(torch.bfloat16)}[src]
    p0, s0, b0 = Q.quantize(q, wn)
    p1, s1, b1 = ctx.quantize_matrix(q, wt)
    assert np.array_equal(p0, p1)
    assert np.array_equal(s0.view(np.uint8), s1.view(np.uint8))
    if b0 is not None:
        assert np.array_equal(b0.view(np.uint16), b1.view(np.uint16))
    d0 = Q.dequantize(q, p0, s0, b0, 512)
    d1 = ctx.dequantize_matrix(q, p1, s1, b1, 512)
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))


@pytest.mark.parametrize("name", ["qint8", "int4", "mxfp8", "mxfp4", "nvfp4"])
def test_quantize_golden(ctx, flux2b, name, golden):
    # committed fixture produced by the C oracle (tools/make_golden.py): the device packer reproduces it bit for bit
    q = flux2b.QUANT[name]
    w = golden["quant_w_f16"]
    p1, s1, b1 = ctx.quantize_matrix(q, w)
    assert np.array_equal(p1, golden[f"quant_{name}_packed"])
    assert np.array_equal(s1.view(np.uint8), golden[f"quant_{name}_scales"].view(np.uint8))
    if b1 is not None:
        assert np.array_equal(b1.view(np.uint16), golden[f"quant_{name}_biases"].view(np.uint16))


# ------------------------------------------------------------------ scheduler math / latent plumbing: bit-exact
def test_euler_scale_noise_repaint_bit_exact(ctx):
    from oracle import flux2_oracle as O
    g = torch.Generator().manual_seed(11)
    n = 4096 * 128
    x, v, u, e, m = (torch.randn(n, generator=g) for _ in range(5))
    m = m.sigmoid()
    s0, s1 = 0.91796875, 0.75
    got = ctx.euler_step(x.clone().numpy(), v.numpy(), s0, s1)
    dt = np.float32(s1) - np.float32(s0)
    assert np.array_equal(got, (x + torch.tensor(dt) * v).numpy())
    got = ctx.euler_step(x.clone().cuda(), v.cuda(), s0, s1, pred_uncond=u.cuda(), cfg=3.5).cpu()
    want = x + torch.tensor(dt) * (u + 3.5 * (v - u))
    assert torch.allclose(got, want, rtol=0, atol=1e-6)
    assert np.array_equal(ctx.scale_noise(x.numpy(), e.numpy(), 0.0), x.numpy())       # Flux2CoreTests.swift:405-415
    assert np.allclose(ctx.scale_noise(x.numpy(), e.numpy(), 0.3), O.FlowMatchEulerScheduler.scale_noise(x, 0.3, e).numpy(), atol=1e-6)
    got = ctx.repaint_blend(x.clone().numpy(), v.numpy(), e.numpy(), m.numpy(), 0.4)
    assert np.allclose(got, O.repaint_blend(x, v, e, m, 0.4).numpy(), atol=1e-6)
    assert np.array_equal(ctx.repaint_blend(x.clone().numpy(), v.numpy(), e.numpy(), np.ones(n, np.float32), 0.4), x.numpy())


def test_latent_plumbing_bit_exact(ctx):
    from oracle import flux2_oracle as O
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 128, 24, 40, generator=g)
    seq = ctx.pack_patchified_to_sequence(x.numpy())
    assert np.array_equal(seq, O.pack_patchified_to_sequence(x).numpy())
    assert np.array_equal(ctx.unpack_sequence_to_patchified(seq, 24 * 16, 40 * 16), x.numpy())
    lat = ctx.unpatchify_latents(x.numpy())
    assert np.array_equal(lat, O.unpatchify_latents(x).numpy())
    assert np.array_equal(ctx.pack_latents_to_patchified(lat), x.numpy())
    mean, var = torch.randn(128, generator=g) * 0.1, 1 + 0.1 * torch.rand(128, generator=g)
    d = ctx.bn_latents(x.numpy(), mean.numpy(), var.numpy(), 1e-4, True)
    assert np.allclose(d, O.denormalize_latents_bn(x, mean, var).numpy(), rtol=2e-7, atol=1e-7)
    nrm = ctx.bn_latents(d, mean.numpy(), var.numpy(), 1e-4, False)
    assert np.allclose(nrm, x.numpy(), atol=2e-6)


def test_empty_inputs_are_noops(ctx):
    assert ctx.euler_step(np.zeros(0, np.float32), np.zeros(0, np.float32), 1.0, 0.5).size == 0
    a = torch.zeros(0, 64, dtype=torch.bfloat16).cuda()
    w = torch.zeros(32, 64, dtype=torch.bfloat16).cuda()
    assert ctx.op_gemm(a, w, epilogue=1).shape == (0, 32)
