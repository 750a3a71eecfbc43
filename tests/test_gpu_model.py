"""GPU parity tests of the whole path through the C ABI: DiT forward (all variants), quantized linears, LoRA merge,
denoise loop + step hook, VAE decoder, generate — against the oracle on identical random-init weights and inputs, and
against the committed golden fixtures.

Tolerances (BASELINE.json north_star): quantized packing / scales bit-exact; per-block activations rel-L2 <= 2e-3
(bf16 compute; the residual stream is fp32 on the device, 16-bit only at GEMM / attention operands); final latents
cosine >= 0.999 after N steps.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

TOL_BLOCK = 2e-3      # per-block residual stream, rel-L2 (north_star)
TOL_OUT = 3e-3        # model output [S_img, 128]: one more bf16 GEMM operand rounding after the last block (measured 2.3 - 2.4e-3)
TOL_QUANT_BLOCK = 3e-3  # W-only quantized path: dequantized weights are rounded once more to the 16-bit operand type


def cosine(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float(torch.nn.functional.cosine_similarity(a, b, dim=0))


def tiny_cfg(O, guidance=True, layers=(2, 2), heads=2, joint=256):
    return O.DiTConfig(num_layers=layers[0], num_single_layers=layers[1], num_attention_heads=heads,
                       joint_attention_dim=joint, guidance_embeds=guidance)


def make_ctx(flux2b, cfg, W, quant=0, opts=None, dtype=torch.bfloat16, vae=None, VW=None):
    o = {"record_blocks": 1}
    o.update(opts or {})
    ctx = flux2b.Context(dit=cfg, vae=vae, quant=quant, options=o)
    if W is not None:
        ctx.load_weights(W, dtype=dtype)
    if VW is not None:
        ctx.load_weights(VW)
    ctx.finalize()
    return ctx


def dit_inputs(O, cfg, S_img, S_txt, seed=42):
    side = int(math.isqrt(S_img))
    hidden = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(seed))
    enc = torch.randn(1, S_txt, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(seed + 1))
    t = torch.tensor([0.7])
    gd = torch.tensor([4.0]) if cfg.guidance_embeds else None
    return hidden, enc, t, gd, O.image_position_ids(side * 16, side * 16), O.text_position_ids(S_txt)


def run_and_compare(flux2b, O, cfg, W, opts, S_img=64, S_txt=128, dtype=torch.bfloat16, tol_block=TOL_BLOCK, tol_out=TOL_OUT):
    ctx = make_ctx(flux2b, cfg, W, opts=opts, dtype=dtype)
    hidden, enc, t, gd, img_ids, txt_ids = dit_inputs(O, cfg, S_img, S_txt)
    out = ctx.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), gd.numpy() if gd is not None else None,
                          img_ids.numpy(), txt_ids.numpy())
    rec = []
    ref = O.dit_forward(W, cfg, hidden, enc, t, gd, img_ids, txt_ids, record=rec)
    errs = [rel_l2(ctx.block_output(i, S_txt + S_img, cfg.inner_dim), r) for i, r in enumerate(rec)]
    e_out = rel_l2(out, ref)
    print(f"opts={opts} block rel-L2 max {max(errs):.2e} out {e_out:.2e} cos {cosine(out, ref):.6f}")
    assert max(errs) < tol_block, errs
    assert e_out < tol_out, e_out
    assert cosine(out, ref) >= 0.999
    ctx.close()
    return out


# ------------------------------------------------------------------ DiT forward
@pytest.mark.parametrize("opts", [
    {"fuse_qk_rope": 0, "fuse_swiglu": 0, "attn_variant": 1, "gemm_cta_group": 1},
    {"fuse_qk_rope": 1, "fuse_swiglu": 1, "attn_variant": 1, "gemm_cta_group": 1},
    {"fuse_qk_rope": 1, "fuse_swiglu": 1, "attn_variant": 2, "gemm_cta_group": 2},
    {},
])
def test_dit_forward_tiny_vs_oracle(flux2b, opts):
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O)
    run_and_compare(flux2b, O, cfg, O.random_dit_weights(cfg, seed=0), opts)


@pytest.mark.parametrize("S_img,S_txt,quant", [(100, 256, "bf16"), (256, 512, "bf16"), (81, 256, "qint8")])
def test_dit_forward_grouped_streams(flux2b, S_img, S_txt, quant):
    """Double-stream blocks with text and image rows in one launch per operation (option group_streams, taken when S_txt is a
    multiple of 256): same bits as the per-stream launches, and the oracle's tolerance."""
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, layers=(2, 1))
    W = O.random_dit_weights(cfg, seed=5, round_to=torch.float16 if quant != "bf16" else torch.bfloat16)
    hidden, enc, t, gd, img_ids, txt_ids = dit_inputs(O, cfg, S_img, S_txt)
    img_ids = O.image_position_ids(16, 16 * S_img) if int(math.isqrt(S_img)) ** 2 != S_img else img_ids
    outs, launches = [], []
    for g in (1, 0):
        ctx = make_ctx(flux2b, cfg, W, quant=flux2b.QUANT[quant], opts={"group_streams": g},
                       dtype=torch.float16 if quant != "bf16" else torch.bfloat16)
        l0 = ctx.launch_count()
        outs.append(ctx.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), gd.numpy(), img_ids.numpy(), txt_ids.numpy()))
        launches.append(ctx.launch_count() - l0)
        if g == 1 and quant == "bf16":
            rec = []
            ref = O.dit_forward(W, cfg, hidden, enc, t, gd, img_ids, txt_ids, record=rec)
            errs = [rel_l2(ctx.block_output(i, S_txt + S_img, cfg.inner_dim), r) for i, r in enumerate(rec)]
            assert max(errs) < TOL_BLOCK and rel_l2(outs[0], ref) < TOL_OUT
        ctx.close()
    assert launches[0] == launches[1] - 6 * cfg.num_layers     # 7 instead of 13 launches per double block
    assert np.array_equal(outs[0], outs[1])


def test_dit_forward_f16_compute(flux2b):
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False)
    W = O.random_dit_weights(cfg, seed=0, round_to=torch.float16)
    run_and_compare(flux2b, O, cfg, W, {"compute_f16": 1}, dtype=torch.float16, tol_block=3e-4, tol_out=6e-4)


def test_dit_forward_ragged_sequence(flux2b):
    # sequence lengths that are not multiples of any tile (S_txt = 77, S_img = 9x9 = 81)
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1), heads=3)
    run_and_compare(flux2b, O, cfg, O.random_dit_weights(cfg, seed=3), {}, S_img=81, S_txt=77)


def test_dit_forward_golden(flux2b, golden):
    from oracle import flux2_oracle as O
    import make_golden as MG
    cfg, W, hidden, enc = MG.tiny_inputs()
    ctx = make_ctx(flux2b, cfg, W)
    out = ctx.dit_forward(golden["dit_hidden"], golden["dit_enc"], np.array([0.7], np.float32), np.array([4.0], np.float32),
                          O.image_position_ids(MG.HW, MG.HW).numpy(), O.text_position_ids(MG.S_TXT).numpy())
    for i in range(golden["dit_blocks"].shape[0]):
        assert rel_l2(ctx.block_output(i, MG.S_TXT + MG.S_IMG, cfg.inner_dim), golden["dit_blocks"][i]) < TOL_BLOCK
    assert rel_l2(out, golden["dit_out"]) < TOL_OUT
    # denoise fixture: 3 Euler steps
    x = golden["dit_hidden"].copy()
    ctx.denoise(x, golden["dit_enc"], [float(s) for s in golden["denoise_sigmas"]], MG.HW, MG.HW, guidance=4.0)
    assert cosine(x, golden["denoise_out"]) >= 0.999 and rel_l2(x, golden["denoise_out"]) < TOL_OUT
    ctx.close()


def test_dit_forward_batch_and_device_pointers(flux2b):
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    W = O.random_dit_weights(cfg, seed=1)
    ctx = make_ctx(flux2b, cfg, W)
    h0, enc0, t, _, img_ids, txt_ids = dit_inputs(O, cfg, 64, 64, seed=10)
    h1, enc1, _, _, _, _ = dit_inputs(O, cfg, 64, 64, seed=20)
    hidden, enc = torch.cat([h0, h1]), torch.cat([enc0, enc1])
    tt = torch.tensor([0.7, 0.3])
    # host pointers, B = 2, bf16 text embeddings
    out = ctx.dit_forward(hidden.numpy(), enc.to(torch.bfloat16), tt.numpy(), None, img_ids.numpy(), txt_ids.numpy())
    for b in range(2):
        ref = O.dit_forward(W, cfg, hidden[b:b + 1], enc[b:b + 1].to(torch.bfloat16).float(), tt[b:b + 1], None, img_ids, txt_ids)
        assert rel_l2(out[b], ref[0]) < TOL_OUT
    # device pointers give the same bits as host pointers
    out_d = ctx.dit_forward(hidden.cuda(), enc.to(torch.bfloat16).cuda(), tt.cuda(), None, img_ids.cuda(), txt_ids.cuda())
    ctx.synchronize()
    assert np.array_equal(out_d.cpu().numpy(), out)
    ctx.close()


def test_dit_kv_extract_and_cached(flux2b):
    # klein-9b-kv path (Flux2Transformer.swift:346-546): [txt | ref | img] with the ref->output block mask, then cached K/V
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 2))
    W = O.random_dit_weights(cfg, seed=2)
    ctx = make_ctx(flux2b, cfg, W)
    hidden, enc, t, _, img_ids, txt_ids = dit_inputs(O, cfg, 64, 64)
    ref_hidden = torch.randn(1, 36, 128, generator=torch.Generator().manual_seed(77))
    ref_ids = O.reference_position_ids([6], [6])
    out1 = ctx.dit_forward_kv_extract(hidden.numpy(), ref_hidden.numpy(), enc.numpy(), t.numpy(), None, img_ids.numpy(),
                                      ref_ids.numpy(), txt_ids.numpy())
    want1, cache = O.dit_forward(W, cfg, hidden, enc, t, None, img_ids, txt_ids, kv_mode=1, ref_hidden=ref_hidden, ref_ids=ref_ids)
    assert rel_l2(out1, want1) < TOL_OUT
    t2 = torch.tensor([0.4])
    out2 = ctx.dit_forward_kv_cached(hidden.numpy(), enc.numpy(), t2.numpy(), None, img_ids.numpy(), txt_ids.numpy())
    want2 = O.dit_forward(W, cfg, hidden, enc, t2, None, img_ids, txt_ids, kv_mode=2, kv_cache=cache)
    assert rel_l2(out2, want2) < TOL_OUT
    ctx.close()


def test_dit_errors_mirror_flux2error(flux2b):
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    W = O.random_dit_weights(cfg, seed=1)
    ctx = flux2b.Context(dit=cfg)
    hidden, enc, t, _, img_ids, txt_ids = dit_inputs(O, cfg, 16, 16)
    with pytest.raises(flux2b.Flux2Error) as e:      # forward before weights: modelNotLoaded
        ctx.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), None, img_ids.numpy(), txt_ids.numpy())
    assert e.value.case == "modelNotLoaded"
    Wm = dict(W)
    del Wm["transformerBlocks.0.attn.toK.weight"]
    ctx.load_weights(Wm, dtype=torch.bfloat16)
    with pytest.raises(flux2b.Flux2Error) as e:      # missing tensor: weightLoadingFailed, names the key
        ctx.finalize()
    assert e.value.case == "weightLoadingFailed" and "toK" in str(e.value)
    ctx.set_tensor("transformerBlocks.0.attn.toK.weight", torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(flux2b.Flux2Error) as e:      # wrong shape
        ctx.finalize()
    assert e.value.case == "weightLoadingFailed"
    with pytest.raises(flux2b.Flux2Error):
        ctx.set_option("no_such_option", 1)
    with pytest.raises(flux2b.Flux2Error) as e:      # cached forward without an extraction pass: generationFailed
        c2 = make_ctx(flux2b, cfg, W)
        c2.dit_forward_kv_cached(hidden.numpy(), enc.numpy(), t.numpy(), None, img_ids.numpy(), txt_ids.numpy())
    assert e.value.case == "generationFailed"
    ctx.close()


# ------------------------------------------------------------------ quantized linears
@pytest.mark.parametrize("name", ["qint8", "int4", "mxfp8", "mxfp4", "nvfp4"])
def test_quantized_dit_packing_bit_exact_and_forward(flux2b, name):
    from oracle import flux2_oracle as O
    from oracle import quant_oracle as Q
    q = flux2b.QUANT[name]
    bits, group, has_b, _ = flux2b.quant_params(q)
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    # the reference holds f16 weights when quantize(model:) runs (WeightLoader.swift:256-268)
    W = O.random_dit_weights(cfg, seed=4, round_to=torch.float16)
    ctx = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16)
    Wd = {}
    for k, w in W.items():
        if w.dim() != 2:
            Wd[k] = w
            continue
        base = k[:-len(".weight")]
        p0, s0, b0 = Q.quantize(q, w.half().numpy())
        # every Linear is quantized and exported in MLX's layout, bit for bit
        assert np.array_equal(ctx.get_tensor(base + ".weight"), p0), k
        assert np.array_equal(ctx.get_tensor(base + ".scales").view(np.uint8), s0.view(np.uint8)), k
        if has_b:
            assert np.array_equal(ctx.get_tensor(base + ".biases").view(np.uint16), b0.view(np.uint16)), k
        Wd[k] = torch.from_numpy(Q.dequantize(q, p0, s0, b0, w.shape[1]))
    hidden, enc, t, _, img_ids, txt_ids = dit_inputs(O, cfg, 64, 64)
    out = ctx.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), None, img_ids.numpy(), txt_ids.numpy())
    rec = []
    ref = O.dit_forward(Wd, cfg, hidden, enc, t, None, img_ids, txt_ids, record=rec)
    errs = [rel_l2(ctx.block_output(i, 128, cfg.inner_dim), r) for i, r in enumerate(rec)]
    print(f"{name}: block rel-L2 max {max(errs):.2e}, out {rel_l2(out, ref):.2e}")
    assert max(errs) < TOL_QUANT_BLOCK and rel_l2(out, ref) < 2 * TOL_QUANT_BLOCK
    # pre-quantized hand-over (PrequantizedCheckpoint.swift:290-387): same packed tensors in -> same output bits
    ctx2 = flux2b.Context(dit=cfg, quant=q, options={"record_blocks": 1})
    for k, w in W.items():
        if w.dim() != 2:
            ctx2.set_tensor(k, w)
            continue
        base = k[:-len(".weight")]
        for suffix in (".weight", ".scales") + ((".biases",) if has_b else ()):
            ctx2.set_tensor(base + suffix, ctx.get_tensor(base + suffix))
    ctx2.finalize()
    out2 = ctx2.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), None, img_ids.numpy(), txt_ids.numpy())
    assert np.array_equal(out, out2)
    ctx.close(); ctx2.close()


def test_lora_merge_dense_bit_exact(flux2b):
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    for dt, code_view in ((torch.float16, np.float16), (torch.bfloat16, None)):
        W = O.random_dit_weights(cfg, seed=5, round_to=dt)
        ctx = make_ctx(flux2b, cfg, W, dtype=dt)
        g = torch.Generator().manual_seed(6)
        key = "transformerBlocks.0.attn.toQ"
        A, B = torch.randn(16, 256, generator=g) * 0.02, torch.randn(256, 16, generator=g) * 0.02
        ctx.merge_lora(key, A, B, 0.75)
        got = ctx.get_tensor(key + ".weight")
        want = O.lora_merge(W[key + ".weight"], A, B, 0.75, dt)
        if dt == torch.float16:
            assert np.array_equal(got.view(np.uint16), want.numpy().view(np.uint16))
        else:
            assert np.array_equal(got, want.view(torch.uint16).numpy())
        # the merged weight is what the next forward uses
        hidden, enc, t, _, img_ids, txt_ids = dit_inputs(O, cfg, 64, 64)
        out = ctx.dit_forward(hidden.numpy(), enc.numpy(), t.numpy(), None, img_ids.numpy(), txt_ids.numpy())
        W2 = dict(W); W2[key + ".weight"] = want.float()
        # (activations are bf16 in both cases: only the stored weight dtype differs)
        assert rel_l2(out, O.dit_forward(W2, cfg, hidden, enc, t, None, img_ids, txt_ids)) < 4e-3   # merged weights: measured 3.25e-3 (the LoRA delta enlarges the output the rounding acts on)
        ctx.close()


@pytest.mark.parametrize("name", ["qint8", "nvfp4", "mxfp8"])
def test_lora_merge_quantized_bit_exact(flux2b, name):
    # dequantized -> f16, + s*(B A) in f16, quantized with the layer's own parameters (WeightLoader.swift:792-822)
    from oracle import flux2_oracle as O
    from oracle import quant_oracle as Q
    q = flux2b.QUANT[name]
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    W = O.random_dit_weights(cfg, seed=5, round_to=torch.float16)
    ctx = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16)
    key = "singleTransformerBlocks.0.attn.toOut"
    w = W[key + ".weight"]
    g = torch.Generator().manual_seed(6)
    A, B = torch.randn(8, w.shape[1], generator=g) * 0.02, torch.randn(w.shape[0], 8, generator=g) * 0.02
    ctx.merge_lora(key, A, B, 1.0)
    p0, s0, b0 = Q.quantize(q, w.half().numpy())
    deq = torch.from_numpy(Q.dequantize(q, p0, s0, b0, w.shape[1])).half()
    merged = O.lora_merge(deq, A, B, 1.0, torch.float16)
    p1, s1, b1 = Q.quantize(q, merged.numpy())
    assert np.array_equal(ctx.get_tensor(key + ".weight"), p1)
    assert np.array_equal(ctx.get_tensor(key + ".scales").view(np.uint8), s1.view(np.uint8))
    if b1 is not None:
        assert np.array_equal(ctx.get_tensor(key + ".biases").view(np.uint16), b1.view(np.uint16))
    ctx.close()


# ------------------------------------------------------------------ denoise loop + hook
def test_denoise_loop_variants(flux2b):
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 2))
    W = O.random_dit_weights(cfg, seed=7)
    ctx = make_ctx(flux2b, cfg, W)
    HW, S_img, S_txt = 128, 64, 64
    lat = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(42))
    enc = torch.randn(1, S_txt, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43))
    enc_u = torch.randn(1, S_txt, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(44))
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(4, S_img)
    # (1) plain T2I
    x = lat.numpy().copy()
    ctx.denoise(x, enc.numpy(), sched.sigmas, HW, HW)
    want = O.denoise(W, cfg, lat, enc, sched.sigmas, HW, HW)
    assert cosine(x, want) >= 0.999 and rel_l2(x, want) < TOL_OUT
    # (2) classical CFG (klein *Base*, Flux2Pipeline.swift:1958-1973)
    x = lat.numpy().copy()
    ctx.denoise(x, enc.numpy(), sched.sigmas, HW, HW, enc_uncond=enc_u.numpy(), cfg_scale=3.0)
    want = O.denoise(W, cfg, lat, enc, sched.sigmas, HW, HW, enc_uncond=enc_u, cfg_scale=3.0)
    assert cosine(x, want) >= 0.999 and rel_l2(x, want) < 1.5 * TOL_OUT
    # (3) I2I with reference tokens [output | refs] (:1696-1767)
    ref_lat = torch.randn(1, 36, 128, generator=torch.Generator().manual_seed(45))
    ref_ids = O.reference_position_ids([6], [6])
    x = lat.numpy().copy()
    ctx.denoise(x, enc.numpy(), sched.sigmas, HW, HW, ref_latents=ref_lat.numpy(), ref_ids=ref_ids.numpy())
    want = O.denoise(W, cfg, lat, enc, sched.sigmas, HW, HW, ref_latents=ref_lat, ref_ids=ref_ids)
    assert cosine(x, want) >= 0.999 and rel_l2(x, want) < TOL_OUT
    # (3b) the klein-9b-kv loop (:1565-1644): extract at step 0, cached K / V afterwards; is_i2i is reported to the hook
    seen_kv = []
    x = lat.numpy().copy()
    ctx.denoise(x, enc.numpy(), sched.sigmas, HW, HW, ref_latents=ref_lat.numpy(), ref_ids=ref_ids.numpy(), kv_cache=True,
                hook=lambda sc, v: seen_kv.append((sc.step_idx, sc.is_i2i)) and None)
    want = O.denoise(W, cfg, lat, enc, sched.sigmas, HW, HW, ref_latents=ref_lat, ref_ids=ref_ids, kv_cache=True)
    assert cosine(x, want) >= 0.999 and rel_l2(x, want) < TOL_OUT
    assert seen_kv == [(0, 1), (1, 1), (2, 1), (3, 1)]
    # ... and equals the manual composition of the public KV calls, bit for bit
    y = lat.numpy().copy()
    ids, tids = O.image_position_ids(HW, HW).numpy(), O.text_position_ids(S_txt).numpy()
    for i in range(4):
        ts = np.array([sched.sigmas[i]], np.float32)
        if i == 0:
            pred = ctx.dit_forward_kv_extract(y, ref_lat.numpy(), enc.numpy(), ts, None, ids, ref_ids.numpy(), tids)
        else:
            pred = ctx.dit_forward_kv_cached(y, enc.numpy(), ts, None, ids, tids)
        ctx.euler_step(y, pred, sched.sigmas[i], sched.sigmas[i + 1])
    assert np.array_equal(x, y)
    with pytest.raises(flux2b.Flux2Error):   # no classical-CFG branch in the KV loop
        ctx.denoise(lat.numpy().copy(), enc.numpy(), sched.sigmas, HW, HW, ref_latents=ref_lat.numpy(), ref_ids=ref_ids.numpy(),
                    kv_cache=True, enc_uncond=enc_u.numpy(), cfg_scale=2.0)
    # (4) Flux2StepHook: RePaint blend after every step incl. the last (sigma_next == 0), Flux2MaskedInpaintingChain.swift:399-403
    x0 = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(46))
    eps = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(47))
    mask = (torch.rand(1, S_img, 1, generator=torch.Generator().manual_seed(48)) > 0.5).float().expand(1, S_img, 128).contiguous()
    seen = []

    def hook(sc, view):
        seen.append((sc.step_idx, sc.total_steps, sc.sigma, sc.sigma_next, sc.height, sc.width, sc.is_i2i))
        cur = torch.from_numpy(view.copy()).reshape(1, S_img, 128)
        view[:] = O.repaint_blend(cur, x0, eps, mask, sc.sigma_next).reshape(-1).numpy()

    x = lat.numpy().copy()
    ctx.denoise(x, enc.numpy(), sched.sigmas, HW, HW, hook=hook)
    want = O.denoise(W, cfg, lat, enc, sched.sigmas, HW, HW,
                     hook=lambda i, n, s, sn, xx: O.repaint_blend(xx, x0, eps, mask, sn))
    assert [s[0] for s in seen] == [0, 1, 2, 3] and seen[-1][3] == 0.0 and seen[0][1] == 4 and seen[0][4:] == (HW, HW, 0)
    assert cosine(x, want) >= 0.999 and rel_l2(x, want) < TOL_OUT
    # (5) a hook that cancels -> Flux2Error.generationCancelled
    with pytest.raises(flux2b.Flux2Error) as e:
        ctx.denoise(lat.numpy().copy(), enc.numpy(), sched.sigmas, HW, HW, hook=lambda sc, v: 1)
    assert e.value.case == "generationCancelled"
    # (6) loop == manual composition of the public calls, bit for bit
    x = lat.numpy().copy()
    ctx.denoise(x, enc.numpy(), sched.sigmas, HW, HW)
    y = lat.numpy().copy()
    ids, tids = O.image_position_ids(HW, HW).numpy(), O.text_position_ids(S_txt).numpy()
    for i in range(4):
        pred = ctx.dit_forward(y, enc.numpy(), np.array([sched.sigmas[i]], np.float32), None, ids, tids)
        ctx.euler_step(y, pred, sched.sigmas[i], sched.sigmas[i + 1])
    assert np.array_equal(x, y)
    ctx.close()


# ------------------------------------------------------------------ VAE decoder
@pytest.mark.parametrize("small,h,w", [(True, 8, 8), (False, 8, 8), (True, 6, 10)])
def test_vae_decode_vs_oracle(flux2b, small, h, w):
    from oracle import flux2_oracle as O
    vcfg = O.vae_small_decoder() if small else O.VAEConfig()
    VW = O.random_vae_weights(vcfg, seed=1)
    ctx = flux2b.Context(vae=vcfg)
    ctx.load_weights(VW)
    ctx.finalize()
    z = torch.randn(1, 32, h, w, generator=torch.Generator().manual_seed(7))
    out = ctx.vae_decode(z.numpy())
    ref = O.vae_decode(VW, vcfg, z)
    assert out.shape == (1, 3, 8 * h, 8 * w)
    e = rel_l2(out, ref)
    print(f"vae small={small} {h}x{w}: rel-L2 {e:.2e}")
    assert e < 3.5e-3  # f16 activations through ~30 convolutions and 30 GroupNorms (measured <= 2.7e-3)
    u8 = ctx.vae_decode_u8(z.numpy())
    want = O.postprocess_vae_output(ref).numpy()
    d = np.abs(u8[0].astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 3 and d.mean() < 0.6
    ctx.close()


def test_vae_golden_and_batch(flux2b, golden):
    from oracle import flux2_oracle as O
    vcfg = O.vae_small_decoder()
    VW = O.random_vae_weights(vcfg, seed=1)
    ctx = flux2b.Context(vae=vcfg)
    ctx.load_weights(VW)
    ctx.finalize()
    out = ctx.vae_decode(golden["vae_z"])
    assert rel_l2(out, golden["vae_out"]) < 3.5e-3
    z2 = np.concatenate([golden["vae_z"], golden["vae_z"][:, :, ::-1].copy()])
    out2 = ctx.vae_decode(z2)
    assert np.array_equal(out2[0], out[0])
    assert rel_l2(out2[1], O.vae_decode(VW, vcfg, torch.from_numpy(z2[1:2]))[0]) < 3.5e-3
    with pytest.raises(flux2b.Flux2Error) as e:
        flux2b.Context(vae=vcfg).vae_decode(golden["vae_z"])
    assert e.value.case == "modelNotLoaded"
    ctx.close()


def test_generate_end_to_end(flux2b):
    # denoise -> unpack -> BN denorm (eps 1e-4) -> unpatchify -> decode -> uint8 (Flux2Pipeline.swift:1933-2098, 2425-2468)
    from oracle import flux2_oracle as O
    cfg = tiny_cfg(O, guidance=False, layers=(1, 2))
    vcfg = O.vae_small_decoder()
    W, VW = O.random_dit_weights(cfg, seed=0), O.random_vae_weights(vcfg, seed=1)
    ctx = make_ctx(flux2b, cfg, W, vae=vcfg, VW=VW)
    H, Wd, S_txt = 64, 96, 64
    S_img = (H // 16) * (Wd // 16)
    lat = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(42))
    enc = torch.randn(1, S_txt, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43))
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(2, S_img)
    x = lat.numpy().copy()
    rgb = ctx.generate(x, enc.numpy(), sched.sigmas, H, Wd)
    ref_lat = O.denoise(W, cfg, lat, enc, sched.sigmas, H, Wd)
    ref_rgb = O.postprocess_vae_output(O.latents_to_image(VW, vcfg, ref_lat, H, Wd)).numpy()
    assert rgb.shape == (H, Wd, 3) and cosine(x, ref_lat) >= 0.999
    d = np.abs(rgb.astype(np.int32) - ref_rgb.astype(np.int32))
    print(f"generate: image mean |delta| {d.mean():.3f}/255 max {d.max()}")
    assert d.mean() < 1.0
    ctx.close()


# ------------------------------------------------------------------ BASELINE.json configs[0]: Klein 4B, 1 Euler step at 256x256
def test_klein4b_256_one_euler_step_vs_oracle(flux2b):
    from oracle import flux2_oracle as O
    cfg = O.klein_4b()
    torch.set_num_threads(os.cpu_count() or 1)
    W = O.random_dit_weights(cfg, seed=0)
    ctx = make_ctx(flux2b, cfg, W, opts={"keep_raw_weights": 0})
    S_img, S_txt = 256, 512
    lat = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(42))
    enc = torch.randn(1, S_txt, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43))
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(4, S_img)
    sig = sched.sigmas[:2]
    ids, tids = O.image_position_ids(256, 256), O.text_position_ids(S_txt)
    pred = ctx.dit_forward(lat.numpy(), enc.numpy(), np.array([sig[0]], np.float32), None, ids.numpy(), tids.numpy())
    rec = []
    ref = O.dit_forward(W, cfg, lat, enc, torch.tensor([sig[0]]), None, ids, tids, record=rec)
    errs = [rel_l2(ctx.block_output(i, S_txt + S_img, cfg.inner_dim), r) for i, r in enumerate(rec)]
    print(f"klein4b@256: block rel-L2 max {max(errs):.2e} (first {errs[0]:.2e}, last {errs[-1]:.2e}), pred {rel_l2(pred, ref):.2e}")
    assert max(errs) < TOL_BLOCK
    x = lat.numpy().copy()
    ctx.euler_step(x, pred, sig[0], sig[1])
    want = lat + torch.tensor(np.float32(sig[1]) - np.float32(sig[0])) * ref
    assert cosine(x, want) >= 0.999 and rel_l2(x, want) < TOL_OUT
    ctx.close()
