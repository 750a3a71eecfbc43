"""GPU tests of the pre-quantized checkpoint row through the C ABI: Flux2PrequantizedCheckpoint.save / load
(Loading/PrequantizedCheckpoint.swift:225-387) — export, validate-before-touch load, bit-identical weights and outputs."""
import numpy as np
import pytest
import torch

from test_gpu_model import dit_inputs, make_ctx, tiny_cfg
from test_safetensors_cpu import meta, write_safetensors

pytestmark = pytest.mark.gpu


def _args(O, cfg):
    hidden, enc, t, gd, img_ids, txt_ids = dit_inputs(O, cfg, 64, 64)
    return (hidden.numpy(), enc.numpy(), t.numpy(), gd.numpy() if gd is not None else None, img_ids.numpy(), txt_ids.numpy())


@pytest.mark.parametrize("name,native", [("int4", 0), ("qint8", 0), ("mxfp8", 0), ("nvfp4", 0), ("nvfp4", 1)])
def test_export_and_reload_bit_identical(flux2b, tmp_path, name, native):
    from oracle import flux2_oracle as O
    q = flux2b.QUANT[name]
    _, _, has_b, _ = flux2b.quant_params(q)
    cfg = tiny_cfg(O, guidance=True, layers=(1, 2))
    W = O.random_dit_weights(cfg, seed=4, round_to=torch.float16)
    ctx = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16, opts={"native_mx": native})
    out = ctx.dit_forward(*_args(O, cfg))
    p = str(tmp_path / "mlx-prequantized" / name)
    import os
    os.makedirs(p)
    p = os.path.join(p, "transformer.safetensors")       # <source>/mlx-prequantized/<quant>/<component>.safetensors (:50-60)
    ctx.save_prequantized(p, "tiny-model", "w.safetensors:1:2")
    assert flux2b.prequantized_is_valid(p, name, "tiny-model", "w.safetensors:1:2")
    assert not os.path.exists(os.path.join(os.path.dirname(p), ".tmp-transformer.safetensors"))
    # writer pinned against the format's reference implementation: the official safetensors library opens the file, sees the
    # metadata of PrequantizedCheckpoint.swift:177-185 and the same bytes get_tensor returns
    try:
        from safetensors import safe_open
    except ImportError:
        safe_open = None
    if safe_open is not None:
        with safe_open(p, framework="np") as f:
            md = f.metadata()
            assert md["format"] == "flux2-mlx-prequantized-v1" and md["quantization"] == name and md["source"] == "tiny-model"
            assert md["source_fingerprint"] == "w.safetensors:1:2" and md["component"] == "transformer"
            keys = set(f.keys())
            k0 = next(k for k, w in W.items() if w.dim() == 2)
            base0 = k0[:-len(".weight")]
            assert {base0 + ".weight", base0 + ".scales"} <= keys
            a = f.get_tensor(base0 + ".weight")
            assert a.dtype == np.uint32 and np.array_equal(a, ctx.get_tensor(base0 + ".weight"))

    ctx2 = flux2b.Context(dit=cfg, quant=q, options={"record_blocks": 1, "native_mx": native})
    assert ctx2.load_prequantized(p, "tiny-model", "w.safetensors:1:2")
    ctx2.finalize()
    for k, w in W.items():
        if w.dim() != 2:
            assert np.array_equal(ctx2.get_tensor(k), ctx.get_tensor(k))
            continue
        base = k[:-len(".weight")]
        for suffix in (".weight", ".scales") + ((".biases",) if has_b else ()):
            a, b = ctx.get_tensor(base + suffix), ctx2.get_tensor(base + suffix)
            assert a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8)), base + suffix
    assert np.array_equal(out, ctx2.dit_forward(*_args(O, cfg)))
    ctx.close(); ctx2.close()


def test_load_validates_before_touching_the_context(flux2b, tmp_path):
    from oracle import flux2_oracle as O
    q = flux2b.QUANT["mxfp4"]
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    W = O.random_dit_weights(cfg, seed=4, round_to=torch.float16)
    ctx = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16)
    p = str(tmp_path / "transformer.safetensors")
    ctx.save_prequantized(p, "tiny-model", "fp", lora_baked=True)
    blob = open(p, "rb").read()

    def fresh(quant=q, cfg_=cfg):
        return flux2b.Context(dit=cfg_, quant=quant)

    # a LoRA-baked export loads, with a loud warning
    c = fresh()
    assert c.load_prequantized(p, "tiny-model", "fp") and "BAKED IN" in flux2b.last_error()
    c.close()
    # wrong source / stale fingerprint / other quantization / other architecture: not applied, nothing handed over
    for kw, why in ((dict(source_name="other-model"), "source"), (dict(source_name="tiny-model", source_fingerprint="new"), "stale")):
        c = fresh()
        assert not c.load_prequantized(p, **kw) and why in flux2b.last_error()
        with pytest.raises(flux2b.Flux2Error):
            c.get_tensor("xEmbedder.weight")
        c.close()
    c = fresh(quant=flux2b.QUANT["nvfp4"])
    assert not c.load_prequantized(p) and "quantization" in flux2b.last_error()
    c.close()
    c = fresh(cfg_=tiny_cfg(O, guidance=False, layers=(2, 1)))
    assert not c.load_prequantized(p) and "key set mismatch" in flux2b.last_error()
    c.close()
    c = fresh(cfg_=tiny_cfg(O, guidance=False, layers=(1, 1), joint=512))
    assert not c.load_prequantized(p) and "tensor mismatch at contextEmbedder" in flux2b.last_error()
    c.close()
    # truncated payload behind a valid header (:99-141)
    open(p, "wb").write(blob[:-64])
    c = fresh()
    assert not c.load_prequantized(p) and "truncated" in flux2b.last_error()
    # ... and the standard path still works on the untouched context
    c.load_weights(W, dtype=torch.float16)
    c.finalize()
    assert np.array_equal(c.dit_forward(*_args(O, cfg)), ctx.dit_forward(*_args(O, cfg)))
    c.close()
    # bf16 models are not exported (:234-237)
    b = make_ctx(flux2b, cfg, W, dtype=torch.float16)
    with pytest.raises(flux2b.Flux2Error) as e:
        b.save_prequantized(p)
    assert e.value.case == "invalidConfiguration"
    b.close(); ctx.close()


def test_foreign_writer_layout_and_generic_loader(flux2b, tmp_path):
    """a file laid out by another writer (pretty-printed header, reversed key order, bf16 / f32 float parameters) loads;
    flux2b_load_safetensors hands over any file that already uses the Swift module keys"""
    from oracle import flux2_oracle as O
    q = flux2b.QUANT["nvfp4"]
    cfg = tiny_cfg(O, guidance=False, layers=(1, 1))
    W = O.random_dit_weights(cfg, seed=4, round_to=torch.float16)
    ctx = make_ctx(flux2b, cfg, W, quant=q, dtype=torch.float16)
    T = {}
    for k, w in W.items():
        if w.dim() != 2:
            T[k] = w.numpy().astype(np.float32)          # float parameters may differ in precision (:366-372)
            continue
        base = k[:-len(".weight")]
        T[base + ".weight"], T[base + ".scales"] = ctx.get_tensor(base + ".weight"), ctx.get_tensor(base + ".scales").view(np.uint8)
    p = str(tmp_path / "foreign.safetensors")
    write_safetensors(p, T, meta(source="tiny"), shuffle=True)
    c = flux2b.Context(dit=cfg, quant=q)
    assert c.load_prequantized(p, "tiny")
    c.finalize()
    assert np.array_equal(c.dit_forward(*_args(O, cfg)), ctx.dit_forward(*_args(O, cfg)))
    c.close()
    c = flux2b.Context(dit=cfg, quant=q)
    assert c.load_safetensors(p) == len(T)
    c.finalize()
    assert np.array_equal(c.dit_forward(*_args(O, cfg)), ctx.dit_forward(*_args(O, cfg)))
    c.close(); ctx.close()
