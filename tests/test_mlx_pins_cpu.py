"""CPU: the oracle against REAL MLX outputs (tests/golden/mlx_pins.npz, written by tools/mlx_fixtures.py on a machine with
`mlx`). Absent file = "parity unpinned" (reported as a skip): the oracle's quantizer / norm / attention rules are restated from
MLX's published semantics and have never been compared with MLX itself (oracle/quant_oracle.c header)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINS = os.path.join(ROOT, "tests", "golden", "mlx_pins.npz")

pytestmark = pytest.mark.skipif(not os.path.exists(PINS), reason="parity unpinned: tests/golden/mlx_pins.npz absent — run tools/mlx_fixtures.py "
                                                                  "on a machine with mlx (mlx-swift 0.31.6 / MLX 0.31.x) and commit the file")


@pytest.fixture(scope="module")
def pins():
    return dict(np.load(PINS, allow_pickle=False))


@pytest.mark.parametrize("tag", ["golden", "gauss"])
@pytest.mark.parametrize("name", ["qint8", "int4", "mxfp8", "mxfp4", "nvfp4"])
def test_quantizers_bit_exact_vs_mlx(pins, tag, name):
    from oracle import quant_oracle as Q
    q = Q.QUANT[name]
    w = pins[f"quant_{tag}_w"]
    p, s, b = Q.quantize(q, w)
    assert np.array_equal(p, pins[f"quant_{tag}_{name}_packed"].view(np.uint32)), "packed codes differ from mx.quantize"
    ms = pins[f"quant_{tag}_{name}_scales"]
    if s.dtype == np.uint8:
        assert np.array_equal(s, ms.view(np.uint8).reshape(s.shape))
    else:
        assert np.array_equal(s.astype(np.float32), ms.astype(np.float32)), "scales differ"
        assert np.array_equal(b.astype(np.float32), pins[f"quant_{tag}_{name}_biases"].astype(np.float32)), "biases differ"
    d = Q.dequantize(q, p, s, b, w.shape[1])
    np.testing.assert_array_equal(d.astype(np.float16).astype(np.float32), pins[f"quant_{tag}_{name}_dequant"].astype(np.float16).astype(np.float32))
    # QuantizedLinear forward: x · dequant(W)^T
    x = pins[f"qmm_{tag}_{name}_x"]
    np.testing.assert_allclose(x @ d.T, pins[f"qmm_{tag}_{name}_y"], rtol=2e-3, atol=2e-3)


def test_norms_attention_conv_vs_mlx(pins):
    from oracle import flux2_oracle as O
    t = torch.from_numpy
    np.testing.assert_allclose(O.rms_norm(t(pins["rms_x"]), t(pins["rms_w"])).numpy(), pins["rms_y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.layer_norm(t(pins["ln_x"])).numpy(), pins["ln_y"], rtol=1e-5, atol=1e-5)
    y = O.sdpa(t(pins["sdpa_q"]), t(pins["sdpa_k"]), t(pins["sdpa_v"]))
    np.testing.assert_allclose(y.numpy(), pins["sdpa_y"], rtol=1e-4, atol=1e-5)
    ym = O.sdpa(t(pins["sdpa_q"]), t(pins["sdpa_k"]), t(pins["sdpa_v"]), mask=t(pins["sdpa_mask"]))
    np.testing.assert_allclose(ym.numpy(), pins["sdpa_y_masked"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(O.linear(t(pins["lin_x"]), t(pins["lin_w"]).float()).numpy(), pins["lin_y"], rtol=1e-5, atol=1e-5)
    assert "float32" in str(pins["lin_y_dtype"])   # fp32 activations x f16 weights promote to fp32
    np.testing.assert_allclose(O.conv2d_nhwc(t(pins["conv_x"]), t(pins["conv_w"]), None, 1).numpy(), pins["conv_y"], rtol=1e-4, atol=1e-5)
    W = {"d.conv.weight": t(pins["conv_w"]), "d.conv.bias": torch.zeros(pins["conv_w"].shape[0])}
    np.testing.assert_allclose(O.downsample2d(W, "d", t(pins["conv_x"])).numpy(), pins["conv_y_s2"], rtol=1e-4, atol=1e-5)
    xs = t(pins["silu_x"])
    np.testing.assert_allclose(torch.nn.functional.silu(xs).numpy(), pins["silu_y"], rtol=1e-5, atol=1e-6)


def test_uint8_cast_vs_mlx(pins):
    from oracle import flux2_oracle as O
    x = torch.from_numpy(pins["u8_in"]).reshape(1, 1, 1, -1).expand(1, 3, 1, -1)
    got = O.postprocess_vae_output(x).numpy()[0, :, 0]
    assert np.array_equal(got, pins["u8_out"]), "MLX asType(.uint8) does not truncate as restated (SURVEY §9.17): flip the single rounding switch"
