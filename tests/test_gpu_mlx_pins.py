"""GPU: the CUDA packers against REAL MLX outputs (tests/golden/mlx_pins.npz, tools/mlx_fixtures.py). Skipped ("parity unpinned")
while the file is absent — see tests/test_mlx_pins_cpu.py."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINS = os.path.join(ROOT, "tests", "golden", "mlx_pins.npz")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(PINS), reason="parity unpinned: tests/golden/mlx_pins.npz absent "
                                                                                    "(tools/mlx_fixtures.py needs a machine with mlx)")]


@pytest.mark.parametrize("tag", ["golden", "gauss"])
@pytest.mark.parametrize("name", ["qint8", "int4", "mxfp8", "mxfp4", "nvfp4"])
def test_device_packers_bit_exact_vs_mlx(flux2b, tag, name):
    pins = dict(np.load(PINS, allow_pickle=False))
    ctx = flux2b.Context()
    q = flux2b.QUANT[name]
    w = pins[f"quant_{tag}_w"]
    p, s, b = ctx.quantize_matrix(q, w)
    assert np.array_equal(p.view(np.uint32), pins[f"quant_{tag}_{name}_packed"].view(np.uint32))
    ms = pins[f"quant_{tag}_{name}_scales"]
    if s.dtype == np.uint8:
        assert np.array_equal(s, ms.view(np.uint8).reshape(s.shape))
    else:
        assert np.array_equal(s.astype(np.float32), ms.astype(np.float32))
        assert np.array_equal(b.astype(np.float32), pins[f"quant_{tag}_{name}_biases"].astype(np.float32))
    ctx.close()
