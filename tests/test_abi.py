"""CPU: the C-ABI library loads, exports every symbol include/flux2b.h declares, has no CPU path, and its host-side
logic (scheduler, position ids) agrees with the oracle."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "flux2b.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(flux2b_[a-z0-9_]+)\s*\(", src))
    names -= {"flux2b_step_hook"}
    return sorted(names)


def test_header_symbols_exported(flux2b):
    L = flux2b.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert len(declared_symbols()) >= 50


def test_python_binding_lists_every_header_symbol(flux2b):
    assert sorted(flux2b.EXPORTS) == declared_symbols()


def test_no_cpu_fallback(flux2b, have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(flux2b.Flux2Error) as e:
        flux2b.Context()
    assert e.value.case == "noDevice"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "flux-2-swift-mlx_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".swift")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "quant_oracle.so" not in txt, f


@pytest.mark.parametrize("steps,seq,strength", [(4, 4096, 1.0), (4, 256, 1.0), (28, 16384, 1.0), (50, 4096, 0.5),
                                                (20, 1024, 0.3), (8, None, 1.0), (50, 4301, 0.01)])
def test_scheduler_matches_oracle(flux2b, steps, seq, strength):
    from oracle import flux2_oracle as O
    a, b = flux2b.FlowMatchEulerScheduler(), O.FlowMatchEulerScheduler()
    ta = a.set_timesteps(steps, seq, strength)
    tb = b.set_timesteps(steps, seq, strength)
    assert ta == tb and len(a.sigmas) == len(b.sigmas)
    np.testing.assert_allclose(a.sigmas, b.sigmas, rtol=2e-6, atol=1e-7)


def test_custom_sigmas(flux2b):
    s = flux2b.FlowMatchEulerScheduler()
    turbo = [1.0, 0.6509, 0.4374, 0.2932, 0.1893, 0.1108, 0.0495, 0.00031]
    s.set_custom_sigmas(turbo)
    assert len(s.sigmas) == 9 and s.sigmas[-1] == 0.0
    s.set_custom_sigmas([1.0, 0.5, 0.0])
    assert len(s.sigmas) == 3


@pytest.mark.parametrize("seq,steps", [(256, 4), (4096, 4), (4300, 50), (4301, 50), (16384, 28)])
def test_mu_matches_oracle(flux2b, seq, steps):
    from oracle import flux2_oracle as O
    assert abs(flux2b.compute_empirical_mu(seq, steps) - O.compute_empirical_mu(seq, steps)) < 1e-6


def test_position_ids_match_oracle(flux2b):
    from oracle import flux2_oracle as O
    assert np.array_equal(flux2b.image_position_ids(1024, 768), O.image_position_ids(1024, 768).numpy())
    assert np.array_equal(flux2b.text_position_ids(512), O.text_position_ids(512).numpy())
    assert np.array_equal(flux2b.reference_position_ids([4, 3], [5, 2]), O.reference_position_ids([4, 3], [5, 2]).numpy())


def test_quant_table(flux2b):
    # TransformerQuantization -> (bits, groupSize, mode): QuantizationConfig.swift:51-60; test :64-85, :1119-1124
    expect = {"qint8": (8, 64, True), "int4": (4, 64, True), "mxfp8": (8, 32, False), "mxfp4": (4, 32, False), "nvfp4": (4, 16, False)}
    for name, (bits, group, has_b) in expect.items():
        b, g, h, _ = flux2b.quant_params(flux2b.QUANT[name])
        assert (b, g, h) == (bits, group, has_b)
    with pytest.raises(flux2b.Flux2Error):
        flux2b.quant_params(0)


def test_te_config_is_validated_before_the_device(flux2b):
    """flux2b_te_create: configuration errors are Flux2Error.invalidConfiguration and are raised without a GPU."""
    from oracle import flux2_oracle as O
    for bad in (O.TEConfig(head_dim=80), O.TEConfig(num_heads=32, num_kv_heads=5), O.TEConfig(hidden_size=2561)):
        with pytest.raises(flux2b.Flux2Error) as e:
            flux2b.TextEncoder(bad)
        assert e.value.case == "invalidConfiguration"
