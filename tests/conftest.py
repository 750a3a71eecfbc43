import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "flux-2-swift-mlx_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def flux2b():
    import flux2b as m
    return m


@pytest.fixture(scope="session")
def have_gpu(flux2b):
    return flux2b.device_count() > 0


def rel_l2(a, b):
    import torch
    a = torch.as_tensor(a).detach().cpu().double().flatten()
    b = torch.as_tensor(b).detach().cpu().double().flatten()
    e = float((a - b).norm() / (b.norm() + 1e-30))
    # F2B_PARITY_LOG=<file>: every measured relative error with the test that asked for it (tools/parity_margins.py turns the log
    # into profiles/r02_parity_margins.md: measured error vs the bound written in the test)
    log = os.environ.get("F2B_PARITY_LOG")
    if log:
        with open(log, "a") as f:
            f.write(f"{os.environ.get('PYTEST_CURRENT_TEST', '?').split(' ')[0]}\t{e:.6e}\n")
    return e


@pytest.fixture(scope="session")
def golden():
    """tests/golden/golden.npz — oracle outputs committed by tools/make_golden.py."""
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden.npz")))
