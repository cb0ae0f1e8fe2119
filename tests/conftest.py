import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference algorithm (oracle/rest_oracle.c) -- the checker, never the product."""
    from oracle.api import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def oracle_blas():
    """Same oracle with its BLAS calls routed to the scipy-bundled OpenBLAS (what the reference links)."""
    from oracle.api import Oracle
    o = Oracle()
    if not o.load_openblas():
        pytest.skip("no OpenBLAS found for the oracle")
    return o


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def rt():
    import rest_tensors_b200
    return rest_tensors_b200


@pytest.fixture(scope="session")
def ctx():
    """device context on cuda:0 (gpu tests only)"""
    import torch
    from rest_tensors_b200.device import Context
    assert torch.cuda.is_available(), "gpu test started without a CUDA device"
    return Context(0)


def rel_err(x, y):
    """norm-wise relative error ||x-y||_inf / ||y||_inf"""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    den = np.max(np.abs(y)) if y.size else 0.0
    num = np.max(np.abs(x - y)) if y.size else 0.0
    return num / den if den > 0 else num


def assert_close_1e10(x, y, what=""):
    """The north-star bar: 1e-10 relative, norm-wise and element-wise |x-y| <= 1e-10*(|y| + 1e-3*||y||_inf)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    assert x.shape == y.shape, f"{what}: shape {x.shape} vs {y.shape}"
    if y.size == 0:
        return
    assert np.all(np.isfinite(x)), f"{what}: non-finite values"
    ymax = np.max(np.abs(y))
    assert rel_err(x, y) <= 1e-10, f"{what}: norm-wise rel err {rel_err(x, y):.3e} > 1e-10"
    bound = 1e-10 * (np.abs(y) + 1e-3 * ymax)
    bad = np.abs(x - y) > bound
    assert not bad.any(), f"{what}: {bad.sum()} elements beyond the element-wise 1e-10 bound"
