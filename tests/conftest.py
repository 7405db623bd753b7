"""pytest configuration: the `gpu` marker, import paths, shared helpers.

`-m "not gpu"` tests: oracle vs golden vectors, host logic, C-ABI symbol export (no GPU needed).
`-m gpu` tests: CUDA path (through the C-ABI) vs the oracle and the golden vectors on a B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cv-lite-object-detection_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]
    return load


def assert_close(actual, expected, rtol=1e-5, atol_floor=1e-5, what=""):
    """north_star tolerance: |a-b| <= rtol*max(atol_floor/rtol... i.e. 1e-5 relative with an
    absolute floor of 1e-5*max(1,|b|) for targets that cancel to ~0."""
    a = np.asarray(actual, dtype=np.float64)
    b = np.asarray(expected, dtype=np.float64)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = rtol * np.maximum(1.0, np.abs(b)) if atol_floor else rtol * np.abs(b)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError("%s: %d/%d outside tolerance; worst at %s: got %r want %r"
                             % (what, int(bad.sum()), bad.size, i, a[i], b[i]))
