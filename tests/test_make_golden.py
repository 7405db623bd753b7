"""The committed recipe reproduces the committed fixtures: `oracle/make_golden.py` is run as a whole into a temporary
directory (reference source under the TF stub) and every array is compared with tests/golden byte for byte.  Needs
/root/reference (skipped on the GPU box, where the reference does not exist)."""
import os

import numpy as np
import pytest

from oracle import make_golden as M
from oracle import ref_loader as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.skipif(not R.available(), reason="reference tree not present")
def test_recipe_runs_and_reproduces_every_fixture(tmp_path):
    saved = M.OUT
    try:
        M.main(str(tmp_path))
    finally:
        M.OUT = saved
    made = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    assert made == sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    for f in made:
        a, b = np.load(os.path.join(GOLDEN, f)), np.load(os.path.join(tmp_path, f))
        assert sorted(a.files) == sorted(b.files), f
        for k in a.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, (f, k)
            if f == "grad.npz" and k.endswith(("_fd", "_fd0", "_fd1", "_fd2", "_fd3", "_fd4")):
                # finite differences of float64 sums: reproducible to the last bits of the quotient, not beyond
                assert np.allclose(a[k], b[k], rtol=1e-9, atol=1e-9), (f, k)
            else:
                assert a[k].tobytes() == b[k].tobytes(), (f, k)
