"""The framework-free DLPack capsule reader (densehead/_dlpack.py) against torch CPU tensors: pointer, shape, strides,
dtype and device must match what torch reports; the capsule stays usable afterwards."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import PKG  # noqa: F401


def test_capsule_view_matches_torch():
    from densehead import _dlpack
    t = torch.arange(2 * 3 * 5, dtype=torch.float32).reshape(2, 3, 5)
    v = _dlpack.view(torch.utils.dlpack.to_dlpack(t))
    assert v.ptr == t.data_ptr() and v.shape == (2, 3, 5) and v.dtype == "float32" and v.device == ("cpu", 0)
    assert v.is_contiguous()
    s = t[:, 1:, ::2]
    v = _dlpack.view(torch.utils.dlpack.to_dlpack(s))
    assert v.ptr == s.data_ptr() and v.shape == tuple(s.shape) and v.strides == tuple(s.stride()) and not v.is_contiguous()
    i = torch.zeros(4, dtype=torch.int32)
    assert _dlpack.view(i).dtype == "int32"          # via __dlpack__()
    cap = torch.utils.dlpack.to_dlpack(t)
    _dlpack.view(cap)
    back = torch.utils.dlpack.from_dlpack(cap)       # looking inside does not consume the capsule
    assert back.data_ptr() == t.data_ptr()
    with pytest.raises(ValueError):
        _dlpack.view(cap)                            # ... but a consumed capsule is refused
