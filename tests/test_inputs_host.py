"""Host-side input handling of the Python mirror that needs no GPU: class-id validation (the reference raises
IndexError for a class outside [0, num_classes), FCOS/fcos.py:281-283) and how `to_device` decides between the
zero-copy DLPack route and the host route."""
import numpy as np
import pytest

from densehead import _batch, _tensors


def test_check_classes_accepts_valid_and_ignores_padding():
    boxes = np.zeros((2, 4, 5), np.float32)
    boxes[0, :2, 4] = [0, 19]
    boxes[1, :1, 4] = [7]
    boxes[1, 1:, 4] = 99  # padding rows beyond nbox are never read
    _batch.check_classes(boxes, np.array([2, 1], np.int32), 20)


@pytest.mark.parametrize("bad", [-1.0, 20.0, 1e9, float("nan")])
def test_check_classes_raises_index_error_like_the_reference(bad):
    boxes = np.zeros((1, 4, 5), np.float32)
    boxes[0, 1, 4] = bad
    with pytest.raises(IndexError):
        _batch.check_classes(boxes, np.array([3], np.int32), 20)


def test_check_classes_skips_device_inputs():
    _batch.check_classes(object(), None, 20)  # not an ndarray: the kernels flag it instead (dh_get_status)


class _FakeEager:
    """Looks like a TensorFlow EagerTensor: has .numpy() AND the DLPack protocol."""

    def __init__(self, device_type):
        self.device_type = device_type

    def numpy(self):
        return np.zeros(3, np.float32)

    def __dlpack__(self, stream=None):
        raise AssertionError("not expected to be consumed in this test")

    def __dlpack_device__(self):
        return (self.device_type, 0)


def test_dlpack_route_is_chosen_by_where_the_data_lives():
    assert _tensors._on_cuda(_FakeEager(2)) and _tensors._on_cuda(_FakeEager(13))
    assert not _tensors._on_cuda(_FakeEager(1))           # kDLCPU: .numpy() is the cheap route
    assert not _tensors._on_cuda(np.zeros(3))             # no protocol at all
    assert _tensors._tf_gpu_capsule(np.zeros(3)) is None  # not a TensorFlow object
