"""Tensors cross the boundary zero-copy via DLPack: device-resident capsules / `__dlpack__` objects (what
`tf.experimental.dlpack.to_dlpack` produces on the reference side) reach the kernels without a copy, and the raw-pointer
reader of densehead/_dlpack.py agrees with the framework about where the data lives."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import synth  # noqa: E402


class Foreign:
    """A tensor from 'another framework': speaks only the DLPack protocol."""

    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None):
        return self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_device_capsules_are_consumed_without_a_copy():
    import densehead as dh
    from densehead import _dlpack, _tensors
    boxes, nbox = synth.config_boxes("fcos_voc", 4, synth.seed_for(8, 1))
    pred = [torch.from_numpy(p).cuda() for p in synth.fcos_predictions(4, 512, 20, 3)]
    bt, nt = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
    v = _dlpack.view(torch.utils.dlpack.to_dlpack(pred[0]))
    assert v.ptr == pred[0].data_ptr() and v.device == ("cuda", 0) and v.dtype == "float32" and v.is_contiguous()
    assert _tensors.to_device(torch.utils.dlpack.to_dlpack(pred[0]), torch.float32).data_ptr() == pred[0].data_ptr()
    assert _tensors.to_device(Foreign(pred[0]), torch.float32).data_ptr() == pred[0].data_ptr()
    want = dh.fcos.encode_loss_batch(bt, nt, [512, 512], 20, [512, 512], pred)
    got = dh.fcos.encode_loss_batch(torch.utils.dlpack.to_dlpack(bt), Foreign(nt), [512, 512], 20, [512, 512],
                                    [torch.utils.dlpack.to_dlpack(p) if i % 2 else Foreign(p) for i, p in enumerate(pred)])
    assert all(torch.equal(a, b) for a, b in zip(want, got))
    outs, cnt = dh.fcos.format_data_batch(Foreign(bt), torch.utils.dlpack.to_dlpack(nt), [512, 512], 20, [512, 512])
    outs2, cnt2 = dh.fcos.format_data_batch(boxes, nbox, [512, 512], 20, [512, 512])
    assert torch.equal(cnt, cnt2) and all(torch.equal(a, b) for a, b in zip(outs, outs2))


class EagerLike(Foreign):
    """Like a TensorFlow EagerTensor on the GPU: speaks DLPack AND has .numpy() (which would be a D2H copy)."""

    def numpy(self):
        raise AssertionError("a device-resident tensor must cross through DLPack, not through the host")


def test_device_tensors_with_a_numpy_method_still_cross_zero_copy():
    import densehead as dh
    from densehead import _tensors
    boxes, nbox = synth.config_boxes("fcos_voc", 2, synth.seed_for(8, 2))
    pred = [torch.from_numpy(p).cuda() for p in synth.fcos_predictions(2, 512, 20, 4)]
    bt, nt = torch.from_numpy(boxes).cuda(), torch.from_numpy(nbox).cuda()
    assert _tensors.to_device(EagerLike(pred[0]), torch.float32).data_ptr() == pred[0].data_ptr()
    want = dh.fcos.encode_loss_batch(bt, nt, [512, 512], 20, [512, 512], pred)
    got = dh.fcos.encode_loss_batch(EagerLike(bt), EagerLike(nt), [512, 512], 20, [512, 512], [EagerLike(p) for p in pred])
    assert all(torch.equal(a, b) for a, b in zip(want, got))
    tg, _ = dh.fcos.format_data_batch(bt, nt, [512, 512], 20, [512, 512])
    a = dh.fcos.model_loss_batch(tg, pred)
    b = dh.fcos.model_loss_batch([EagerLike(t) for t in tg], [EagerLike(p) for p in pred])
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_calls_on_a_side_stream_are_ordered_on_it():
    """`stream=`: staging copies, allocations and kernels all run on the caller's stream (they used to be issued
    on the current stream while the kernels ran on the other one)."""
    import densehead as dh
    boxes, nbox = synth.config_boxes("fcos_voc", 4, synth.seed_for(8, 3))
    pred = [torch.from_numpy(p).cuda() for p in synth.fcos_predictions(4, 512, 20, 5)]
    want = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], 20, [512, 512], pred)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for _ in range(20):  # host inputs: staged on `side`, consumed on `side`
        got = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], 20, [512, 512], pred, stream=side)
        side.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(want, got))
