"""Label preparation / result formatting kernels vs the golden vectors and the oracle: bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dense_head_ref as O  # noqa: E402


def _dh():
    import densehead
    return densehead


def test_box_converts_golden(golden):
    dh = _dh()
    z = golden("prep")
    raw = z["raw"]
    assert np.array_equal(dh.prep.swap_xy(raw).cpu().numpy(), z["swap_xy"])
    assert np.array_equal(dh.prep.convert_to_xywh(raw).cpu().numpy(), z["to_xywh"])
    assert np.array_equal(dh.prep.convert_to_corners(raw).cpu().numpy(), z["to_corners"])
    assert np.array_equal(dh.prep.flip_boxes_horizontal(raw).cpu().numpy(), z["flipped"])
    assert dh.prep.swap_xy(np.zeros((0, 4), np.float32)).shape == (0, 4)


def test_prepare_labels_ragged_padded_and_flip(golden):
    dh = _dh()
    z = golden("prep")
    raw = z["raw"]
    rng = np.random.default_rng(5)
    counts = [7, 0, 21, 12]
    boxes = [raw[sum(counts[:i]):sum(counts[:i + 1])] for i in range(4)]
    classes = [rng.integers(0, 20, size=c).astype(np.float32) for c in counts]
    flip = [1, 0, 0, 1]
    lab, nb = dh.prep.prepare_labels(boxes, classes, flip=flip)
    lab, nb = lab.cpu().numpy(), nb.cpu().numpy()
    assert nb.tolist() == counts and lab.shape == (4, 24, 5)
    for i in range(4):
        want = O.prepare_labels(boxes[i], classes[i], flip=bool(flip[i]))
        assert np.array_equal(lab[i, :counts[i]], want) and not lab[i, counts[i]:].any()
    # the same through the padded-input form, truncated to 8 boxes per image
    pad = np.zeros((4, 21, 4), np.float32); pc = np.zeros((4, 21), np.float32)
    for i in range(4):
        pad[i, :counts[i]], pc[i, :counts[i]] = boxes[i], classes[i]
    lab2, nb2 = dh.prep.prepare_labels(pad, pc, flip=flip, nbox=counts, max_boxes=8)
    assert nb2.cpu().numpy().tolist() == [7, 0, 8, 8]
    from densehead import _capi
    assert _capi.status(0) & _capi.DH_STATUS_TRUNCATED  # the kernel says that it dropped boxes (and the read clears the bit)
    assert not _capi.status(0) & _capi.DH_STATUS_TRUNCATED
    with pytest.raises(ValueError):  # host lists: the wrapper knows the counts
        dh.prep.prepare_labels(boxes, classes, max_boxes=8)
    assert np.array_equal(lab2.cpu().numpy()[2], lab[2, :8])
    # feeds the encoders directly
    outs, cnt = dh.fcos.format_data_batch(lab2, nb2, [512, 512], 20, [512, 512])
    want, wcnt = O.fcos_format_data(lab[3, :8], [512, 512], 20)
    assert cnt[3].tolist() == wcnt and all(np.array_equal(outs[l][3].cpu().numpy(), want[l]) for l in range(5))


def test_format_detections_vs_oracle():
    dh = _dh()
    rng = np.random.default_rng(6)
    rows = rng.uniform(0, 600, size=(3, 50, 6)).astype(np.float32)
    rows[..., 5] = rng.integers(0, 80, size=(3, 50))
    n_keep = [50, 0, 17]
    ratios = np.array([[1.5, 0.75], [1.0, 1.0], [0.3333, 2.1]], np.float32)
    b, s, l = (t.cpu().numpy() for t in dh.prep.format_detections(rows, n_keep, ratios))
    for i in range(3):
        wb, ws, wl = O.format_detections(rows[i, :n_keep[i]], ratios[i, 0], ratios[i, 1])
        assert np.array_equal(b[i, :n_keep[i]], wb) and np.array_equal(s[i, :n_keep[i]], ws) and np.array_equal(l[i, :n_keep[i]], wl)
        assert not b[i, n_keep[i]:].any() and np.all(l[i, n_keep[i]:] == -1)
