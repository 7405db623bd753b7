"""Target maps sent back through the detector: the computation inside the reference's `show_heatmap`
(FCOS/train_fcos_center_voc.py:13-121) -- SURVEY section 8 row f-4.

CPU: the oracle's restatement recovers the ground-truth boxes it was given (a property the reference's figure relies on).
GPU: `densehead.fcos.ground_truth_detections` against the oracle, bit for bit, and the same property at the full batch."""
import numpy as np
import pytest

from oracle import dense_head_ref as O
from oracle import synth


def _gt_rectangles(boxes, n, side):
    """(x1, y1, w, h) pixels of the n valid rows of one image's (cy, cx, h, w, class) normalised labels."""
    g = boxes[:n].astype(np.float64)
    return np.stack([(g[:, 1] - g[:, 3] / 2) * side, (g[:, 0] - g[:, 2] / 2) * side, g[:, 3] * side, g[:, 2] * side], axis=1)


def _assert_every_detection_is_a_gt_box(rect, gt, tol=2e-2):
    for r in rect:
        assert np.min(np.abs(gt - r[None]).max(axis=1)) < tol, (r, gt)


@pytest.mark.parametrize("center", [True, False])
def test_oracle_round_trip_recovers_ground_truth(center):
    boxes, nbox = synth.config_boxes("fcos_voc", 4, synth.seed_for(8, 40))
    for b in range(4):
        maps, _ = O.fcos_center_format_data(boxes[b, :nbox[b]], [384, 384], 20, [384, 384])
        rect, sc = O.fcos_ground_truth_detections(maps, 20, (384, 384), 384, 384, center=center)
        assert len(rect) >= 1 and np.all(sc == 1.0)  # (a cell shared by two classes yields its box once per class)
        _assert_every_detection_is_a_gt_box(rect, _gt_rectangles(boxes[b], nbox[b], 384))
    # source image of another size: rectangles scale with the reference's (swapped) ratios
    maps, _ = O.fcos_center_format_data(boxes[0, :nbox[0]], [384, 384], 20, [384, 384])
    a, _ = O.fcos_ground_truth_detections(maps, 20, (384, 384), 384, 384)
    b2, _ = O.fcos_ground_truth_detections(maps, 20, (768, 192), 384, 384)
    np.testing.assert_allclose(b2[:, 0], np.where(a[:, 0] > 0, a[:, 0] * 0.5, 0), rtol=1e-6)
    np.testing.assert_allclose(b2[:, 1], np.where(a[:, 1] > 0, a[:, 1] * 2.0, 0), rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("encoder", ["center", "fcos"])
@pytest.mark.parametrize("center", [True, False])
def test_gpu_round_trip_matches_the_oracle(encoder, center):
    torch = pytest.importorskip("torch")
    import densehead as dh
    B, side = 6, 384
    boxes, nbox = synth.config_boxes("fcos_voc", B, synth.seed_for(8, 41))
    enc = dh.fcos.format_data_batch
    maps, _ = enc(boxes, nbox, [side, side], 20, [side, side], mode=encoder)
    shapes = [(384, 384), (480, 640), (768, 192), (100, 100), (384, 384), (999, 1333)]
    rect, sc, valid = dh.fcos.ground_truth_detections(maps, 20, shapes, side, side, center=center)
    rect, sc, valid = rect.cpu().numpy(), sc.cpu().numpy(), valid.cpu().numpy()
    host = [m.cpu().numpy() for m in maps]
    for b in range(B):
        want_r, want_s = O.fcos_ground_truth_detections([m[b] for m in host], 20, shapes[b], side, side, center=center)
        assert valid[b] == len(want_r)
        assert np.array_equal(rect[b, :valid[b]], want_r) and np.array_equal(sc[b, :valid[b]], want_s)
        assert not rect[b, valid[b]:].any() and not sc[b, valid[b]:].any()


@pytest.mark.gpu
def test_gpu_round_trip_recovers_ground_truth_at_the_full_batch():
    torch = pytest.importorskip("torch")
    import densehead as dh
    B, side = 256, 512
    boxes, nbox = synth.config_boxes("fcos_voc", B, synth.seed_for(8, 42))
    maps, _ = dh.fcos.format_data_batch(boxes, nbox, [side, side], 20, [side, side], mode="center")
    rect, sc, valid = dh.fcos.ground_truth_detections(maps, 20, [(side, side)] * B, side, side, center=True)
    rect, valid = rect.cpu().numpy(), valid.cpu().numpy()
    assert np.all(valid >= 1)
    for b in range(B):
        _assert_every_detection_is_a_gt_box(rect[b, :valid[b]], _gt_rectangles(boxes[b], nbox[b], side))
    assert valid.sum() > 0.8 * nbox.sum()  # only boxes sharing a centre cell or overlapping > 0.75 may go missing
