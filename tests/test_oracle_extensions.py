"""The extension modes BASELINE's north_star names and the reference does not have -- GIoU loss,
Gaussian heat map, min-area FCOS tie-break.  There is no reference behaviour to pin them to: their specification is the
oracle function, and these tests tie that function to independent statements of the same thing (torchvision's
generalized_box_iou, closed forms, brute-force per-cell rules, finite differences)."""
import numpy as np
import pytest

from oracle import dense_head_ref as O
from oracle import synth


# ---- GIoU -----------------------------------------------------------------------------------------------------------
def _tblr_case(seed, hw=(7, 9)):
    rng = np.random.default_rng(seed)
    t = rng.uniform(0.3, 5.0, size=hw + (4,)).astype(np.float32)
    p = rng.uniform(0.3, 5.0, size=hw + (4,)).astype(np.float32)
    m = (rng.random(hw) < 0.6).astype(np.float32)
    return t, p, m


def test_giou_loss_matches_torchvision_generalized_box_iou():
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision.ops")
    t, p, m = _tblr_case(3)
    hh, ww = m.shape
    gy, gx = np.meshgrid(np.arange(hh, dtype=np.float64), np.arange(ww, dtype=np.float64), indexing="ij")

    def corners(v):  # (x1, y1, x2, y2)
        v = v.astype(np.float64)
        return np.stack([gx - v[..., 2], gy - v[..., 0], gx + v[..., 3], gy + v[..., 1]], -1).reshape(-1, 4)
    g = torch.diag(tv.generalized_box_iou(torch.from_numpy(corners(t)), torch.from_numpy(corners(p)))).numpy()
    want = float(((1.0 - g) * m.reshape(-1)).sum())
    assert abs(float(O.giou_loss(t, p, m)) - want) <= 1e-5 * max(1.0, abs(want))


def test_giou_loss_closed_forms():
    one = np.ones((1, 1), np.float32)
    same = np.array([[[1.0, 2.0, 1.5, 0.5]]], np.float32)
    assert abs(float(O.giou_loss(same, same, one))) < 1e-6                      # identical boxes: GIoU = 1
    inner = np.array([[[0.5, 1.0, 0.75, 0.25]]], np.float32)                    # nested: enclosing box = the outer one
    iou = (1.5 * 1.0) / (3.0 * 2.0)
    assert abs(float(O.giou_loss(same, inner, one)) - (1.0 - iou)) < 1e-6
    # disjoint boxes cannot be written as tblr around one point with positive distances; negative distances can:
    a = np.array([[[1.0, 1.0, 1.0, 1.0]]], np.float32)                          # [-1, 1] x [-1, 1]
    b = np.array([[[-2.0, 4.0, -2.0, 4.0]]], np.float32)                        # [2, 4] x [2, 4]
    # IoU = 0, C = 5 * 5, union = 4 + 4 -> GIoU = -(25 - 8) / 25
    assert abs(float(O.giou_loss(a, b, one)) - (1.0 + 17.0 / 25.0)) < 1e-6
    assert float(O.giou_loss(a, b, np.zeros((1, 1), np.float32))) == 0.0


def test_giou_float64_loss_and_gradient():
    t, p, m = _tblr_case(11)
    tt = np.concatenate([t, (m[..., None] > 0).astype(np.float32)], -1)          # one class channel marks the positives
    pp = np.concatenate([p, np.zeros_like(m)[..., None]], -1)
    kw = dict(reg_ch=4, cen_mode=0, reg_mode=2, pos_rule="gt0")
    c, r, e = O.dense_loss_f64(tt, pp, **kw)
    assert abs(r - float(O.giou_loss(t, p, m))) <= 1e-5 * max(1.0, abs(r))
    g = O.dense_loss_grad(tt, pp, weights=(0.0, 1.0, 0.0), **kw)
    h = 1e-6
    rng = np.random.default_rng(0)
    for _ in range(60):
        ix = (int(rng.integers(0, t.shape[0])), int(rng.integers(0, t.shape[1])), int(rng.integers(0, 4)))
        q1, q2 = pp.astype(np.float64).copy(), pp.astype(np.float64).copy()
        q1[ix] += h
        q2[ix] -= h
        fd = (O.dense_loss_f64(tt, q1, **kw)[1] - O.dense_loss_f64(tt, q2, **kw)[1]) / (2 * h)
        assert abs(fd - g[ix]) <= 2e-5 * max(1.0, abs(fd)), (ix, fd, g[ix])


def test_fcos_model_loss_giou_is_the_sum_over_levels():
    boxes, nbox = synth.make_boxes(1, 256, 10, 6, 8.0, 150.0, 5)
    tg, _ = O.fcos_format_data(boxes[0, :nbox[0]], [256, 256], 6)
    pred = [np.abs(p[0]) + np.float32(0.3) for p in synth.fcos_predictions(1, 256, 6, 9)]
    _, reg, _ = O.fcos_model_loss(tg, pred, reg_type="giou")
    want = sum(float(O.giou_loss(t[..., :4], p[..., :4], (t[..., 5:].max(-1) >= 1).astype(np.float32))) for t, p in zip(tg, pred))
    assert abs(float(reg) - want) <= 1e-5 * max(1.0, want)


# ---- FCOS min-area tie-break ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fcos_min_area_is_the_per_cell_rule(seed):
    """Brute force: a cell's channels 0..4 are those the SMALLEST covering box would paint alone; classes are the union."""
    rng = np.random.default_rng(100 + seed)                                     # 14 boxes crowded around the image centre
    hw = np.exp(rng.uniform(np.log(20.0), np.log(120.0), size=(14, 2)))
    g = np.concatenate([rng.uniform(90.0, 166.0, size=(14, 2)), hw], axis=1) / 256.0
    g = np.concatenate([g, rng.integers(0, 5, size=(14, 1))], axis=1).astype(np.float32)
    g[3, 2:4] = g[9, 2:4]                                                        # and one exact area tie
    got, cnt = O.fcos_format_data(g, [256, 256], 5, order="min_area")
    ref, cnt_ref = O.fcos_format_data(g, [256, 256], 5)
    assert cnt == cnt_ref
    alone = [O.fcos_format_data(g[k:k + 1], [256, 256], 5)[0] for k in range(len(g))]
    area = (g[:, 2] * np.float32(256)) * (g[:, 3] * np.float32(256))
    overlaps = 0
    for l in range(5):
        cover = np.stack([a[l][..., 5:].max(-1) > 0 for a in alone])            # [n, H, W]
        assert np.array_equal(got[l][..., 5:], ref[l][..., 5:])                  # classes do not depend on the order
        for i, j in zip(*np.nonzero(cover.any(0))):
            ks = np.nonzero(cover[:, i, j])[0]
            k = ks[np.lexsort((-ks, area[ks]))[0]]                               # smallest area; equal areas: higher index
            assert np.array_equal(got[l][i, j, :5], alone[k][l][i, j, :5]), (l, i, j)
            overlaps += len(ks) > 1
    assert overlaps > 0
    assert any(not np.array_equal(a, b) for a, b in zip(got, ref))               # and it differs from the reference's rule


# ---- Gaussian heat ----------------------------------------------------------------------------------------------------
def test_gaussian_heat_map_properties():
    boxes, nbox = synth.make_boxes(1, 512, 40, 3, 8.0, 400.0, 5)
    g = boxes[0, :nbox[0]]
    got = O.centernet_gaussian_format_data(g, [512, 512], 3, stride=4)
    ref = O.centernet_format_data(g, [512, 512], 3, stride=4)
    assert np.array_equal(got[..., :4], ref[..., :4]) and np.array_equal(got[..., 5:], ref[..., 5:])
    heat = got[..., 4]
    assert np.array_equal(heat > 0, ref[..., 4] > 0)                             # same footprints
    assert heat.max() == 1.0 and heat.min() >= 0.0
    # one box alone: closed form, separable, symmetric about the integer mean, 1 at the four cells around it
    k = int(np.argmax(g[:, 2] * g[:, 3]))
    one = O.centernet_gaussian_format_data(g[k:k + 1], [512, 512], 3, stride=4)[..., 4]
    ys, xs = np.nonzero(one)
    y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
    mu_y, mu_x = (y0 + y1) // 2, (x0 + x1) // 2
    std = max(1.0, float(np.sqrt(((g[k, 2] * g[k, 3]) * np.float32(128)) * np.float32(128), dtype=np.float32)))
    gy, gx = np.arange(y0, y1) + 0.5, np.arange(x0, x1) + 0.5
    want = np.exp(-(((gy - mu_y) ** 2 - 0.25)[:, None] + ((gx - mu_x) ** 2 - 0.25)[None, :]) / (2 * std * std))
    want[mu_y - y0, mu_x - x0] = 1.0
    assert np.allclose(one[y0:y1, x0:x1], want.astype(np.float32), rtol=1e-6, atol=0)
    assert np.allclose(one[mu_y - 1:mu_y + 1, mu_x - 1:mu_x + 1], 1.0)
    # overlaps: the maximum over the boxes
    singles = np.stack([O.centernet_gaussian_format_data(g[q:q + 1], [512, 512], 3, stride=4)[..., 4] for q in range(len(g))])
    assert np.array_equal(heat, singles.max(0))
