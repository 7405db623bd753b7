"""The oracle's analytic loss gradients (oracle.dense_loss_grad) against central finite differences of the float64
loss they differentiate, and that float64 loss against the float32 reference-order loss the golden vectors pin."""
import numpy as np
import pytest

from oracle import dense_head_ref as O
from oracle import synth


def _case(seed, ch, reg_ch, cen_mode, hw=(6, 7)):
    rng = np.random.default_rng(seed)
    cls0 = reg_ch + (1 if cen_mode else 0)
    t = np.zeros(hw + (ch,), np.float64)
    p = rng.normal(-1.0, 1.5, size=hw + (ch,))
    pos = rng.random(hw) < 0.3
    t[..., cls0:][pos, rng.integers(0, ch - cls0, size=pos.sum())] = 1.0
    if reg_ch:
        t[..., :4][pos] = rng.uniform(0.3, 4.0, size=(pos.sum(), 4))
        p[..., :4] = rng.uniform(0.3, 4.0, size=hw + (4,))
    if cen_mode:
        t[..., reg_ch][pos] = rng.uniform(0.2, 1.0, size=pos.sum())
    return t, p


@pytest.mark.parametrize("cfg", [dict(ch=9, reg_ch=4, cen_mode=1, reg_mode=0, pos_rule="ge1"),
                                 dict(ch=9, reg_ch=4, cen_mode=2, reg_mode=0, pos_rule="ge1"),
                                 dict(ch=9, reg_ch=4, cen_mode=1, reg_mode=1, pos_rule="ge1"),
                                 dict(ch=8, reg_ch=4, cen_mode=0, reg_mode=0, pos_rule="gt0"),
                                 dict(ch=5, reg_ch=0, cen_mode=0, reg_mode=0, pos_rule="gt0")])
@pytest.mark.parametrize("gamma", [2.0, 1.5])
def test_analytic_gradient_matches_finite_differences(cfg, gamma):
    t, p = _case(7 + cfg["ch"], cfg["ch"], cfg["reg_ch"], cfg["cen_mode"])
    w = (1.3, 0.7, 2.1)
    kw = dict(reg_ch=cfg["reg_ch"], cen_mode=cfg["cen_mode"], reg_mode=cfg["reg_mode"], pos_rule=cfg["pos_rule"], gamma=gamma)
    g = O.dense_loss_grad(t, p, weights=w, **kw)

    def scalar(q):
        c, r, e = O.dense_loss_f64(t, q, **kw)
        return w[0] * c + w[1] * r + w[2] * e
    rng = np.random.default_rng(1)
    idx = [tuple(rng.integers(0, s) for s in p.shape) for _ in range(60)]
    idx += [(i, j, c) for (i, j) in zip(*np.nonzero(t[..., -1] + t[..., -2] > 0)) for c in range(min(p.shape[-1], 5))][:40]
    h = 1e-6
    for ix in idx:
        q1, q2 = p.copy(), p.copy()
        q1[ix] += h
        q2[ix] -= h
        fd = (scalar(q1) - scalar(q2)) / (2 * h)
        assert abs(fd - g[ix]) <= 2e-5 * max(1.0, abs(fd)), (ix, fd, g[ix])


def test_float64_loss_agrees_with_reference_order_float32_loss():
    boxes, nbox = synth.make_boxes(1, 256, 10, 6, 8.0, 150.0, 5)
    tg, _ = O.fcos_format_data(boxes[0, :nbox[0]], [256, 256], 6)
    pred = synth.fcos_predictions(1, 256, 6, 9)
    want = O.fcos_model_loss(tg, [p[0] for p in pred])
    got = np.zeros(3)
    for t, p in zip(tg, pred):
        got += O.dense_loss_f64(t, p[0], reg_ch=4, cen_mode=1, pos_rule="ge1")
    assert np.allclose(got, [float(v) for v in want], rtol=2e-6)
