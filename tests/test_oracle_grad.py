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


# ---- pinned by the reference itself: central differences THROUGH the reference's own loss functions -----------------
# tests/golden/grad.npz (oracle/make_golden.py: grad) holds d(w . (cls, reg, cen)) / d pred at sampled elements,
# obtained by finite differences of fcos.model_loss (FCOS/fcos.py:464-496), its fcos_center / fcos_center_v1 copies,
# RetinaNet.train_loss (RetinaNet/retinanet_module.py:403-426) and the CenterNet model_loss pair, the reference's code
# evaluated in float64 under the stub.  The oracle's analytic gradient must reproduce them.
import os  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grad.npz"))
W_GOLD = tuple(float(v) for v in GOLD["weights"])


def _fd_close(got, want, what):
    err = np.abs(got - want)
    assert np.all(err <= 2e-6 * np.maximum(1.0, np.abs(want))), "%s: max err %g (want %g)" % (what, err.max(), want[err.argmax()])


def golden_fcos_case():
    g, seed = GOLD["fcos_g"], int(GOLD["fcos_seed"])
    pred = [p[0].astype(np.float64) for p in synth.fcos_predictions(1, 256, 20, seed)]
    for p in pred:
        p[..., :4] = np.abs(p[..., :4]) + 0.3
    return g, pred


FCOS_VARIANTS = {  # name -> (encoder, loss kwargs)
    "fcos_l1": (lambda g: O.fcos_format_data(g, [256, 256], 20)[0], dict(cen_mode=1, reg_mode=0)),
    "fcos_iou": (lambda g: O.fcos_format_data(g, [256, 256], 20)[0], dict(cen_mode=1, reg_mode=1)),
    "center_focal": (lambda g: O.fcos_center_format_data(g, [256, 256], 20)[0], dict(cen_mode=2, reg_mode=0)),
    "center_l1": (lambda g: O.fcos_center_format_data(g, [256, 256], 20)[0], dict(cen_mode=1, reg_mode=0)),
    "v1": (lambda g: O.fcos_center_v1_format_data(g, [256, 256], 20)[0], dict(cen_mode=2, reg_mode=0)),
}


@pytest.mark.parametrize("name", sorted(FCOS_VARIANTS))
def test_fcos_gradients_match_differences_through_the_reference(name):
    g, pred = golden_fcos_case()
    enc, kw = FCOS_VARIANTS[name]
    tg = enc(g)
    for l in range(5):
        grad = O.dense_loss_grad(tg[l], pred[l], weights=W_GOLD, reg_ch=4, pos_rule="ge1", **kw)
        _fd_close(grad.reshape(-1)[GOLD["%s_idx%d" % (name, l)]], GOLD["%s_fd%d" % (name, l)], "%s level %d" % (name, l))


def test_retina_gradients_match_differences_through_the_reference():
    g, seed = GOLD["retina_g"], int(GOLD["retina_seed"])
    lab, _ = O.retina_format_data(g, [128, 128], 20)
    pred = synth.retina_predictions(1, 128, 20, seed)
    for l in range(5):
        grad = O.dense_loss_grad(np.stack(lab[l]), pred[l][0].astype(np.float64), weights=(W_GOLD[0], W_GOLD[1], 0.0), reg_ch=4,
                                 cen_mode=0, pos_rule="gt0")
        _fd_close(grad.reshape(-1)[GOLD["retina_idx%d" % l]], GOLD["retina_fd%d" % l], "retina level %d" % l)


@pytest.mark.parametrize("name", ["cn_s8", "cn_hg"])
def test_centernet_gradients_match_differences_through_the_reference(name):
    boxes, nbox, seed = GOLD["cn_boxes"], GOLD["cn_nbox"], int(GOLD["cn_seed"])
    yp = synth.centernet_s8_predictions(2, 256, 8, 5, 3, seed).astype(np.float64)
    if name == "cn_s8":
        yt = np.stack([O.centernet_s8_format_data(boxes[b, :nbox[b]], [32, 64, 128, 256, 512], [256, 256], 3)[0] for b in range(2)])
    else:
        yt = np.stack([O.centernet_hourglass_format_data(boxes[b, :nbox[b]], [256, 256], 3)[0] for b in range(2)])
        yp = np.ascontiguousarray(yp[:, :, :, 0, :])
    grad = O.dense_loss_grad(yt, yp, weights=W_GOLD, reg_ch=4, cen_mode=0, pos_rule="gt0")
    _fd_close(grad.reshape(-1)[GOLD[name + "_idx"]], GOLD[name + "_fd"], name)


# ---- a second, independent check: torch autograd over a torch restatement of the reference's loss formulas -----------
def _torch_losses(t, p, reg_ch, cen_mode, reg_mode, pos_rule, alpha=0.25, gamma=2.0, delta=1.0):
    """FCOS/fcos.py:380-462 written with torch ops in float64 (focal in the reference's stable form, smooth-L1 without
    the -delta/2 term, -log IoU on the integer grid); returns (cls, reg, cen)."""
    import torch
    cls0 = reg_ch + (1 if cen_mode else 0)

    def focal(y, x):
        s = torch.sigmoid(x)
        ce = torch.log1p(torch.exp(-x.abs()))
        return (y * alpha * ce * (1 - s) ** gamma + s ** gamma * (1 - y) * (1 - alpha) * ce
                + (1 - y) * (1 - alpha) * torch.clamp(x, min=0) * s ** gamma - y * alpha * torch.clamp(x, max=0) * (1 - s) ** gamma).sum()

    def sl1(a, b, m):
        d = (a - b).abs()
        return (torch.where(d < delta, 0.5 * d * d, d) * m).sum()
    obj = t[..., cls0:].max(dim=-1).values
    m = (obj >= 1).double() if pos_rule == "ge1" else (obj > 0).double()
    cls = focal(t[..., cls0:], p[..., cls0:])
    reg = torch.zeros((), dtype=torch.float64)
    if reg_ch:
        if reg_mode == 0:
            reg = sl1(t[..., :4], p[..., :4], m[..., None])
        else:
            hh, ww = p.shape[-3], p.shape[-2]
            gy, gx = torch.meshgrid(torch.arange(hh, dtype=torch.float64), torch.arange(ww, dtype=torch.float64), indexing="ij")
            tb = (gy - t[..., 0], gx - t[..., 2], gy + t[..., 1], gx + t[..., 3])
            pb = (gy - p[..., 0], gx - p[..., 2], gy + p[..., 1], gx + p[..., 3])
            ih = torch.clamp(torch.minimum(tb[2], pb[2]) - torch.maximum(tb[0], pb[0]), min=0)
            iw = torch.clamp(torch.minimum(tb[3], pb[3]) - torch.maximum(tb[1], pb[1]), min=0)
            inter = ih * iw
            union = (tb[2] - tb[0]) * (tb[3] - tb[1]) + (pb[2] - pb[0]) * (pb[3] - pb[1]) - inter
            reg = (-torch.log(inter / (union + 1e-12) + 1e-12) * m).sum()
    cen = torch.zeros((), dtype=torch.float64)
    if cen_mode == 1:
        cen = sl1(t[..., reg_ch], torch.sigmoid(p[..., reg_ch]), 1.0)
    elif cen_mode == 2:
        cen = focal(t[..., reg_ch], p[..., reg_ch])
    return cls, reg, cen


@pytest.mark.parametrize("cfg", [dict(ch=9, reg_ch=4, cen_mode=1, reg_mode=0, pos_rule="ge1"),
                                 dict(ch=9, reg_ch=4, cen_mode=2, reg_mode=0, pos_rule="ge1"),
                                 dict(ch=9, reg_ch=4, cen_mode=1, reg_mode=1, pos_rule="ge1"),
                                 dict(ch=8, reg_ch=4, cen_mode=0, reg_mode=0, pos_rule="gt0")])
def test_analytic_gradient_matches_torch_autograd(cfg):
    torch = pytest.importorskip("torch")
    t, p = _case(31 + cfg["ch"], cfg["ch"], cfg["reg_ch"], cfg["cen_mode"], hw=(9, 11))
    w = (1.3, 0.7, 2.1)
    tt, pp = torch.from_numpy(t), torch.from_numpy(p).requires_grad_(True)
    kw = {k: v for k, v in cfg.items() if k != "ch"}
    c, r, e = _torch_losses(tt, pp, **kw)
    (w[0] * c + w[1] * r + w[2] * e).backward()
    g = O.dense_loss_grad(t, p, weights=w, **kw)
    want = pp.grad.numpy()
    assert np.all(np.abs(g - want) <= 1e-9 * np.maximum(1.0, np.abs(want)))
    # and the restated loss values are the oracle's float64 loss
    assert np.allclose([float(c.detach()), float(r.detach()), float(e.detach())], O.dense_loss_f64(t, p, **kw), rtol=1e-12)
