"""Freeze golden vectors by executing the reference's own source under the TF stub.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

TEST INFRASTRUCTURE ONLY.  The reference ships no tests or fixtures, so
these files are the pin: every array below is the output of an UNMODIFIED reference function
(`oracle/ref_loader.py` imports it from where it lies; nothing is copied) on seeded synthetic
inputs from `oracle/synth.py`, cast to float32.  Random predictions are not stored -- tests
regenerate them from the recorded seed -- only the reference's results are.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader as R  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SCALES = [32, 64, 128, 256, 512]


def f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32)


def scalar(x):
    return np.float32(np.asarray(x.numpy() if hasattr(x, "numpy") else x))


class _FixedModel:
    def __init__(self, outputs):
        self.outputs = outputs

    def __call__(self, x, training=None):
        return self.outputs


def kat(tf):
    """Appendix B of SURVEY.md: hand-sized known-answer inputs."""
    fcos, fc, fv1 = R.load("FCOS", "fcos"), R.load("FCOS", "fcos_center"), R.load("FCOS", "fcos_center_v1")
    g = np.array([(200, 200, 60, 60, 3), (203, 197, 40, 40, 7), (100, 400, 300, 200, 1),
                  (37, 45, 20, 10, 0), (500, 500, 50, 50, 2)], dtype=np.float32)
    g[:, :4] /= 512
    img = tf.cast([512, 512], tf.float32)
    d = {"g": g}
    for name, fn, kw in (("fcos", fcos.format_data, {}), ("center", fc.format_data, {}),
                         ("center_only", fc.format_data, {"center_only": True}), ("v1", fv1.format_data, {})):
        outs, cnt = fn(tf.constant(g), img, 20, img_pad=[512, 512], **kw)
        d[name + "_counts"] = np.array(cnt, dtype=np.int32)
        for l, o in enumerate(outs):
            d["%s_L%d" % (name, l)] = f32(o)
    d["focal"] = scalar(fcos.focal_loss(np.array([0, 1, 0, 1, 1.]), tf.constant([-3, -.5, 0, .5, 3.])))
    d["smooth_l1"] = scalar(fcos.smooth_l1_loss(tf.constant([[0, 1, 2, 3.]]), tf.constant([[.5, 1, 4, 2.2]]), mask=1.0))
    yt = np.zeros((2, 2, 4)); yt[1, 1] = (1, 2, 1.5, .5)
    mask = np.zeros((2, 2), dtype=np.float32); mask[1, 1] = 1
    d["iou_loss"] = scalar(fcos.iou_loss(yt, tf.constant(np.ones((2, 2, 4), dtype=np.float32)), mask))
    ut = R.load("RetinaNet", "utils")
    d["iou_b1"] = np.array([[100, 60, 40, 30], [10, 10, 4, 4]], dtype=np.float32)
    d["iou_b2"] = np.array([[96, 56, 45.254833, 22.627417], [104, 64, 32, 32], [10, 10, 4, 4]], dtype=np.float32)
    d["iou"] = f32(ut.compute_iou(d["iou_b1"], d["iou_b2"]))
    rn = R.retinanet(80)
    g2 = np.array([(100, 60, 40, 30, 3), (104, 64, 44, 30, 7)], dtype=np.float32)
    g2[:, :4] /= 256
    outs, n = rn.format_data(tf.constant(g2), tf.cast([256, 256], tf.float32))
    d["retina_g"], d["retina_pairs"] = g2, np.int64(n)
    d["retina_anchor_dims"] = np.array([[np.asarray(a) for a in lv] for lv in rn.anchor_boxes], dtype=np.float32)
    for a in range(9):
        d["retina_L0_A%d" % a] = f32(outs[0][a])
    d["retina_upper_sum"] = np.float32(sum(float(np.abs(outs[l][a]).sum()) for l in range(1, 5) for a in range(9)))
    d["nms_dets"] = np.array([[0, 0, 10, 10, .9], [1, 1, 11, 11, .8], [20, 20, 30, 30, .7], [0, 0, 10, 10, .9]], dtype=np.float32)
    d["nms_keep"] = np.asarray(rn.cpu_nms(d["nms_dets"], .5), dtype=np.int64)
    s8 = R.load("CenterNet", "tf_centernet_resnet_s8")
    bb = np.array([[0, 0, 10, 10, 90, 1], [1, 1, 10, 10, 80, 1], [20, 20, 10, 10, 70, 1], [0, 0, 10, 10, 85, 2]], dtype=np.float64)
    d["cnms_in"] = bb
    d["cnms_hard"] = np.array(s8.nms(bb.copy(), .5))
    d["cnms_soft"] = np.array(s8.nms(bb.copy(), .5, method="soft-nms"))
    g3 = np.array([(200, 200, 160, 120, 0), (210, 215, 40, 40, 0), (100, 400, 300, 200, 0)], dtype=np.float32)
    g3[:, :4] /= 512
    d["s8_g"] = g3
    d["s8"] = f32(s8.format_data(tf.constant(g3), SCALES, [512, 512], 1, stride=8)[0])
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **d)


def fcos_family(tf):
    mods = {"fcos": R.load("FCOS", "fcos"), "center": R.load("FCOS", "fcos_center"), "v1": R.load("FCOS", "fcos_center_v1")}
    d = {}
    cases = [  # (tag, side, img_dim, img_pad, nmax, classes, lo, hi, seed)
        ("c1", 512, (512., 512.), (512, 512), 20, 20, 8.0, 0.6 * 512, synth.seed_for(1, 0)),
        ("c1b", 512, (512., 512.), (512, 512), 20, 20, 8.0, 0.6 * 512, synth.seed_for(1, 1)),
        ("s384", 384, (384., 384.), (384, 384), 20, 20, 6.0, 330.0, synth.seed_for(1, 2)),
        ("pad", 512, (448., 512.), (512, 640), 20, 20, 6.0, 400.0, synth.seed_for(1, 3)),
        ("tiny", 512, (512., 512.), (512, 512), 30, 20, 2.0, 14.0, synth.seed_for(1, 4)),   # degenerate footprints
        ("coco", 640, (640., 640.), (640, 640), 100, 80, 8.0, 0.6 * 640, synth.seed_for(1, 5)),
    ]
    for tag, side, img_dim, img_pad, nmax, classes, lo, hi, seed in cases:
        boxes, nbox = synth.make_boxes(1, side, nmax, classes, lo, hi, seed, full=(tag in ("tiny", "coco")))
        g = boxes[0, :nbox[0]]
        d[tag + "_g"] = g
        d[tag + "_meta"] = np.array([img_dim[0], img_dim[1], img_pad[0], img_pad[1], classes], dtype=np.float64)
        img = tf.cast(list(img_dim), tf.float32)
        for name, fn, kw in (("fcos", mods["fcos"].format_data, {}), ("center", mods["center"].format_data, {}),
                             ("center_only", mods["center"].format_data, {"center_only": True}),
                             ("v1", mods["v1"].format_data, {})):
            outs, cnt = fn(tf.constant(g), img, classes, img_pad=list(img_pad), **kw)
            d["%s_%s_counts" % (tag, name)] = np.array(cnt, dtype=np.int32)
            for l, o in enumerate(outs):
                d["%s_%s_L%d" % (tag, name, l)] = f32(o)
    np.savez_compressed(os.path.join(OUT, "fcos_encode.npz"), **d)


def retina(tf):
    d = {}
    rn = R.retinanet(80)
    cases = [("s256", 256, 12, 8.0, 150.0, synth.seed_for(3, 0), None),
             ("c3", 640, 100, 8.0, 0.6 * 640, synth.seed_for(3, 1), None),
             ("thr4", 384, 30, 8.0, 300.0, synth.seed_for(3, 2), 0.4)]
    for tag, side, nmax, lo, hi, seed, thr in cases:
        boxes, nbox = synth.make_boxes(1, side, nmax, 80, lo, hi, seed, full=(tag == "c3"))
        g = boxes[0, :nbox[0]]
        kw = {} if thr is None else {"iou_thresh": thr}
        outs, n = rn.format_data(tf.constant(g), tf.cast([side, side], tf.float32), **kw)
        d[tag + "_g"] = g
        d[tag + "_meta"] = np.array([side, 0.5 if thr is None else thr], dtype=np.float64)
        d[tag + "_pairs"] = np.int64(n)
        for l in range(5):
            d["%s_L%d" % (tag, l)] = np.stack([f32(outs[l][a]) for a in range(9)])
    # custom anchors as the training script uses (train_retinanet_coco.py:343)
    rn2 = R.retinanet(80, anchor_sizes=[20.0, 40.0, 80.0, 160.0, 320.0])
    boxes, nbox = synth.make_boxes(1, 320, 25, 80, 8.0, 250.0, synth.seed_for(3, 3))
    g = boxes[0, :nbox[0]]
    outs, n = rn2.format_data(tf.constant(g), tf.cast([320, 320], tf.float32))
    d["a20_g"], d["a20_pairs"], d["a20_meta"] = g, np.int64(n), np.array([320, 0.5])
    d["a20_anchor_dims"] = np.array([[np.asarray(a) for a in lv] for lv in rn2.anchor_boxes], dtype=np.float32)
    for l in range(5):
        d["a20_L%d" % l] = np.stack([f32(outs[l][a]) for a in range(9)])
    np.savez_compressed(os.path.join(OUT, "retina_encode.npz"), **d)


def centernet(tf):
    s8, hg, cn = (R.load("CenterNet", m) for m in ("tf_centernet_resnet_s8", "tf_centernet_hourglass", "tf_centernet"))
    d = {}
    cases = [("c2s8", 512, (512, 512), (512, 512), 150, 1, 8, synth.seed_for(2, 0)),
             ("c2s4", 512, (512, 512), (512, 512), 150, 1, 4, synth.seed_for(2, 1)),
             ("pad", 512, (448, 448), (512, 512), 60, 3, 8, synth.seed_for(2, 2)),
             ("s16", 384, (384, 384), (384, 384), 40, 5, 16, synth.seed_for(2, 3))]
    for tag, side, img_dim, img_pad, nmax, classes, stride, seed in cases:
        boxes, nbox = synth.make_boxes(1, side, nmax, classes, 8.0, min(400.0, 0.8 * side), seed, full=tag.startswith("c2"))
        g = boxes[0, :nbox[0]]
        d[tag + "_g"] = g
        d[tag + "_meta"] = np.array([img_dim[0], img_dim[1], img_pad[0], img_pad[1], classes, stride], dtype=np.float64)
        d[tag + "_s8"] = f32(s8.format_data(tf.constant(g), SCALES, list(img_dim), classes, img_pad=list(img_pad), stride=stride)[0])
        d[tag + "_hg"] = f32(hg.format_data(tf.constant(g), list(img_dim), classes, img_pad=list(img_pad), stride=stride)[0])
        d[tag + "_cn"] = f32(cn.format_data(tf.constant(g), list(img_dim), classes, img_pad=list(img_pad), stride=stride))
    np.savez_compressed(os.path.join(OUT, "centernet_encode.npz"), **d)


def losses(tf):
    fcos, fc, fv1 = R.load("FCOS", "fcos"), R.load("FCOS", "fcos_center"), R.load("FCOS", "fcos_center_v1")
    d = {}
    for t in range(3):
        seed = synth.seed_for(5, t)
        boxes, nbox = synth.make_boxes(1, 512, 20, 20, 8.0, 300.0, seed)
        g = boxes[0, :nbox[0]]
        img = tf.cast([512, 512], tf.float32)
        pred = synth.fcos_predictions(1, 512, 20, seed)
        y_pred = [tf.constant(p) for p in pred]
        d["fcos%d_g" % t], d["fcos%d_seed" % t] = g, np.int64(seed)
        tg, _ = fcos.format_data(tf.constant(g), img, 20)
        d["fcos%d_l1" % t] = np.array([scalar(v) for v in fcos.model_loss(tg, y_pred, None)])
        d["fcos%d_iou" % t] = np.array([scalar(v) for v in fcos.model_loss(tg, y_pred, None, reg_type="iou")])
        tg, _ = fc.format_data(tf.constant(g), img, 20)
        d["fcos%d_center_focal" % t] = np.array([scalar(v) for v in fc.model_loss(tg, y_pred, cen_type="focal")])
        d["fcos%d_center_l1" % t] = np.array([scalar(v) for v in fc.model_loss(tg, y_pred)])
        tg, _ = fv1.format_data(tf.constant(g), img, 20)
        d["fcos%d_v1" % t] = np.array([scalar(v) for v in fv1.model_loss(tg, y_pred)])
    rn = R.retinanet(80)
    for t in range(2):
        seed = synth.seed_for(5, 10 + t)
        boxes, nbox = synth.make_boxes(1, 256, 12, 80, 8.0, 150.0, seed)
        g = boxes[0, :nbox[0]]
        lab, _ = rn.format_data(tf.constant(g), tf.cast([256, 256], tf.float32))
        pred = synth.retina_predictions(1, 256, 80, seed)
        rn.model = _FixedModel([[tf.constant(p[:, a]) for a in range(9)] for p in pred])
        d["retina%d_g" % t], d["retina%d_seed" % t] = g, np.int64(seed)
        d["retina%d_loss" % t] = np.array([scalar(v) for v in rn.train_loss(None, lab)])
    s8, hg = R.load("CenterNet", "tf_centernet_resnet_s8"), R.load("CenterNet", "tf_centernet_hourglass")
    seed = synth.seed_for(5, 20)
    boxes, nbox = synth.make_boxes(2, 512, 60, 3, 8.0, 400.0, seed)
    d["cn_boxes"], d["cn_nbox"], d["cn_seed"] = boxes, nbox, np.int64(seed)
    yt = np.stack([f32(s8.format_data(tf.constant(boxes[b, :nbox[b]]), SCALES, [512, 512], 3)[0]) for b in range(2)])
    yp = synth.centernet_s8_predictions(2, 512, 8, 5, 3, seed)
    d["cn_s8_loss"] = np.array([scalar(v) for v in s8.model_loss(tf.constant(yt), tf.constant(yp))])
    yt = np.stack([f32(hg.format_data(tf.constant(boxes[b, :nbox[b]]), [512, 512], 3)[0]) for b in range(2)])
    d["cn_hg_loss"] = np.array([scalar(v) for v in hg.model_loss(tf.constant(yt), tf.constant(yp[:, :, :, 0, :]))])
    # stand-alone loss functions on flat random data
    rng = np.random.default_rng(synth.seed_for(5, 30))
    x = rng.normal(-2, 3, size=(64, 40)).astype(np.float32)
    y = (rng.uniform(size=(64, 40)) < 0.1).astype(np.float32)
    d["flat_x"], d["flat_y"] = x, y
    d["flat_focal"] = scalar(fcos.focal_loss(y, tf.constant(x)))
    d["flat_focal_a4g15"] = scalar(fcos.focal_loss(y, tf.constant(x), alpha=0.4, gamma=1.5))
    a = rng.normal(0, 2, size=(16, 16, 4)).astype(np.float32)
    b = rng.normal(0, 2, size=(16, 16, 4)).astype(np.float32)
    m = (rng.uniform(size=(16, 16)) < 0.3).astype(np.float32)
    d["sl1_a"], d["sl1_b"], d["sl1_m"] = a, b, m
    d["sl1"] = scalar(fcos.smooth_l1_loss(a, tf.constant(b), mask=m))
    d["sl1_d2"] = scalar(fcos.smooth_l1_loss(a, tf.constant(b), mask=m, delta=2.0))
    d["iou_l"] = scalar(fcos.iou_loss(np.abs(a), tf.constant(np.abs(b)), m))
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **d)


GRAD_W = (1.0, 0.7, 1.3)  # weights of (cls, reg, cen) in the differentiated scalar


def _fd_gradient(loss_fn, preds, picks, h=1e-6):
    """Central differences of `loss_fn(preds) -> float` (float64) at the picked elements: picks[m] = flat indices
    into preds[m].  Returns one float64 array per map."""
    out = []
    for m, idx in enumerate(picks):
        flat = preds[m].reshape(-1)
        g = np.zeros(len(idx), dtype=np.float64)
        for n, e in enumerate(idx):
            keep = flat[e]
            flat[e] = keep + h
            up = loss_fn(preds)
            flat[e] = keep - h
            dn = loss_fn(preds)
            flat[e] = keep
            g[n] = (up - dn) / (2 * h)
        out.append(g)
    return out


def _grad_picks(targets, ch, cls0, rng, n_pos_rows=6, n_other=24):
    """Per map: every channel of a few positive rows (they carry the regression / centerness / labelled-class
    derivatives) plus random other elements (label-0 focal and all-location centerness terms)."""
    picks = []
    for t in targets:
        rows = t.reshape(-1, ch)
        pos = np.nonzero(rows[:, cls0:].max(axis=1) > 0)[0]
        if len(pos) > n_pos_rows:
            pos = rng.choice(pos, n_pos_rows, replace=False)
        idx = [r * ch + c for r in pos for c in range(ch)]
        idx += list(rng.integers(0, rows.size, size=min(n_other, rows.size)))
        picks.append(np.unique(np.asarray(idx, dtype=np.int64)))
    return picks


def grad(tf):
    """d(w . (cls, reg, cen)) / d pred by central differences THROUGH THE REFERENCE'S OWN LOSS FUNCTIONS
    (fcos.model_loss FCOS/fcos.py:464-496 and its fcos_center / fcos_center_v1 copies, RetinaNet.train_loss
    RetinaNet/retinanet_module.py:403-426, the CenterNet model_loss pair) -- the callers obtain this gradient from
    tf.GradientTape (FCOS/train_fcos.py:152-176); TensorFlow's autodiff is not available here, finite differences of
    the reference's code are.  The stub's `tf.float32` is switched to float64 for the duration, so the reference's
    formulas are evaluated in double precision and a step of 1e-6 gives derivatives good to ~1e-9."""
    fcos, fc, fv1 = R.load("FCOS", "fcos"), R.load("FCOS", "fcos_center"), R.load("FCOS", "fcos_center_v1")
    s8, hg = R.load("CenterNet", "tf_centernet_resnet_s8"), R.load("CenterNet", "tf_centernet_hourglass")
    rn = R.retinanet(20)
    d = {"weights": np.array(GRAD_W)}
    rng = np.random.default_rng(synth.seed_for(7, 0))
    w = GRAD_W
    img = tf.cast([256, 256], tf.float32)
    seed = synth.seed_for(7, 1)
    boxes, nbox = synth.make_boxes(1, 256, 12, 20, 8.0, 180.0, seed)
    g = boxes[0, :nbox[0]]
    d["fcos_g"], d["fcos_seed"] = g, np.int64(seed)
    targets = {"fcos": fcos.format_data(tf.constant(g), img, 20)[0], "center": fc.format_data(tf.constant(g), img, 20)[0],
               "v1": fv1.format_data(tf.constant(g), img, 20)[0]}
    rseed = synth.seed_for(7, 2)
    rboxes, rnbox = synth.make_boxes(1, 128, 8, 20, 8.0, 100.0, rseed)
    rg = rboxes[0, :rnbox[0]]
    rlab, _ = rn.format_data(tf.constant(rg), tf.cast([128, 128], tf.float32))
    d["retina_g"], d["retina_seed"] = rg, np.int64(rseed)
    cseed = synth.seed_for(7, 3)
    cboxes, cnbox = synth.make_boxes(2, 256, 30, 3, 8.0, 200.0, cseed)
    d["cn_boxes"], d["cn_nbox"], d["cn_seed"] = cboxes, cnbox, np.int64(cseed)
    yt_s8 = np.stack([np.asarray(s8.format_data(tf.constant(cboxes[b, :cnbox[b]]), SCALES, [256, 256], 3)[0]) for b in range(2)])
    yt_hg = np.stack([np.asarray(hg.format_data(tf.constant(cboxes[b, :cnbox[b]]), [256, 256], 3)[0]) for b in range(2)])
    with R.float64_mode():
        # ---- FCOS family: five model_loss variants on the same predictions ---------------------------------
        base = [p.astype(np.float64) for p in synth.fcos_predictions(1, 256, 20, seed)]
        for p in base:
            p[..., :4] = np.abs(p[..., :4]) + 0.3  # tblr distances are positive in a trained head (IoU loss)
        variants = (("fcos_l1", fcos.model_loss, targets["fcos"], (None,), {}),
                    ("fcos_iou", fcos.model_loss, targets["fcos"], (None,), {"reg_type": "iou"}),
                    ("center_focal", fc.model_loss, targets["center"], (), {"cen_type": "focal"}),
                    ("center_l1", fc.model_loss, targets["center"], (), {}),
                    ("v1", fv1.model_loss, targets["v1"], (), {}))
        for name, fn, tg, args, kw in variants:
            preds = [p.copy() for p in base]
            tgt = [np.asarray(t, dtype=np.float64) for t in tg]

            def loss(ps, fn=fn, tgt=tgt, args=args, kw=kw):
                out = fn(tgt, [tf.constant(p) for p in ps], *args, **kw)
                return sum(wk * float(np.asarray(v.numpy() if hasattr(v, "numpy") else v)) for wk, v in zip(w, out))
            picks = _grad_picks(tgt, 25, 5, rng)
            fd = _fd_gradient(loss, preds, picks)
            for l in range(5):
                d["%s_idx%d" % (name, l)], d["%s_fd%d" % (name, l)] = picks[l], fd[l]
        # ---- RetinaNet.train_loss -----------------------------------------------------------------------------
        rbase = [p.astype(np.float64) for p in synth.retina_predictions(1, 128, 20, rseed)]
        rt = [[np.asarray(rlab[l][a], dtype=np.float64) for a in range(9)] for l in range(5)]

        def rloss(ps):
            rn.model = _FixedModel([[tf.constant(p[:, a]) for a in range(9)] for p in ps])
            out = rn.train_loss(None, rt)
            return sum(wk * float(np.asarray(v.numpy() if hasattr(v, "numpy") else v)) for wk, v in zip(w, out))
        picks = _grad_picks([np.stack(rt[l]) for l in range(5)], 24, 4, rng, n_pos_rows=5, n_other=20)
        fd = _fd_gradient(rloss, rbase, picks)
        for l in range(5):
            d["retina_idx%d" % l], d["retina_fd%d" % l] = picks[l], fd[l]
        # ---- CenterNet s8 / hourglass model_loss (batched) -----------------------------------------------------
        yp = synth.centernet_s8_predictions(2, 256, 8, 5, 3, cseed).astype(np.float64)
        for name, fn, yt, pr in (("cn_s8", s8.model_loss, yt_s8, yp), ("cn_hg", hg.model_loss, yt_hg, np.ascontiguousarray(yp[:, :, :, 0, :]))):
            tgt = yt.astype(np.float64)

            def closs(ps, fn=fn, tgt=tgt):
                out = fn(tf.constant(tgt), tf.constant(ps[0]))
                return sum(wk * float(np.asarray(v.numpy() if hasattr(v, "numpy") else v)) for wk, v in zip(w, out))
            picks = _grad_picks([tgt], 7, 4, rng, n_pos_rows=12, n_other=60)
            fd = _fd_gradient(closs, [pr.copy()], picks)
            d[name + "_idx"], d[name + "_fd"] = picks[0], fd[0]
    np.savez_compressed(os.path.join(OUT, "grad.npz"), **d)


def decode_nms(tf):
    fcos, fv1 = R.load("FCOS", "fcos"), R.load("FCOS", "fcos_center_v1")
    s8 = R.load("CenterNet", "tf_centernet_resnet_s8")
    rn = R.retinanet(80)
    d = {}
    seed = synth.seed_for(4, 0)
    p = synth.fcos_predictions(1, 128, 20, seed)[0][0]           # [16,16,25]
    d["p2c_seed"] = np.int64(seed)
    d["p2c_fcos"] = f32(fcos.prediction_to_corners(p[..., :4], 8))
    d["p2c_v1"] = f32(fv1.prediction_to_corners(p[..., :4], 64, 8))
    d["p2c_retina"] = f32(rn.prediction_to_corners(p[..., :4], rn.anchor_boxes[0][4], 8))
    yp = synth.centernet_s8_predictions(1, 128, 8, 5, 3, seed)[0]
    d["p2c_s8"] = f32(s8.prediction_to_corners(yp[..., :4], SCALES, 8))
    # FCOS decode front end: capture what infer_fcos.image_detections hands to the TF NMS op
    captured = {}

    class _Image:
        @staticmethod
        def combined_non_max_suppression(boxes, scores, *a, **k):
            captured["boxes"], captured["scores"], captured["args"] = np.asarray(boxes), np.asarray(scores), (a, k)
            return None
    stub = R.tf()
    saved_image = stub.image  # the stub's own tf.image (data_preprocess.random_flip_horizontal needs it in prep())
    stub.image = _Image
    ns = R.functions_only("FCOS", "infer_fcos", ("image_detections",))
    for center in (False, True):
        heads = synth.fcos_predictions(1, 128, 20, seed + 1)
        ns["image_detections"](None, _FixedModel([tf.constant(h) for h in heads]), 20, center=center)
        d["fcos_dec_boxes_c%d" % center] = f32(captured["boxes"][0, :, 0, :])
        d["fcos_dec_scores_c%d" % center] = f32(captured["scores"][0])
    d["fcos_dec_seed"] = np.int64(seed + 1)
    stub.image = saved_image
    # RetinaNet image_detections (decode + threshold + class-agnostic NMS); unique scores enforced
    t = 0
    for s in range(50):
        sd = synth.seed_for(4, 10 + s)
        pr = synth.retina_predictions(1, 128, 80, sd, logit_sigma=2.5)
        rn.model = _FixedModel([[tf.constant(x[:, a]) for a in range(9)] for x in pr])
        dets = rn.image_detections(None, iou_thresh=0.5, cls_thresh=0.05)
        flat_scores = np.concatenate([(1 / (1 + np.exp(-x[..., 4:].astype(np.float32)))).max(-1).reshape(-1) for x in pr])
        cand = flat_scores[flat_scores >= 0.05]
        if len(np.unique(cand)) != len(cand):
            continue                                   # tie -> reference order is unspecified (unstable argsort)
        d["retina_det%d_seed" % t], d["retina_det%d" % t] = np.int64(sd), f32(dets)
        t += 1
        if t == 2:
            break
    assert t == 2
    for n in (64, 700, 3000):
        dets = synth.nms_candidates(n, 640, synth.seed_for(4, 60) + n)
        d["cpu_nms_%d" % n] = np.asarray(rn.cpu_nms(dets, 0.5), dtype=np.int64)
        d["cpu_nms_%d_t3" % n] = np.asarray(rn.cpu_nms(dets, 0.3), dtype=np.int64)
        bb = dets.astype(np.float64)
        bb = np.stack([bb[:, 1], bb[:, 0], bb[:, 3] - bb[:, 1], bb[:, 2] - bb[:, 0],
                       np.floor(bb[:, 4] * 100), bb[:, 5]], axis=1)       # (x, y, w, h, int score, class)
        for method in ("nms", "soft-nms"):
            rows = np.array(s8.nms(bb.copy(), 0.5, method=method)).reshape(-1, 6)
            d["cnms_%d_%s" % (n, method)] = rows[np.argsort(rows[:, 5], kind="stable")]   # class-ascending canon
    np.savez_compressed(os.path.join(OUT, "decode_nms.npz"), **d)


def prep(tf):
    """Label preparation helpers (FCOS/utils.py, FCOS/data_preprocess.py) executed from the reference."""
    utils = R.load("FCOS", "utils")
    dp = R.load("FCOS", "data_preprocess")
    rng = np.random.default_rng(synth.seed_for(6, 1))
    lo = rng.uniform(0.0, 0.7, size=(40, 2)).astype(np.float32)
    raw = np.concatenate([lo, lo + rng.uniform(0.02, 0.3, size=(40, 2)).astype(np.float32)], axis=1)   # xmin, ymin, xmax, ymax
    d = {"raw": raw}
    d["swap_xy"] = f32(utils.swap_xy(tf.constant(raw)))
    d["to_xywh"] = f32(utils.convert_to_xywh(tf.constant(raw)))
    d["to_corners"] = f32(utils.convert_to_corners(tf.constant(raw)))
    img = tf.constant(np.zeros((4, 6, 3), np.float32))
    _, flipped = dp.random_flip_horizontal(img, tf.constant(raw), p_flip=1.0)
    _, kept = dp.random_flip_horizontal(img, tf.constant(raw), p_flip=-1.0)
    d["flipped"] = f32(flipped)
    assert np.array_equal(f32(kept), raw)
    d["labels_flipped"] = f32(utils.convert_to_xywh(utils.swap_xy(flipped)))
    d["labels_plain"] = f32(utils.convert_to_xywh(utils.swap_xy(tf.constant(raw))))
    np.savez_compressed(os.path.join(OUT, "prep.npz"), **d)


def hourglass4(tf):
    """The 4-scale encoder that CenterNet/train_hourglass_voc.py:95-153 keeps inline in train(), sliced out of the
    reference file and executed unmodified (oracle/ref_loader.hourglass_inline_encoder)."""
    from oracle import dense_head_ref as O
    d = {}
    for t, raw_dims in enumerate([320, 288]):
        img_dims = raw_dims if raw_dims % 64 == 0 else (raw_dims // 64 + 1) * 64          # :87-92
        pad = int((img_dims - raw_dims) / 2.0)
        boxes, nbox = synth.make_boxes(1, raw_dims, 25, 5, 6.0, 0.95 * raw_dims, synth.seed_for(6, 20 + t))
        g = boxes[0, :nbox[0]]
        bb = np.stack([g[:, 1] - g[:, 3] / 2, g[:, 0] - g[:, 2] / 2, g[:, 1] + g[:, 3] / 2, g[:, 0] + g[:, 2] / 2], 1).astype(np.float32)
        ref = R.hourglass_inline_encoder([{"objects": {"bbox": bb, "label": g[:, 4].astype(np.int64)}}], 5, raw_dims, img_dims, pad)[0]
        xywh = O.convert_to_xywh(bb)            # what the reference derives from the dataset boxes (:104)
        d["hg4_%d_labels" % t] = np.stack([xywh[:, 1], xywh[:, 0], xywh[:, 3], xywh[:, 2], g[:, 4]], 1).astype(np.float32)
        d["hg4_%d_dims" % t] = np.array([raw_dims, img_dims])
        d["hg4_%d_map" % t] = f32(ref)
    np.savez_compressed(os.path.join(OUT, "hourglass4.npz"), **d)


SPARSE_LABELS = [(1, "person"), (2, "bicycle"), (3, "car"), (7, "zebra"), (9, "apple")]
SPARSE_OBJECTS = [  # filename, img_width, img_height, category id, x_lower, y_lower, box_width, box_height
    ("a.jpg", 640, 480, 3, 100.25, 50.5, 30.0, 45.75), ("a.jpg", 640, 480, 1, 300.0, 200.0, 120.0, 100.0),
    ("a.jpg", 640, 480, 7, 301.5, 201.25, 40.0, 30.0),                       # overlaps the previous one: both emit
    ("a.jpg", 640, 480, 9, 600.0, 440.0, 90.0, 90.0),                        # runs off the canvas: cut by the slice
    ("b.jpg", 500, 375, 7, 10.2, 20.7, 30.3, 15.1), ("b.jpg", 500, 375, 2, 480.2, 360.7, 30.3, 35.1),
    ("b.jpg", 500, 375, 1, 100.0, 100.0, 31.25, 23.4375),                    # exactly 28 wide on the canvas: NOT < 28, slot 1
    ("b.jpg", 500, 375, 3, 50.0, 50.0, -4.0, 10.0),                          # negative width: skipped
    ("b.jpg", 500, 375, 9, 200.0, 200.0, 0.5, 0.5),                          # empty footprint
    ("c.jpg", 448, 448, 2, -6.0, 10.0, 3.0, 20.0),                           # negative corner: NumPy slices wrap around
    ("c.jpg", 448, 448, 3, -30.0, -40.0, 60.0, 70.0), ("c.jpg", 448, 448, 1, 20.0, 30.0, 230.0, 60.0),
    ("d.jpg", 896, 224, 9, 100.0, 20.0, 500.0, 150.0),                       # slot 4 (the last, unconditional)
]


def sparse_fcos(tf):
    """The offline formatter /format_COCO_annotations_fcos.py run unmodified on a small annotation table
    (oracle/ref_loader.offline_fcos_formatter).  Per image: the kernel-side inputs (rows (x_lower, y_lower, box_width,
    box_height, label) with the label the script's tables produce, source dims), the number of entries, and either the
    entries themselves or -- for the image with a slot-4 box -- their SHA-256."""
    import hashlib
    import pandas as pd
    labels = pd.DataFrame(SPARSE_LABELS, columns=["id", "name"])
    objects = pd.DataFrame(SPARSE_OBJECTS, columns=["filename", "img_width", "img_height", "id", "x_lower", "y_lower",
                                                    "box_width", "box_height"])
    out = R.offline_fcos_formatter(objects, labels)
    names = ["objectness"] + sorted(labels["name"])
    d = {"files": np.array([f for f, _, _ in out])}
    for f, dims, sp in out:
        key = f.split(".")[0]
        sub = objects[objects.filename == f]
        lab = [names.index(labels[labels.id == i].iloc[0]["name"]) for i in sub.id]
        d[key + "_objects"] = np.column_stack([sub.x_lower, sub.y_lower, sub.box_width, sub.box_height, lab]).astype(np.float64)
        d[key + "_src_dims"] = np.array([sub.img_width.iloc[0], sub.img_height.iloc[0]], np.float64)
        d[key + "_img_dims"] = np.array(dims)
        d[key + "_dense_shape"] = np.array(sp.dense_shape)
        assert all(len(i) == 4 or (len(i) == 5 and i[2] == i[3]) for i in sp.indices)
        idx = np.array([i if len(i) == 4 else [i[0], i[1], i[2], i[4]] for i in sp.indices], np.int32)  # (:171: one index too many)
        val = np.array([float(v) for v in sp.values], np.float32)
        d[key + "_nnz"] = np.array(len(val))
        if len(val) <= 200000:
            d[key + "_indices"], d[key + "_values"] = idx, val
        d[key + "_sha"] = np.array(hashlib.sha256(idx.tobytes() + val.tobytes()).hexdigest())
    np.savez_compressed(os.path.join(OUT, "sparse_fcos.npz"), **d)


def main(out_dir=None):
    """Rewrite every fixture (into `out_dir` when given: tests/test_make_golden.py runs the whole recipe into a
    temporary directory and compares the bytes of the arrays with the committed files)."""
    global OUT
    if not R.available():
        raise SystemExit("reference tree not found at %s" % R.REF_ROOT)
    if out_dir is not None:
        OUT = out_dir
    os.makedirs(OUT, exist_ok=True)
    tf = R.tf()
    for fn in (kat, fcos_family, retina, centernet, losses, grad, decode_nms, prep, hourglass4, sparse_fcos):
        fn(tf)
        print("wrote", fn.__name__)
    for f in sorted(os.listdir(OUT)):
        print("%9d  %s" % (os.path.getsize(os.path.join(OUT, f)), f))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
