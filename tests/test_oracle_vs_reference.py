"""Live cross-check of the CPU oracle against the reference's own source executed under the TF stub.
Needs /root/reference (the build container); skipped on the GPU box.  The frozen outputs of the same
reference run are what tests/test_oracle_golden.py checks everywhere."""
import numpy as np
import pytest

from oracle import dense_head_ref as O
from oracle import ref_loader as R
from oracle import synth

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")
SCALES = [32, 64, 128, 256, 512]


def _eq(a, b):
    return np.array_equal(np.asarray(a, np.float32), np.asarray(b, np.float64).astype(np.float32))


@pytest.mark.parametrize("trial", range(8))
def test_fcos_family_random(trial):
    tf = R.tf()
    fcos, fc, fv1 = R.load("FCOS", "fcos"), R.load("FCOS", "fcos_center"), R.load("FCOS", "fcos_center_v1")
    side = [512, 384, 640][trial % 3]
    boxes, nbox = synth.make_boxes(1, side, 20, 20, 4.0, 0.9 * side, 7000 + trial)
    g = boxes[0, :nbox[0]]
    dim = [side - 64.0 * (trial % 2), float(side)]
    pad = [side, side + 128 * (trial % 2)]
    lab, img = tf.constant(g), tf.cast(dim, tf.float32)
    for ref_fn, ora_fn, kw in ((fcos.format_data, O.fcos_format_data, {}), (fc.format_data, O.fcos_center_format_data, {}),
                               (fc.format_data, O.fcos_center_format_data, {"center_only": True}),
                               (fv1.format_data, O.fcos_center_v1_format_data, {})):
        ro, rn = ref_fn(lab, img, 20, img_pad=pad, **kw)
        oo, on = ora_fn(g, dim, 20, img_pad=pad, **kw)
        assert rn == on and all(_eq(a, b) for a, b in zip(oo, ro))


@pytest.mark.parametrize("trial", range(3))
def test_retina_random(trial):
    tf = R.tf()
    rn_ = R.retinanet(80)
    side = [256, 384, 320][trial]
    boxes, nbox = synth.make_boxes(1, side, 15, 80, 8.0, 0.6 * side, 7100 + trial)
    g = boxes[0, :nbox[0]]
    ro, rn = rn_.format_data(tf.constant(g), tf.cast([side, side], tf.float32), iou_thresh=[0.5, 0.4, 0.6][trial])
    oo, on = O.retina_format_data(g, [side, side], 80, iou_thresh=[0.5, 0.4, 0.6][trial])
    assert rn == on
    for l in range(5):
        for a in range(9):
            assert _eq(oo[l][a], ro[l][a])


@pytest.mark.parametrize("trial", range(6))
def test_centernet_random(trial):
    tf = R.tf()
    s8, hg, cn = (R.load("CenterNet", m) for m in ("tf_centernet_resnet_s8", "tf_centernet_hourglass", "tf_centernet"))
    stride = [8, 4, 16][trial % 3]
    boxes, nbox = synth.make_boxes(1, 512, 60, 3, 8.0, 400, 7200 + trial)
    g = boxes[0, :nbox[0]]
    dim, pad = ([512, 512], [512, 512]) if trial < 3 else ([448, 448], [512, 512])
    assert _eq(O.centernet_s8_format_data(g, SCALES, dim, 3, img_pad=pad, stride=stride)[0],
               s8.format_data(tf.constant(g), SCALES, dim, 3, img_pad=pad, stride=stride)[0])
    assert _eq(O.centernet_hourglass_format_data(g, dim, 3, img_pad=pad, stride=stride)[0],
               hg.format_data(tf.constant(g), dim, 3, img_pad=pad, stride=stride)[0])
    assert _eq(O.centernet_format_data(g, dim, 3, img_pad=pad, stride=stride),
               cn.format_data(tf.constant(g), dim, 3, img_pad=pad, stride=stride))


def test_losses_and_nms_random():
    tf = R.tf()
    fcos = R.load("FCOS", "fcos")
    rn_ = R.retinanet(80)
    s8 = R.load("CenterNet", "tf_centernet_resnet_s8")
    boxes, nbox = synth.make_boxes(1, 512, 20, 20, 6.0, 400, 7300)
    g = boxes[0, :nbox[0]]
    tg, _ = fcos.format_data(tf.constant(g), tf.cast([512, 512], tf.float32), 20)
    pred = synth.fcos_predictions(1, 512, 20, 7300)
    for reg_type in ("l1", "iou"):
        r = fcos.model_loss(tg, [tf.constant(p) for p in pred], None, reg_type=reg_type)
        o = O.fcos_model_loss([x.astype(np.float32) for x in tg], [p[0] for p in pred], reg_type=reg_type)
        for a, b in zip(o, r):
            b = float(np.asarray(b.numpy() if hasattr(b, "numpy") else b))
            assert abs(float(a) - b) <= 1e-5 * max(1.0, abs(b))
    d = synth.nms_candidates(800, 640, 7301)
    assert np.array_equal(rn_.cpu_nms(d, 0.5), O.cpu_nms(d, 0.5))
    bb = d.astype(np.float64)
    bb = np.stack([bb[:, 1], bb[:, 0], bb[:, 3] - bb[:, 1], bb[:, 2] - bb[:, 0], np.floor(bb[:, 4] * 100), bb[:, 5]], axis=1)
    for method in ("nms", "soft-nms"):
        rr = np.array(s8.nms(bb.copy(), 0.5, method=method)).reshape(-1, 6)
        oo, _ = O.centernet_nms(bb, 0.5, method=method)
        assert np.allclose(rr[np.argsort(rr[:, 5], kind="stable")], oo, rtol=1e-12)


@pytest.mark.parametrize("trial", range(6))
def test_hourglass4_inline_encoder_random(trial):
    raw_dims = [320, 288, 416, 224, 352, 256][trial]
    img_dims = raw_dims if raw_dims % 64 == 0 else (raw_dims // 64 + 1) * 64
    pad = int((img_dims - raw_dims) / 2.0)
    boxes, nbox = synth.make_boxes(1, raw_dims, 25, 5, 6.0, 0.95 * raw_dims, 7400 + trial)
    g = boxes[0, :nbox[0]]
    bb = np.stack([g[:, 1] - g[:, 3] / 2, g[:, 0] - g[:, 2] / 2, g[:, 1] + g[:, 3] / 2, g[:, 0] + g[:, 2] / 2], 1).astype(np.float32)
    ref = R.hourglass_inline_encoder([{"objects": {"bbox": bb, "label": g[:, 4].astype(np.int64)}}], 5, raw_dims, img_dims, pad)[0]
    xywh = O.convert_to_xywh(bb)
    gl = np.stack([xywh[:, 1], xywh[:, 0], xywh[:, 3], xywh[:, 2], g[:, 4]], 1).astype(np.float32)
    assert _eq(O.hourglass4_format_data(gl, raw_dims, img_dims, 5), ref)


def test_label_prep_against_reference_utils():
    tf = R.tf()
    utils, dp = R.load("FCOS", "utils"), R.load("FCOS", "data_preprocess")
    rng = np.random.default_rng(3)
    lo = rng.uniform(0, 0.7, size=(30, 2)).astype(np.float32)
    raw = np.concatenate([lo, lo + rng.uniform(0.02, 0.3, size=(30, 2)).astype(np.float32)], axis=1)
    assert _eq(O.swap_xy(raw), utils.swap_xy(tf.constant(raw)))
    assert _eq(O.convert_to_xywh(raw), utils.convert_to_xywh(tf.constant(raw)))
    assert _eq(O.convert_to_corners(raw), utils.convert_to_corners(tf.constant(raw)))
    _, fl = dp.random_flip_horizontal(tf.constant(np.zeros((2, 2, 3), np.float32)), tf.constant(raw), p_flip=1.0)
    assert _eq(O.flip_boxes_horizontal(raw), fl)


def test_offline_sparse_formatter_random():
    """/format_COCO_annotations_fcos.py run unmodified on a random annotation table (oracle/ref_loader.offline_fcos_formatter)."""
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(11)
    labels = pd.DataFrame([(3, "cat"), (5, "bus"), (8, "kite"), (13, "oven")], columns=["id", "name"])
    names = ["objectness"] + sorted(labels["name"])
    rows = []
    for f, (w, h) in (("p.jpg", (640, 427)), ("q.jpg", (333, 500)), ("r.jpg", (448, 448))):
        for _ in range(5):
            rows.append((f, w, h, int(rng.choice(labels["id"])), round(float(rng.uniform(-5, w * 0.9)), 2), round(float(rng.uniform(-5, h * 0.9)), 2),
                         round(float(rng.uniform(1, 90)), 2), round(float(rng.uniform(1, 90)), 2)))
    objects = pd.DataFrame(rows, columns=["filename", "img_width", "img_height", "id", "x_lower", "y_lower", "box_width", "box_height"])
    with np.errstate(invalid="ignore", divide="ignore"):
        out = R.offline_fcos_formatter(objects, labels)
    assert len(out) == 3
    for f, dims, sp in out:
        sub = objects[objects.filename == f]
        lab = [names.index(labels[labels.id == i].iloc[0]["name"]) for i in sub.id]
        idx, val = O.fcos_sparse_format(np.column_stack([sub.x_lower, sub.y_lower, sub.box_width, sub.box_height, lab]),
                                        (sub.img_width.iloc[0], sub.img_height.iloc[0]), dims)
        ref_idx = np.array([i if len(i) == 4 else [i[0], i[1], i[2], i[4]] for i in sp.indices], np.int32).reshape(-1, 4)
        ref_val = np.array([float(v) for v in sp.values], np.float32)
        assert np.array_equal(idx, ref_idx) and np.array_equal(val, ref_val, equal_nan=True)
