"""The CPU oracle (`oracle/dense_head_ref.py`) against golden vectors frozen from the reference's
own source (`oracle/make_golden.py`).  Runs everywhere, no GPU, no /root/reference needed."""
import numpy as np
import pytest

from oracle import dense_head_ref as O
from oracle import synth
from conftest import assert_close

SCALES = [32, 64, 128, 256, 512]


def _rel(a, b):
    return abs(float(a) - float(b)) / max(1e-12, abs(float(b)))


def test_kat_fcos_family(golden):
    k = golden("kat")
    g = k["g"]
    for name, fn, kw in (("fcos", O.fcos_format_data, {}), ("center", O.fcos_center_format_data, {}),
                         ("center_only", O.fcos_center_format_data, {"center_only": True}),
                         ("v1", O.fcos_center_v1_format_data, {})):
        outs, cnt = fn(g, [512, 512], 20, img_pad=[512, 512], **kw)
        assert cnt == k[name + "_counts"].tolist()
        for l, o in enumerate(outs):
            assert np.array_equal(o, k["%s_L%d" % (name, l)]), (name, l)
    # Appendix B spot values
    l1 = O.fcos_format_data(g, [512, 512], 20)[0][1]
    assert l1[12, 12, :5].tolist() == [1.875, 1.875, 1.875, 1.875, 1.0]
    assert np.nonzero(l1[12, 12, 5:])[0].tolist() == [3, 7]


def test_kat_losses_iou_nms(golden):
    k = golden("kat")
    assert _rel(O.focal_loss([0, 1, 0, 1, 1.], [-3, -.5, 0, .5, 3.]), k["focal"]) < 1e-6
    assert _rel(O.focal_loss([0, 1, 0, 1, 1.], [-3, -.5, 0, .5, 3.]), 0.241320652524) < 1e-6
    assert _rel(O.smooth_l1_loss([[0, 1, 2, 3.]], [[.5, 1, 4, 2.2]]), 2.445) < 1e-6
    yt = np.zeros((2, 2, 4), np.float32); yt[1, 1] = (1, 2, 1.5, .5)
    m = np.zeros((2, 2), np.float32); m[1, 1] = 1
    assert _rel(O.iou_loss(yt, np.ones((2, 2, 4), np.float32), m), k["iou_loss"]) < 1e-6
    assert np.array_equal(O.compute_iou(k["iou_b1"], k["iou_b2"]), k["iou"])
    assert np.array_equal(O.cpu_nms(k["nms_dets"], .5), k["nms_keep"])
    assert O.cpu_nms(k["nms_dets"], .5).tolist() == [0, 2]
    rows, src = O.centernet_nms(k["cnms_in"], .5)
    assert np.array_equal(rows, k["cnms_hard"]) and src.tolist() == [0, 2, 3]
    rows, _ = O.centernet_nms(k["cnms_in"], .5, method="soft-nms")
    assert np.allclose(rows, k["cnms_soft"], rtol=1e-12)


def test_kat_retina_and_centernet(golden):
    k = golden("kat")
    assert np.array_equal(O.retina_anchor_dims(), k["retina_anchor_dims"])
    outs, n = O.retina_format_data(k["retina_g"], [256, 256], 80)
    assert n == int(k["retina_pairs"]) == 40
    for a in range(9):
        assert np.array_equal(outs[0][a], k["retina_L0_A%d" % a])
    assert sum(float(np.abs(outs[l][a]).sum()) for l in range(1, 5) for a in range(9)) == 0.0
    out, n = O.centernet_s8_format_data(k["s8_g"], SCALES, [512, 512], 1)
    assert n == 3 and np.array_equal(out, k["s8"])
    with pytest.raises(ValueError):
        O.centernet_s8_format_data(np.array([[.5, .5, 1., 1., 0]], np.float32), SCALES, [512, 512], 1)


@pytest.mark.parametrize("tag", ["c1", "c1b", "s384", "pad", "tiny", "coco"])
def test_fcos_family_golden(golden, tag):
    z = golden("fcos_encode")
    g, meta = z[tag + "_g"], z[tag + "_meta"]
    img_dim, img_pad, classes = [meta[0], meta[1]], [int(meta[2]), int(meta[3])], int(meta[4])
    for name, fn, kw in (("fcos", O.fcos_format_data, {}), ("center", O.fcos_center_format_data, {}),
                         ("center_only", O.fcos_center_format_data, {"center_only": True}),
                         ("v1", O.fcos_center_v1_format_data, {})):
        outs, cnt = fn(g, img_dim, classes, img_pad=img_pad, **kw)
        assert cnt == z["%s_%s_counts" % (tag, name)].tolist()
        for l, o in enumerate(outs):
            assert np.array_equal(o, z["%s_%s_L%d" % (tag, name, l)]), (tag, name, l)


@pytest.mark.parametrize("tag", ["s256", "c3", "thr4", "a20"])
def test_retina_golden(golden, tag):
    z = golden("retina_encode")
    side, thr = int(z[tag + "_meta"][0]), float(z[tag + "_meta"][1])
    dims = z["a20_anchor_dims"] if tag == "a20" else None
    if tag == "a20":
        assert np.array_equal(O.retina_anchor_dims(anchor_sizes=[20., 40., 80., 160., 320.]), dims)
    outs, n = O.retina_format_data(z[tag + "_g"], [side, side], 80, anchor_dims=dims, iou_thresh=thr)
    assert n == int(z[tag + "_pairs"])
    for l in range(5):
        assert np.array_equal(np.stack(outs[l]), z["%s_L%d" % (tag, l)]), (tag, l)


@pytest.mark.parametrize("tag", ["c2s8", "c2s4", "pad", "s16"])
def test_centernet_golden(golden, tag):
    z = golden("centernet_encode")
    g, m = z[tag + "_g"], z[tag + "_meta"]
    img_dim, img_pad, classes, stride = [int(m[0]), int(m[1])], [int(m[2]), int(m[3])], int(m[4]), int(m[5])
    out, n = O.centernet_s8_format_data(g, SCALES, img_dim, classes, img_pad=img_pad, stride=stride)
    assert n == len(g) and np.array_equal(out, z[tag + "_s8"])
    out, n = O.centernet_hourglass_format_data(g, img_dim, classes, img_pad=img_pad, stride=stride)
    assert np.array_equal(out, z[tag + "_hg"])
    out = O.centernet_format_data(g, img_dim, classes, img_pad=img_pad, stride=stride)
    assert np.array_equal(out, z[tag + "_cn"])


def test_losses_golden(golden):
    z = golden("losses")
    for t in range(3):
        g, seed = z["fcos%d_g" % t], int(z["fcos%d_seed" % t])
        pred = [p[0] for p in synth.fcos_predictions(1, 512, 20, seed)]
        tg, _ = O.fcos_format_data(g, [512, 512], 20)
        assert_close(O.fcos_model_loss(tg, pred), z["fcos%d_l1" % t], what="fcos l1")
        assert_close(O.fcos_model_loss(tg, pred, reg_type="iou"), z["fcos%d_iou" % t], what="fcos iou")
        tg, _ = O.fcos_center_format_data(g, [512, 512], 20)
        assert_close(O.fcos_model_loss(tg, pred, cen_type="focal"), z["fcos%d_center_focal" % t], what="center focal")
        assert_close(O.fcos_model_loss(tg, pred), z["fcos%d_center_l1" % t], what="center l1")
        tg, _ = O.fcos_center_v1_format_data(g, [512, 512], 20)
        assert_close(O.fcos_model_loss(tg, pred, cen_type="focal"), z["fcos%d_v1" % t], what="v1")
    for t in range(2):
        g, seed = z["retina%d_g" % t], int(z["retina%d_seed" % t])
        lab, _ = O.retina_format_data(g, [256, 256], 80)
        pred = synth.retina_predictions(1, 256, 80, seed)
        assert_close(O.retina_train_loss(lab, [[p[0, a] for a in range(9)] for p in pred]),
                     z["retina%d_loss" % t], what="retina loss")
    boxes, nbox, seed = z["cn_boxes"], z["cn_nbox"], int(z["cn_seed"])
    yp = synth.centernet_s8_predictions(2, 512, 8, 5, 3, seed)
    yt = np.stack([O.centernet_s8_format_data(boxes[b, :nbox[b]], SCALES, [512, 512], 3)[0] for b in range(2)])
    assert_close(O.centernet_s8_model_loss(yt, yp), z["cn_s8_loss"], what="s8 loss")
    yt = np.stack([O.centernet_hourglass_format_data(boxes[b, :nbox[b]], [512, 512], 3)[0] for b in range(2)])
    assert_close(O.centernet_hourglass_model_loss(yt, yp[:, :, :, 0, :]), z["cn_hg_loss"], what="hg loss")
    assert_close(O.focal_loss(z["flat_y"], z["flat_x"]), z["flat_focal"], what="flat focal")
    assert_close(O.focal_loss(z["flat_y"], z["flat_x"], alpha=0.4, gamma=1.5), z["flat_focal_a4g15"], what="focal a/g")
    assert_close(O.smooth_l1_loss(z["sl1_a"], z["sl1_b"], z["sl1_m"]), z["sl1"], what="sl1")
    assert_close(O.smooth_l1_loss(z["sl1_a"], z["sl1_b"], z["sl1_m"], delta=2.0), z["sl1_d2"], what="sl1 d2")
    assert_close(O.iou_loss(np.abs(z["sl1_a"]), np.abs(z["sl1_b"]), z["sl1_m"]), z["iou_l"], what="iou loss")


def test_decode_golden(golden):
    z = golden("decode_nms")
    seed = int(z["p2c_seed"])
    p = synth.fcos_predictions(1, 128, 20, seed)[0][0]
    assert np.array_equal(O.fcos_prediction_to_corners(p[..., :4], 8), z["p2c_fcos"])
    assert np.array_equal(O.fcos_center_v1_prediction_to_corners(p[..., :4], 64, 8), z["p2c_v1"])
    assert np.array_equal(O.retina_prediction_to_corners(p[..., :4], O.retina_anchor_dims()[0, 4], 8), z["p2c_retina"])
    yp = synth.centernet_s8_predictions(1, 128, 8, 5, 3, seed)[0]
    assert np.array_equal(O.centernet_s8_prediction_to_corners(yp[..., :4], SCALES, 8), z["p2c_s8"])
    heads = [h[0] for h in synth.fcos_predictions(1, 128, 20, int(z["fcos_dec_seed"]))]
    for center in (0, 1):
        boxes, scores = O.fcos_decode_scores(heads, 20, center=bool(center))
        assert np.array_equal(boxes, z["fcos_dec_boxes_c%d" % center])
        assert_close(scores, z["fcos_dec_scores_c%d" % center], rtol=1e-6, what="fcos scores")
    for t in range(2):
        pr = synth.retina_predictions(1, 128, 80, int(z["retina_det%d_seed" % t]), logit_sigma=2.5)
        dets, _ = O.retina_image_detections([[x[0, a] for a in range(9)] for x in pr])
        assert np.array_equal(dets, z["retina_det%d" % t])


@pytest.mark.parametrize("n", [64, 700, 3000])
def test_nms_golden(golden, n):
    z = golden("decode_nms")
    dets = synth.nms_candidates(n, 640, synth.seed_for(4, 60) + n)
    assert np.array_equal(O.cpu_nms(dets, 0.5), z["cpu_nms_%d" % n])
    assert np.array_equal(O.cpu_nms(dets, 0.3), z["cpu_nms_%d_t3" % n])
    bb = dets.astype(np.float64)
    bb = np.stack([bb[:, 1], bb[:, 0], bb[:, 3] - bb[:, 1], bb[:, 2] - bb[:, 0], np.floor(bb[:, 4] * 100), bb[:, 5]], axis=1)
    for method in ("nms", "soft-nms"):
        rows, _ = O.centernet_nms(bb, 0.5, method=method)
        assert np.allclose(rows, z["cnms_%d_%s" % (n, method)], rtol=1e-12, atol=0), method


def test_combined_nms_against_torchvision():
    """`combined_nms` is parity-unpinned (TF op absent); its greedy core must agree with
    torchvision's batched NMS on per-class keeps when no cap binds."""
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(5)
    dets = synth.nms_candidates(300, 640, 77)
    scores = np.zeros((300, 4), np.float32)
    cls = rng.integers(0, 4, size=300)
    scores[np.arange(300), cls] = dets[:, 4]
    ob, os_, oc, valid, flat = O.combined_nms(dets[:, :4], scores, 1000, 1000, 0.5, 0.01)
    keep = tv.ops.batched_nms(torch.from_numpy(dets[:, [1, 0, 3, 2]].copy()), torch.from_numpy(dets[:, 4].copy()),
                              torch.from_numpy(cls), 0.5).numpy()
    assert sorted((flat[:valid] // 4).tolist()) == sorted(keep.tolist())
    assert np.all(np.diff(os_[:valid]) <= 0)
    ob2, os2, oc2, v2, _ = O.combined_nms(dets[:, :4], scores, 5, 12, 0.5, 0.05)
    assert v2 == 12 and np.array_equal(os2[:12], np.sort(os2[:12])[::-1])
    for c in range(4):
        assert (oc2[:v2] == c).sum() <= 5


def test_label_prep_golden(golden):
    """f-3: swap_xy / convert_to_xywh / convert_to_corners / horizontal box flip / label assembly against the
    reference's FCOS/utils.py and FCOS/data_preprocess.py outputs frozen in prep.npz."""
    z = golden("prep")
    raw = z["raw"]
    assert np.array_equal(O.swap_xy(raw), z["swap_xy"])
    assert np.array_equal(O.convert_to_xywh(raw), z["to_xywh"])
    assert np.array_equal(O.convert_to_corners(raw), z["to_corners"])
    assert np.array_equal(O.flip_boxes_horizontal(raw), z["flipped"])
    cls = np.arange(len(raw), dtype=np.float32) % 7
    assert np.array_equal(O.prepare_labels(raw, cls)[:, :4], z["labels_plain"])
    assert np.array_equal(O.prepare_labels(raw, cls, flip=True)[:, :4], z["labels_flipped"])
    assert np.array_equal(O.prepare_labels(raw, cls)[:, 4], cls)
    # f-4: detect_bboxes post-processing (retinanet_module.py:559-569), restated: columns 0/2 scale with w_ratio
    rows = np.array([[10, 20, 30, 40, .9, 3], [1, 2, 3, 4, .5, 0]], np.float32)
    b, s, l = O.format_detections(rows, 2.0, 0.5)
    assert np.array_equal(b, np.array([[10, 20, 20, 60], [1, 2, 2, 6]], np.float32))
    assert s.tolist() == [np.float32(.9), .5] and l.tolist() == [3, 0]


def test_hourglass4_encoder_golden(golden):
    """f-2: the inline 4-scale encoder of CenterNet/train_hourglass_voc.py:99-153 (frozen reference output)."""
    z = golden("hourglass4")
    for t in range(2):
        raw_dims, img_dims = (int(v) for v in z["hg4_%d_dims" % t])
        got = O.hourglass4_format_data(z["hg4_%d_labels" % t], raw_dims, img_dims, 5)
        assert np.array_equal(got, z["hg4_%d_map" % t])
        assert got[..., 4].sum() > 0 and got.shape == (img_dims // 8, img_dims // 8, 4, 10)
    # BCE against the closed form (float64)
    x = np.linspace(-9, 9, 41).astype(np.float32); zl = (np.arange(41) % 3 == 0).astype(np.float32)
    want = np.sum(np.logaddexp(0, x.astype(np.float64)) - x.astype(np.float64) * zl)
    assert abs(float(O.sigmoid_bce_sum(zl, x)) - want) <= 1e-6 * want
