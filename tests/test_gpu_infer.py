"""Inference path on the GPU (decode, pre-NMS top-k, NMS) vs golden vectors and the CPU oracle.
Kept indices / selected candidates are compared bit-exactly on identical inputs; decoded scores
(sigmoid) within 1e-6."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dense_head_ref as O  # noqa: E402
from oracle import synth  # noqa: E402
from conftest import assert_close  # noqa: E402

SCALES = [32, 64, 128, 256, 512]


def _dh():
    import densehead
    return densehead


@pytest.fixture(autouse=True, params=["lazy", "mask+sweep", "mask+sweep, bitonic sort"])
def nms_kernel(request):
    """Every test runs against both NMS implementations (DH_OPT_NMS_KERNEL): the lazy one-CTA-per-image sweep and
    the all-SM mask matrix + block sweep, and against both score sorts (DH_OPT_NMS_SORT: bucket sort, bitonic network);
    they must agree bit for bit with the oracle."""
    dh = _dh()
    dh.set_option(0, 6, 1 if request.param == "lazy" else 2)
    dh.set_option(0, 10, 1 if "bitonic" in request.param else 0)
    yield request.param
    dh.set_option(0, 6, 0)
    dh.set_option(0, 10, 0)


def test_prediction_to_corners_golden(golden):
    dh = _dh()
    z = golden("decode_nms")
    seed = int(z["p2c_seed"])
    p = synth.fcos_predictions(1, 128, 20, seed)[0][0]
    assert np.array_equal(dh.fcos.prediction_to_corners(p[..., :4], 8).cpu().numpy(), z["p2c_fcos"])
    assert np.array_equal(dh.fcos.prediction_to_corners(p, 8).cpu().numpy(), z["p2c_fcos"])  # full rows, first 4 used
    assert np.array_equal(dh.fcos.prediction_to_corners_center_v1(p[..., :4], 64, 8).cpu().numpy(), z["p2c_v1"])
    head = dh.retinanet.RetinaNetHead(80)
    assert np.array_equal(head.prediction_to_corners(p[..., :4], head.anchor_boxes[0][4], 8).cpu().numpy(), z["p2c_retina"])
    yp = synth.centernet_s8_predictions(1, 128, 8, 5, 3, seed)[0]
    assert np.array_equal(dh.centernet.prediction_to_corners_s8(yp[..., :4], SCALES, 8).cpu().numpy(), z["p2c_s8"])
    assert np.array_equal(dh.centernet.prediction_to_corners(p[..., :4], 8).cpu().numpy(), z["p2c_fcos"])


def test_fcos_decode_golden(golden):
    dh = _dh()
    z = golden("decode_nms")
    heads = synth.fcos_predictions(1, 128, 20, int(z["fcos_dec_seed"]))
    for center in (0, 1):
        boxes, scores = dh.fcos.decode_batch(heads, 20, [128, 128], center=bool(center))
        assert np.array_equal(boxes[0].cpu().numpy(), z["fcos_dec_boxes_c%d" % center])
        assert_close(scores[0].cpu().numpy(), z["fcos_dec_scores_c%d" % center], rtol=2e-6, atol_floor=0, what="fcos scores")


@pytest.mark.parametrize("t", [0, 1])
def test_retina_image_detections_golden(golden, t):
    dh = _dh()
    z = golden("decode_nms")
    pr = synth.retina_predictions(1, 128, 80, int(z["retina_det%d_seed" % t]), logit_sigma=2.5)
    head = dh.retinanet.RetinaNetHead(80)
    got = head.image_detections(head_outputs=pr).cpu().numpy()
    want = z["retina_det%d" % t]
    assert got.shape == want.shape
    assert np.array_equal(got[:, :4], want[:, :4]) and np.array_equal(got[:, 5], want[:, 5])
    assert_close(got[:, 4], want[:, 4], rtol=2e-6, atol_floor=0, what="retina scores")


@pytest.mark.parametrize("n", [64, 700, 3000])
def test_cpu_nms_golden(golden, n):
    dh = _dh()
    z = golden("decode_nms")
    k = golden("kat")
    head = dh.retinanet.RetinaNetHead(80)
    assert head.cpu_nms(k["nms_dets"], .5).tolist() == k["nms_keep"].tolist() == [0, 2]
    dets = synth.nms_candidates(n, 640, synth.seed_for(4, 60) + n)
    assert np.array_equal(head.cpu_nms(dets, 0.5), z["cpu_nms_%d" % n])
    assert np.array_equal(head.cpu_nms(dets, 0.3), z["cpu_nms_%d_t3" % n])
    assert head.cpu_nms(np.zeros((0, 6), np.float32), 0.5).shape == (0,)


def test_nms_batch_ragged_and_large():
    dh = _dh()
    from densehead import infer
    rng = np.random.default_rng(3)
    counts = [5000, 0, 1, 777, 4097]
    dets = np.zeros((len(counts), 5000, 6), np.float32)
    for b, c in enumerate(counts):
        if c:
            dets[b, :c] = synth.nms_candidates(c, 640, 900 + b)
    keep, n_keep = infer.nms(dets, 0.5, n_valid=np.array(counts, np.int32))
    for b, c in enumerate(counts):
        want = O.cpu_nms(dets[b, :c], 0.5) if c else np.zeros(0, np.int64)
        assert int(n_keep[b]) == len(want)
        assert np.array_equal(keep[b, :len(want)].cpu().numpy(), want)
    # score ties: order is by index (stable), as the oracle defines it
    d = synth.nms_candidates(400, 640, 5)
    d[:, 4] = np.round(d[:, 4] * 20) / 20
    keep, n_keep = infer.nms(d[None], 0.5)
    assert np.array_equal(keep[0, :int(n_keep[0])].cpu().numpy(), O.cpu_nms(d, 0.5))
    with pytest.raises(ValueError):
        infer.nms(np.zeros((1, 20000, 6), np.float32), 0.5)


def test_select_topk_vs_oracle():
    from densehead import infer
    rng = np.random.default_rng(11)
    n, segs = 9000, [0, 5000, 5003, 5003, 8000, 9000]
    dets = rng.normal(size=(3, n, 6)).astype(np.float32)
    dets[..., 4] = np.round(rng.uniform(0, 1, size=(3, n)) * 200) / 200       # many exact ties
    dets[2, :, 4] = 0.01                                                      # nothing passes in image 2
    for k, thr, incl in ((1000, 0.05, True), (7, 0.5, False), (6000, 0.2, True)):
        out, src = infer.select_topk(dets, segs, k, thr, score_inclusive=incl, with_source=True)
        out, src = out.cpu().numpy(), src.cpu().numpy()
        for b in range(3):
            want = O.select_topk(dets[b, :, 4], segs, k, thr, incl)
            for s, w in enumerate(want):
                got = src[b, s * k:(s + 1) * k]
                assert np.array_equal(got[:len(w)], w), (k, b, s)
                assert np.all(got[len(w):] == -1) and np.all(np.isneginf(out[b, s * k + len(w):(s + 1) * k, 4]))
                assert np.array_equal(out[b, s * k:s * k + len(w)], dets[b, w])


def test_per_class_nms_vs_oracle_and_torchvision():
    from densehead import infer
    tv = pytest.importorskip("torchvision")
    dets = synth.nms_candidates(1500, 640, 77, classes=6)
    n, c = len(dets), 6
    scores = np.zeros((n, c), np.float32)
    scores[np.arange(n), dets[:, 5].astype(int)] = dets[:, 4]
    for mpc, mt in ((0, 0), (5, 12), (100, 100)):
        keep, n_keep = infer.nms(dets[None], 0.5, mode=infer.NMS_PER_CLASS, min_score=0.05, score_inclusive=False, num_classes=c,
                                 max_per_class=mpc, max_total=mt, max_out=n)
        got = keep[0, :int(n_keep[0])].cpu().numpy()
        _, os_, oc, valid, flat = O.combined_nms(dets[:, :4], scores, mpc or n, mt or n, 0.5, 0.05)
        assert np.array_equal(got, flat[:valid] // c), (mpc, mt)
    keep, n_keep = infer.nms(dets[None], 0.5, mode=infer.NMS_PER_CLASS, min_score=0.01, score_inclusive=False, num_classes=c, max_out=n)
    tvk = tv.ops.batched_nms(torch.from_numpy(dets[:, [1, 0, 3, 2]].copy()), torch.from_numpy(dets[:, 4].copy()),
                             torch.from_numpy(dets[:, 5].astype(np.int64)), 0.5).numpy()
    assert sorted(keep[0, :int(n_keep[0])].cpu().numpy().tolist()) == sorted(tvk.tolist())


def test_c4_shaped_detection_pipelines():
    """C4: COCO-shaped heads (640 x 640, 80 classes), 1000 pre-NMS candidates per level, IoU 0.5.  The decode front ends
    are compared with the ORACLE's decode of the same 640 x 640 heads (1e-5 relative: sigmoid and the box arithmetic; labels
    equal wherever the two best class scores are not within rounding of each other); selection + NMS are then checked
    bit-exactly against the oracle run on the GPU-decoded rows (a float that differs in its last bit may legitimately
    fall on the other side of a top-k cut)."""
    dh = _dh()
    B = 3
    pr = synth.retina_predictions(B, 640, 80, synth.seed_for(4, 80), logit_sigma=2.5)
    dets = dh.retinanet.decode_batch(pr, 80, [640, 640])
    for b in range(B):
        want = O.retina_decode_dets([[p[b, a] for a in range(9)] for p in pr])
        got = dets[b].cpu().numpy()
        assert got.shape == want.shape == (76725, 6)
        assert np.all(np.abs(got[:, :5] - want[:, :5]) <= 1e-5 * np.maximum(1.0, np.abs(want[:, :5]))), "retina decode, image %d" % b
        differ = np.nonzero(got[:, 5] != want[:, 5])[0]
        assert len(differ) <= 8  # first-argmax ties within float32 rounding of the two sigmoids
        for r in differ:
            row = np.concatenate([p[b].reshape(-1, 84) for p in pr])[r, 4:]
            top = np.sort(1.0 / (1.0 + np.exp(-row.astype(np.float64))))[-2:]
            assert top[1] - top[0] <= 2e-7 * top[1], (b, r)
    cand, keep, n_keep = dh.retinanet.detect_batch(pr, 80, [640, 640], pre_nms_topk=1000)
    lens = [9 * (640 // s) ** 2 for s in (8, 16, 32, 64, 128)]
    seg = np.concatenate([[0], np.cumsum(lens)])
    for b in range(B):
        wcand, wkeep, _ = O.retina_detect_from_dets(dets[b].cpu().numpy(), seg, pre_nms_topk=1000)
        got_c = cand[b].cpu().numpy()
        got_c = got_c[np.isfinite(got_c[:, 4])]
        assert np.array_equal(got_c, wcand)
        # candidate slots are level-aligned (k per level), the oracle's are dense: compare kept rows
        kept_rows = cand[b].index_select(0, keep[b, :int(n_keep[b])].long()).cpu().numpy()
        assert np.array_equal(kept_rows, wcand[wkeep])
    heads = synth.fcos_predictions(B, 640, 80, synth.seed_for(4, 81))
    for h in heads:
        h[..., 5:] = h[..., 5:] * 2.5 + 6.9          # N(-4.6, 2.5): plenty of candidates above 0.05
    boxes, scores = dh.fcos.decode_batch(heads, 80, [640, 640])
    ob, os_, oc, nv = dh.fcos.detect_batch(heads, 80, [640, 640], pre_nms_topk=1000)
    seg = np.concatenate([[0], np.cumsum([(640 // s) ** 2 * 80 for s in (8, 16, 32, 64, 128)])])
    for b in range(B):
        wbx, wsc = O.fcos_decode_scores([h[b] for h in heads], 80)
        assert np.all(np.abs(boxes[b].cpu().numpy() - wbx) <= 1e-5 * np.maximum(1.0, np.abs(wbx))), "fcos decode boxes, image %d" % b
        assert np.all(np.abs(scores[b].cpu().numpy() - wsc) <= 1e-5 * np.maximum(1e-2, np.abs(wsc))), "fcos decode scores, image %d" % b
    for b in range(B):
        wb, ws, wc, wv, _ = O.fcos_detect_from_scores(boxes[b].cpu().numpy(), scores[b].cpu().numpy(), seg, pre_nms_topk=1000)
        assert int(nv[b]) == wv
        assert np.array_equal(ob[b].cpu().numpy(), wb) and np.array_equal(os_[b].cpu().numpy(), ws)
        assert np.array_equal(oc[b].cpu().numpy(), wc)


@pytest.mark.parametrize("shape", [(256, 20, 300), (640, 80, 1000)], ids=["256px-20cls", "640px-80cls"])
@pytest.mark.parametrize("case", ["plain", "quantised ties", "saturated", "few pass", "threshold edge"])
def test_fcos_select_logit_space_equals_exact_scoring(case, shape):
    """dh_fcos_detect selects candidates on raw logits when it can (no sigmoid per element) -- by default with a
    thread-block cluster per long level (DH_OPT_FCOS_SELECT=3), an estimate + one streaming pass + per-level finish (=4;
    the default, 0, picks between these two by batch size) or one CTA per (image, level) (=2); the candidate rows and
    the final detections must be bit-identical to the path that scores every (location, class) pair (=1), including
    where distinct logits collide after the float32 sigmoid and ties are cut by index.  The 640-pixel shape spreads
    level 0 over a whole cluster and has a level whose rows are not float4-aligned."""
    dh = _dh()
    size, classes, topk = shape
    B = 2
    heads = synth.fcos_predictions(B, size, classes, synth.seed_for(4, 90))
    for h in heads:
        x = h[..., 5:] * 2.5 + 6.9
        if case == "quantised ties":
            x = np.round(x * 8) / 8
        elif case == "saturated":
            x = x + 14.0                      # thousands of scores round to exactly 1.0 or within a few ulps of it
        elif case == "few pass":
            x = x - (6.0 if size == 256 else 9.0)
        elif case == "threshold edge":
            x = np.where(np.abs(x + 2.9444) < 0.2, np.float32(-2.9444389) + np.round((x + 2.9444) * 1e7) * np.float32(2.4e-7), x)
        h[..., 5:] = x.astype(np.float32)
    out = []
    for mode in (3, 1, 2, 4, 0):
        dh.set_option(0, 7, mode)
        try:
            out.append([t.cpu().numpy() for t in dh.fcos.detect_batch(heads, classes, [size, size], pre_nms_topk=topk, with_candidates=True)])
        finally:
            dh.set_option(0, 7, 0)
    for other in (out[0], out[2], out[3], out[4]):
        for a, b in zip(other, out[1]):
            assert np.array_equal(a, b)
    assert np.isfinite(out[0][4][..., 4]).any() or case == "few pass"


def test_fcos_select_cluster_small_batches_and_levels():
    """The cluster selector's host-side plan and the streaming pre-select's chunk table: one level only, levels shorter
    than a slice / a chunk, a single image."""
    dh = _dh()
    for size, classes, strides, B in ((640, 80, [8], 1), (128, 3, [8, 16, 32, 64, 128], 3), (512, 20, [16, 32], 2)):
        heads = synth.fcos_predictions(B, size, classes, synth.seed_for(4, 91), strides=strides)
        for h in heads:
            h[..., 5:] = h[..., 5:] * 2.5 + 6.9
        out = []
        for mode in (1, 3, 4):
            dh.set_option(0, 7, mode)
            try:
                out.append([t.cpu().numpy() for t in dh.fcos.detect_batch(heads, classes, [size, size], pre_nms_topk=200, with_candidates=True,
                                                                           strides=strides)])
            finally:
                dh.set_option(0, 7, 0)
        for other in out[1:]:
            for a, b in zip(other, out[0]):
                assert np.array_equal(a, b)


def test_fcos_select_large_topk_uses_clusters_by_default():
    """pre_nms_topk above 1024 does not fit the streaming pre-select: the default falls back to the cluster selector
    (small batch) -- same candidates and detections as scoring every pair."""
    dh = _dh()
    heads = synth.fcos_predictions(2, 640, 80, synth.seed_for(4, 92))
    for h in heads:
        h[..., 5:] = h[..., 5:] * 2.5 + 6.9
    out = []
    for mode in (1, 0, 4):
        dh.set_option(0, 7, mode)
        try:
            out.append([t.cpu().numpy() for t in dh.fcos.detect_batch(heads, 80, [640, 640], pre_nms_topk=3000, with_candidates=True)])
        finally:
            dh.set_option(0, 7, 0)
    for other in out[1:]:
        for a, b in zip(other, out[0]):
            assert np.array_equal(a, b)


def test_compute_iou_and_bboxes_iou(golden):
    dh = _dh()
    k = golden("kat")
    assert np.array_equal(dh.retinanet.compute_iou(k["iou_b1"], k["iou_b2"]).cpu().numpy(), k["iou"])
    rng = np.random.default_rng(2)
    b1 = rng.uniform(0, 300, size=(37, 4)).astype(np.float32); b2 = rng.uniform(0, 300, size=(211, 4)).astype(np.float32)
    assert np.array_equal(dh.retinanet.compute_iou(b1, b2).cpu().numpy(), O.compute_iou(b1, b2))
    c1 = np.sort(rng.uniform(0, 100, size=(50, 2, 2)), axis=1).reshape(50, 4)[:, [0, 1, 2, 3]]
    c2 = np.sort(rng.uniform(0, 100, size=(50, 2, 2)), axis=1).reshape(50, 4)
    assert np.allclose(dh.centernet.bboxes_iou(c1, c2).cpu().numpy(), O.bboxes_iou(c1, c2), rtol=1e-14, atol=0)
    assert np.allclose(dh.centernet.bboxes_iou(c1[:1], c2).cpu().numpy(), O.bboxes_iou(c1[:1], c2), rtol=1e-14, atol=0)


@pytest.mark.parametrize("n", [64, 700, 3000])
def test_centernet_nms_golden(golden, n):
    dh = _dh()
    z = golden("decode_nms")
    k = golden("kat")
    rows = dh.centernet.nms(k["cnms_in"].copy(), .5)
    assert np.array_equal(np.array(rows), k["cnms_hard"][np.argsort(k["cnms_hard"][:, 5], kind="stable")])
    dets = synth.nms_candidates(n, 640, synth.seed_for(4, 60) + n)
    bb = dets.astype(np.float64)
    bb = np.stack([bb[:, 1], bb[:, 0], bb[:, 3] - bb[:, 1], bb[:, 2] - bb[:, 0], np.floor(bb[:, 4] * 100), bb[:, 5]], axis=1)
    keep_in = bb.copy()
    for method in ("nms", "soft-nms"):
        got = np.array(dh.centernet.nms(bb, 0.5, method=method)).reshape(-1, 6)
        want = z["cnms_%d_%s" % (n, method)]
        assert got.shape == want.shape, method
        assert np.allclose(got, want, rtol=1e-11, atol=0), method
        assert np.array_equal(got[:, :4], want[:, :4]) and np.array_equal(got[:, 5], want[:, 5])
    assert np.array_equal(bb, keep_in)   # input untouched
    assert dh.centernet.nms(np.zeros((0, 6)), 0.5) == []
