"""The extension modes (GIoU loss, Gaussian heat map, min-area FCOS tie-break: BASELINE's north_star names them, the
reference does not have them -- SURVEY.md section 0) against their specifications in the oracle, and the corners of the
fused kernel's target-row machinery (pair list overflow -> dense visit, rows matched by many boxes, a bad class id)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dense_head_ref as O  # noqa: E402
from oracle import synth  # noqa: E402
from conftest import assert_close  # noqa: E402

W = (1.25, 0.5, 2.0)


def _dh():
    import densehead
    return densehead


def _crowded(seed, n=14, side=256, classes=5):
    rng = np.random.default_rng(seed)
    hw = np.exp(rng.uniform(np.log(20.0), np.log(120.0), size=(n, 2)))
    g = np.concatenate([rng.uniform(0.35 * side, 0.65 * side, size=(n, 2)), hw], axis=1) / side
    g = np.concatenate([g, rng.integers(0, classes, size=(n, 1))], axis=1).astype(np.float32)
    g[3, 2:4] = g[9, 2:4]  # an exact area tie
    boxes = np.zeros((1, (n + 3) & ~3, 5), np.float32)
    boxes[0, :n] = g
    return boxes, np.array([n], np.int32), g


@pytest.mark.parametrize("seed", [1, 2])
def test_fcos_min_area_encode_bit_exact(seed):
    dh = _dh()
    boxes, nbox, g = _crowded(100 + seed)
    outs, cnt = dh.fcos.format_data_batch(boxes, nbox, [256, 256], 5, [256, 256], mode="min_area")
    want, wcnt = O.fcos_format_data(g, [256, 256], 5, order="min_area")
    ref, _ = O.fcos_format_data(g, [256, 256], 5)
    assert cnt[0].tolist() == wcnt
    differs = False
    for l in range(5):
        got = outs[l][0].cpu().numpy()
        assert np.array_equal(np.delete(got, 4, axis=-1), np.delete(want[l], 4, axis=-1)), "level %d" % l
        assert np.allclose(got[..., 4], want[l][..., 4], rtol=1e-6, atol=0)   # centerness: float64 sqrt on both sides
        differs = differs or not np.array_equal(got, ref[l])
    assert differs
    # and through the fused loss: equal to the unfused loss over the materialised min-area targets
    pred = synth.fcos_predictions(1, 256, 5, 7)
    pi, tot, _ = dh.fcos.encode_loss_batch(boxes, nbox, [256, 256], 5, [256, 256], pred, mode="min_area")
    upi, utot = dh.fcos.model_loss_batch(outs, pred)
    assert_close(pi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what="min-area fused vs unfused")


@pytest.mark.parametrize("stride", [4, 8])
def test_centernet_gaussian_encode(stride):
    dh = _dh()
    boxes, nbox = synth.make_boxes(3, 512, 40, 3, 8.0, 400.0, synth.seed_for(2, 70))
    nbox[1] = 0
    out, _ = dh.centernet.format_data_batch(boxes, nbox, [512, 512], 3, [512, 512], stride=stride, mode="gaussian")
    for b in range(3):
        want = O.centernet_gaussian_format_data(boxes[b, :nbox[b]], [512, 512], 3, stride=stride)
        got = out[b].cpu().numpy()
        assert np.array_equal(np.delete(got, 4, axis=-1), np.delete(want, 4, axis=-1)), "image %d" % b
        assert np.array_equal(got[..., 4] > 0, want[..., 4] > 0)
        assert np.allclose(got[..., 4], want[..., 4], rtol=2e-6, atol=0), "heat of image %d" % b   # exp: libm vs the GPU's
    # both encoder kernels agree bit for bit
    from densehead import _capi
    res = []
    for kern in (1, 2):
        dh.set_option(0, _capi.DH_OPT_ENCODE_KERNEL, kern)
        try:
            o, _ = dh.centernet.format_data_batch(boxes, nbox, [512, 512], 3, [512, 512], stride=stride, mode="gaussian")
        finally:
            dh.set_option(0, _capi.DH_OPT_ENCODE_KERNEL, 0)
        res.append(o.clone())
    assert torch.equal(res[0], res[1])
    # fused loss over the Gaussian targets == unfused loss over the materialised ones
    rng = np.random.default_rng(64)
    hw = 512 // stride
    yp = rng.normal(-2.0, 1.5, size=(3, hw, hw, 8)).astype(np.float32)
    yp[..., :4] = rng.uniform(0.2, 5.0, size=(3, hw, hw, 4)).astype(np.float32)
    pi, tot, _ = dh.centernet.encode_loss_batch(boxes, nbox, [512, 512], 3, [512, 512], yp, stride=stride, mode="gaussian")
    for b in range(3):
        want = O.dense_loss_f64(out[b].cpu().numpy(), yp[b], reg_ch=4, cen_mode=1, pos_rule="ge1")
        assert_close(pi[b, :3].cpu().numpy(), np.array(want), 1e-5, what="gaussian fused loss, image %d" % b)


def test_giou_loss_forward_and_gradient():
    dh = _dh()
    from densehead import losses
    B, C = 3, 20
    boxes, nbox = synth.config_boxes("fcos_voc", B, synth.seed_for(6, 30))
    pred = synth.fcos_predictions(B, 512, C, 66)
    for p in pred:
        p[..., :4] = np.abs(p[..., :4]) + np.float32(0.3)
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [512, 512], C, [512, 512])
    pi, tot, _, grads = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], pred, reg_type="giou", weights=W)
    upi, utot, ugrads = dh.fcos.model_loss_batch(tg, pred, "giou", "l1", weights=W)
    for b in range(B):
        want = O.fcos_model_loss([t[b].cpu().numpy() for t in tg], [p[b] for p in pred], reg_type="giou")
        assert_close(pi[b, :3].cpu().numpy(), np.array([float(v) for v in want]), 1e-5, what="fused giou image %d" % b)
        assert_close(upi[b, :3].cpu().numpy(), np.array([float(v) for v in want]), 1e-5, what="unfused giou image %d" % b)
    for l in range(5):
        want = O.dense_loss_grad(tg[l].cpu().numpy(), pred[l], weights=W, reg_ch=4, cen_mode=1, reg_mode=2, pos_rule="ge1")
        for name, g in (("fused", grads[l]), ("unfused", ugrads[l])):
            err = np.abs(g.cpu().numpy() - want)
            assert np.all(err <= 2e-5 * np.maximum(1.0, np.abs(want))), "%s level %d: %g" % (name, l, err.max())
    # the stand-alone wrapper
    t, p = tg[1][0, ..., :4].cpu().numpy(), pred[1][0, ..., :4]
    m = (tg[1][0, ..., 5:].cpu().numpy().max(-1) >= 1).astype(np.float32)
    assert_close(float(losses.giou_loss(t, p, m)), float(O.giou_loss(t, p, m)), 1e-5, what="giou_loss")
    assert float(losses.giou_loss(t, t, m)) <= 1e-4 * max(1.0, float(m.sum()))


@pytest.mark.parametrize("thr", [0.02, -1.0])
def test_retina_fused_loss_dense_targets(thr):
    """An IoU threshold near zero (or below it: every pair matches) floods the chunk's pair list: the fused kernel then
    visits every row, and rows matched by dozens of boxes take the whole-list path.  Same sums as the unfused loss
    over the materialised targets, same pair counts as the encoder."""
    dh = _dh()
    B, C = 2, 20
    boxes, nbox = synth.make_boxes(B, 256, 40, C, 8.0, 200.0, synth.seed_for(6, 31))
    pred = synth.retina_predictions(B, 256, C, 67)
    lab, pairs_e = dh.retinanet.format_data_batch(boxes, nbox, [256, 256], C, [256, 256], iou_thresh=thr)
    pi, tot, pairs = dh.retinanet.encode_loss_batch(boxes, nbox, [256, 256], C, [256, 256], pred, iou_thresh=thr)
    upi, utot = dh.retinanet.loss_batch(lab, pred)
    assert torch.equal(pairs, pairs_e)
    assert_close(pi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what="thr %g" % thr)
    assert torch.equal(pi[:, 3], upi[:, 3])
    for b in range(B):
        want_lab, want_pairs = O.retina_format_data(boxes[b, :nbox[b]], [256, 256], C, iou_thresh=thr)
        assert int(pairs[b]) == want_pairs


def test_fused_loss_rows_matched_by_many_boxes():
    """Nine near-identical GT boxes: every positive anchor is matched by all of them (a run longer than a lane keeps)."""
    dh = _dh()
    C = 20
    base = np.array([0.5, 0.5, 0.25, 0.25], np.float32)
    g = np.stack([np.concatenate([base + np.float32(1e-3 * k), [np.float32(k % C)]]) for k in range(12)]).astype(np.float32)
    boxes = g[None]
    nbox = np.array([12], np.int32)
    pred = synth.retina_predictions(1, 256, C, 68)
    lab, pairs_e = dh.retinanet.format_data_batch(boxes, nbox, [256, 256], C, [256, 256])
    pi, tot, pairs = dh.retinanet.encode_loss_batch(boxes, nbox, [256, 256], C, [256, 256], pred)
    upi, _ = dh.retinanet.loss_batch(lab, pred)
    want_lab, want_pairs = O.retina_format_data(g, [256, 256], C)
    assert int(pairs[0]) == want_pairs == int(pairs_e[0]) and want_pairs > 0
    assert_close(pi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what="many boxes per row")
    want = O.retina_train_loss(want_lab, [[p[0, a] for a in range(9)] for p in pred])
    assert_close(pi[0, :2].cpu().numpy(), np.array([float(want[0]), float(want[1])]), 1e-5, what="vs oracle")


def test_bad_class_id_is_flagged_not_written():
    """Host labels: IndexError like the reference (FCOS/fcos.py:281-283).  Device labels: the kernels drop the box and set
    DH_STATUS_BAD_CLASS; nothing outside the row's class channels is written."""
    dh = _dh()
    from densehead import _capi
    boxes, nbox = synth.config_boxes("fcos_voc", 2, synth.seed_for(6, 32))
    bad = boxes.copy()
    bad[1, 0, 4] = 20.0  # == num_classes
    with pytest.raises(IndexError):
        dh.fcos.format_data_batch(bad, nbox, [512, 512], 20, [512, 512])
    with pytest.raises(IndexError):
        dh.retinanet.format_data_batch(bad, nbox, [512, 512], 20, [512, 512])
    _capi.status(0)  # clear
    bd, nd = torch.from_numpy(bad).cuda(), torch.from_numpy(nbox).cuda()
    outs, _ = dh.fcos.format_data_batch(bd, nd, [512, 512], 20, [512, 512])
    assert _capi.status(0, reset=False) & _capi.DH_STATUS_BAD_CLASS
    with pytest.raises(IndexError):
        _capi.raise_for_status(0)
    assert _capi.status(0) == 0
    dropped = bad.copy()
    dropped[1, 0] = dropped[1, nbox[1] - 1]  # the same image without the bad box
    n2 = nbox.copy()
    n2[1] -= 1
    want, _ = dh.fcos.format_data_batch(dropped, n2, [512, 512], 20, [512, 512])
    for a, b in zip(outs, want):
        assert torch.equal(a, b)
    routs, rp = dh.retinanet.format_data_batch(bd, nd, [512, 512], 20, [512, 512])
    assert _capi.status(0) & _capi.DH_STATUS_BAD_CLASS
    rwant, rpw = dh.retinanet.format_data_batch(dropped, n2, [512, 512], 20, [512, 512])
    # (the RetinaNet regression winner is the highest GT index, so moving the last box to slot 0 changes ties: compare
    # the class channels and the pair counts)
    for a, b in zip(routs, rwant):
        assert torch.equal(a[..., 4:], b[..., 4:])
    assert torch.equal(rp, rpw)
