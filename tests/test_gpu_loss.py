"""CUDA losses (unfused over materialised targets, and fused encode+loss) vs the golden values frozen
from the reference and vs the CPU oracle.  Tolerance: 1e-5 relative (north_star) on every loss scalar;
positive counts exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dense_head_ref as O  # noqa: E402
from oracle import synth  # noqa: E402
from conftest import assert_close  # noqa: E402

SCALES = [32, 64, 128, 256, 512]
RTOL = 1e-5


def _dh():
    import densehead
    return densehead


def _f(*vals):
    return np.array([float(v) for v in vals])


def test_kat_losses(golden):
    dh = _dh()
    k = golden("kat")
    assert_close(float(dh.fcos.focal_loss([0, 1, 0, 1, 1.], [-3, -.5, 0, .5, 3.])), k["focal"], RTOL, what="focal KAT")
    assert_close(float(dh.fcos.smooth_l1_loss([[0, 1, 2, 3.]], [[.5, 1, 4, 2.2]], mask=1.0)), 2.445, RTOL, what="sl1 KAT")
    yt = np.zeros((2, 2, 4), np.float32); yt[1, 1] = (1, 2, 1.5, .5)
    m = np.zeros((2, 2), np.float32); m[1, 1] = 1
    assert_close(float(dh.fcos.iou_loss(yt, np.ones((2, 2, 4), np.float32), m)), k["iou_loss"], RTOL, what="iou KAT")


def test_flat_losses_golden(golden):
    dh = _dh()
    z = golden("losses")
    assert_close(float(dh.fcos.focal_loss(z["flat_y"], z["flat_x"])), z["flat_focal"], RTOL, what="flat focal")
    assert_close(float(dh.fcos.focal_loss(z["flat_y"], z["flat_x"], alpha=0.4, gamma=1.5)), z["flat_focal_a4g15"], RTOL, what="focal a/g")
    assert_close(float(dh.fcos.smooth_l1_loss(z["sl1_a"], z["sl1_b"], z["sl1_m"])), z["sl1"], RTOL, what="sl1")
    assert_close(float(dh.fcos.smooth_l1_loss(z["sl1_a"], z["sl1_b"], z["sl1_m"], delta=2.0)), z["sl1_d2"], RTOL, what="sl1 d2")
    assert_close(float(dh.fcos.iou_loss(np.abs(z["sl1_a"]), np.abs(z["sl1_b"]), z["sl1_m"])), z["iou_l"], RTOL, what="iou")
    # scalar mask on a non-multiple-of-4 array (the centerness call pattern, fcos.py:484-486)
    a = np.linspace(-2, 2, 37, dtype=np.float32); b = np.zeros(37, np.float32)
    assert_close(float(dh.fcos.smooth_l1_loss(a, b, mask=1.0)), O.smooth_l1_loss(a, b), RTOL, what="sl1 scalar mask")


@pytest.mark.parametrize("t", [0, 1, 2])
def test_fcos_model_loss_golden(golden, t):
    dh = _dh()
    z = golden("losses")
    g, seed = z["fcos%d_g" % t], int(z["fcos%d_seed" % t])
    pred = synth.fcos_predictions(1, 512, 20, seed)
    tg, _ = dh.fcos.format_data(g, [512, 512], 20)
    assert_close(_f(*dh.fcos.model_loss(tg, pred, None)), z["fcos%d_l1" % t], RTOL, what="fcos l1")
    assert_close(_f(*dh.fcos.model_loss(tg, pred, None, reg_type="iou")), z["fcos%d_iou" % t], RTOL, what="fcos iou")
    assert float(dh.fcos.model_loss(tg, pred, None, cen_type="focal")[2]) == 0.0   # reference quirk Q27
    tg, _ = dh.fcos.format_data_center(g, [512, 512], 20)
    assert_close(_f(*dh.fcos.model_loss_center(tg, pred, cen_type="focal")), z["fcos%d_center_focal" % t], RTOL, what="center focal")
    assert_close(_f(*dh.fcos.model_loss_center(tg, pred)), z["fcos%d_center_l1" % t], RTOL, what="center l1")
    tg, _ = dh.fcos.format_data_center_v1(g, [512, 512], 20)
    assert_close(_f(*dh.fcos.model_loss_center_v1(tg, pred)), z["fcos%d_v1" % t], RTOL, what="v1")
    # fused: same numbers without materialising targets
    boxes = np.zeros((1, 20, 5), np.float32); boxes[0, :len(g)] = g
    for mode, reg, cen, key in (("fcos", "l1", "l1", "fcos%d_l1"), ("fcos", "iou", "l1", "fcos%d_iou"),
                                ("center", "l1", "focal", "fcos%d_center_focal"), ("center_v1", "l1", "focal", "fcos%d_v1")):
        pi, tot, cnt = dh.fcos.encode_loss_batch(boxes, [len(g)], [512, 512], 20, [512, 512], pred, mode=mode,
                                                 reg_type=reg, cen_type=cen)
        assert_close(tot[:3].cpu().numpy(), z[key % t], RTOL, what="fused " + key % t)
        assert torch.equal(pi[0], tot)


@pytest.mark.parametrize("t", [0, 1])
def test_retina_loss_golden(golden, t):
    dh = _dh()
    z = golden("losses")
    g, seed = z["retina%d_g" % t], int(z["retina%d_seed" % t])
    head = dh.retinanet.RetinaNetHead(80)
    lab, pairs = head.format_data(g, [256, 256])
    pred = synth.retina_predictions(1, 256, 80, seed)
    x_pred = [[torch.from_numpy(p[:, a]).cuda() for a in range(9)] for p in pred]   # reference layout [1,Hl,Wl,84]
    assert_close(_f(*head.loss(x_pred, lab)), z["retina%d_loss" % t], RTOL, what="retina loss")
    boxes = np.zeros((1, 12, 5), np.float32); boxes[0, :len(g)] = g
    pi, tot, pr = dh.retinanet.encode_loss_batch(boxes, [len(g)], [256, 256], 80, [256, 256], pred)
    assert_close(tot[:2].cpu().numpy(), z["retina%d_loss" % t], RTOL, what="retina fused")
    assert int(pr[0]) == pairs


def test_centernet_loss_golden(golden):
    dh = _dh()
    z = golden("losses")
    boxes, nbox, seed = z["cn_boxes"], z["cn_nbox"], int(z["cn_seed"])
    yp = synth.centernet_s8_predictions(2, 512, 8, 5, 3, seed)
    yt, _ = dh.centernet.format_data_batch(boxes, nbox, [512, 512], 3, [512, 512], stride=8, mode="s8", box_scales=SCALES)
    assert_close(_f(*dh.centernet.model_loss_s8(yt, yp)), z["cn_s8_loss"], RTOL, what="s8 loss")
    pi, tot, st = dh.centernet.encode_loss_batch(boxes, nbox, [512, 512], 3, [512, 512], yp, stride=8, mode="s8", box_scales=SCALES)
    assert_close(tot[:2].cpu().numpy(), z["cn_s8_loss"], RTOL, what="s8 fused")
    yp2 = np.ascontiguousarray(yp[:, :, :, 0, :])
    yt, _ = dh.centernet.format_data_batch(boxes, nbox, [512, 512], 3, [512, 512], stride=8, mode="hourglass")
    assert_close(_f(*dh.centernet.model_loss_hourglass(yt, yp2)), z["cn_hg_loss"], RTOL, what="hg loss")
    pi, tot, st = dh.centernet.encode_loss_batch(boxes, nbox, [512, 512], 3, [512, 512], yp2, stride=8, mode="hourglass")
    assert_close(tot[:2].cpu().numpy(), z["cn_hg_loss"], RTOL, what="hg fused")


def test_batched_losses_vs_oracle_and_determinism():
    """C1-shaped FCOS batch and a RetinaNet batch: per-image sums equal the oracle's, the batch total is
    the sum of the per-image rows, fused == unfused, and two runs are bit-identical."""
    dh = _dh()
    boxes, nbox = synth.config_boxes("fcos_voc", 8, synth.seed_for(5, 70))
    pred = synth.fcos_predictions(8, 512, 20, 123)
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [512, 512], 20, [512, 512])
    pi, tot = dh.fcos.model_loss_batch(tg, pred)
    pi2, tot2 = dh.fcos.model_loss_batch(tg, pred)
    assert torch.equal(pi, pi2) and torch.equal(tot, tot2)
    fpi, ftot, _ = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], 20, [512, 512], pred)
    assert_close(fpi.cpu().numpy(), pi.cpu().numpy(), 2e-6, what="fused vs unfused per image")
    assert_close(tot.cpu().numpy(), pi.double().sum(0).cpu().numpy(), 1e-6, what="total = sum(per image)")
    for b in (0, 5):
        want_t, _ = O.fcos_format_data(boxes[b, :nbox[b]], [512, 512], 20)
        want = O.fcos_model_loss(want_t, [p[b] for p in pred])
        assert_close(pi[b, :3].cpu().numpy(), _f(*want), RTOL, what="fcos b%d" % b)
        npos = sum(int((x[..., 5:].max(-1) >= 1).sum()) for x in want_t)
        assert int(pi[b, 3]) == npos == int(fpi[b, 3])
    boxes, nbox = synth.make_boxes(3, 320, 40, 80, 8.0, 250.0, synth.seed_for(5, 71))
    pred = synth.retina_predictions(3, 320, 80, 321)
    lab, _ = dh.retinanet.format_data_batch(boxes, nbox, [320, 320], 80, [320, 320])
    pi, tot = dh.retinanet.loss_batch(lab, pred)
    fpi, ftot, _ = dh.retinanet.encode_loss_batch(boxes, nbox, [320, 320], 80, [320, 320], pred)
    assert_close(fpi.cpu().numpy(), pi.cpu().numpy(), 2e-6, what="retina fused vs unfused")
    want_l, _ = O.retina_format_data(boxes[1, :nbox[1]], [320, 320], 80)
    want = O.retina_train_loss(want_l, [[p[1, a] for a in range(9)] for p in pred])
    assert_close(pi[1, :2].cpu().numpy(), _f(*want), RTOL, what="retina b1")


def test_loss_linearity_property():
    """Size-independent property at C5's full batch: loss(batch) rows are independent of batch composition.  With
    uniform scheduler chunks (DH_OPT_FUSED_TAIL = 0) any sub-batch reproduces the same per-image rows bit for bit;
    with the tiered tail (default: the last images of a launch are cut into finer chunks) an image's float32
    partial sums are grouped by its position in the batch, so the rows agree to rounding (2e-6), counts exactly."""
    dh = _dh()
    from densehead import _capi
    boxes, nbox = synth.config_boxes("fcos_voc", 256, synth.seed_for(5, 72))
    pred = [torch.from_numpy(p).cuda() for p in synth.fcos_predictions(256, 512, 20, 77)]
    sub = slice(64, 96)
    spred = [p[sub].contiguous() for p in pred]
    try:
        dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, 0)
        pi0, tot0, _ = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], 20, [512, 512], pred)
        spi0, _, _ = dh.fcos.encode_loss_batch(boxes[sub], nbox[sub], [512, 512], 20, [512, 512], spred)
        assert torch.equal(spi0, pi0[sub])
    finally:
        dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, 1)
    pi, tot, _ = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], 20, [512, 512], pred)
    spi, _, _ = dh.fcos.encode_loss_batch(boxes[sub], nbox[sub], [512, 512], 20, [512, 512], spred)
    assert torch.allclose(spi, pi[sub], rtol=2e-6, atol=0) and torch.equal(spi[:, 3], pi[sub, 3])
    assert torch.allclose(pi, pi0, rtol=2e-6, atol=0) and torch.allclose(tot, tot0, rtol=2e-6, atol=0)
    assert torch.isfinite(tot).all() and float(tot[3]) == float(pi[:, 3].sum())


def test_fused_loss_edge_cases():
    """Images without boxes, 256 boxes per image (the capacity), > 128 classes (falls back to the
    shared-memory target-tile kernel), predictions at a 4-byte-aligned (not 16-byte) address (scalar loads), and the
    stream+correct kernel against the target-tile kernel (DH_OPT_FUSED_LOSS_KERNEL)."""
    dh = _dh()
    C = 20
    # no boxes at all: pure zero-label loss, equal to the unfused loss over all-zero targets
    pred = synth.fcos_predictions(2, 256, C, 5)
    boxes, nbox = np.zeros((2, 4, 5), np.float32), np.zeros((2,), np.int32)
    pi, tot, _ = dh.fcos.encode_loss_batch(boxes, nbox, [256, 256], C, [256, 256], pred)
    zeros = [np.zeros_like(p) for p in pred]
    upi, utot = dh.fcos.model_loss_batch(zeros, pred)
    assert_close(pi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what="no boxes") and float(tot[3]) == 0
    # capacity: 256 boxes in one image
    boxes, nbox = synth.make_boxes(1, 512, 256, C, 4.0, 200.0, synth.seed_for(5, 90), full=True)
    assert int(nbox[0]) == 256
    pred = synth.fcos_predictions(1, 512, C, 6)
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [512, 512], C, [512, 512])
    upi, _ = dh.fcos.model_loss_batch(tg, pred)
    for opt in (0, 1):
        dh.set_option(0, 5, opt)
        try:
            fpi, _, _ = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], pred)
        finally:
            dh.set_option(0, 5, 0)
        assert_close(fpi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what="256 boxes, kernel %d" % opt)
        assert int(fpi[0, 3]) == int(upi[0, 3])
    # > 128 classes: the bitmask of the stream+correct kernel does not fit, the tile kernel takes over
    Cb = 150
    boxes, nbox = synth.make_boxes(2, 256, 12, Cb, 8.0, 150.0, synth.seed_for(5, 91))
    pred = synth.fcos_predictions(2, 256, Cb, 7)
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [256, 256], Cb, [256, 256])
    upi, _ = dh.fcos.model_loss_batch(tg, pred)
    fpi, _, _ = dh.fcos.encode_loss_batch(boxes, nbox, [256, 256], Cb, [256, 256], pred)
    assert_close(fpi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what="150 classes")
    with pytest.raises(ValueError):
        dh.fcos.encode_loss_batch(boxes, nbox, [256, 256], Cb, [256, 256], pred, weights=(1, 1, 1))
    # misaligned predictions (RetinaNet, C = 80: the vector path needs 16-byte alignment)
    boxes, nbox = synth.make_boxes(2, 256, 12, 80, 8.0, 150.0, synth.seed_for(5, 92))
    pred = synth.retina_predictions(2, 256, 80, 8)
    aligned = [torch.from_numpy(p).cuda() for p in pred]
    shifted = []
    for p in aligned:
        buf = torch.empty(p.numel() + 1, device="cuda")
        buf[1:].copy_(p.reshape(-1))
        shifted.append(buf[1:].view(p.shape))
        assert shifted[-1].data_ptr() % 16 == 4
    a = dh.retinanet.encode_loss_batch(boxes, nbox, [256, 256], 80, [256, 256], aligned)
    b = dh.retinanet.encode_loss_batch(boxes, nbox, [256, 256], 80, [256, 256], shifted)
    assert_close(b[0].cpu().numpy(), a[0].cpu().numpy(), 2e-6, what="misaligned") and torch.equal(a[2], b[2])


@pytest.mark.parametrize("classes", [20, 12, 28, 1])
def test_retina_fused_loss_row_lengths(classes):
    """Row lengths whose float4 count shares a factor with the warp size (C = 20 -> 6 float4 per row, 12 -> 4, 28 -> 8)
    and a non-vector row (C = 1 -> 5 floats): fused == unfused, with gradients."""
    dh = _dh()
    boxes, nbox = synth.make_boxes(2, 256, 12, classes, 8.0, 150.0, synth.seed_for(5, 93))
    pred = synth.retina_predictions(2, 256, classes, 9)
    lab, _ = dh.retinanet.format_data_batch(boxes, nbox, [256, 256], classes, [256, 256])
    upi, utot, ug = dh.retinanet.loss_batch(lab, pred, weights=(1.0, 1.0))
    fpi, ftot, _, fg = dh.retinanet.encode_loss_batch(boxes, nbox, [256, 256], classes, [256, 256], pred, weights=(1.0, 1.0))
    assert_close(fpi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what="C=%d" % classes)
    for a, b in zip(fg, ug):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("classes", [80, 20, 12])
@pytest.mark.parametrize("logits", ["prior N(-4.6, 1)", "wide N(-2, 3)", "all below -0.7", "one above -0.7 per image"])
def test_retina_fused_forward_small_logit_form(logits, classes):
    """The forward gamma == 2 kernel evaluates the label-0 focal term of a warp batch whose logits are all <= -0.7 as
    e^3 P(e) and any other batch through sigmoid and softplus: every mix of the two must agree with the loss over
    materialised targets (which never uses the polynomial) and with the float64 oracle to the 1e-5 bar."""
    dh = _dh()
    B = 3
    boxes, nbox = synth.make_boxes(B, 320, 30, classes, 8.0, 250.0, synth.seed_for(5, 94))
    pred = synth.retina_predictions(B, 320, classes, 10, logit_sigma=1.0)
    rng = np.random.default_rng(12)
    for p in pred:
        if logits.startswith("wide"):
            p[..., 4:] = rng.normal(-2.0, 3.0, size=p[..., 4:].shape).astype(np.float32)
        elif logits.startswith("all below"):
            p[..., 4:] = np.minimum(p[..., 4:], np.float32(-0.7))
        elif logits.startswith("one above"):
            p[..., 4:] = np.minimum(p[..., 4:], np.float32(-0.7))
            p[:, 0, 0, 0, 4] = 3.0
    pred[0][0, 1, 2, 3, 5] = np.float32(-0.7)          # exactly the switch point
    pred[0][0, 1, 2, 3, 6] = np.nextafter(np.float32(-0.7), np.float32(0))
    lab, _ = dh.retinanet.format_data_batch(boxes, nbox, [320, 320], classes, [320, 320])
    upi, utot = dh.retinanet.loss_batch(lab, pred)[:2]
    fpi, ftot = dh.retinanet.encode_loss_batch(boxes, nbox, [320, 320], classes, [320, 320], pred)[:2]
    assert_close(fpi.cpu().numpy(), upi.cpu().numpy(), 2e-6, what=logits)
    want = np.zeros(2)
    for l in range(5):
        y = lab[l].cpu().numpy().astype(np.float64)
        x = pred[l].astype(np.float64)
        s = 1.0 / (1.0 + np.exp(-x[..., 4:]))
        sp_pos, sp_neg = np.logaddexp(0.0, -x[..., 4:]), np.logaddexp(0.0, x[..., 4:])
        yl = y[..., 4:]
        want[0] += float(np.sum(yl * 0.25 * (1 - s) ** 2 * sp_pos + (1 - yl) * 0.75 * s ** 2 * sp_neg))
    assert_close(ftot[:1].cpu().numpy(), want[:1], RTOL, what=logits + " vs float64")


def test_bench_configuration_parity():
    """The configuration bench.py times (BASELINE configs[4] on RetinaNet-COCO shapes: 256 images of 640 x 640, 80 classes,
    <= 100 boxes, 9 anchors x 5 levels; the same generators): three sampled images of dh_retina_encode_loss against the oracle
    (1e-5 relative on the sums, positive and pair counts exact), the per-image rows against the total, the gradient
    variant's sums against the forward ones and its level-0 gradient of one image against the oracle's analytic gradient."""
    dh = _dh()
    B, side, C = 256, 640, 80
    boxes, nbox = synth.config_boxes("retina_coco", B, synth.seed_for(5, 100))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1000)
    pred = []
    for h in (80, 40, 20, 10, 5):
        p = torch.empty((B, 9, h, h, C + 4), device="cuda")
        p[..., :4].uniform_(-1, 2, generator=gen)
        p[..., 4:].normal_(-4.595, 1.0, generator=gen)
        pred.append(p)
    pi, tot, pairs = dh.retinanet.encode_loss_batch(boxes, nbox, [side, side], C, [side, side], pred)
    pi_h, tot_h, pairs_h = pi.cpu().numpy(), tot.cpu().numpy(), pairs.cpu().numpy()
    for b in (0, 97, 255):
        lab, n_pairs = O.retina_format_data(boxes[b, :nbox[b]], [side, side], C)
        want = O.retina_train_loss(lab, [[p[b, a].cpu().numpy() for a in range(9)] for p in pred])
        n_pos = sum(int((m[..., 4:].max(-1) > 0).sum()) for lv in lab for m in lv)
        assert_close(pi_h[b, :2], _f(*want), RTOL, what="bench shape, image %d" % b)
        assert int(pi_h[b, 3]) == n_pos and int(pairs_h[b]) == n_pairs
    s = pi_h.astype(np.float64).sum(axis=0)
    assert np.all(np.abs(s[:2] - tot_h[:2]) <= 1e-6 * np.abs(s[:2])) and s[3] == tot_h[3]
    gpi, gtot, gpairs, grads = dh.retinanet.encode_loss_batch(boxes, nbox, [side, side], C, [side, side], pred, weights=(1.0, 1.0))
    assert torch.allclose(gpi, pi, rtol=2e-6, atol=0) and torch.equal(gpairs, pairs) and torch.equal(gpi[:, 3], pi[:, 3])
    b = 97
    lab, _ = O.retina_format_data(boxes[b, :nbox[b]], [side, side], C)
    want = O.dense_loss_grad(np.stack(lab[0]), pred[0][b].cpu().numpy(), weights=(1.0, 1.0, 0.0), reg_ch=4, cen_mode=0, pos_rule="gt0")
    err = np.abs(grads[0][b].cpu().numpy() - want)
    assert np.all(err <= 2e-5 * np.maximum(1.0, np.abs(want))), err.max()


def test_retina_loss_batch_of_one_returns_gradients_too():
    """`loss_batch(..., weights=)` used to drop the gradients for a one-image batch."""
    dh = _dh()
    boxes, nbox = synth.make_boxes(1, 256, 12, 20, 8.0, 150.0, synth.seed_for(5, 95))
    pred = synth.retina_predictions(1, 256, 20, 9)
    lab, _ = dh.retinanet.format_data_batch(boxes, nbox, [256, 256], 20, [256, 256])
    out = dh.retinanet.loss_batch(lab, pred, weights=(1.0, 0.5))
    assert len(out) == 3 and len(out[2]) == 5
    pi, tot = dh.retinanet.loss_batch(lab, pred)
    assert torch.equal(out[0], pi) and torch.equal(out[1], tot)
    want = O.dense_loss_grad(lab[2][0].cpu().numpy(), pred[2][0], weights=(1.0, 0.5, 0.0), reg_ch=4, cen_mode=0, pos_rule="gt0")
    assert np.all(np.abs(out[2][2][0].cpu().numpy() - want) <= 2e-5 * np.maximum(1.0, np.abs(want)))


@pytest.mark.parametrize("batch", [16, 32, 40, 64, 96])
def test_chunk_plans_agree_at_the_bench_shapes(batch):
    """The host plans the fused kernel's chunks differently per batch size (equal chunks of up to 18 tiles for small
    launches, tiers of 16 / 8 / 4 / 2 tiles otherwise); whatever the plan, the per-image sums are those of the uniform
    8-tile plan (DH_OPT_FUSED_TAIL = 0) to float32 summation order, and the positive counts are the same integers.  (A
    plan-specific bug -- rows of a chunk's 17th and 18th tile never resolved -- once passed every test that only ran the
    tiered plans.)"""
    dh = _dh()
    from densehead import _capi
    boxes, nbox = synth.config_boxes("retina_coco", batch, synth.seed_for(5, 400 + batch))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(batch)
    pred = []
    for h in (80, 40, 20, 10, 5):
        p = torch.empty((batch, 9, h, h, 84), device="cuda")
        p[..., :4].uniform_(-1, 2, generator=gen)
        p[..., 4:].normal_(-4.595, 1.0, generator=gen)
        pred.append(p)
    res = {}
    for tail in (1, 0):
        dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, tail)
        try:
            pi, tot, pairs = dh.retinanet.encode_loss_batch(boxes, nbox, [640, 640], 80, [640, 640], pred)
        finally:
            dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, 1)
        res[tail] = (pi.cpu().numpy(), tot.cpu().numpy(), pairs.cpu().numpy())
    assert np.array_equal(res[1][0][:, 3], res[0][0][:, 3]) and np.array_equal(res[1][2], res[0][2])
    assert np.allclose(res[1][0][:, :3], res[0][0][:, :3], rtol=2e-6, atol=1e-4)
    assert np.allclose(res[1][1][:3], res[0][1][:3], rtol=2e-6)
    # and one image of the batch against the oracle
    b = batch // 2
    lab, n_pairs = O.retina_format_data(boxes[b, :nbox[b]], [640, 640], 80)
    want = O.retina_train_loss(lab, [[p[b, a].cpu().numpy() for a in range(9)] for p in pred])
    got = res[1][0][b]
    assert int(res[1][2][b]) == n_pairs
    assert_close(got[:2], np.array([float(want[0]), float(want[1])]), 1e-5, what="image %d of %d" % (b, batch))



@pytest.mark.parametrize("family,batch", [("fcos", 256), ("centernet", 64), ("centernet", 32)])
def test_chunk_plans_agree_for_the_other_policies(family, batch):
    """Same check as above for the FCOS and CenterNet policies at batch sizes where the planner picks equal chunks (FCOS-VOC
    at 256 images: 512 chunks of 12 tiles; CenterNet-s8 stride 4 at 64 / 32 images: 18-tile chunks)."""
    dh = _dh()
    from densehead import _capi
    rng = np.random.default_rng(batch)
    if family == "fcos":
        boxes, nbox = synth.config_boxes("fcos_voc", batch, synth.seed_for(1, 600))
        pred = [rng.normal(-3.0, 1.0, size=(batch, 512 // s, 512 // s, 25)).astype(np.float32) for s in (8, 16, 32, 64, 128)]
        pred = [torch.from_numpy(p).cuda() for p in pred]
        call = lambda: dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], 20, [512, 512], pred)
    else:
        boxes, nbox = synth.config_boxes("centernet_crowdhuman", batch, synth.seed_for(2, 600))
        pred = torch.from_numpy(rng.normal(-3.0, 1.0, size=(batch, 128, 128, 5, 5)).astype(np.float32)).cuda()
        call = lambda: dh.centernet.encode_loss_batch(boxes, nbox, [512, 512], 1, [512, 512], pred, stride=4, mode="s8",
                                                      box_scales=[32, 64, 128, 256, 512])
    res = {}
    for tail in (1, 0):
        dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, tail)
        try:
            out = call()
        finally:
            dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, 1)
        res[tail] = (out[0].cpu().numpy(), out[1].cpu().numpy())
    assert np.array_equal(res[1][0][:, 3], res[0][0][:, 3])
    assert np.allclose(res[1][0][:, :3], res[0][0][:, :3], rtol=2e-6, atol=1e-4)
    assert np.allclose(res[1][1][:3], res[0][1][:3], rtol=2e-6)
