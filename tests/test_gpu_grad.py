"""Loss gradients on the GPU (dh_*_grad: SURVEY 8f-1) against the oracle's analytic float64 gradients; the loss sums
returned by the same call must equal the forward-only entry points bit for bit."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dense_head_ref as O  # noqa: E402
from oracle import synth  # noqa: E402

SCALES = [32, 64, 128, 256, 512]
W = (1.25, 0.5, 2.0)


def _dh():
    import densehead
    return densehead


def _close(got, want, what, rtol=2e-5):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want)
    tol = rtol * np.maximum(1.0, np.abs(want)) * 1.0
    assert np.all(err <= np.maximum(tol, 2e-6 * np.abs(want).max())), "%s: max err %g at %r (want %g)" % (
        what, err.max(), np.unravel_index(err.argmax(), err.shape), want.flat[err.argmax()])


@pytest.mark.parametrize("mode,reg,cen", [("fcos", "l1", "l1"), ("fcos", "iou", "l1"), ("center", "l1", "focal"), ("center_v1", "l1", "focal")])
def test_fcos_gradients_fused_and_unfused(mode, reg, cen):
    dh = _dh()
    B, C = 3, 20
    boxes, nbox = synth.config_boxes("fcos_voc", B, synth.seed_for(6, 10))
    pred = synth.fcos_predictions(B, 512, C, 61)
    if reg == "iou":
        for p in pred:
            p[..., :4] = np.abs(p[..., :4]) + 0.3      # tblr distances are positive in a trained head
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [512, 512], C, [512, 512], mode=mode)
    pi, tot, cnt, grads = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], pred, mode=mode, reg_type=reg,
                                                   cen_type=cen, weights=W)
    pi0, tot0, _ = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], pred, mode=mode, reg_type=reg, cen_type=cen)
    # (the forward and the gradient kernel batch 7 and 5 float4s per lane, so a few small logits take the polynomial form of
    # the label-0 term in one and the general form in the other: a few 1e-7, not bit for bit; counts are exact)
    assert torch.allclose(pi, pi0, rtol=2e-6, atol=0) and torch.allclose(tot, tot0, rtol=2e-6, atol=0) and torch.equal(pi[:, 3], pi0[:, 3])
    upi, utot, ugrads = dh.fcos.model_loss_batch(tg, pred, reg, cen, weights=W)
    cen_mode = 1 if cen == "l1" else 2
    for l in range(5):
        want = O.dense_loss_grad(tg[l].cpu().numpy(), pred[l], weights=W, reg_ch=4, cen_mode=cen_mode,
                                 reg_mode=1 if reg == "iou" else 0, pos_rule="ge1")
        _close(grads[l].cpu().numpy(), want, "fused level %d" % l)
        _close(ugrads[l].cpu().numpy(), want, "unfused level %d" % l)


def test_retina_gradients_fused_and_unfused():
    dh = _dh()
    B, C = 2, 80
    boxes, nbox = synth.make_boxes(B, 320, 40, C, 8.0, 250.0, synth.seed_for(6, 11))
    pred = synth.retina_predictions(B, 320, C, 62)
    lab, _ = dh.retinanet.format_data_batch(boxes, nbox, [320, 320], C, [320, 320])
    pi, tot, pairs, grads = dh.retinanet.encode_loss_batch(boxes, nbox, [320, 320], C, [320, 320], pred, weights=(W[0], W[1]))
    pi0, tot0, _ = dh.retinanet.encode_loss_batch(boxes, nbox, [320, 320], C, [320, 320], pred)
    # the forward-only gamma == 2 kernel evaluates the label-0 term of small logits as e^3 P(e) (6.4e-7 relative), the
    # gradient kernel through sigmoid and softplus: the loss values agree to a few 1e-7, not bit for bit
    assert torch.allclose(pi, pi0, rtol=2e-6, atol=0) and torch.allclose(tot, tot0, rtol=2e-6, atol=0)
    upi, utot, ugrads = dh.retinanet.loss_batch(lab, pred, weights=(W[0], W[1]))
    for l in range(5):
        want = O.dense_loss_grad(lab[l].cpu().numpy(), pred[l], weights=(W[0], W[1], 0.0), reg_ch=4, cen_mode=0, pos_rule="gt0")
        _close(grads[l].cpu().numpy(), want, "fused level %d" % l)
        _close(ugrads[l].cpu().numpy(), want, "unfused level %d" % l)
    # other focal parameters take the generic pow path
    _, _, _, g15 = dh.retinanet.encode_loss_batch(boxes, nbox, [320, 320], C, [320, 320], pred, alpha=0.4, gamma=1.5, weights=(1.0, 1.0))
    want = O.dense_loss_grad(lab[0].cpu().numpy(), pred[0], weights=(1.0, 1.0, 0.0), alpha=0.4, gamma=1.5)
    _close(g15[0].cpu().numpy(), want, "gamma 1.5", rtol=5e-5)


@pytest.mark.parametrize("mode", ["s8", "hourglass", "falloff"])
def test_centernet_gradients(mode):
    dh = _dh()
    B, C = 2, 3
    boxes, nbox = synth.make_boxes(B, 512, 60, C, 8.0, 400, synth.seed_for(6, 12))
    if mode == "s8":
        yp = synth.centernet_s8_predictions(B, 512, 8, 5, C, 63)
        kw = dict(stride=8, mode="s8", box_scales=SCALES)
    elif mode == "hourglass":
        yp = np.ascontiguousarray(synth.centernet_s8_predictions(B, 512, 8, 5, C, 63)[:, :, :, 0, :])
        kw = dict(stride=8, mode="hourglass")
    else:
        rng = np.random.default_rng(64)
        yp = rng.normal(-2.0, 1.5, size=(B, 64, 64, C + 5)).astype(np.float32)
        yp[..., :4] = rng.uniform(0.2, 5.0, size=(B, 64, 64, 4)).astype(np.float32)
        kw = dict(stride=8, mode="falloff")
    yt, _ = dh.centernet.format_data_batch(boxes, nbox, [512, 512], C, [512, 512], **kw)
    pi, tot, st, grad = dh.centernet.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], yp, weights=W, **kw)
    pi0, tot0, _ = dh.centernet.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], yp, **kw)
    assert torch.allclose(pi, pi0, rtol=2e-6, atol=0) and torch.allclose(tot, tot0, rtol=2e-6, atol=0) and torch.equal(pi[:, 3], pi0[:, 3])
    if mode == "falloff":
        want = O.dense_loss_grad(yt.cpu().numpy(), yp, weights=W, reg_ch=4, cen_mode=1, pos_rule="ge1")
    else:
        want = O.dense_loss_grad(yt.cpu().numpy(), yp, weights=W, reg_ch=4, cen_mode=0, pos_rule="gt0")
    _close(grad.cpu().numpy(), want, mode)


def test_autograd_wrapper_and_grad_only_call():
    dh = _dh()
    from densehead import losses
    B, C = 2, 20
    boxes, nbox = synth.config_boxes("fcos_voc", B, synth.seed_for(6, 13))
    pred = [torch.from_numpy(p).cuda().requires_grad_(True) for p in synth.fcos_predictions(B, 512, C, 65)]

    def run(weights):
        _, tot, _, grads = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], [p.detach() for p in pred], weights=weights)
        return tot, grads
    loss = losses.WeightedLoss.apply(run, W, *pred)
    (loss * 0.5).backward()
    _, tot, _, grads = dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], C, [512, 512], [p.detach() for p in pred], weights=W)
    lv = float(loss.detach())
    assert abs(lv - float((tot[:3].cpu() * torch.tensor(W)).sum())) <= 1e-5 * abs(lv)
    for p, g in zip(pred, grads):
        assert torch.allclose(p.grad, 0.5 * g)


# ---- against the gradients pinned by the reference itself (tests/golden/grad.npz: central differences through the
#      reference's own model_loss / train_loss in float64, oracle/make_golden.py: grad) ----------------------------------
import os  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grad.npz"))
W_GOLD = tuple(float(v) for v in GOLD["weights"])


def _fd_close(got, want, what):
    err = np.abs(np.asarray(got, np.float64) - want)
    assert np.all(err <= 2e-5 * np.maximum(1.0, np.abs(want))), "%s: max err %g (want %g)" % (what, err.max(), want[err.argmax()])


@pytest.mark.parametrize("name,mode,reg,cen", [("fcos_l1", "fcos", "l1", "l1"), ("fcos_iou", "fcos", "iou", "l1"),
                                               ("center_focal", "center", "l1", "focal"), ("center_l1", "center", "l1", "l1"),
                                               ("v1", "center_v1", "l1", "focal")])
def test_fcos_gradients_against_the_reference_pinned_golden(name, mode, reg, cen):
    dh = _dh()
    g, seed = GOLD["fcos_g"], int(GOLD["fcos_seed"])
    pred = synth.fcos_predictions(1, 256, 20, seed)
    for p in pred:
        p[..., :4] = np.abs(p[..., :4]) + np.float32(0.3)
    boxes = np.zeros((1, (len(g) + 3) & ~3, 5), np.float32)
    boxes[0, :len(g)] = g
    nbox = np.array([len(g)], np.int32)
    _, _, _, grads = dh.fcos.encode_loss_batch(boxes, nbox, [256, 256], 20, [256, 256], pred, mode=mode, reg_type=reg, cen_type=cen,
                                              weights=W_GOLD)
    tg, _ = dh.fcos.format_data_batch(boxes, nbox, [256, 256], 20, [256, 256], mode=mode)
    _, _, ugrads = dh.fcos.model_loss_batch(tg, pred, reg, cen, weights=W_GOLD)
    for l in range(5):
        idx, want = GOLD["%s_idx%d" % (name, l)], GOLD["%s_fd%d" % (name, l)]
        _fd_close(grads[l].cpu().numpy().reshape(-1)[idx], want, "fused %s level %d" % (name, l))
        _fd_close(ugrads[l].cpu().numpy().reshape(-1)[idx], want, "unfused %s level %d" % (name, l))


def test_retina_gradients_against_the_reference_pinned_golden():
    dh = _dh()
    g, seed = GOLD["retina_g"], int(GOLD["retina_seed"])
    pred = synth.retina_predictions(1, 128, 20, seed)
    boxes = np.zeros((1, (len(g) + 3) & ~3, 5), np.float32)
    boxes[0, :len(g)] = g
    nbox = np.array([len(g)], np.int32)
    _, _, _, grads = dh.retinanet.encode_loss_batch(boxes, nbox, [128, 128], 20, [128, 128], pred, weights=(W_GOLD[0], W_GOLD[1]))
    for l in range(5):
        _fd_close(grads[l].cpu().numpy().reshape(-1)[GOLD["retina_idx%d" % l]], GOLD["retina_fd%d" % l], "retina level %d" % l)


@pytest.mark.parametrize("name,mode", [("cn_s8", "s8"), ("cn_hg", "hourglass")])
def test_centernet_gradients_against_the_reference_pinned_golden(name, mode):
    dh = _dh()
    boxes, nbox, seed = GOLD["cn_boxes"], GOLD["cn_nbox"], int(GOLD["cn_seed"])
    yp = synth.centernet_s8_predictions(2, 256, 8, 5, 3, seed)
    kw = dict(stride=8, mode="s8", box_scales=SCALES)
    if mode == "hourglass":
        yp = np.ascontiguousarray(yp[:, :, :, 0, :])
        kw = dict(stride=8, mode="hourglass")
    b4 = np.zeros((2, (boxes.shape[1] + 3) & ~3, 5), np.float32)
    b4[:, :boxes.shape[1]] = boxes
    _, _, _, grad = dh.centernet.encode_loss_batch(b4, nbox, [256, 256], 3, [256, 256], yp, weights=W_GOLD, **kw)
    _fd_close(grad.cpu().numpy().reshape(-1)[GOLD[name + "_idx"]], GOLD[name + "_fd"], name)


def test_gradient_does_not_depend_on_the_chunk_plan():
    """32 COCO-shaped images: the planner picks one 18-tile chunk per CTA (DH_OPT_FUSED_TAIL = 1) or uniform 8-tile chunks
    (= 0); the gradient of every element is the same bits either way, and so are the positive counts."""
    import densehead as dh
    from densehead import _capi
    B = 32
    boxes, nbox = synth.config_boxes("retina_coco", B, synth.seed_for(5, 700))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    pred = []
    for h in (80, 40, 20, 10, 5):
        p = torch.empty((B, 9, h, h, 84), device="cuda")
        p[..., :4].uniform_(-1, 2, generator=gen)
        p[..., 4:].normal_(-4.595, 1.0, generator=gen)
        pred.append(p)
    res = {}
    for tail in (1, 0):
        dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, tail)
        try:
            pi, tot, pairs, grads = dh.retinanet.encode_loss_batch(boxes, nbox, [640, 640], 80, [640, 640], pred, weights=(1.0, 1.0))
        finally:
            dh.set_option(0, _capi.DH_OPT_FUSED_TAIL, 1)
        res[tail] = (pi.clone(), pairs.clone(), [g.clone() for g in grads])
    assert torch.equal(res[1][0][:, 3], res[0][0][:, 3]) and torch.equal(res[1][1], res[0][1])
    assert all(torch.equal(a, b) for a, b in zip(res[1][2], res[0][2]))
    assert torch.allclose(res[1][0][:, :3], res[0][0][:, :3], rtol=2e-6, atol=1e-4)
