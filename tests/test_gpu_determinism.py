"""Run-to-run determinism as a race detector (compute-sanitizer is not available on the GPU pool): every kernel family
is run repeatedly on the same inputs and must reproduce its first result bit for bit."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import synth  # noqa: E402

SC = [32, 64, 128, 256, 512]
REPS = 12


def _dh():
    import densehead
    return densehead


def _same(a, b):
    if isinstance(a, (list, tuple)):
        return all(_same(x, y) for x, y in zip(a, b))
    if a is None:
        return b is None
    return torch.equal(a, b)


def _valid(keep, n_keep):
    """keep[b, n_keep[b]:] is unspecified (the caller's buffer is not cleared): mask it out before comparing."""
    idx = torch.arange(keep.shape[1], device=keep.device).unsqueeze(0)
    return torch.where(idx < n_keep.unsqueeze(1), keep, torch.full_like(keep, -1))


def _repeat(fn):
    first = fn()
    torch.cuda.synchronize()
    for _ in range(REPS):
        again = fn()
        assert _same(first, again)


def test_encoders_and_losses_are_deterministic():
    dh = _dh()
    boxes, nbox = synth.config_boxes("retina_coco", 6, synth.seed_for(7, 1))
    pred = [torch.from_numpy(p).cuda() for p in synth.retina_predictions(6, 640, 80, 3)]
    _repeat(lambda: dh.retinanet.format_data_batch(boxes, nbox, [640, 640], 80, [640, 640]))
    _repeat(lambda: dh.retinanet.encode_loss_batch(boxes, nbox, [640, 640], 80, [640, 640], pred, weights=(1.0, 1.0)))
    lab, _ = dh.retinanet.format_data_batch(boxes, nbox, [640, 640], 80, [640, 640])
    _repeat(lambda: dh.retinanet.loss_batch(lab, pred, weights=(1.0, 1.0)))
    boxes, nbox = synth.config_boxes("centernet_crowdhuman", 8, synth.seed_for(7, 2))
    _repeat(lambda: dh.centernet.format_data_batch(boxes, nbox, [512, 512], 1, [512, 512], stride=4, mode="s8", box_scales=SC))
    boxes, nbox = synth.config_boxes("fcos_voc", 16, synth.seed_for(7, 3))
    fp = [torch.from_numpy(p).cuda() for p in synth.fcos_predictions(16, 512, 20, 4)]
    _repeat(lambda: dh.fcos.format_data_batch(boxes, nbox, [512, 512], 20, [512, 512]))
    _repeat(lambda: dh.fcos.encode_loss_batch(boxes, nbox, [512, 512], 20, [512, 512], fp, weights=(1.0, 1.0, 1.0)))


@pytest.mark.parametrize("nms_kernel", [1, 2])
def test_detection_pipelines_are_deterministic(nms_kernel):
    dh = _dh()
    dh.set_option(0, 6, nms_kernel)
    try:
        heads = [torch.from_numpy(p).cuda() for p in synth.retina_predictions(4, 640, 80, synth.seed_for(7, 4), logit_sigma=2.5)]
        def retina():
            cand, keep, n_keep, rows = dh.retinanet.detect_batch(heads, 80, [640, 640], pre_nms_topk=1000, with_rows=True)
            return cand, _valid(keep, n_keep), n_keep, rows
        _repeat(retina)
        fh = synth.fcos_predictions(4, 640, 80, synth.seed_for(7, 5))
        for h in fh:
            h[..., 5:] = h[..., 5:] * 2.5 + 6.9
        fh = [torch.from_numpy(h).cuda() for h in fh]
        _repeat(lambda: dh.fcos.detect_batch(fh, 80, [640, 640], pre_nms_topk=1000, with_candidates=True))
        d = torch.from_numpy(synth.nms_candidates(5000, 640, 9)[None]).cuda()
        from densehead import infer
        def nms(**kw):
            keep, n_keep = infer.nms(d, 0.5, **kw)
            return _valid(keep, n_keep), n_keep
        _repeat(nms)
        _repeat(lambda: nms(mode=infer.NMS_PER_CLASS, min_score=0.05, score_inclusive=False, num_classes=80,
                            max_per_class=100, max_total=100, max_out=100))
    finally:
        dh.set_option(0, 6, 0)
