"""CUDA target encoders (through the C ABI) vs the CPU oracle and the golden vectors frozen from the
reference.  Integer decisions (which cells are positive, class bits, pair counts) and float32
targets must be bit-identical to `reference_map.astype(float32)`; the float64-derived channels
(centerness, RetinaNet offsets, fall-off heat) are held to 1e-6 relative."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dense_head_ref as O  # noqa: E402
from oracle import synth  # noqa: E402
from conftest import assert_close  # noqa: E402

SCALES = [32, 64, 128, 256, 512]
FCOS_ORACLE = {
    "fcos": (O.fcos_format_data, {}),
    "center": (O.fcos_center_format_data, {}),
    "center_only": (O.fcos_center_format_data, {"center_only": True}),
    "center_v1": (O.fcos_center_v1_format_data, {}),
}


def _dh():
    import densehead
    return densehead


def _same(got, want, what, soft_channels=()):
    """bit-exact everywhere except `soft_channels` (float64-derived), which get 1e-6 relative."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.array_equal(got != 0, want != 0), "%s: support (positive mask / class bits) differs" % what
    hard = [c for c in range(got.shape[-1]) if c not in soft_channels]
    assert np.array_equal(got[..., hard], want[..., hard]), "%s: exact channels differ" % what
    for c in soft_channels:
        assert_close(got[..., c], want[..., c], rtol=1e-6, atol_floor=0, what="%s ch%d" % (what, c))


@pytest.mark.parametrize("mode", ["fcos", "center", "center_only", "center_v1"])
@pytest.mark.parametrize("tag", ["c1", "c1b", "s384", "pad", "tiny", "coco"])
def test_fcos_family_golden(golden, tag, mode):
    dh = _dh()
    z = golden("fcos_encode")
    g, meta = z[tag + "_g"], z[tag + "_meta"]
    img_dim, img_pad, classes = [meta[0], meta[1]], [int(meta[2]), int(meta[3])], int(meta[4])
    name = {"center_v1": "v1"}.get(mode, mode)
    fn = {"fcos": dh.fcos.format_data, "center": dh.fcos.format_data_center,
          "center_only": lambda *a, **k: dh.fcos.format_data_center(*a, center_only=True, **k),
          "center_v1": dh.fcos.format_data_center_v1}[mode]
    outs, cnt = fn(g, img_dim, classes, img_pad=img_pad)
    assert cnt == z["%s_%s_counts" % (tag, name)].tolist()
    for l, o in enumerate(outs):
        _same(o.cpu().numpy(), z["%s_%s_L%d" % (tag, name, l)], "%s/%s L%d" % (tag, mode, l),
              soft_channels=(4,) if mode == "fcos" else ())


@pytest.mark.parametrize("mode", ["fcos", "center", "center_only", "center_v1"])
def test_fcos_batch_vs_oracle(mode):
    """C1-shaped batch (8 x 512^2, 20 classes, <= 20 boxes) and ragged/empty images."""
    dh = _dh()
    boxes, nbox = synth.config_boxes("fcos_voc", 8, synth.seed_for(1, 40))
    nbox[3] = 0  # an image without GT
    outs, cnt = dh.fcos.format_data_batch(boxes, nbox, [512, 512], 20, [512, 512], mode=mode)
    fn, kw = FCOS_ORACLE[mode]
    for b in range(8):
        want, wcnt = fn(boxes[b, :nbox[b]], [512, 512], 20, **kw)
        assert cnt[b].tolist() == wcnt
        for l in range(5):
            _same(outs[l][b].cpu().numpy(), want[l], "%s b%d L%d" % (mode, b, l), soft_channels=(4,) if mode == "fcos" else ())


def test_fcos_per_image_dims_and_options():
    """img_dim differs per image inside one padded batch; TMA-store and st.global paths agree."""
    dh = _dh()
    boxes, nbox = synth.make_boxes(5, 384, 24, 7, 6.0, 300.0, 4242)
    dims = np.array([[384, 384], [320, 384], [384, 352], [300, 300], [384, 384]], dtype=np.float32)
    outs, _ = dh.fcos.format_data_batch(boxes, nbox, dims, 7, [384, 384])
    ref = [o.clone() for o in outs]
    for b in range(5):
        want, _ = O.fcos_format_data(boxes[b, :nbox[b]], dims[b], 7, img_pad=[384, 384])
        for l in range(5):
            _same(outs[l][b].cpu().numpy(), want[l], "dims b%d L%d" % (b, l), soft_channels=(4,))
    try:
        dh.set_option(0, 1, 0)  # DH_OPT_TMA_STORE = 0
        outs2, _ = dh.fcos.format_data_batch(boxes, nbox, dims, 7, [384, 384])
        dh.set_option(0, 2, 4096)  # small tiles: many tiles per CTA, exercises buffer recycling
        dh.set_option(0, 1, 1)
        outs3, _ = dh.fcos.format_data_batch(boxes, nbox, dims, 7, [384, 384])
        dh.set_option(0, 3, 1)
        outs4, _ = dh.fcos.format_data_batch(boxes, nbox, dims, 7, [384, 384])
    finally:
        dh.set_option(0, 1, 1), dh.set_option(0, 2, 49152), dh.set_option(0, 3, 4)
    for a, b2, c, d in zip(ref, outs2, outs3, outs4):
        assert torch.equal(a, b2) and torch.equal(a, c) and torch.equal(a, d)


def test_output_buffers_are_fully_overwritten():
    dh = _dh()
    boxes, nbox = synth.make_boxes(3, 256, 10, 4, 8.0, 200.0, 7)
    shapes = dh.fcos.level_shapes([256, 256], [8, 16, 32, 64, 128])
    out = [torch.full((3, h, w, 9), float("nan"), device="cuda") for h, w in shapes]
    dh.fcos.format_data_batch(boxes, nbox, [256, 256], 4, [256, 256], out=out)
    for o in out:
        assert not torch.isnan(o).any()


@pytest.mark.parametrize("tag", ["s256", "c3", "thr4", "a20"])
def test_retina_golden(golden, tag):
    dh = _dh()
    z = golden("retina_encode")
    side, thr = int(z[tag + "_meta"][0]), float(z[tag + "_meta"][1])
    head = dh.retinanet.RetinaNetHead(80, anchor_sizes=[20., 40., 80., 160., 320.] if tag == "a20" else None)
    outs, n = head.format_data(z[tag + "_g"], [side, side], iou_thresh=thr)
    assert n == int(z[tag + "_pairs"])
    for l in range(5):
        got = torch.stack(outs[l]).cpu().numpy()
        _same(got, z["%s_L%d" % (tag, l)], "%s L%d" % (tag, l), soft_channels=(0, 1, 2, 3))


def test_retina_batch_vs_oracle():
    dh = _dh()
    boxes, nbox = synth.make_boxes(4, 320, 40, 80, 8.0, 250.0, synth.seed_for(3, 40))
    nbox[1] = 0
    outs, pairs = dh.retinanet.format_data_batch(boxes, nbox, [320, 320], 80, [320, 320])
    for b in range(4):
        want, wp = O.retina_format_data(boxes[b, :nbox[b]], [320, 320], 80)
        assert int(pairs[b]) == wp
        for l in range(5):
            _same(outs[l][b].cpu().numpy(), np.stack(want[l]), "retina b%d L%d" % (b, l), soft_channels=(0, 1, 2, 3))


@pytest.mark.parametrize("tag", ["c2s8", "c2s4", "pad", "s16"])
def test_centernet_golden(golden, tag):
    dh = _dh()
    z = golden("centernet_encode")
    g, m = z[tag + "_g"], z[tag + "_meta"]
    img_dim, img_pad, classes, stride = [int(m[0]), int(m[1])], [int(m[2]), int(m[3])], int(m[4]), int(m[5])
    out, n = dh.centernet.format_data_s8(g, SCALES, img_dim, classes, img_pad=img_pad, stride=stride)
    assert n == len(g)
    _same(out.cpu().numpy(), z[tag + "_s8"], tag + " s8")
    out, n = dh.centernet.format_data_hourglass(g, img_dim, classes, img_pad=img_pad, stride=stride)
    _same(out.cpu().numpy(), z[tag + "_hg"], tag + " hg")
    out = dh.centernet.format_data(g, img_dim, classes, img_pad=img_pad, stride=stride)
    _same(out.cpu().numpy(), z[tag + "_cn"], tag + " cn", soft_channels=(4,))


def test_centernet_batch_c2_shape_and_errors():
    """C2: 32 x 512^2, stride 4, <= 150 boxes, 1 class; plus the ValueError the reference raises."""
    dh = _dh()
    boxes, nbox = synth.config_boxes("centernet_crowdhuman", 32, synth.seed_for(2, 40))
    out, status = dh.centernet.format_data_batch(boxes, nbox, [512, 512], 1, [512, 512], stride=4, mode="s8", box_scales=SCALES)
    assert int(status[0]) == 0
    for b in (0, 7, 31):
        want, _ = O.centernet_s8_format_data(boxes[b, :nbox[b]], SCALES, [512, 512], 1, stride=4)
        _same(out[b].cpu().numpy(), want, "c2 b%d" % b)
    out, _ = dh.centernet.format_data_batch(boxes, nbox, [512, 512], 1, [512, 512], stride=4, mode="falloff")
    for b in (0, 31):
        want = O.centernet_format_data(boxes[b, :nbox[b]], [512, 512], 1, stride=4)
        _same(out[b].cpu().numpy(), want, "c2 falloff b%d" % b, soft_channels=(4,))
    with pytest.raises(ValueError):
        dh.centernet.format_data_s8(np.array([[.5, .5, 1., 1., 0]], np.float32), SCALES, [512, 512], 1)
    big = np.array([[[.5, .5, 1., 1., 0]]], np.float32)
    _, status = dh.centernet.format_data_batch(big, [1], [512, 512], 1, [512, 512], mode="s8", box_scales=SCALES)
    assert int(status[0]) == 1


def test_full_size_properties():
    """Size-independent checks at BASELINE's full sizes: idempotence (same input -> same bytes),
    positive-row counts equal the oracle's on sampled images, batch order independence."""
    dh = _dh()
    boxes, nbox = synth.config_boxes("retina_coco", 64, synth.seed_for(3, 60))
    a, pa = dh.retinanet.format_data_batch(boxes, nbox, [640, 640], 80, [640, 640])
    b, pb = dh.retinanet.format_data_batch(boxes, nbox, [640, 640], 80, [640, 640])
    assert torch.equal(pa, pb) and all(torch.equal(x, y) for x, y in zip(a, b))
    perm = np.random.default_rng(1).permutation(64)
    c, pc = dh.retinanet.format_data_batch(boxes[perm], nbox[perm], [640, 640], 80, [640, 640])
    inv = torch.as_tensor(perm, device="cuda")
    assert torch.equal(pc, pa[inv]) and all(torch.equal(x, y[inv]) for x, y in zip(c, a))
    for bi in (0, 63):
        want, wp = O.retina_format_data(boxes[bi, :nbox[bi]], [640, 640], 80)
        assert int(pa[bi]) == wp
        for l in range(5):
            _same(a[l][bi].cpu().numpy(), np.stack(want[l]), "c3 b%d L%d" % (bi, l), soft_channels=(0, 1, 2, 3))
    # checksum of checksums: every level sum equals the oracle's for a FCOS batch of 256
    boxes, nbox = synth.config_boxes("fcos_voc", 256, synth.seed_for(5, 60))
    outs, cnt = dh.fcos.format_data_batch(boxes, nbox, [512, 512], 20, [512, 512])
    for bi in (0, 100, 255):
        want, wc = O.fcos_format_data(boxes[bi, :nbox[bi]], [512, 512], 20)
        assert cnt[bi].tolist() == wc
        for l in range(5):
            _same(outs[l][bi].cpu().numpy(), want[l], "c5 b%d L%d" % (bi, l), soft_channels=(4,))


@pytest.mark.parametrize("family,batch", [("fcos", 256), ("centernet", 80), ("centernet", 256), ("retina", 16), ("retina", 3)])
def test_both_encoder_kernels_agree_at_every_chunking_rule(family, batch):
    """The direct-store kernel cuts its chunks by output size and boxes per image (14 chunks per SM from 128 MB on, 3 for
    images with more than 64 boxes, 4 below 24 MB, 8 otherwise); whatever the rule, it writes the bytes the tile streamer
    writes (which the golden and oracle tests above pin)."""
    dh = _dh()
    from densehead import _capi
    if family == "fcos":
        boxes, nbox = synth.config_boxes("fcos_voc", batch, synth.seed_for(1, 500 + batch))
        call = lambda: dh.fcos.format_data_batch(boxes, nbox, [512, 512], 20, [512, 512])[0]
    elif family == "centernet":
        boxes, nbox = synth.config_boxes("centernet_crowdhuman", batch, synth.seed_for(2, 500 + batch))
        call = lambda: [dh.centernet.format_data_batch(boxes, nbox, [512, 512], 1, [512, 512], stride=4, mode="s8", box_scales=SCALES)[0]]
    else:
        boxes, nbox = synth.config_boxes("retina_coco", batch, synth.seed_for(3, 500 + batch))
        call = lambda: dh.retinanet.format_data_batch(boxes, nbox, [640, 640], 80, [640, 640])[0]
    outs = {}
    for kern in (1, 2):
        dh.set_option(0, _capi.DH_OPT_ENCODE_KERNEL, kern)
        try:
            outs[kern] = [o.clone() for o in call()]
        finally:
            dh.set_option(0, _capi.DH_OPT_ENCODE_KERNEL, 0)
    assert all(torch.equal(a, b) for a, b in zip(outs[1], outs[2]))
    assert all(torch.equal(a, b) for a, b in zip(outs[1], call()))  # and the default choice
