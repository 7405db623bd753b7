"""Seeded synthetic inputs for the dense-head path.

TEST / BENCH INFRASTRUCTURE.  Pure NumPy; shared by `tests/`, `oracle/make_golden.py` and
`bench.py` so that the CUDA path, the oracle and the frozen golden vectors all see the same
bytes for a given (config, seed).

Boxes are float32 rows `(cy, cx, h, w, class)` normalised by the image side, padded to
`[B, Nmax, 5]` with `nbox[B]`.  The generator keeps, per image: pairwise-distinct areas (the
reference's `np.argsort` is unstable on ties), and every box edge / centre / sigma-shrunk edge
at least `1e-3` cells away from a cell boundary at the finest stride, so integer decisions do
not hinge on the last float32 bit.
"""
import numpy as np

CONFIGS = {
    # name: (image side, Nmax, classes, min box side, max box side)
    "fcos_voc": (512, 20, 20, 8.0, 0.6 * 512),
    "centernet_crowdhuman": (512, 150, 1, 8.0, 400.0),
    "retina_coco": (640, 100, 80, 8.0, 0.6 * 640),
    "fcos_coco": (640, 100, 80, 8.0, 0.6 * 640),
}


def seed_for(config_index, batch_index=0):
    return 20240000 + 100 * config_index + batch_index


def _margin_ok(cy, cx, h, w, side, fine_stride=4.0, eps=1e-3):
    vals = []
    for c, d in ((cy, h), (cx, w)):
        for k in (0.5, 0.125):           # full box edges, sigma=.25 shrunk edges
            vals += [(c - k * d) * side / fine_stride, (c + k * d) * side / fine_stride]
        vals.append(c * side / fine_stride)
        vals.append(c * side / fine_stride + 0.5)
    v = np.asarray(vals, dtype=np.float64)
    return bool(np.all(np.abs(v - np.round(v)) > eps))


def make_boxes(batch, side, nmax, classes, lo, hi, seed, full=False):
    """-> (boxes float32 [B, Nmax, 5], nbox int32 [B])"""
    rng = np.random.default_rng(seed)
    boxes = np.zeros((batch, nmax, 5), dtype=np.float32)
    nbox = np.zeros((batch,), dtype=np.int32)
    for b in range(batch):
        n = nmax if full else int(rng.integers(1, nmax + 1))
        areas = set()
        k = 0
        while k < n:
            hh = float(np.exp(rng.uniform(np.log(lo), np.log(hi))))
            ww = float(np.exp(rng.uniform(np.log(lo), np.log(hi))))
            cy = float(rng.uniform(hh / 2, side - hh / 2))
            cx = float(rng.uniform(ww / 2, side - ww / 2))
            row = np.array([cy / side, cx / side, hh / side, ww / side], dtype=np.float32)
            area = float((row[2] * np.float32(side)) * (row[3] * np.float32(side)))
            if area in areas or not _margin_ok(*[float(v) for v in row], side):
                continue
            areas.add(area)
            boxes[b, k, :4] = row
            boxes[b, k, 4] = float(rng.integers(0, classes))
            k += 1
        nbox[b] = n
    return boxes, nbox


def config_boxes(name, batch, seed, full=False):
    side, nmax, classes, lo, hi = CONFIGS[name]
    return make_boxes(batch, side, nmax, classes, lo, hi, seed, full=full)


def level_shapes(side, strides=(8, 16, 32, 64, 128)):
    return [(int(side / s), int(side / s)) for s in strides]


def fcos_predictions(batch, side, classes, seed, strides=(8, 16, 32, 64, 128)):
    """Per-level `[B, Hl, Wl, C+5]` float32: regs ~ U(0,4), centerness logit ~ N(0,1),
    class logits ~ N(-4.595, 1) (the focal prior, FCOS/fcos.py:12-13)."""
    rng = np.random.default_rng(seed + 7)
    out = []
    for hl, wl in level_shapes(side, strides):
        p = np.empty((batch, hl, wl, classes + 5), dtype=np.float32)
        p[..., :4] = rng.uniform(0, 4, size=p[..., :4].shape)
        p[..., 4] = rng.normal(0, 1, size=p[..., 4].shape)
        p[..., 5:] = rng.normal(-4.595, 1, size=p[..., 5:].shape)
        out.append(p)
    return out


def retina_predictions(batch, side, classes, seed, n_anchors=9, strides=(8, 16, 32, 64, 128),
                       logit_sigma=1.0):
    """Per-level `[B, A, Hl, Wl, C+4]` float32 (our packed layout; `[:, a]` is the reference's
    per-anchor head `x_pred[level][a]`)."""
    rng = np.random.default_rng(seed + 11)
    out = []
    for hl, wl in level_shapes(side, strides):
        p = np.empty((batch, n_anchors, hl, wl, classes + 4), dtype=np.float32)
        p[..., :4] = rng.uniform(-1, 2, size=p[..., :4].shape)
        p[..., 4:] = rng.normal(-4.595, logit_sigma, size=p[..., 4:].shape)
        out.append(p)
    return out


def centernet_s8_predictions(batch, side, stride, n_scales, classes, seed):
    rng = np.random.default_rng(seed + 13)
    h = int(side / stride)
    p = np.empty((batch, h, h, n_scales, classes + 4), dtype=np.float32)
    p[..., :4] = rng.uniform(0, 1, size=p[..., :4].shape)
    p[..., 4:] = rng.normal(-4.595, 1, size=p[..., 4:].shape)
    return p


def nms_candidates(n, side, seed, classes=80, clusters=40):
    """`[n, 6]` float32 (y1, x1, y2, x2, score, label): boxes jittered around a few cluster
    centres so suppression chains are non-trivial; scores unique."""
    rng = np.random.default_rng(seed + 17)
    cen = rng.uniform(0.1 * side, 0.9 * side, size=(clusters, 2))
    sz = np.exp(rng.uniform(np.log(16), np.log(0.4 * side), size=(clusters, 2)))
    which = rng.integers(0, clusters, size=n)
    c = cen[which] + rng.normal(0, 6.0, size=(n, 2))
    s = sz[which] * np.exp(rng.normal(0, 0.15, size=(n, 2)))
    dets = np.zeros((n, 6), dtype=np.float32)
    dets[:, 0], dets[:, 1] = c[:, 0] - s[:, 0] / 2, c[:, 1] - s[:, 1] / 2
    dets[:, 2], dets[:, 3] = c[:, 0] + s[:, 0] / 2, c[:, 1] + s[:, 1] / 2
    score = rng.uniform(0.05, 1.0, size=n).astype(np.float32)
    order = np.argsort(score)
    score[order] = np.linspace(0.05, 0.999, n, dtype=np.float32)   # unique by construction
    dets[:, 4] = score
    dets[:, 5] = (which % classes).astype(np.float32)
    return dets
