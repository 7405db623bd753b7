"""CPU oracle for the dense-head path: a NumPy restatement of the reference's algorithms.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module, and only as the checker /
the timed CPU baseline -- never on the product path.  The product (`densehead`, the C-ABI in
`include/densehead.h`) fails loudly if the CUDA library is missing; it never falls back here.

Parity pin: this file is checked against the reference's own source executed under the
TensorFlow stub (`oracle/ref_loader.py`) by `tests/test_oracle_vs_reference.py` (live, when
/root/reference exists) and against the frozen outputs of that same run in `tests/golden/`
(always).  The reference ships no tests or golden vectors of its own.
`combined_nms` is the one exception: its arithmetic lives in TensorFlow's C++ kernel
(`tf.image.combined_non_max_suppression`, version unpinned, absent from /root/reference), so
that routine is "parity unpinned" -- it restates the op's documented behaviour and is
cross-checked against torchvision's batched NMS only.

Conventions (all citations are relative to /root/reference):
  * GT rows are `(cy, cx, h, w, class)` normalised by `img_dim` (FCOS/fcos.py:139).
  * Arithmetic is float32 in the reference's operation order where the reference is float32
    (NumPy >= 2 scalar rules, TF tensors float32) and float64 where the reference is float64
    (centerness, RetinaNet regression targets, CenterNet power fall-off); maps are returned
    as float32 == `reference_map.astype(np.float32)`.
  * `int()` on a float truncates toward zero.
  * Equal-area ties are ordered by original index (stable sort); the reference's
    `np.argsort` is unstable there, so generators keep areas distinct.
"""
import numpy as np

F = np.float32
_HALF = F(0.5)

DEFAULT_STRIDES = (8, 16, 32, 64, 128)
DEFAULT_B_DIM = (32, 64, 128, 256)


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _labels(gt_labels):
    g = np.array(gt_labels, dtype=np.float32, copy=True)
    return g.reshape(-1, 5)


def _level_of(gt_h, gt_w, b_dim, n_levels):
    """FCOS/fcos.py:168-179 -- one pyramid level per GT from max(w, h)."""
    d = np.maximum(gt_w, gt_h)
    lvl = np.full(d.shape, n_levels - 1, dtype=np.int64)
    assigned = np.zeros(d.shape, dtype=bool)
    for n in range(n_levels - 1):
        lo_ok = np.ones(d.shape, bool) if n == 0 else d >= F(b_dim[n - 1])
        m = lo_ok & (d < F(b_dim[n])) & ~assigned
        lvl[m] = n
        assigned |= m
    # last level: d >= b_dim[-1]; anything unassigned and below b_dim[-1] cannot exist.
    return lvl


def _pixel_corners(row, hi, wi):
    """FCOS/fcos.py:211-215 -- (y0, x0, y1, x1) in pixels, float32."""
    cy, cx, h, w = row[0], row[1], row[2], row[3]
    return ((cy - _HALF * h) * hi, (cx - _HALF * w) * wi,
            (cy + _HALF * h) * hi, (cx + _HALF * w) * wi)


def _ratio64(a, b):
    """FCOS/fcos.py:262-271 -- (min+1e-8)/(max+1e-8) in float64."""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return (np.minimum(a, b) + 1.0e-8) / (np.maximum(a, b) + 1.0e-8)


# --------------------------------------------------------------------------------------
# a1  FCOS/fcos.py:136-378
# --------------------------------------------------------------------------------------
def fcos_format_data(gt_labels, img_dim, num_classes, img_pad=None, strides=None, b_dim=None, order="reference"):
    """Per-level FCOS targets `[Hl, Wl, C+5]` (t, b, l, r, centerness, multi-hot classes)
    and the per-level GT counts.  Follows FCOS/fcos.py:136-378 (Appendix A.1 of SURVEY.md).

    `order="min_area"` is an EXTENSION and the specification of DH_FCOS_MIN_AREA: the reference paints the boxes of a level
    in ASCENDING area order, so where footprints overlap the largest box supplies channels 0..4 (:202-209) although its
    own comment (:185-188) -- and the FCOS paper -- want the smallest; this mode paints in DESCENDING area order (stable:
    of two equal areas the higher index is painted later), everything else unchanged.
    """
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    b_dim = list(DEFAULT_B_DIM if b_dim is None else b_dim)
    g = _labels(gt_labels)
    hi, wi = F(img_dim[0]), F(img_dim[1])
    pad = (img_dim[0], img_dim[1]) if img_pad is None else img_pad
    gt_h, gt_w = g[:, 2] * hi, g[:, 3] * wi
    lvl = _level_of(gt_h, gt_w, b_dim, len(strides))
    area = gt_h * gt_w                                          # :202-204
    outs, counts = [], []
    for n, s in enumerate(strides):
        hl, wl = int(float(pad[0]) / s), int(float(pad[1]) / s)
        out = np.zeros((hl, wl, num_classes + 5), dtype=np.float32)
        idx = np.nonzero(lvl == n)[0]
        counts.append(int(idx.size))
        if idx.size:
            if order == "min_area":
                idx = idx[np.argsort(-area[idx], kind="stable")]    # descending: smallest painted last (extension)
            else:
                idx = idx[np.argsort(area[idx], kind="stable")]     # ascending: largest painted last
        sf = F(s)
        h_ratio, w_ratio = hi / sf, wi / sf                     # :162-163
        for k in idx:
            row = g[k]
            y0, x0, y1, x1 = _pixel_corners(row, hi, wi)
            y0s, x0s, y1s, x1s = y0 / sf, x0 / sf, y1 / sf, x1 / sf
            half_h, half_w = row[2] / F(2), row[3] / F(2)
            y_low = max(0, int((row[0] - half_h) * h_ratio) + 1)    # :217-225
            x_low = max(0, int((row[1] - half_w) * w_ratio) + 1)
            y_upp = min(int((row[0] + half_h) * h_ratio) + 1, hl)
            x_upp = min(int((row[1] + half_w) * w_ratio) + 1, wl)
            y_cen = min(int(0.5 * (y_low + y_upp)), hl - 1)          # :227-230
            x_cen = min(int(0.5 * (x_low + x_upp)), wl - 1)
            live_y, live_x = y_upp > y_low, x_upp > x_low
            if live_y:
                ys = slice(y_low, y_upp)
                gy = np.arange(y_low, y_upp, dtype=np.float32) + _HALF
                t = np.maximum(F(0), gy - y0s)
                b = np.maximum(F(0), y1s - gy)
            else:
                ys = slice(y_cen, y_cen + 1)                          # :320-355 / :356-374
                t = np.maximum(F(0), F(y_cen + 0.5) - y0s).reshape(1)
                b = np.maximum(F(0), (y1s - F(y_cen)) - _HALF).reshape(1)
            if live_x:
                xs = slice(x_low, x_upp)
                gx = np.arange(x_low, x_upp, dtype=np.float32) + _HALF
                l = np.maximum(F(0), gx - x0s)
                r = np.maximum(F(0), x1s - gx)
            else:
                xs = slice(x_cen, x_cen + 1)                          # :284-319 / :356-374
                l = np.maximum(F(0), F(x_cen + 0.5) - x0s).reshape(1)
                r = np.maximum(F(0), (x1s - F(x_cen)) - _HALF).reshape(1)
            if ys.start < 0 or xs.start < 0:
                continue  # the reference would wrap a negative index; boxes outside the image are out of contract
            ny, nx = t.size, l.size
            out[ys, xs, 0] = np.broadcast_to(t[:, None], (ny, nx))
            out[ys, xs, 1] = np.broadcast_to(b[:, None], (ny, nx))
            out[ys, xs, 2] = np.broadcast_to(l[None, :], (ny, nx))
            out[ys, xs, 3] = np.broadcast_to(r[None, :], (ny, nx))
            q_y = _ratio64(t, b) if live_y else np.ones(1)
            q_x = _ratio64(l, r) if live_x else np.ones(1)
            cen = np.sqrt(q_y[:, None] * q_x[None, :])              # :273-274 (float64)
            out[ys, xs, 4] = cen.astype(np.float32)
            out[y_cen, x_cen, 4] = 1.0                                # :279-280
            out[ys, xs, 5 + int(row[4])] = 1.0                        # :281-283
        outs.append(out)
    return outs, counts


# --------------------------------------------------------------------------------------
# a2  FCOS/fcos_center.py:149-279
# --------------------------------------------------------------------------------------
def fcos_center_format_data(gt_labels, img_dim, num_classes, img_pad=None, b_dim=None,
                            strides=None, center_only=False):
    """3x3 (or centre-only) assignment with 1 / .5 / .25 centre scores and unclipped tblr."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    b_dim = list(DEFAULT_B_DIM if b_dim is None else b_dim)
    g = _labels(gt_labels)
    hi, wi = F(img_dim[0]), F(img_dim[1])
    pad = (img_dim[0], img_dim[1]) if img_pad is None else img_pad
    gt_h, gt_w = g[:, 2] * hi, g[:, 3] * wi
    lvl = _level_of(gt_h, gt_w, b_dim, len(strides))
    area = gt_h * gt_w
    offs = (0,) if center_only else (-1, 0, 1)
    outs, counts = [], []
    for n, s in enumerate(strides):
        hl, wl = int(float(pad[0]) / s), int(float(pad[1]) / s)
        out = np.zeros((hl, wl, num_classes + 5), dtype=np.float32)
        idx = np.nonzero(lvl == n)[0]
        counts.append(int(idx.size))
        if idx.size:
            idx = idx[np.argsort(area[idx], kind="stable")]
        sf = F(s)
        h_ratio, w_ratio = hi / sf, wi / sf
        for k in idx:
            row = g[k]
            y0, x0, y1, x1 = _pixel_corners(row, hi, wi)
            y0s, x0s, y1s, x1s = y0 / sf, x0 / sf, y1 / sf, x1 / sf
            y_cen = int(row[0] * h_ratio + _HALF)                    # fcos_center.py:231-232
            x_cen = int(row[1] * w_ratio + _HALF)
            for dx in offs:
                j = x_cen - dx
                if j < 0 or j >= wl:
                    continue
                for dy in offs:
                    i = y_cen - dy
                    if i < 0 or i >= hl:
                        continue
                    score = 1.0 if (dx == 0 and dy == 0) else (0.25 if (dx != 0 and dy != 0) else 0.5)
                    if score >= out[i, j, 4]:                        # :263-265
                        out[i, j, 4] = score
                    out[i, j, 0] = F(i + 0.5) - y0s                  # :267-273 (unclipped)
                    out[i, j, 1] = (y1s - F(i)) - _HALF
                    out[i, j, 2] = F(j + 0.5) - x0s
                    out[i, j, 3] = (x1s - F(j)) - _HALF
                    out[i, j, 5 + int(row[4])] = 1.0
        outs.append(out)
    return outs, counts


# --------------------------------------------------------------------------------------
# a3  FCOS/fcos_center_v1.py:149-258
# --------------------------------------------------------------------------------------
def fcos_center_v1_format_data(gt_labels, img_dim, num_classes, img_pad=None, b_dim=None,
                               strides=None, center_only=False):
    """Centre-cell only; regs = (off_y, off_x, h/box_sc, w/box_sc).  `center_only` is accepted
    and unused, as in the reference."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    b_dim = list(DEFAULT_B_DIM if b_dim is None else b_dim)
    g = _labels(gt_labels)
    hi, wi = F(img_dim[0]), F(img_dim[1])
    pad = (img_dim[0], img_dim[1]) if img_pad is None else img_pad
    gt_h, gt_w = g[:, 2] * hi, g[:, 3] * wi
    lvl = _level_of(gt_h, gt_w, b_dim, len(strides))
    area = gt_h * gt_w
    outs, counts = [], []
    for n, s in enumerate(strides):
        hl, wl = int(float(pad[0]) / s), int(float(pad[1]) / s)
        out = np.zeros((hl, wl, num_classes + 5), dtype=np.float32)
        idx = np.nonzero(lvl == n)[0]
        counts.append(int(idx.size))
        if idx.size:
            idx = idx[np.argsort(area[idx], kind="stable")]
        sf = F(s)
        box_sc = F(b_dim[n]) if n < len(strides) - 1 else max(hi, wi)   # fcos_center_v1.py:182-196
        for k in idx:
            row = g[k]
            box_h, box_w = row[2] * hi, row[3] * wi
            raw_y, raw_x = row[0] * hi, row[1] * wi
            i, j = int(raw_y / sf), int(raw_x / sf)
            if not (0 <= i < hl and 0 <= j < wl):
                continue
            out[i, j, 0] = (raw_y - F(i * s)) / sf
            out[i, j, 1] = (raw_x - F(j * s)) / sf
            out[i, j, 2] = box_h / box_sc
            out[i, j, 3] = box_w / box_sc
            out[i, j, 4] = 1.0
            out[i, j, 5 + int(row[4])] = 1.0
        outs.append(out)
    return outs, counts


# --------------------------------------------------------------------------------------
# a4-a6  RetinaNet/retinanet_module.py:205-365, RetinaNet/utils.py:42-83
# --------------------------------------------------------------------------------------
def retina_anchor_dims(anchor_sizes=None, aspect_ratios=None, anchor_scales=None):
    """float32 `[5, 9, 2]` (h, w) table -- RetinaNet/retinanet_module.py:201-219.
    Index a = 3*ratio_idx + scale_idx."""
    sizes = [32.0, 64.0, 128.0, 256.0, 512.0] if anchor_sizes is None else list(anchor_sizes)
    ratios = [0.5, 1.0, 2.0] if aspect_ratios is None else list(aspect_ratios)
    scales = [2 ** x for x in (0, 1 / 3, 2 / 3)] if anchor_scales is None else list(anchor_scales)
    if len(sizes) != 5:
        raise ValueError("anchor_sizes must be of dimension 5.")
    if len(scales) != 3:
        raise ValueError("anchor_scales must be of dimension 3.")
    table = []
    for area in sorted(x ** 2 for x in sizes):
        level = []
        for ratio in ratios:
            ah = np.sqrt(F(area / ratio))            # tf.math.sqrt on a float32 tensor
            aw = F(area) / ah
            for sc in scales:
                level.append((F(sc) * ah, F(sc) * aw))
        table.append(level)
    return np.array(table, dtype=np.float32)


def compute_iou(boxes1, boxes2):
    """Pairwise IoU of centre-size boxes, float32 -- RetinaNet/utils.py:42-83."""
    b1 = _f32(boxes1).reshape(-1, 4)
    b2 = _f32(boxes2).reshape(-1, 4)
    lo1, hi1 = b1[:, :2] - b1[:, 2:] / F(2), b1[:, :2] + b1[:, 2:] / F(2)
    lo2, hi2 = b2[:, :2] - b2[:, 2:] / F(2), b2[:, :2] + b2[:, 2:] / F(2)
    side = np.maximum(F(0), np.minimum(hi1[:, None, :], hi2[None]) - np.maximum(lo1[:, None, :], lo2[None]))
    inter = side[..., 0] * side[..., 1]
    union = np.maximum((b1[:, 2] * b1[:, 3])[:, None] + (b2[:, 2] * b2[:, 3])[None] - inter, F(1e-8))
    return np.clip(inter / union, F(0), F(1))


def retina_format_data(gt_labels, img_dim, num_classes, anchor_dims=None, iou_thresh=0.5,
                       img_pad=None, strides=None):
    """`out[level][anchor]` float32 `[Hl, Wl, C+4]` maps and the positive-pair count.
    RetinaNet/retinanet_module.py:251-365 (Appendix A.4)."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    dims = retina_anchor_dims() if anchor_dims is None else _f32(anchor_dims)
    g = _labels(gt_labels)
    hi, wi = F(img_dim[0]), F(img_dim[1])
    pad = (img_dim[0], img_dim[1]) if img_pad is None else img_pad
    gpx = g.copy()
    gpx[:, :4] = g[:, :4] * np.array([hi, wi, hi, wi], dtype=np.float32)   # :274-278
    thr = F(iou_thresh)
    n_pairs = 0
    outs = []
    for n, s in enumerate(strides):
        hl, wl = int(float(pad[0]) / s), int(float(pad[1]) / s)
        ii, jj = np.meshgrid(np.arange(hl), np.arange(wl), indexing="ij")
        cy = (ii.reshape(-1) * s).astype(np.float32)
        cx = (jj.reshape(-1) * s).astype(np.float32)
        level = []
        for a in range(dims.shape[1]):
            ah, aw = dims[n, a]
            out = np.zeros((hl, wl, num_classes + 4), dtype=np.float32)
            if len(gpx):
                anc = np.stack([cy, cx, np.full_like(cy, ah), np.full_like(cx, aw)], axis=1)
                pos = compute_iou(gpx[:, :4], anc) > thr                   # :297-302
                n_pairs += int(pos.sum())
                flat = out.reshape(-1, num_classes + 4)
                for k in np.nonzero(pos.any(axis=1))[0]:                   # ascending gt: last wins
                    p = np.nonzero(pos[k])[0]
                    gy, gx, gh, gw = (np.float64(v) for v in gpx[k, :4])
                    flat[p, 0] = ((cy[p].astype(np.float64) - gy) / np.float64(ah)).astype(np.float32)
                    flat[p, 1] = ((cx[p].astype(np.float64) - gx) / np.float64(aw)).astype(np.float32)
                    flat[p, 2] = np.float32(gh / np.float64(ah))
                    flat[p, 3] = np.float32(gw / np.float64(aw))
                    flat[p, 4 + int(gpx[k, 4])] = 1.0
            level.append(out)
        outs.append(level)
    return outs, n_pairs


# --------------------------------------------------------------------------------------
# a16-a18  CenterNet encoders
# --------------------------------------------------------------------------------------
def _ascending_area(g, hi, wi):
    if len(g) <= 1:
        return np.arange(len(g))
    return np.argsort((g[:, 2] * hi) * (g[:, 3] * wi), kind="stable")


def centernet_s8_format_data(gt_labels, box_scales, img_dim, num_classes, img_pad=None, stride=8):
    """`[H, W, S, C+4]` one-hot-centre targets -- CenterNet/tf_centernet_resnet_s8.py:243-330.
    Raises ValueError when a box is not smaller than the largest scale (`min([])` there)."""
    g = _labels(gt_labels)
    hi, wi = F(img_dim[0]), F(img_dim[1])
    pad = (img_dim[0], img_dim[1]) if img_pad is None else img_pad
    sf = F(stride)
    h_max, w_max = int(float(pad[1]) / stride), int(float(pad[0]) / stride)      # :259-262 (index-swapped)
    pad_y = F(int((float(pad[1]) - float(img_dim[1])) / 2.0))
    pad_x = F(int((float(pad[0]) - float(img_dim[0])) / 2.0))
    scales = [F(v) for v in box_scales]
    out = np.zeros((h_max, w_max, len(scales), num_classes + 4), dtype=np.float32)
    for k in _ascending_area(g, hi, wi):
        row = g[k]
        y0, x0, y1, x1 = _pixel_corners(row, hi, wi)
        bh, bw = y1 - y0, x1 - x0
        d = max(bh, bw)
        fits = [n for n, sc in enumerate(scales) if d < sc]
        if not fits:
            raise ValueError("box side %r is not below the largest box scale" % float(d))
        n_sc = fits[0]
        yc, xc = (y0 + y1) / F(2), (x0 + x1) / F(2)
        i, j = int((pad_y + yc) / sf), int((pad_x + xc) / sf)                      # :310-313
        out[i, j, n_sc, 0] = ((pad_y + yc) - F(i * stride)) / sf
        out[i, j, n_sc, 1] = ((pad_x + xc) - F(j * stride)) / sf
        out[i, j, n_sc, 2] = bh / scales[n_sc]
        out[i, j, n_sc, 3] = bw / scales[n_sc]
        out[i, j, n_sc, 4 + int(row[4])] = 1.0
    return out, len(g)


def centernet_hourglass_format_data(gt_labels, img_dim, num_classes, img_pad=None, stride=8):
    """`[H, W, C+4]` centre-cell targets with FCOS-style offsets --
    CenterNet/tf_centernet_hourglass.py:379-456."""
    g = _labels(gt_labels)
    hi, wi = F(img_dim[0]), F(img_dim[1])
    pad = (img_dim[0], img_dim[1]) if img_pad is None else img_pad
    sf = F(stride)
    h_max, w_max = int(float(pad[1]) / stride), int(float(pad[0]) / stride)
    pad_y = F(int((float(pad[1]) - float(img_dim[1])) / 2.0))
    pad_x = F(int((float(pad[0]) - float(img_dim[0])) / 2.0))
    out = np.zeros((h_max, w_max, num_classes + 4), dtype=np.float32)
    for k in _ascending_area(g, hi, wi):
        row = g[k]
        y0, x0, y1, x1 = _pixel_corners(row, hi, wi)
        yc, xc = (y0 + y1) / F(2), (x0 + x1) / F(2)
        i, j = int((pad_y + yc) / sf), int((pad_x + xc) / sf)
        out[i, j, 0] = F(i + 0.5) - (pad_y + y0) / sf                               # :445-449
        out[i, j, 1] = ((pad_y + y1) / sf - F(i)) - _HALF
        out[i, j, 2] = F(j + 0.5) - (pad_x + x0) / sf
        out[i, j, 3] = ((pad_x + x1) / sf - F(j)) - _HALF
        out[i, j, 4 + int(row[4])] = 1.0
    return out, len(g)


def _falloff64(coords, mu, spread=8.0):
    """CenterNet/tf_centernet.py:6-19 without the max-normaliser (applied by the caller)."""
    return 1.0 / np.power(np.asarray(coords, dtype=np.float64) - float(mu), spread)


def centernet_gaussian_format_data(gt_labels, img_dim, num_classes, img_pad=None, stride=8, sigma=0.25):
    """EXTENSION and the specification of DH_CENTERNET_GAUSSIAN (BASELINE's north_star asks for "Gaussian heatmap
    rendering"; the reference renders none).  `centernet_format_data` with channel 4 rendered by the
    reference's commented-out `gaussian_dist_2d` (CenterNet/tf_centernet.py:30-40): exp(-d^2 / (2 std^2)) per live axis on
    the footprint grid (cells at z + 0.5, integer mean), divided by its maximum over the footprint, with
    std = max(1, sqrt(h * w * h_ratio * w_ratio)) as :203-205 compute it before the override to 8.0 (float32), the
    exponent in float64, the centre cell forced to 1 (:261-262).  Where footprints overlap the heat is the MAXIMUM over
    the boxes (the canonical CenterNet splat -- order-free); channels 0..3 and the classes as in `centernet_format_data`."""
    return centernet_format_data(gt_labels, img_dim, num_classes, img_pad, stride, sigma, heat="gaussian")


def centernet_format_data(gt_labels, img_dim, num_classes, img_pad=None, stride=8, sigma=0.25, heat="falloff"):
    """Single-level `[H, W, C+5]` map: sigma-shrunk footprint, tblr of the full box, inverse-power
    fall-off heat (`tmp_std` forced to 8.0) -- CenterNet/tf_centernet.py:152-342."""
    heat_mode = heat
    g = _labels(gt_labels)
    hi, wi = F(img_dim[0]), F(img_dim[1])
    pad = (img_dim[0], img_dim[1]) if img_pad is None else img_pad
    sf = F(stride)
    h_ratio, w_ratio = hi / sf, wi / sf
    hl, wl = int(float(pad[0]) / stride), int(float(pad[1]) / stride)
    clip_h, clip_w = int(float(img_dim[0]) / stride), int(float(img_dim[1]) / stride)   # :222-223
    sg = F(sigma)
    out = np.zeros((hl, wl, num_classes + 5), dtype=np.float32)
    for k in _ascending_area(g, hi, wi):
        row = g[k]
        y0, x0, y1, x1 = _pixel_corners(row, hi, wi)
        y0s, x0s, y1s, x1s = y0 / sf, x0 / sf, y1 / sf, x1 / sf
        y_cen, x_cen = int(row[0] * h_ratio), int(row[1] * w_ratio)                     # :208-209
        y_low = max(0, 1 + int((row[0] - sg * row[2] / F(2)) * h_ratio))               # :211-223
        x_low = max(0, 1 + int((row[1] - sg * row[3] / F(2)) * w_ratio))
        y_upp = min(1 + int((row[0] + sg * row[2] / F(2)) * h_ratio), clip_h)
        x_upp = min(1 + int((row[1] + sg * row[3] / F(2)) * w_ratio), clip_w)
        live_y, live_x = y_upp > y_low, x_upp > x_low
        if live_y:
            ys = slice(y_low, y_upp)
            gy = np.arange(y_low, y_upp, dtype=np.float32) + _HALF
            t, b = np.maximum(F(0), gy - y0s), np.maximum(F(0), y1s - gy)
            mu_y = int(0.5 * (y_low + y_upp))
            fy = _falloff64(gy, mu_y)
        else:
            ys = slice(y_cen, y_cen + 1)
            t = np.maximum(F(0), F(y_cen + 0.5) - y0s).reshape(1)
            b = np.maximum(F(0), (y1s - F(y_cen)) - _HALF).reshape(1)
            mu_y, fy = y_cen, np.ones(1)
        if live_x:
            xs = slice(x_low, x_upp)
            gx = np.arange(x_low, x_upp, dtype=np.float32) + _HALF
            l, r = np.maximum(F(0), gx - x0s), np.maximum(F(0), x1s - gx)
            mu_x = int(0.5 * (x_low + x_upp))
            fx = _falloff64(gx, mu_x)
        else:
            xs = slice(x_cen, x_cen + 1)
            l = np.maximum(F(0), F(x_cen + 0.5) - x0s).reshape(1)
            r = np.maximum(F(0), (x1s - F(x_cen)) - _HALF).reshape(1)
            mu_x, fx = x_cen, np.ones(1)
        if not (0 <= ys.start < hl and 0 <= xs.start < wl):
            continue
        ny, nx = t.size, l.size
        out[ys, xs, 0] = np.broadcast_to(t[:, None], (ny, nx))
        out[ys, xs, 1] = np.broadcast_to(b[:, None], (ny, nx))
        out[ys, xs, 2] = np.broadcast_to(l[None, :], (ny, nx))
        out[ys, xs, 3] = np.broadcast_to(r[None, :], (ny, nx))
        if heat_mode == "gaussian":
            std = np.maximum(F(1), np.sqrt(((row[2] * row[3]) * h_ratio) * w_ratio, dtype=np.float32))   # :203-205
            d2 = np.zeros((ny, nx), dtype=np.float64)
            if live_y:
                d2 = d2 + ((gy.astype(np.float64) - mu_y) ** 2 - 0.25)[:, None]
            if live_x:
                d2 = d2 + ((gx.astype(np.float64) - mu_x) ** 2 - 0.25)[None, :]
            heat = np.exp(-d2 / (2.0 * float(std) * float(std)))
            if 0 <= mu_y - ys.start < ny and 0 <= mu_x - xs.start < nx:
                heat[mu_y - ys.start, mu_x - xs.start] = 1.0
            out[ys, xs, 4] = np.maximum(out[ys, xs, 4], heat.astype(np.float32))
            out[ys, xs, 5 + int(row[4])] = 1.0
            continue
        if live_y or live_x:
            heat = fy[:, None] * fx[None, :]
            heat = heat / heat.max()                                                    # :9, :18
        else:
            heat = np.ones((1, 1))                                                      # :337-338
        out[ys, xs, 4] = heat.astype(np.float32)
        out[mu_y, mu_x, 4] = 1.0                                                        # :261-262
        out[ys, xs, 5 + int(row[4])] = 1.0
    return out


# --------------------------------------------------------------------------------------
# a7-a10  losses (FCOS/fcos.py:380-496 and its copies)
# --------------------------------------------------------------------------------------
def _sigmoid32(x):
    with np.errstate(over="ignore"):
        return F(1) / (F(1) + np.exp(-x))


def focal_loss(labels, logits, alpha=0.25, gamma=2.0):
    """Sum-reduced focal loss in the reference's stable form -- FCOS/fcos.py:443-462."""
    y, x = _f32(labels), _f32(logits)
    al, gm = F(alpha), F(gamma)
    soft = np.log(F(1) + np.exp(-np.abs(x)))
    sg = _sigmoid32(x)
    p_pos, p_neg = np.power(F(1) - sg, gm), np.power(sg, gm)
    absterm = y * al * soft * p_pos + p_neg * ((F(1) - y) * (F(1) - al) * soft)
    xneg = y * al * np.minimum(x, F(0)) * p_pos
    xpos = (F(1) - y) * (F(1) - al) * (np.maximum(x, F(0)) * p_neg)
    return F(np.sum(absterm + xpos - xneg, dtype=np.float32))


def smooth_l1_loss(xy_true, xy_pred, mask=1.0, delta=1.0):
    """`sum(mask * where(|d| < delta, d^2/2, |d|))` -- FCOS/fcos.py:380-391 (no -delta/2 term)."""
    d = _f32(xy_true) - _f32(xy_pred)
    ad = np.abs(d)
    per = np.where(ad < F(delta), _HALF * d * d, ad)
    m = _f32(mask)
    m = m.reshape(m.shape + (1,)) if m.ndim else m.reshape(1)
    return F(np.sum(per * m, dtype=np.float32))


def iou_loss(xy_true, xy_pred, mask):
    """`sum(-log(iou + 1e-12) * mask)` on an integer grid (no +.5) -- FCOS/fcos.py:393-441."""
    yt, yp, m = _f32(xy_true), _f32(xy_pred), _f32(mask)
    hh, ww = yp.shape[0], yp.shape[1]
    gx, gy = np.meshgrid(np.arange(ww, dtype=np.float32), np.arange(hh, dtype=np.float32))
    tb = (gy - yt[..., 0], gx - yt[..., 2], gy + yt[..., 1], gx + yt[..., 3])
    pb = (gy - yp[..., 0], gx - yp[..., 2], gy + yp[..., 1], gx + yp[..., 3])
    ih = np.maximum(F(0), np.minimum(tb[2], pb[2]) - np.maximum(tb[0], pb[0]))
    iw = np.maximum(F(0), np.minimum(tb[3], pb[3]) - np.maximum(tb[1], pb[1]))
    inter = iw * ih
    union = ((tb[2] - tb[0]) * (tb[3] - tb[1]) + (pb[2] - pb[0]) * (pb[3] - pb[1])) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (union + F(1e-12))
        return F(np.sum(F(-1) * np.log(iou + F(1e-12)) * m, dtype=np.float32))


def giou_loss(xy_true, xy_pred, mask):
    """EXTENSION -- the reference has no GIoU loss (BASELINE's north_star names "IoU/GIoU losses").
    `sum((1 - GIoU) * mask)` on exactly the box construction of `iou_loss` (FCOS/fcos.py:393-441: tblr distances around
    the integer grid point, `iou = inter / (union + 1e-12)`), with GIoU = IoU - (C - union) / (C + 1e-12), C the area of
    the smallest box enclosing both (Rezatofighi et al. 2019).  float32 in the written order; this function IS the
    specification of DH_REG_GIOU."""
    yt, yp, m = _f32(xy_true), _f32(xy_pred), _f32(mask)
    hh, ww = yp.shape[0], yp.shape[1]
    gx, gy = np.meshgrid(np.arange(ww, dtype=np.float32), np.arange(hh, dtype=np.float32))
    tb = (gy - yt[..., 0], gx - yt[..., 2], gy + yt[..., 1], gx + yt[..., 3])
    pb = (gy - yp[..., 0], gx - yp[..., 2], gy + yp[..., 1], gx + yp[..., 3])
    ih = np.maximum(F(0), np.minimum(tb[2], pb[2]) - np.maximum(tb[0], pb[0]))
    iw = np.maximum(F(0), np.minimum(tb[3], pb[3]) - np.maximum(tb[1], pb[1]))
    inter = iw * ih
    union = ((tb[2] - tb[0]) * (tb[3] - tb[1]) + (pb[2] - pb[0]) * (pb[3] - pb[1])) - inter
    ac = (np.maximum(tb[2], pb[2]) - np.minimum(tb[0], pb[0])) * (np.maximum(tb[3], pb[3]) - np.minimum(tb[1], pb[1]))
    with np.errstate(divide="ignore", invalid="ignore"):
        giou = inter / (union + F(1e-12)) - (ac - union) / (ac + F(1e-12))
        per = np.where(m != 0, (F(1) - giou) * m, F(0))
        return F(np.sum(per, dtype=np.float32))


def fcos_model_loss(y_true, y_pred, reg_type="l1", cen_type="l1", pos_rule="ge1"):
    """(cls, reg, cen) summed over levels -- FCOS/fcos.py:464-496; `cen_type="focal"` is the
    fcos_center / fcos_center_v1 variant (fcos_center.py:365-399, fcos_center_v1.py:294-317).
    `y_pred[l]` is `[Hl, Wl, C+5]` (the reference indexes `y_pred[l][0]` on a batch of one)."""
    cls = reg = cen = F(0)
    for yt, yp in zip(y_true, y_pred):
        yt, yp = _f32(yt), _f32(yp)
        obj = yt[..., 5:].max(axis=-1)
        mask = (obj >= 1).astype(np.float32) if pos_rule == "ge1" else (obj > 0).astype(np.float32)
        cls = cls + focal_loss(yt[..., 5:], yp[..., 5:])
        if cen_type.lower() == "l1":
            cen = cen + smooth_l1_loss(yt[..., 4], _sigmoid32(yp[..., 4]), mask=1.0)
        elif cen_type.lower() == "focal":
            cen = cen + focal_loss(yt[..., 4], yp[..., 4])
        if reg_type == "iou":
            reg = reg + iou_loss(yt[..., :4], yp[..., :4], mask)
        elif reg_type == "giou":  # extension
            reg = reg + giou_loss(yt[..., :4], yp[..., :4], mask)
        else:
            reg = reg + smooth_l1_loss(yt[..., :4], yp[..., :4], mask=mask)
    return F(cls), F(reg), F(cen)


def retina_train_loss(x_label, x_pred):
    """(cls, reg) over 5x9 maps; positives are `max(class) > 0` --
    RetinaNet/retinanet_module.py:403-426 (model forward removed)."""
    cls = reg = F(0)
    for lt, lp in zip(x_label, x_pred):
        for yt, yp in zip(lt, lp):
            yt, yp = _f32(yt), _f32(yp)
            mask = (yt[..., 4:].max(axis=-1) > 0).astype(np.float32)
            cls = cls + focal_loss(yt[..., 4:], yp[..., 4:])
            reg = reg + smooth_l1_loss(yt[..., :4], yp[..., :4], mask=mask)
    return F(cls), F(reg)


def centernet_s8_model_loss(y_true, y_pred):
    """Batched per-scale (cls, reg) -- CenterNet/tf_centernet_resnet_s8.py:368-385.
    `y_true`, `y_pred`: `[B, H, W, S, C+4]`."""
    yt, yp = _f32(y_true), _f32(y_pred)
    cls = reg = F(0)
    for n in range(yp.shape[3]):
        mask = (yt[:, :, :, n, 4:].max(axis=-1) > 0).astype(np.float32)
        cls = cls + focal_loss(yt[:, :, :, n, 4:], yp[:, :, :, n, 4:])
        reg = reg + smooth_l1_loss(yt[:, :, :, n, :4], yp[:, :, :, n, :4], mask=mask)
    return F(cls), F(reg)


def centernet_hourglass_model_loss(y_true, y_pred):
    """(cls, reg) -- CenterNet/tf_centernet_hourglass.py:492-505."""
    yt, yp = _f32(y_true), _f32(y_pred)
    mask = (yt[..., 4:].max(axis=-1) > 0).astype(np.float32)
    return (focal_loss(yt[..., 4:], yp[..., 4:]),
            smooth_l1_loss(yt[..., :4], yp[..., :4], mask=mask))


# --------------------------------------------------------------------------------------
# f-2  hourglass 4-scale encoder (inline in CenterNet/train_hourglass_voc.py:95-153) and its losses
# --------------------------------------------------------------------------------------
def hourglass4_format_data(gt_labels, raw_dims, img_dims, num_classes, stride=8):
    """The inline encoder of CenterNet/train_hourglass_voc.py:99-153 -> float32 [img_dims/8, img_dims/8, 4, C+5]:
    channels (h_off, w_off, h_reg, w_reg, objectness, classes...).  `gt_labels` rows are (cy, cx, h, w, class) normalised
    by the unpadded square side `raw_dims`; the image sits at offset pad = int((img_dims - raw_dims) / 2) in the padded
    `img_dims` square (:91-93).  Scales are img_dims / (8, 4, 2, 1) (:96-97); a box goes to the first scale that exceeds
    BOTH its sides, else the last (:123-138); ascending-area paint order (:109-115), last painter wins the five
    regression / objectness channels, classes OR (:148-150).  float32 arithmetic (NumPy >= 2 scalar rules)."""
    g = _labels(gt_labels)
    n_map = int(img_dims / stride)
    out = np.zeros((n_map, n_map, 4, num_classes + 5), dtype=np.float32)
    scales = [img_dims / (2 ** x) for x in range(4)][::-1]
    pad = int((img_dims - raw_dims) / 2.0)
    areas = g[:, 3] * g[:, 2] * F(100)                      # w * h * 100 (:110-111)
    for k in np.argsort(areas, kind="stable"):
        cy, cx, h, w, c = g[k]
        x_cen, y_cen = F(pad + cx * F(raw_dims)), F(pad + cy * F(raw_dims))
        bw, bh = F(w * F(raw_dims)), F(h * F(raw_dims))
        if bw < 0 or bh < 0:
            continue
        sc = 3
        for n in range(3):
            if bw < scales[n] and bh < scales[n]:
                sc = n
                break
        box_scale = F(scales[sc])
        i, j = int(y_cen / F(stride)), int(x_cen / F(stride))
        if not (0 <= i < n_map and 0 <= j < n_map):
            continue                                        # the reference would raise / wrap; out of contract
        out[i, j, sc, :5] = [(y_cen - F(i * stride)) / F(stride), (x_cen - F(j * stride)) / F(stride), bh / box_scale, bw / box_scale, 1.0]
        out[i, j, sc, 5 + int(c)] = 1.0
    return out


def sigmoid_bce_sum(labels, logits):
    """`reduce_sum(tf.nn.sigmoid_cross_entropy_with_logits)` -- CenterNet/tf_hourglass_net.py:347-349, :381-382:
    max(x, 0) - x*z + log(1 + exp(-|x|)) in float32."""
    z, x = _f32(labels), _f32(logits)
    return F(np.sum(np.maximum(x, F(0)) - x * z + np.log(F(1) + np.exp(-np.abs(x))), dtype=np.float32))


def hourglass4_model_loss(bboxes, masks, outputs, loss_type="sigmoid"):
    """(cls, reg) -- CenterNet/tf_hourglass_net.py:372-388: sigmoid BCE (or focal) summed over channels 4: (objectness
    + classes), L1 `sum(|bboxes[..., :4] - outputs[..., :4]| * masks[..., None])`."""
    yt, yp, m = _f32(bboxes), _f32(outputs), _f32(masks)
    cls = sigmoid_bce_sum(yt[..., 4:], yp[..., 4:]) if loss_type == "sigmoid" else focal_loss(yt[..., 4:], yp[..., 4:])
    reg = F(np.sum(np.abs(yt[..., :4] - yp[..., :4]) * m[..., None], dtype=np.float32))
    return cls, reg


# --------------------------------------------------------------------------------------
# f-3 / f-4  label preparation before the path, result formatting after it
# --------------------------------------------------------------------------------------
def swap_xy(boxes):
    """FCOS/utils.py:6-14 (same in RetinaNet/utils.py, CenterNet/utils.py)."""
    b = _f32(boxes)
    return np.stack([b[..., 1], b[..., 0], b[..., 3], b[..., 2]], axis=-1)


def convert_to_xywh(boxes):
    """FCOS/utils.py:16-27 -- [(lo + hi) / 2, hi - lo] in float32."""
    b = _f32(boxes)
    return np.concatenate([(b[..., :2] + b[..., 2:]) / F(2), b[..., 2:] - b[..., :2]], axis=-1)


def convert_to_corners(boxes):
    """FCOS/utils.py:29-40 -- [c - size / 2, c + size / 2] in float32."""
    b = _f32(boxes)
    return np.concatenate([b[..., :2] - b[..., 2:] / F(2), b[..., :2] + b[..., 2:] / F(2)], axis=-1)


def flip_boxes_horizontal(boxes):
    """The box half of random_flip_horizontal, FCOS/data_preprocess.py:36-39."""
    b = _f32(boxes)
    return np.stack([F(1) - b[..., 2], b[..., 1], F(1) - b[..., 0], b[..., 3]], axis=-1)


def prepare_labels(bboxes, classes, flip=False):
    """One image: dataset boxes (xmin, ymin, xmax, ymax) + class ids -> [n, 5] (cy, cx, h, w, class) --
    random_flip_horizontal (when flipped), swap_xy, convert_to_xywh (FCOS/data_preprocess.py:121-131), then the
    concat with the class column (FCOS/train_fcos.py:131-135)."""
    b = _f32(bboxes).reshape(-1, 4)
    if flip:
        b = flip_boxes_horizontal(b)
    b = convert_to_xywh(swap_xy(b))
    return np.concatenate([b, _f32(classes).reshape(-1, 1)], axis=1).astype(np.float32)


def format_detections(rows, w_ratio, h_ratio):
    """RetinaNet/retinanet_module.py:559-569 after image_detections: `swap_xy(dets[:, :4] * [wr, hr, wr, hr])` (the
    product is float64 because the ratio array is), scores, integer labels."""
    r = _f32(rows).reshape(-1, 6)
    scaled = r[:, :4].astype(np.float64) * np.array([w_ratio, h_ratio, w_ratio, h_ratio], dtype=np.float64)
    boxes = np.stack([scaled[:, 1], scaled[:, 0], scaled[:, 3], scaled[:, 2]], axis=-1).astype(np.float32)
    return boxes, r[:, 4].copy(), r[:, 5].astype(np.int32)


# --------------------------------------------------------------------------------------
# f-1  loss gradients (what tf.GradientTape yields for the formulas above; FCOS/train_fcos.py:152-176)
# --------------------------------------------------------------------------------------
# Parity unpinned: TensorFlow's autodiff is not available here.  These are the analytic derivatives of the
# reference's loss expressions in float64; tests/test_oracle_grad.py checks them against central finite
# differences of a float64 evaluation of the same expressions.
def dense_loss_f64(target, pred, reg_ch=4, cen_mode=0, reg_mode=0, pos_rule="gt0", alpha=0.25, gamma=2.0, delta=1.0,
                   mask=None):
    """(cls, reg, cen) of one [H, W, ch] (or [..., ch]) map in float64: channels [0, reg_ch) boxes, then an optional
    centerness channel (cen_mode 1 smooth-L1 of sigmoid, 2 focal, 3 ignored), then classes."""
    t, p = np.asarray(target, np.float64), np.asarray(pred, np.float64)
    cls0 = reg_ch + (1 if cen_mode else 0)
    y, x = t[..., cls0:], p[..., cls0:]
    sg = 1.0 / (1.0 + np.exp(-x))
    sp_pos, sp_neg = np.logaddexp(0.0, x), np.logaddexp(0.0, -x)
    cls = np.sum(y * alpha * (1 - sg) ** gamma * sp_neg + (1 - y) * (1 - alpha) * sg ** gamma * sp_pos)
    if mask is None:
        obj = t[..., cls0:].max(axis=-1)
        mask = (obj >= 1) if pos_rule == "ge1" else (obj > 0)
    m = np.asarray(mask, np.float64)
    reg = 0.0
    if reg_ch:
        if reg_mode == 0:
            d = t[..., :4] - p[..., :4]
            reg = np.sum(np.where(np.abs(d) < delta, 0.5 * d * d, np.abs(d)) * m[..., None])
        else:
            hh, ww = p.shape[-3], p.shape[-2]
            gx, gy = np.meshgrid(np.arange(ww, dtype=np.float64), np.arange(hh, dtype=np.float64))
            tb = (gy - t[..., 0], gx - t[..., 2], gy + t[..., 1], gx + t[..., 3])
            pb = (gy - p[..., 0], gx - p[..., 2], gy + p[..., 1], gx + p[..., 3])
            ih = np.maximum(0, np.minimum(tb[2], pb[2]) - np.maximum(tb[0], pb[0]))
            iw = np.maximum(0, np.minimum(tb[3], pb[3]) - np.maximum(tb[1], pb[1]))
            inter = iw * ih
            union = (tb[2] - tb[0]) * (tb[3] - tb[1]) + (pb[2] - pb[0]) * (pb[3] - pb[1]) - inter
            with np.errstate(divide="ignore", invalid="ignore"):
                if reg_mode == 2:  # 1 - GIoU (extension, see giou_loss)
                    ac = (np.maximum(tb[2], pb[2]) - np.minimum(tb[0], pb[0])) * (np.maximum(tb[3], pb[3]) - np.minimum(tb[1], pb[1]))
                    per = 1.0 - (inter / (union + 1e-12) - (ac - union) / (ac + 1e-12))
                else:
                    per = -np.log(inter / (union + 1e-12) + 1e-12)
            reg = np.sum(np.where(m > 0, per, 0.0) * m)
    cen = 0.0
    if cen_mode == 1:
        d = t[..., reg_ch] - 1.0 / (1.0 + np.exp(-p[..., reg_ch]))
        cen = np.sum(np.where(np.abs(d) < delta, 0.5 * d * d, np.abs(d)))
    elif cen_mode == 2:
        yc, xc = t[..., reg_ch], p[..., reg_ch]
        sc = 1.0 / (1.0 + np.exp(-xc))
        cen = np.sum(yc * alpha * (1 - sc) ** gamma * np.logaddexp(0.0, -xc) + (1 - yc) * (1 - alpha) * sc ** gamma * np.logaddexp(0.0, xc))
    return float(cls), float(reg), float(cen)


def _focal_grad64(y, x, alpha, gamma):
    sg = 1.0 / (1.0 + np.exp(-x))
    om = 1.0 - sg
    sp_pos, sp_neg = np.logaddexp(0.0, x), np.logaddexp(0.0, -x)
    d_pos = -(gamma * sg * om ** gamma * sp_neg + om ** (gamma + 1))
    d_neg = gamma * sg ** gamma * om * sp_pos + sg ** (gamma + 1)
    return y * alpha * d_pos + (1 - y) * (1 - alpha) * d_neg


def _sl1_grad64(y, x, delta):
    d = x - y
    return np.where(np.abs(d) < delta, d, np.sign(d))


def dense_loss_grad(target, pred, weights=(1.0, 1.0, 1.0), reg_ch=4, cen_mode=0, reg_mode=0, pos_rule="gt0", alpha=0.25,
                    gamma=2.0, delta=1.0, mask=None):
    """d(w_cls*cls + w_reg*reg + w_cen*cen) / d pred of `dense_loss_f64`, float64, same shape as `pred`."""
    t, p = np.asarray(target, np.float64), np.asarray(pred, np.float64)
    w_cls, w_reg, w_cen = (float(v) for v in weights)
    cls0 = reg_ch + (1 if cen_mode else 0)
    g = np.zeros_like(p)
    g[..., cls0:] = w_cls * _focal_grad64(t[..., cls0:], p[..., cls0:], alpha, gamma)
    if mask is None:
        obj = t[..., cls0:].max(axis=-1)
        mask = (obj >= 1) if pos_rule == "ge1" else (obj > 0)
    m = np.asarray(mask, np.float64)
    if reg_ch:
        if reg_mode == 0:
            g[..., :4] = w_reg * m[..., None] * _sl1_grad64(t[..., :4], p[..., :4], delta)
        else:
            hh, ww = p.shape[-3], p.shape[-2]
            gx, gy = np.meshgrid(np.arange(ww, dtype=np.float64), np.arange(hh, dtype=np.float64))
            ty0, tx0, ty1, tx1 = gy - t[..., 0], gx - t[..., 2], gy + t[..., 1], gx + t[..., 3]
            py0, px0, py1, px1 = gy - p[..., 0], gx - p[..., 2], gy + p[..., 1], gx + p[..., 3]
            ih_raw = np.minimum(ty1, py1) - np.maximum(ty0, py0)
            iw_raw = np.minimum(tx1, px1) - np.maximum(tx0, px0)
            ih, iw = np.maximum(0, ih_raw), np.maximum(0, iw_raw)
            inter = iw * ih
            ph, pw = py1 - py0, px1 - px0
            den = (ty1 - ty0) * (tx1 - tx0) + ph * pw - inter + 1e-12
            with np.errstate(divide="ignore", invalid="ignore"):
                k = -1.0 / (inter / den + 1e-12)
                di = [iw * (ih_raw > 0) * (py0 > ty0), iw * (ih_raw > 0) * (py1 < ty1),
                      ih * (iw_raw > 0) * (px0 > tx0), ih * (iw_raw > 0) * (px1 < tx1)]
                da = [pw, pw, ph, ph]
                if reg_mode == 2:  # 1 - GIoU: the enclosing box moves with the prediction where the prediction is the outer edge
                    eh, ew = np.maximum(ty1, py1) - np.minimum(ty0, py0), np.maximum(tx1, px1) - np.minimum(tx0, px0)
                    ac = eh * ew
                    dc = ac + 1e-12
                    uni = den - 1e-12
                    de = [ew * (py0 < ty0), ew * (py1 > ty1), eh * (px0 < tx0), eh * (px1 > tx1)]
                for q in range(4):
                    d_uni = da[q] - di[q]
                    d_iou = (di[q] * den - inter * d_uni) / (den * den)
                    if reg_mode == 2:
                        gq = -(d_iou - ((de[q] - d_uni) * dc - (ac - uni) * de[q]) / (dc * dc))
                    else:
                        gq = k * d_iou
                    g[..., q] = w_reg * np.where(m > 0, gq, 0.0) * m
    if cen_mode == 1:
        sc = 1.0 / (1.0 + np.exp(-p[..., reg_ch]))
        g[..., reg_ch] = w_cen * _sl1_grad64(t[..., reg_ch], sc, delta) * sc * (1 - sc)
    elif cen_mode == 2:
        g[..., reg_ch] = w_cen * _focal_grad64(t[..., reg_ch], p[..., reg_ch], alpha, gamma)
    return g


# --------------------------------------------------------------------------------------
# a11  prediction_to_corners variants
# --------------------------------------------------------------------------------------
def _grid32(h, w, half):
    off = _HALF if half else F(0)
    gx, gy = np.meshgrid(np.arange(w, dtype=np.float32) + off, np.arange(h, dtype=np.float32) + off)
    return gy, gx


def fcos_prediction_to_corners(xy_pred, stride):
    """tblr -> `[y1, x1, y2, x2]` pixels, centres at i+.5 -- FCOS/fcos.py:112-134.
    float32 arithmetic, then `stride *` in float64 like the reference's float64 container."""
    p = _f32(xy_pred)
    gy, gx = _grid32(p.shape[0], p.shape[1], True)
    box = np.stack([gy - p[..., 0], gx - p[..., 2], gy + p[..., 1], gx + p[..., 3]], axis=-1)
    return (stride * box.astype(np.float64)).astype(np.float32)


def retina_prediction_to_corners(xy_pred, anchor_dim, stride):
    """`c = i*s - p*a`, `size = p*a` -- RetinaNet/retinanet_module.py:428-451."""
    p = _f32(xy_pred)
    ah, aw = F(anchor_dim[0]), F(anchor_dim[1])
    gy, gx = _grid32(p.shape[0], p.shape[1], False)
    yc, xc = gy * F(stride) - p[..., 0] * ah, gx * F(stride) - p[..., 1] * aw
    bh, bw = p[..., 2] * ah, p[..., 3] * aw
    return np.stack([yc - bh / F(2), xc - bw / F(2), yc + bh / F(2), xc + bw / F(2)], axis=-1)


def fcos_center_v1_prediction_to_corners(xy_pred, box_sc, stride):
    """FCOS/fcos_center_v1.py:125-147."""
    p = _f32(xy_pred)
    gy, gx = _grid32(p.shape[0], p.shape[1], False)
    yc, xc = (gy + p[..., 0]) * F(stride), (gx + p[..., 1]) * F(stride)
    bh, bw = p[..., 2] * F(box_sc), p[..., 3] * F(box_sc)
    return np.stack([yc - bh / F(2), xc - bw / F(2), yc + bh / F(2), xc + bw / F(2)], axis=-1)


def centernet_s8_prediction_to_corners(xy_pred, box_scales, stride=8):
    """`[H, W, S, 4]` -- CenterNet/tf_centernet_resnet_s8.py:210-241."""
    p = _f32(xy_pred)
    gy, gx = _grid32(p.shape[0], p.shape[1], False)
    out = np.zeros(p.shape[:3] + (4,), dtype=np.float32)
    for n, sc in enumerate(box_scales):
        yc, xc = (gy + p[:, :, n, 0]) * F(stride), (gx + p[:, :, n, 1]) * F(stride)
        bh, bw = p[:, :, n, 2] * F(sc), p[:, :, n, 3] * F(sc)
        out[:, :, n] = np.stack([yc - bh / F(2), xc - bw / F(2), yc + bh / F(2), xc + bw / F(2)], axis=-1)
    return out


# --------------------------------------------------------------------------------------
# a13-a15  NMS
# --------------------------------------------------------------------------------------
def cpu_nms(dets, base_thr):
    """Class-agnostic greedy NMS, keeps `ovr <= thr`, `+1e-8` in the union --
    RetinaNet/retinanet_module.py:453-481.  Returns kept row indices in score order.
    Score ties are ordered by index (stable); the reference's argsort is unstable there."""
    d = np.asarray(dets)
    c0, c1, c2, c3, sc = d[:, 0], d[:, 1], d[:, 2], d[:, 3], d[:, 4]
    areas = (c2 - c0) * (c3 - c1)
    order = np.argsort(-sc, kind="stable")
    keep = []
    while order.size:
        i, rest = order[0], order[1:]
        keep.append(int(i))
        w = np.maximum(0.0, np.minimum(c2[i], c2[rest]) - np.maximum(c0[i], c0[rest]))
        h = np.maximum(0.0, np.minimum(c3[i], c3[rest]) - np.maximum(c1[i], c1[rest]))
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter + 1e-8)
        order = rest[ovr <= base_thr]
    return np.array(keep, dtype=np.int64)


def bboxes_iou(boxes1, boxes2):
    """Corner-box IoU floored at float32 eps -- CenterNet/tf_centernet_resnet_s8.py:22-42."""
    b1, b2 = np.array(boxes1), np.array(boxes2)
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    side = np.maximum(np.minimum(b1[..., 2:], b2[..., 2:]) - np.maximum(b1[..., :2], b2[..., :2]), 0.0)
    inter = side[..., 0] * side[..., 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.maximum(1.0 * inter / (a1 + a2 - inter), np.finfo(np.float32).eps)


def centernet_nms(bboxes, iou_threshold, sigma=0.3, method="nms"):
    """Per-class greedy (soft-)NMS on `(xmin, ymin, w, h, score, class)` rows; returns the kept
    rows as `(x1, y1, x2, y2, score, class)` plus their source row indices.
    CenterNet/tf_centernet_resnet_s8.py:44-85.  Unlike the reference this does not mutate its
    input; classes are visited in ascending order (the reference iterates a `set`)."""
    assert method in ("nms", "soft-nms")
    bb = np.array(bboxes, dtype=np.float64, copy=True)
    if bb.size == 0:
        return np.zeros((0, 6)), np.zeros((0,), dtype=np.int64)
    bb[:, 2] = bb[:, 0] + bb[:, 2]
    bb[:, 3] = bb[:, 1] + bb[:, 3]
    rows, src = [], []
    for c in sorted(set(bb[:, 5].tolist())):
        ids = np.nonzero(bb[:, 5] == c)[0]
        cur = bb[ids].copy()
        while len(cur):
            m = int(np.argmax(cur[:, 4]))
            best = cur[m].copy()
            rows.append(best)
            src.append(int(ids[m]))
            cur = np.delete(cur, m, axis=0)
            ids = np.delete(ids, m)
            if not len(cur):
                break
            iou = bboxes_iou(best[None, :4], cur[:, :4])
            if method == "nms":
                wgt = np.where(iou > iou_threshold, np.float32(0), np.float32(1))
            else:
                wgt = np.exp(-(1.0 * iou ** 2 / sigma))
            cur[:, 4] = cur[:, 4] * wgt
            alive = cur[:, 4] > 0.0
            cur, ids = cur[alive], ids[alive]
    return np.array(rows).reshape(-1, 6), np.array(src, dtype=np.int64)


def combined_nms(boxes, scores, max_output_size_per_class, max_total_size,
                 iou_threshold=0.5, score_threshold=0.05):
    """PARITY UNPINNED restatement of `tf.image.combined_non_max_suppression` as the reference
    calls it (FCOS/infer_fcos.py:58-61: q=1 shared boxes, `clip_boxes=False`,
    `pad_per_class=False`).  boxes `[N, 4]` (y1, x1, y2, x2), scores `[N, C]`.
    Per class: candidates with `score > score_threshold`, score-descending (ties: lower index
    first), greedy suppress `IoU > iou_threshold`, at most `max_output_size_per_class`; then the
    union over classes is sorted by score (ties: lower box index, then lower class -- TensorFlow's own order among
    equal scores is unspecified; target maps sent back through the detector produce nothing but ties) and truncated to
    `max_total_size`.  Returns zero-padded (boxes `[T,4]`, scores `[T]`, classes `[T]`, valid)
    and the flat candidate index (box*C + class) of every kept detection."""
    b, s = _f32(boxes), _f32(scores)
    n, c = s.shape
    y1, x1 = np.minimum(b[:, 0], b[:, 2]), np.minimum(b[:, 1], b[:, 3])
    y2, x2 = np.maximum(b[:, 0], b[:, 2]), np.maximum(b[:, 1], b[:, 3])
    area = (y2 - y1) * (x2 - x1)
    picked = []
    for k in range(c):
        cand = np.nonzero(s[:, k] > F(score_threshold))[0]
        cand = cand[np.argsort(-s[cand, k], kind="stable")]
        kept = []
        for i in cand:
            if len(kept) >= max_output_size_per_class:
                break
            ok = True
            for j in kept:
                ih = max(F(0), min(y2[i], y2[j]) - max(y1[i], y1[j]))
                iw = max(F(0), min(x2[i], x2[j]) - max(x1[i], x1[j]))
                inter = ih * iw
                union = area[i] + area[j] - inter
                if area[i] > 0 and area[j] > 0 and union > 0 and inter / union > F(iou_threshold):
                    ok = False
                    break
            if ok:
                kept.append(int(i))
        picked += [(-float(s[i, k]), i, k) for i in kept]
    picked.sort()
    picked = picked[:max_total_size]
    t = max_total_size
    ob, os_, oc = np.zeros((t, 4), np.float32), np.zeros(t, np.float32), np.zeros(t, np.float32)
    flat = np.full(t, -1, dtype=np.int64)
    for r, (negs, i, k) in enumerate(picked):
        ob[r], os_[r], oc[r], flat[r] = b[i], -negs, k, i * c + k
    return ob, os_, oc, len(picked), flat


# --------------------------------------------------------------------------------------
# a12 / a14  image_detections (head outputs -> detections; model forward removed)
# --------------------------------------------------------------------------------------
def retina_decode_dets(head_outputs, anchor_dims=None, strides=None):
    """The decode front end of RetinaNet/retinanet_module.py:487-520: `head_outputs[level][anchor]` `[Hl, Wl, C+4]` ->
    `[N, 6]` (y1, x1, y2, x2, max score, first-argmax label), order level > anchor > row-major cell, before any threshold."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    dims = retina_anchor_dims() if anchor_dims is None else _f32(anchor_dims)
    flat = []
    for n, level in enumerate(head_outputs):
        for a, m in enumerate(level):
            m = _f32(m)
            box = retina_prediction_to_corners(m[..., :4], dims[n, a], strides[n])
            flat.append(np.concatenate([box, _sigmoid32(m[..., 4:])], axis=-1).reshape(-1, m.shape[-1]))
    flat = np.concatenate(flat, axis=0)
    return np.concatenate([flat[:, :4], flat[:, 4:].max(axis=1)[:, None], flat[:, 4:].argmax(axis=1).astype(np.float32)[:, None]],
                          axis=1).astype(np.float32)


def retina_image_detections(head_outputs, anchor_dims=None, strides=None, iou_thresh=0.5, cls_thresh=0.05):
    """`head_outputs[level][anchor]` `[Hl, Wl, C+4]` -> `[k, 6]` (y1, x1, y2, x2, score, label) in
    kept order -- RetinaNet/retinanet_module.py:483-530."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    dims = retina_anchor_dims() if anchor_dims is None else _f32(anchor_dims)
    flat = []
    for n, level in enumerate(head_outputs):
        for a, m in enumerate(level):
            m = _f32(m)
            box = retina_prediction_to_corners(m[..., :4], dims[n, a], strides[n])
            flat.append(np.concatenate([box, _sigmoid32(m[..., 4:])], axis=-1).reshape(-1, m.shape[-1]))
    flat = np.concatenate(flat, axis=0)
    score = flat[:, 4:].max(axis=1)
    label = flat[:, 4:].argmax(axis=1).astype(np.float32)
    dets = np.concatenate([flat[:, :4], score[:, None], label[:, None]], axis=1).astype(np.float32)
    sel = np.nonzero(dets[:, 4] >= F(cls_thresh))[0]
    dets = dets[sel]
    if len(dets) == 0:
        return dets, sel
    keep = cpu_nms(dets, iou_thresh)
    return dets[keep], sel[keep]


def fcos_decode_scores(head_outputs, num_classes, strides=None, center=False):
    """The (boxes `[N,4]`, scores `[N,C]`) pair FCOS/infer_fcos.py:35-57 hands to combined NMS."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    flat = []
    for n, m in enumerate(head_outputs):
        m = _f32(m).copy()
        m[..., :4] = fcos_prediction_to_corners(m[..., :4], strides[n])
        flat.append(m.reshape(-1, num_classes + 5))
    flat = np.concatenate(flat, axis=0)
    sc = _sigmoid32(flat[:, 5:])
    if center:
        sc = _sigmoid32(flat[:, 4])[:, None] * sc
    return flat[:, :4], sc


def fcos_image_detections(head_outputs, num_classes, center=False, iou_thresh=0.5, cls_thresh=0.05,
                          max_detections=100, max_total_size=100, strides=None):
    """FCOS/infer_fcos.py:27-62 with the third-party NMS replaced by `combined_nms`."""
    boxes, scores = fcos_decode_scores(head_outputs, num_classes, strides, center)
    return combined_nms(boxes, scores, max_detections, max_total_size, iou_thresh, cls_thresh)


def fcos_ground_truth_detections(img_labels, num_classes, image_shape, img_rows=384, img_cols=384, center=True,
                                 strides=None):
    """The computation inside `show_heatmap`, FCOS/train_fcos_center_voc.py:13-121 (heat-map rendering left out): the
    target maps of ONE image, `img_labels[level]` `[Hl, Wl, C+5]`, decoded like predictions -- boxes
    `prediction_to_corners(map[..., :4], stride)` (:54-55), scores `sqrt(class * centerness)` in the maps' float64
    (:58-63) or the class channel (:64-66), combined NMS with 100 / 100 caps at IoU 0.75 / score 0.75 (:86-88, float32 from
    there on), boxes times `[w_ratio, h_ratio, w_ratio, h_ratio]` (:91-94, a float32 TF multiply), `swap_xy`, then
    per box x1 / y1 <= 0 -> 0, `w = x2 - x1`, `h = y2 - y1` (:103-116).  Returns (rectangles `[k, 4]` (x1, y1, w, h),
    scores `[k]`), float32.  PARITY UNPINNED where `combined_nms` is."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    w_ratio, h_ratio = image_shape[0] / img_rows, image_shape[1] / img_cols
    bl, sl = [], []
    for n, m in enumerate(img_labels):
        m = np.asarray(m, dtype=np.float64)
        bl.append(fcos_prediction_to_corners(_f32(m[..., :4]), strides[n]).reshape(-1, 4))
        flat = m.reshape(-1, num_classes + 5)
        sl.append(np.sqrt(flat[:, 5:] * flat[:, 4:5]) if center else flat[:, 5:])
    ob, os_, _, valid, _ = combined_nms(np.concatenate(bl), np.concatenate(sl), 100, 100, 0.75, 0.75)
    det = _f32(ob[:valid]) * np.array([w_ratio, h_ratio, w_ratio, h_ratio], dtype=np.float32)
    x1, y1, x2, y2 = det[:, 1].copy(), det[:, 0].copy(), det[:, 3], det[:, 2]
    x1[x1 <= 0] = 0
    y1[y1 <= 0] = 0
    return np.stack([x1, y1, x2 - x1, y2 - y1], axis=1).astype(np.float32).reshape(-1, 4), _f32(os_[:valid])


# --------------------------------------------------------------------------------------
# f-4  the offline COCO -> sparse FCOS target formatter (/format_COCO_annotations_fcos.py:66-183)
# --------------------------------------------------------------------------------------
def fcos_sparse_scales(img_dims, num_scale=5):
    """`tmp_scale` of the script (:49-56): `int(min(img_dims) / 2**x)`, ascending."""
    return [int(min(img_dims) / (2 ** x)) for x in range(num_scale)][::-1]


def fcos_sparse_format(objects, src_dims, img_dims=(448, 448), num_scale=5):
    """One image of /format_COCO_annotations_fcos.py:66-183.  `objects` `[n, 5]` float64 rows (x_lower, y_lower,
    box_width, box_height, label) in SOURCE-image pixels (`label` is what the script looks up through its label tables,
    1-based: index 0 is "objectness"), `src_dims` = the source image's (img_width, img_height).
    Per object: the box is brought to the `img_dims` canvas by float64 division with the ratios (:86-87, :98-99), the scale
    slot is the first one both sides are strictly below (:101-123), the footprint is the integer rectangle
    `[int(x_lower / w_ratio), int(x_low + width)) x [int(y_lower / h_ratio), int(y_low + height))` (:126-131) cut by
    NumPy's slice rules against an `[img_width, img_height]` array (:147-149), walked x-major (`np.nonzero`, :151).
    `tmp_mask` is never written (:91, :151), so overlapping objects all emit.  Each cell emits, in this order, the six
    entries `[y, x, scale, k]` = (b, t, l, r, centerness, 1) for k = 0..5 with `b = y - y_low`, `t = y_upp - y`,
    `l = x - x_low`, `r = x_upp - x`, `centerness = sqrt(min(l,r)/max(l,r)) * sqrt(min(b,t)/max(b,t))` in float64 (:8-11,
    :158-168), then the class entry, value 1, at `[y, x, scale, label + 4]` -- the script writes that index five long,
    `[y, x, scale, scale, label + 4]` (:171), which `tf.sparse.SparseTensor` cannot take; the duplicate is dropped here
    (documented deviation).  Returns (indices `[nnz, 4]` int32, values `[nnz]` float32 -- the script's mixed int / float64
    list rounded once) and the `dense_shape` the script states (:76-78: `[img_h / 8, img_w / 8, num_scale, n_classes + 4]`
    is the caller's business: it needs the label table)."""
    img_w, img_h = int(img_dims[0]), int(img_dims[1])
    scales = fcos_sparse_scales(img_dims, num_scale)
    w_ratio, h_ratio = np.float64(src_dims[0]) / img_w, np.float64(src_dims[1]) / img_h
    idx, val = [np.zeros((0, 4), np.int32)], [np.zeros(0, np.float32)]
    for o in np.asarray(objects, dtype=np.float64).reshape(-1, 5):
        width, height = o[2] / w_ratio, o[3] / h_ratio
        if width < 0 or height < 0:
            continue
        sc = num_scale - 1
        for k in range(num_scale - 1):
            if width < scales[k] and height < scales[k]:
                sc = k
                break
        x_low = int(o[0] / w_ratio)
        x_upp = int(x_low + width)
        y_low = int(o[1] / h_ratio)
        y_upp = int(y_low + height)
        xs = np.arange(img_w)[x_low:x_upp]
        ys = np.arange(img_h)[y_low:y_upp]
        if len(xs) == 0 or len(ys) == 0:
            continue
        x, y = [a.reshape(-1) for a in np.meshgrid(xs, ys, indexing="ij")]  # x-major
        l, r, b, t = x - x_low, x_upp - x, y - y_low, y_upp - y
        with np.errstate(invalid="ignore", divide="ignore"):
            c = np.sqrt(np.minimum(l, r) / np.maximum(l, r)) * np.sqrt(np.minimum(b, t) / np.maximum(b, t))
        n = len(x)
        entry = np.zeros((n, 7, 4), np.int64)
        entry[:, :, 0], entry[:, :, 1], entry[:, :, 2] = y[:, None], x[:, None], sc
        entry[:, :6, 3] = np.arange(6)
        entry[:, 6, 3] = int(o[4]) + 4
        v = np.stack([b, t, l, r, c, np.ones(n), np.ones(n)], axis=1)
        idx.append(entry.reshape(-1, 4).astype(np.int32))
        val.append(v.reshape(-1).astype(np.float32))
    return np.concatenate(idx), np.concatenate(val)


# --------------------------------------------------------------------------------------
# pre-NMS top-k per pyramid level (an extension: the reference has none)
# --------------------------------------------------------------------------------------
def select_topk(scores, seg_offsets, k, min_score, inclusive=True):
    """Indices (ascending, per segment, concatenated) of the rows whose score passes `min_score` and is among
    the k highest of its segment; ties go to the lower index.  With k >= segment length this is the
    reference's plain threshold."""
    s = np.asarray(scores, dtype=np.float32)
    out = []
    for a, b in zip(seg_offsets[:-1], seg_offsets[1:]):
        seg = s[a:b]
        ok = np.nonzero(seg >= F(min_score) if inclusive else seg > F(min_score))[0]
        if len(ok) > k:
            order = ok[np.argsort(-seg[ok], kind="stable")][:k]
            ok = np.sort(order)
        out.append(ok + a)
    return out


def retina_detect_from_dets(dets, seg_offsets, iou_thresh=0.5, cls_thresh=0.05, pre_nms_topk=None):
    """`dets` [N, 6] as produced by the decode front end -> (candidate rows, kept candidate indices)."""
    d = _f32(dets)
    k = pre_nms_topk if pre_nms_topk is not None else len(d)
    sel = np.concatenate(select_topk(d[:, 4], seg_offsets, k, cls_thresh, True)) if len(d) else np.zeros(0, np.int64)
    cand = d[sel]
    keep = cpu_nms(cand, iou_thresh) if len(cand) else np.zeros(0, np.int64)
    return cand, keep, sel


def fcos_detect_from_scores(boxes, scores, seg_offsets, iou_thresh=0.5, cls_thresh=0.05, max_detections=100,
                            max_total_size=100, pre_nms_topk=None):
    """boxes [N, 4], scores [N, C] (decode front end); per-level top-k over the (location, class) scores,
    then `combined_nms`.  seg_offsets are in units of flattened (location*C + class) entries."""
    b, s = _f32(boxes), _f32(scores)
    n, c = s.shape
    flat = s.reshape(-1)
    if pre_nms_topk is not None:
        sel = np.concatenate(select_topk(flat, seg_offsets, pre_nms_topk, cls_thresh, False))
        masked = np.zeros_like(flat)
        masked[sel] = flat[sel]
        s = masked.reshape(n, c)
    return combined_nms(b, s, max_detections, max_total_size, iou_thresh, cls_thresh)
