"""FCOS dense-head routines: drop-in for FCOS/fcos.py, FCOS/fcos_center.py and
FCOS/fcos_center_v1.py of the reference (paths relative to the reference repository).

`format_data*` keep the reference signatures and return device tensors (torch CUDA tensors, which
TensorFlow ingests zero-copy through DLPack) instead of host float64 arrays; values equal
`reference_map.astype(float32)`.
"""
import ctypes

import numpy as np
import torch

from . import _capi, infer, losses
from ._batch import image_dims, pack_labels, check_classes
from ._tensors import as_host, current_device, stream_ptr, to_device, uses_stream

DEFAULT_STRIDES = [8, 16, 32, 64, 128]
DEFAULT_B_DIM = [32, 64, 128, 256]
MODES = {"fcos": 0, "center": 1, "center_only": 2, "center_v1": 3, "min_area": 4}  # "min_area": extension (DH_FCOS_MIN_AREA)


def level_shapes(img_pad, strides):
    return [(int(img_pad[0] / s), int(img_pad[1] / s)) for s in strides]


@uses_stream
def format_data_batch(boxes, nbox, img_dim, num_classes, img_pad, strides=None, b_dim=None, mode="fcos",
                      out=None, num_targets=None, stream=None):
    """Encode a padded batch on the device.

    boxes [B, Nmax, 5] (cy, cx, h, w, class; normalised), nbox [B], img_dim [B, 2] or [2] (unpadded
    H, W), img_pad (H, W) ints.  Returns (list of n_levels tensors [B, Hl, Wl, C+5], num_targets
    int32 [B, n_levels]) -- both on the device; nothing is synchronised.
    """
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    b_dim = list(DEFAULT_B_DIM if b_dim is None else b_dim)
    if len(b_dim) != len(strides) - 1:
        raise ValueError("b_dim must have len(strides)-1 entries")
    dev = current_device()
    check_classes(boxes, nbox, num_classes)
    boxes_d = to_device(boxes, torch.float32, dev)
    if boxes_d.dim() != 3 or boxes_d.shape[2] != 5:
        raise ValueError("boxes must be [B, Nmax, 5]")
    batch, nmax = int(boxes_d.shape[0]), int(boxes_d.shape[1])
    nbox_d = to_device(nbox, torch.int32, dev)
    dims_d = to_device(image_dims(img_dim, batch) if not isinstance(img_dim, torch.Tensor) or not img_dim.is_cuda
                       else img_dim, torch.float32, dev)
    pad_h, pad_w = int(img_pad[0]), int(img_pad[1])
    shapes = level_shapes((pad_h, pad_w), strides)
    ch = num_classes + 5
    if out is None:
        out = [torch.empty((batch, h, w, ch), dtype=torch.float32, device=dev) for h, w in shapes]
    else:
        for o, (h, w) in zip(out, shapes):
            if tuple(o.shape) != (batch, h, w, ch) or o.dtype != torch.float32 or not o.is_contiguous():
                raise ValueError("out tensors must be contiguous float32 [B, Hl, Wl, C+5]")
    if num_targets is None:
        num_targets = torch.empty((batch, len(strides)), dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().dh_fcos_encode(
        _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax, pad_h, pad_w,
        len(strides), _capi.int_array(strides), _capi.float_array(b_dim), int(num_classes), MODES[mode],
        _capi.ptr_array([o.data_ptr() for o in out]), num_targets.data_ptr(), stream_ptr(stream)), "dh_fcos_encode")
    return out, num_targets


def _single(gt_labels, img_dim, num_classes, img_pad, strides, b_dim, mode):
    g = as_host(gt_labels, np.float32).reshape(-1, 5)
    dim = as_host(img_dim, np.float32).reshape(2)
    pad = [int(v) for v in (as_host(img_pad, np.float64).reshape(2) if img_pad is not None else dim)]
    boxes, nbox = pack_labels([g])
    outs, cnt = format_data_batch(boxes, nbox, dim[None], num_classes, pad, strides, b_dim, mode)
    return [o[0] for o in outs], [int(v) for v in cnt[0].tolist()]


def format_data(gt_labels, img_dim, num_classes, img_pad=None, areas=None, strides=None):
    """FCOS/fcos.py:136 `format_data` (footprint assignment, largest-area-wins, centerness).

    `areas` is accepted for signature compatibility; the reference raises NameError when it is
    passed (fcos.py:145-147 vs :171) and so do we."""
    if areas is not None:
        raise NameError("name 'b_dim' is not defined")
    return _single(gt_labels, img_dim, num_classes, img_pad, strides, None, "fcos")


def format_data_center(gt_labels, img_dim, num_classes, img_pad=None, b_dim=None, strides=None, center_only=False):
    """FCOS/fcos_center.py:149 `format_data` (3x3 / centre-only assignment, scores 1/.5/.25)."""
    return _single(gt_labels, img_dim, num_classes, img_pad, strides, b_dim, "center_only" if center_only else "center")


def format_data_center_v1(gt_labels, img_dim, num_classes, img_pad=None, b_dim=None, strides=None, center_only=False):
    """FCOS/fcos_center_v1.py:149 `format_data` (centre cell, YOLO-style offsets; `center_only` unused)."""
    return _single(gt_labels, img_dim, num_classes, img_pad, strides, b_dim, "center_v1")


# ---- losses -------------------------------------------------------------------------------------
def _as_batched(maps, dev):
    """Accept per-level [Hl, Wl, ch], [1, Hl, Wl, ch] (the reference's `y_pred[l]`) or [B, Hl, Wl, ch]."""
    out = []
    for m in maps:
        t = to_device(m, torch.float32, dev)
        if t.dim() == 3:
            t = t.unsqueeze(0)
        if t.dim() != 4:
            raise ValueError("expected [Hl, Wl, ch] or [B, Hl, Wl, ch] maps")
        out.append(t.contiguous())
    return out


@uses_stream
def model_loss_batch(y_true, y_pred, reg_type="l1", cen_type="l1", alpha=0.25, gamma=2.0, delta=1.0, stream=None, weights=None):
    """Loss over materialised FCOS targets for a batch -> (per_image [B,4], total [4]) = {cls, reg, cen, n_pos}
    (+ per-level gradients when `weights` = (w_cls, w_reg, w_cen) is given)."""
    dev = current_device()
    yt, yp = _as_batched(y_true, dev), _as_batched(y_pred, dev)
    batch, ch = int(yp[0].shape[0]), int(yp[0].shape[-1])
    shapes = [(int(p.shape[1]), int(p.shape[2]), 1) for p in yp]
    cen = {"l1": losses.CEN_SMOOTH_L1, "focal": losses.CEN_FOCAL}.get(cen_type.lower(), losses.CEN_IGNORE)
    reg = losses.reg_mode(reg_type)
    return losses.dense_loss(yt, yp, shapes, batch, ch, 4, cen, reg, losses.POS_GE1, alpha, gamma, delta, stream=stream, weights=weights)


def model_loss(y_true, y_pred, strides=None, reg_type="l1", cen_type="l1", cls_lambda=2.5, reg_lambda=1.0):
    """FCOS/fcos.py:464 `model_loss` -> (cls_loss, reg_loss, cen_loss) device scalars.  `strides`,
    `cls_lambda`, `reg_lambda` are accepted and unused, as in the reference.  With cen_type other than
    "l1" the centerness loss is 0 (fcos.py:483-486); use `model_loss_center` for the focal variant."""
    _, tot = model_loss_batch(y_true, y_pred, reg_type, cen_type if cen_type.lower() == "l1" else "none")
    return tot[0], tot[1], tot[2]


def model_loss_center(y_true, y_pred, reg_type="l1", cen_type="l1"):
    """FCOS/fcos_center.py:365 `model_loss` (centerness smooth-L1 or focal)."""
    _, tot = model_loss_batch(y_true, y_pred, reg_type, "l1" if cen_type.lower() == "l1" else "focal")
    return tot[0], tot[1], tot[2]


def model_loss_center_v1(y_true, y_pred):
    """FCOS/fcos_center_v1.py:294 `model_loss` (focal centerness, smooth-L1 boxes)."""
    _, tot = model_loss_batch(y_true, y_pred, "l1", "focal")
    return tot[0], tot[1], tot[2]


@uses_stream
def encode_loss_batch(boxes, nbox, img_dim, num_classes, img_pad, y_pred, strides=None, b_dim=None, mode="fcos",
                      reg_type="l1", cen_type="l1", alpha=0.25, gamma=2.0, delta=1.0, stream=None, weights=None):
    """Fused target encoding + loss: targets never reach HBM.  y_pred: per-level [B, Hl, Wl, C+5].
    Returns (per_image [B,4], total [4], num_targets [B, n_levels]); with `weights` = (w_cls, w_reg, w_cen) a fourth
    item, the per-level gradients d(w . {cls, reg, cen}) / d y_pred (dh_fcos_encode_loss_grad, same pass)."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    b_dim = list(DEFAULT_B_DIM if b_dim is None else b_dim)
    dev = current_device()
    check_classes(boxes, nbox, num_classes)
    boxes_d = to_device(boxes, torch.float32, dev)
    batch, nmax = int(boxes_d.shape[0]), int(boxes_d.shape[1])
    nbox_d = to_device(nbox, torch.int32, dev)
    dims_d = to_device(image_dims(img_dim, batch) if not isinstance(img_dim, torch.Tensor) or not img_dim.is_cuda
                       else img_dim, torch.float32, dev)
    yp = _as_batched(y_pred, dev)
    shapes = level_shapes((int(img_pad[0]), int(img_pad[1])), strides)
    for p, (h, w) in zip(yp, shapes):
        if tuple(p.shape) != (batch, h, w, num_classes + 5):
            raise ValueError("prediction level has shape %r, expected %r" % (tuple(p.shape), (batch, h, w, num_classes + 5)))
    out_pi = torch.empty((batch, 4), dtype=torch.float32, device=dev)
    out_tot = torch.empty((4,), dtype=torch.float32, device=dev)
    cnt = torch.empty((batch, len(strides)), dtype=torch.int32, device=dev)
    cen = {"l1": losses.CEN_SMOOTH_L1, "focal": losses.CEN_FOCAL}.get(cen_type.lower(), losses.CEN_IGNORE)
    reg = losses.reg_mode(reg_type)
    if weights is not None:
        grads = [torch.empty_like(p) for p in yp]
        _capi.check(_capi.lib().dh_fcos_encode_loss_grad(
            _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax,
            int(img_pad[0]), int(img_pad[1]), len(strides), _capi.int_array(strides), _capi.float_array(b_dim),
            int(num_classes), MODES[mode], _capi.ptr_array([p.data_ptr() for p in yp]), reg, cen, float(alpha),
            float(gamma), float(delta), float(weights[0]), float(weights[1]), float(weights[2]),
            _capi.ptr_array([g.data_ptr() for g in grads]), out_pi.data_ptr(), out_tot.data_ptr(), cnt.data_ptr(),
            stream_ptr(stream)), "dh_fcos_encode_loss_grad")
        return out_pi, out_tot, cnt, grads
    _capi.check(_capi.lib().dh_fcos_encode_loss(
        _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax,
        int(img_pad[0]), int(img_pad[1]), len(strides), _capi.int_array(strides), _capi.float_array(b_dim),
        int(num_classes), MODES[mode], _capi.ptr_array([p.data_ptr() for p in yp]), reg, cen, float(alpha),
        float(gamma), float(delta), out_pi.data_ptr(), out_tot.data_ptr(), cnt.data_ptr(), stream_ptr(stream)),
        "dh_fcos_encode_loss")
    return out_pi, out_tot, cnt


focal_loss = losses.focal_loss
smooth_l1_loss = losses.smooth_l1_loss
iou_loss = losses.iou_loss


# ---- inference ------------------------------------------------------------------------------------
def prediction_to_corners(xy_pred, stride):
    """FCOS/fcos.py:112 -- tblr map [H, W, >=4] -> pixel corners [H, W, 4] (y1, x1, y2, x2)."""
    return infer.prediction_to_corners(xy_pred, 0, stride)


def prediction_to_corners_center_v1(xy_pred, box_sc, stride):
    """FCOS/fcos_center_v1.py:125."""
    return infer.prediction_to_corners(xy_pred, 2, stride, d0=box_sc)


@uses_stream
def sparse_format_batch(objects, nbox, src_dims, img_dims=(448, 448), num_scale=5, stream=None):
    """/format_COCO_annotations_fcos.py:66-183 (the offline COCO -> sparse FCOS targets script) for a batch of images.

    `objects` `[B, n, 5]` float64 rows (x_lower, y_lower, box_width, box_height, label) in source-image pixels -- `label` is
    the script's 1-based class index -- `nbox` `[B]` (or None: every row counts), `src_dims` `[B, 2]` the source images'
    (img_width, img_height).  Returns (indices `[nnz, 4]` int32 = (y, x, scale, channel), values `[nnz]` float32,
    offsets `[B+1]` int64: image b owns entries offsets[b]:offsets[b+1]) on the device, entries in the script's order;
    one host synchronisation (the entry count sizes the outputs)."""
    dev = current_device()
    obj = to_device(objects, torch.float64, dev).contiguous()
    if obj.dim() != 3 or obj.shape[-1] != 5:
        raise ValueError("objects must be [B, n, 5]")
    batch, n = int(obj.shape[0]), int(obj.shape[1])
    nb = to_device(nbox, torch.int32, dev) if nbox is not None else None
    sd = to_device(np.asarray(src_dims, np.float64).reshape(batch, 2) if not isinstance(src_dims, torch.Tensor) else src_dims,
                   torch.float64, dev).contiguous()
    offsets = torch.empty((batch + 1,), dtype=torch.int64, device=dev)

    def call(capacity, idx, val):
        _capi.check(_capi.lib().dh_fcos_sparse_encode(
            _capi.handle(dev.index), obj.data_ptr(), nb.data_ptr() if nb is not None else None, sd.data_ptr(), batch, max(n, 1),
            int(img_dims[0]), int(img_dims[1]), int(num_scale), capacity, idx.data_ptr() if idx is not None else None,
            val.data_ptr() if val is not None else None, offsets.data_ptr(), stream_ptr(stream)), "dh_fcos_sparse_encode")
    if batch == 0 or n == 0:
        return (torch.empty((0, 4), dtype=torch.int32, device=dev), torch.empty((0,), dtype=torch.float32, device=dev),
                torch.zeros((batch + 1,), dtype=torch.int64, device=dev))
    call(0, None, None)
    nnz = int(offsets[-1].item())
    indices = torch.empty((nnz, 4), dtype=torch.int32, device=dev)
    values = torch.empty((nnz,), dtype=torch.float32, device=dev)
    if nnz:
        call(nnz, indices, values)
    return indices, values, offsets


SCORE_MODES = {"cls": 0, "cls_cen": 1, "map": 2, "map_cen": 3}  # DH_FCOS_SCORE_*


def score_mode(center):
    """The reference's `center` flag (False / True) or one of SCORE_MODES ("map", "map_cen": the input is a target map)."""
    if isinstance(center, str):
        if center not in SCORE_MODES:
            raise ValueError("score mode must be one of %s" % sorted(SCORE_MODES))
        return SCORE_MODES[center]
    return 1 if center else 0


@uses_stream
def decode_batch(head_outputs, num_classes, img_pad, strides=None, center=False, stream=None):
    """FCOS/infer_fcos.py:35-57 for a batch: per-level heads [B, Hl, Wl, C+5] -> (boxes [B, N, 4], scores [B, N, C])."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    dev = current_device()
    heads = _as_batched(head_outputs, dev)
    batch = int(heads[0].shape[0])
    n = sum(h * w for h, w in level_shapes(img_pad, strides))
    boxes = torch.empty((batch, n, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((batch, n, num_classes), dtype=torch.float32, device=dev)
    _capi.check(_capi.lib().dh_fcos_decode(
        _capi.handle(dev.index), _capi.ptr_array([h.data_ptr() for h in heads]), batch, int(img_pad[0]), int(img_pad[1]),
        len(strides), _capi.int_array(strides), int(num_classes), score_mode(center), boxes.data_ptr(), scores.data_ptr(),
        stream_ptr(stream)), "dh_fcos_decode")
    return boxes, scores


@uses_stream
def detect_batch(head_outputs, num_classes, img_pad, center=False, iou_thresh=0.5, cls_thresh=0.05, max_detections=100,
                 max_total_size=100, pre_nms_topk=1000, strides=None, with_candidates=False, stream=None):
    """FCOS/infer_fcos.py:27-62 for a batch, one library call (dh_fcos_detect): decode -> per-level top-k of the
    (location, class) scores above cls_thresh -> per-class greedy NMS with the combined-NMS caps.  Returns zero-padded
    (boxes [B, T, 4], scores [B, T], classes [B, T], valid [B]) with T = max_total_size, like
    tf.image.combined_non_max_suppression.

    DEVIATION from the reference, which hands EVERY (location, class) score above `cls_thresh` to the NMS: at most
    `pre_nms_topk` pairs per level go in (the best ones; exact, ties by index), and levels * pre_nms_topk <= 16384.  The
    result equals the reference's whenever no level has more passing scores than that; `pre_nms_topk=None` takes the
    largest value the NMS holds (16384 // levels = 3276 per level for five levels), the closest to the reference."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    dev = current_device()
    heads = _as_batched(head_outputs, dev)
    batch, t = int(heads[0].shape[0]), int(max_total_size)
    longest = max([h * w * int(num_classes) for h, w in level_shapes(img_pad, strides)] + [1])
    if pre_nms_topk is None:
        pre_nms_topk = 16384 // len(strides)
    k = int(min(pre_nms_topk, longest))
    boxes = torch.empty((batch, t, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((batch, t), dtype=torch.float32, device=dev)
    classes = torch.empty((batch, t), dtype=torch.float32, device=dev)
    valid = torch.zeros((batch,), dtype=torch.int32, device=dev)
    cand = torch.empty((batch, len(strides) * k, 6), dtype=torch.float32, device=dev) if with_candidates else None
    _capi.check(_capi.lib().dh_fcos_detect(
        _capi.handle(dev.index), _capi.ptr_array([h.data_ptr() for h in heads]), batch, int(img_pad[0]), int(img_pad[1]),
        len(strides), _capi.int_array(strides), int(num_classes), score_mode(center), float(iou_thresh), float(cls_thresh),
        int(max_detections), t, int(pre_nms_topk), boxes.data_ptr(), scores.data_ptr(), classes.data_ptr(), valid.data_ptr(),
        cand.data_ptr() if cand is not None else None, stream_ptr(stream)), "dh_fcos_detect")
    return (boxes, scores, classes, valid, cand) if with_candidates else (boxes, scores, classes, valid)


@uses_stream
def ground_truth_detections(img_labels, num_classes, image_shapes, img_rows=384, img_cols=384, center=True, iou_thresh=0.75,
                            cls_thresh=0.75, strides=None, stream=None):
    """The computation inside `show_heatmap` (FCOS/train_fcos_center_voc.py:13-121, the copies in the other FCOS training
    scripts) for a batch: target maps `img_labels` (per level [B, Hl, Wl, C+5], what `format_data` returns) are sent back
    through the detector -- boxes from channels 0..3, score sqrt(class * centerness) (`center=True`) or the class channel,
    combined NMS at 0.75 / 0.75 with 100 / 100 caps -- and come out as the rectangles the reference draws: `(x1, y1, w, h)`
    in source-image pixels `[B, 100, 4]`, scores `[B, 100]`, valid `[B]`.  `image_shapes` `[B, 2]` are the source images'
    `.shape[:2]`.  Up to 16384 / n_levels entries per level pass the score threshold into the NMS (a target map has a few
    per box).  Heat-map rendering (tf.image.resize + matplotlib) stays with the caller."""
    strides = list(DEFAULT_STRIDES if strides is None else strides)
    dev = current_device()
    maps = _as_batched(img_labels, dev)
    batch = int(maps[0].shape[0])
    boxes, scores, _, valid = detect_batch(maps, num_classes, (img_rows, img_cols), "map_cen" if center else "map", iou_thresh,
                                           cls_thresh, 100, 100, pre_nms_topk=16384 // len(strides), strides=strides)
    shp = np.asarray(image_shapes, np.float64).reshape(batch, 2)
    ratios = to_device(np.stack([shp[:, 0] / img_rows, shp[:, 1] / img_cols], axis=1).astype(np.float32), torch.float32, dev)
    rect = torch.empty_like(boxes)
    _capi.check(_capi.lib().dh_fcos_rectangles(_capi.handle(dev.index), boxes.data_ptr(), valid.data_ptr(), ratios.data_ptr(), batch,
                                               int(boxes.shape[1]), rect.data_ptr(), stream_ptr(stream)), "dh_fcos_rectangles")
    return rect, scores, valid


def image_detections(image, model, num_classes, center=False, iou_thresh=0.5, cls_thresh=0.05, max_detections=100,
                     max_total_size=100, head_outputs=None, pre_nms_topk=None):
    """FCOS/infer_fcos.py:27 `image_detections`.  `model(image, training=False)` must return the per-level head
    outputs [1, Hl, Wl, C+5] (or pass them as `head_outputs`).  `pre_nms_topk=None` (default): as many candidates per
    level as the NMS holds -- see `detect_batch` for the one deviation from the reference (a cap it does not have)."""
    heads = head_outputs if head_outputs is not None else model(image, training=False)
    dev = current_device()
    hb = _as_batched(heads, dev)
    pad = (int(hb[0].shape[1]) * DEFAULT_STRIDES[0], int(hb[0].shape[2]) * DEFAULT_STRIDES[0])
    return detect_batch(hb, num_classes, pad, center, iou_thresh, cls_thresh, max_detections, max_total_size, pre_nms_topk)
