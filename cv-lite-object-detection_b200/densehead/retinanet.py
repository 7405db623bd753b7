"""RetinaNet dense-head routines: drop-in for the non-Keras methods of
RetinaNet/retinanet_module.py (`RetinaNet` class) and `compute_iou` of RetinaNet/utils.py.

`RetinaNetHead` mirrors the reference class minus the backbone: same constructor arguments for the
anchor configuration, same method names (`get_anchors`, `format_data`, `prediction_to_corners`,
`cpu_nms`, `image_detections`, `train_loss` split into forward + `loss`).
"""
import math

import numpy as np
import torch

from . import _capi, infer, losses
from ._batch import image_dims, pack_labels, check_classes
from ._tensors import as_host, current_device, stream_ptr, to_device, uses_stream

STRIDES = [8, 16, 32, 64, 128]


def anchor_table(anchor_sizes=None, aspect_ratios=None, anchor_scales=None):
    """float32 [5, A, 2] (h, w): RetinaNet/retinanet_module.py:201-219 in the reference's float32
    operation order (`tf.math.sqrt(area / ratio)` on a float32 tensor, `area / h`, `scale * dims`)."""
    sizes = [32.0, 64.0, 128.0, 256.0, 512.0] if anchor_sizes is None else list(anchor_sizes)
    if len(sizes) != 5:
        raise ValueError("anchor_sizes must be of dimension 5.")
    ratios = [0.5, 1.0, 2.0] if aspect_ratios is None else list(aspect_ratios)
    scales = [2 ** x for x in [0, 1 / 3, 2 / 3]] if anchor_scales is None else list(anchor_scales)
    if len(scales) != 3:
        raise ValueError("anchor_scales must be of dimension 3.")
    f = np.float32
    table = np.zeros((5, len(ratios) * len(scales), 2), dtype=np.float32)
    for l, area in enumerate(sorted(x ** 2 for x in sizes)):
        a = 0
        for ratio in ratios:
            h = np.sqrt(f(area / ratio))
            w = f(area) / h
            for sc in scales:
                table[l, a] = (f(sc) * h, f(sc) * w)
                a += 1
    return table


@uses_stream
def format_data_batch(boxes, nbox, img_dim, num_classes, img_pad, anchor_hw=None, iou_thresh=0.5, strides=None,
                      out=None, num_pairs=None, stream=None):
    """Match + encode a padded batch.  Returns (list of 5 tensors [B, A, Hl, Wl, C+4], num_pairs int32 [B])."""
    strides = list(STRIDES if strides is None else strides)
    table = anchor_table() if anchor_hw is None else np.ascontiguousarray(anchor_hw, dtype=np.float32)
    n_levels, n_anchors = table.shape[0], table.shape[1]
    if n_levels != len(strides):
        raise ValueError("anchor table has %d levels for %d strides" % (n_levels, len(strides)))
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
    ch = num_classes + 4
    shapes = [(int(pad_h / s), int(pad_w / s)) for s in strides]
    if out is None:
        out = [torch.empty((batch, n_anchors, h, w, ch), dtype=torch.float32, device=dev) for h, w in shapes]
    if num_pairs is None:
        num_pairs = torch.empty((batch,), dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().dh_retina_encode(
        _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax, pad_h, pad_w,
        n_levels, _capi.int_array(strides), n_anchors, _capi.float_array(table.reshape(-1).tolist()),
        float(iou_thresh), int(num_classes), _capi.ptr_array([o.data_ptr() for o in out]), num_pairs.data_ptr(),
        stream_ptr(stream)), "dh_retina_encode")
    return out, num_pairs


def _pack_levels(x, n_anchors, dev):
    """`x[level][anchor]` maps ([Hl, Wl, ch] or [1, Hl, Wl, ch]) or packed [B, A, Hl, Wl, ch] per level."""
    out = []
    for lv in x:
        if isinstance(lv, (list, tuple)):
            maps = [to_device(m, torch.float32, dev) for m in lv]
            maps = [m[0] if m.dim() == 4 else m for m in maps]
            lv = torch.stack(maps).unsqueeze(0)
        else:
            lv = to_device(lv, torch.float32, dev)
        if lv.dim() != 5 or lv.shape[1] != n_anchors:
            raise ValueError("expected [B, A, Hl, Wl, ch] per level")
        out.append(lv.contiguous())
    return out


@uses_stream
def loss_batch(x_label, x_pred, n_anchors=9, alpha=0.25, gamma=2.0, delta=1.0, stream=None, weights=None):
    """Loss over materialised RetinaNet targets -> (per_image [B,4], total [4]) = {cls, reg, 0, n_pos}."""
    dev = current_device()
    yt, yp = _pack_levels(x_label, n_anchors, dev), _pack_levels(x_pred, n_anchors, dev)
    batch, ch = int(yp[0].shape[0]), int(yp[0].shape[-1])
    shapes = [(int(p.shape[2]), int(p.shape[3]), 1) for p in yp]
    # batch-major packed layout: image stride covers all anchors, so fold anchors into rows as well
    shapes = [(int(p.shape[1]) * int(p.shape[2]), int(p.shape[3]), 1) for p in yp]
    return losses.dense_loss(yt, yp, shapes, batch, ch, 4, losses.CEN_NONE, losses.REG_SMOOTH_L1, losses.POS_GT0,
                             alpha, gamma, delta, stream=stream,
                             weights=None if weights is None else (weights[0], weights[1], 0.0))


@uses_stream
def encode_loss_batch(boxes, nbox, img_dim, num_classes, img_pad, x_pred, anchor_hw=None, iou_thresh=0.5,
                      strides=None, alpha=0.25, gamma=2.0, delta=1.0, stream=None, weights=None):
    """Fused match + encode + loss (targets never reach HBM).  x_pred: per-level [B, A, Hl, Wl, C+4].
    Returns (per_image [B,4], total [4], num_pairs [B]); with `weights` = (w_cls, w_reg) a fourth item, the
    per-level gradients d(w_cls*cls + w_reg*reg) / d x_pred (dh_retina_encode_loss_grad, same pass)."""
    strides = list(STRIDES if strides is None else strides)
    table = anchor_table() if anchor_hw is None else np.ascontiguousarray(anchor_hw, dtype=np.float32)
    n_levels, n_anchors = table.shape[0], table.shape[1]
    dev = current_device()
    check_classes(boxes, nbox, num_classes)
    boxes_d = to_device(boxes, torch.float32, dev)
    batch, nmax = int(boxes_d.shape[0]), int(boxes_d.shape[1])
    nbox_d = to_device(nbox, torch.int32, dev)
    dims_d = to_device(image_dims(img_dim, batch) if not isinstance(img_dim, torch.Tensor) or not img_dim.is_cuda
                       else img_dim, torch.float32, dev)
    yp = _pack_levels(x_pred, n_anchors, dev)
    pad_h, pad_w = int(img_pad[0]), int(img_pad[1])
    for p, s in zip(yp, strides):
        want = (batch, n_anchors, int(pad_h / s), int(pad_w / s), num_classes + 4)
        if tuple(p.shape) != want:
            raise ValueError("prediction level has shape %r, expected %r" % (tuple(p.shape), want))
    out_pi = torch.empty((batch, 4), dtype=torch.float32, device=dev)
    out_tot = torch.empty((4,), dtype=torch.float32, device=dev)
    pairs = torch.empty((batch,), dtype=torch.int32, device=dev)
    if weights is not None:
        grads = [torch.empty_like(p) for p in yp]
        _capi.check(_capi.lib().dh_retina_encode_loss_grad(
            _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax, pad_h, pad_w,
            n_levels, _capi.int_array(strides), n_anchors, _capi.float_array(table.reshape(-1).tolist()),
            float(iou_thresh), int(num_classes), _capi.ptr_array([p.data_ptr() for p in yp]), float(alpha), float(gamma),
            float(delta), float(weights[0]), float(weights[1]), _capi.ptr_array([g.data_ptr() for g in grads]),
            out_pi.data_ptr(), out_tot.data_ptr(), pairs.data_ptr(), stream_ptr(stream)), "dh_retina_encode_loss_grad")
        return out_pi, out_tot, pairs, grads
    _capi.check(_capi.lib().dh_retina_encode_loss(
        _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax, pad_h, pad_w,
        n_levels, _capi.int_array(strides), n_anchors, _capi.float_array(table.reshape(-1).tolist()),
        float(iou_thresh), int(num_classes), _capi.ptr_array([p.data_ptr() for p in yp]), float(alpha), float(gamma),
        float(delta), out_pi.data_ptr(), out_tot.data_ptr(), pairs.data_ptr(), stream_ptr(stream)),
        "dh_retina_encode_loss")
    return out_pi, out_tot, pairs


def compute_iou(boxes1, boxes2):
    """RetinaNet/utils.py:42 -- pairwise float32 IoU matrix [N, M] of centre-size boxes (on the device)."""
    dev = current_device()
    b1 = to_device(boxes1, torch.float32, dev).reshape(-1, 4).contiguous()
    b2 = to_device(boxes2, torch.float32, dev).reshape(-1, 4).contiguous()
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=dev)
    _capi.check(_capi.lib().dh_compute_iou(_capi.handle(dev.index), b1.data_ptr(), int(b1.shape[0]), b2.data_ptr(),
                                           int(b2.shape[0]), out.data_ptr(), stream_ptr(None)), "dh_compute_iou")
    return out


@uses_stream
def decode_batch(head_outputs, num_classes, img_pad, anchor_hw=None, strides=None, stream=None):
    """retinanet_module.py:487-520 for a batch: per-level heads [B, A, Hl, Wl, C+4] -> dets [B, N, 6]
    (y1, x1, y2, x2, max score, first-argmax label), order level > anchor > row-major cell."""
    strides = list(STRIDES if strides is None else strides)
    table = anchor_table() if anchor_hw is None else np.ascontiguousarray(anchor_hw, dtype=np.float32)
    n_anchors = table.shape[1]
    dev = current_device()
    heads = _pack_levels(head_outputs, n_anchors, dev)
    batch = int(heads[0].shape[0])
    pad_h, pad_w = int(img_pad[0]), int(img_pad[1])
    n = sum(n_anchors * int(pad_h / s) * int(pad_w / s) for s in strides)
    dets = torch.empty((batch, n, 6), dtype=torch.float32, device=dev)
    table_d = torch.from_numpy(table).to(dev)
    _capi.check(_capi.lib().dh_retina_decode(
        _capi.handle(dev.index), _capi.ptr_array([h.data_ptr() for h in heads]), batch, pad_h, pad_w, len(strides),
        _capi.int_array(strides), n_anchors, table_d.data_ptr(), int(num_classes), dets.data_ptr(), stream_ptr(stream)),
        "dh_retina_decode")
    return dets


@uses_stream
def detect_batch(head_outputs, num_classes, img_pad, iou_thresh=0.5, cls_thresh=0.05, anchor_hw=None, strides=None,
                 pre_nms_topk=None, with_rows=False, stream=None):
    """retinanet_module.py:483-530 for a batch, one library call (dh_retina_detect): decode -> `score >= cls_thresh`
    (-> optional per-level top-k) -> class-agnostic greedy NMS.  Returns (cand [B, n, 6] candidate rows, keep int32
    [B, n] candidate indices in kept order, n_keep [B]); with_rows=True appends the kept rows [B, n, 6] themselves.
    Without `pre_nms_topk` (the reference has none) at most 16384 / levels candidates per level reach the NMS; an
    image with more raises ValueError."""
    strides = list(STRIDES if strides is None else strides)
    table = anchor_table() if anchor_hw is None else np.ascontiguousarray(anchor_hw, dtype=np.float32)
    n_anchors = table.shape[1]
    dev = current_device()
    heads = _pack_levels(head_outputs, n_anchors, dev)
    batch = int(heads[0].shape[0])
    pad_h, pad_w = int(img_pad[0]), int(img_pad[1])
    lens = [n_anchors * int(pad_h / s) * int(pad_w / s) for s in strides]
    k = int(pre_nms_topk) if pre_nms_topk else infer.NMS_MAX_CANDIDATES // len(strides)
    k = max(min(k, max(lens + [1])), 1)
    n = len(strides) * k
    if n > infer.NMS_MAX_CANDIDATES:
        raise ValueError("pre_nms_topk * levels exceeds the NMS capacity (%d)" % infer.NMS_MAX_CANDIDATES)
    rows = torch.empty((batch, n, 6), dtype=torch.float32, device=dev)
    cand = torch.empty((batch, n, 6), dtype=torch.float32, device=dev)
    keep = torch.empty((batch, n), dtype=torch.int32, device=dev)
    n_keep = torch.zeros((batch,), dtype=torch.int32, device=dev)
    overflow = torch.zeros((batch,), dtype=torch.int32, device=dev) if not pre_nms_topk else None
    table_d = torch.from_numpy(table).to(dev)
    _capi.check(_capi.lib().dh_retina_detect(
        _capi.handle(dev.index), _capi.ptr_array([h.data_ptr() for h in heads]), batch, pad_h, pad_w, len(strides),
        _capi.int_array(strides), n_anchors, table_d.data_ptr(), int(num_classes), float(iou_thresh), float(cls_thresh),
        int(pre_nms_topk or 0), rows.data_ptr(), n, n_keep.data_ptr(), overflow.data_ptr() if overflow is not None else None,
        cand.data_ptr(), keep.data_ptr(), stream_ptr(stream)), "dh_retina_detect")
    if overflow is not None and bool(overflow.any().item()):
        raise ValueError("more than %d candidates above the score threshold on one level exceed the NMS capacity; "
                         "pass pre_nms_topk" % k)
    return (cand, keep, n_keep, rows) if with_rows else (cand, keep, n_keep)


class RetinaNetHead:
    """The reference `RetinaNet` class (RetinaNet/retinanet_module.py:162-569) without its Keras model."""

    def __init__(self, n_classes, id_2_label=None, aspect_ratios=None, anchor_scales=None, anchor_sizes=None):
        self.n_class = n_classes
        self.id_2_label = id_2_label
        self.strides = list(STRIDES)
        self.anchor_table = anchor_table(anchor_sizes, aspect_ratios, anchor_scales)
        self.n_anchors = self.anchor_table.shape[1]
        self.anchor_boxes = [[self.anchor_table[l, a].copy() for a in range(self.n_anchors)] for l in range(5)]

    def get_anchors(self, cnn_shape, level):
        """retinanet_module.py:221-246: per-anchor `[H, W, 4]` grids `(col, row, h, w)` (host, float64)."""
        if level >= 5 or level < 0:
            raise ValueError("level has to be between 0 and 4.")
        gx, gy = np.meshgrid(np.arange(0, cnn_shape[1], dtype=np.float32), np.arange(0, cnn_shape[0], dtype=np.float32))
        base = np.stack([gx, gy, np.ones_like(gx), np.ones_like(gx)], axis=-1).astype(np.float64)
        return [base * np.array([1, 1, d[0], d[1]], dtype=np.float64).reshape(1, 1, 4) for d in self.anchor_boxes[level]]

    def format_data(self, gt_labels, img_dim, iou_thresh=0.50, img_pad=None):
        """retinanet_module.py:251-365.  Returns (`out[level][anchor]` device tensors [Hl, Wl, C+4], n_pairs)."""
        g = as_host(gt_labels, np.float32).reshape(-1, 5)
        dim = as_host(img_dim, np.float32).reshape(2)
        pad = [int(v) for v in (as_host(img_pad, np.float64).reshape(2) if img_pad is not None else dim)]
        boxes, nbox = pack_labels([g])
        outs, pairs = format_data_batch(boxes, nbox, dim[None], self.n_class, pad, self.anchor_table, iou_thresh,
                                        self.strides)
        return [[o[0, a] for a in range(self.n_anchors)] for o in outs], int(pairs[0].item())

    focal_loss = staticmethod(losses.focal_loss)
    smooth_l1_loss = staticmethod(losses.smooth_l1_loss)

    def loss(self, x_pred, x_label):
        """retinanet_module.py:403 `train_loss` without the model forward: `x_pred[level][anchor]` are the
        head outputs ([1, Hl, Wl, C+4]), `x_label` what `format_data` returned.  -> (cls_loss, reg_loss)."""
        _, tot = loss_batch(x_label, x_pred, self.n_anchors)
        return tot[0], tot[1]

    def prediction_to_corners(self, xy_pred, anchor_dim, stride):
        """retinanet_module.py:428 -- [H, W, >=4] regressions -> pixel corners for one anchor shape."""
        return infer.prediction_to_corners(xy_pred, 1, stride, d0=float(anchor_dim[0]), d1=float(anchor_dim[1]))

    def cpu_nms(self, dets, base_thr):
        """retinanet_module.py:453 -- class-agnostic greedy NMS of `[n, >=5]` rows (c0, c1, c2, c3, score, ...).
        Returns the kept row indices (host int64 array, score-descending) like the reference."""
        d = to_device(dets, torch.float32, current_device())
        if d.shape[0] == 0:
            return np.zeros((0,), dtype=np.int64)
        keep, n_keep = infer.nms(d.unsqueeze(0), base_thr)
        return keep[0, :int(n_keep[0])].to(torch.int64).cpu().numpy()

    def image_detections(self, image=None, iou_thresh=0.5, cls_thresh=0.05, head_outputs=None, pre_nms_topk=None):
        """retinanet_module.py:483 -- `[k, 6]` rows (y1, x1, y2, x2, score, label) in kept order for ONE image.
        `self.model(image, training=False)` must yield `out[level][anchor]` heads [1, Hl, Wl, C+4] (or pass
        `head_outputs`).  Returns None when there are no anchors at all, as the reference does."""
        heads = head_outputs if head_outputs is not None else self.model(image, training=False)
        packed = _pack_levels(heads, self.n_anchors, current_device())
        pad = (int(packed[0].shape[2]) * self.strides[0], int(packed[0].shape[3]) * self.strides[0])
        cand, keep, n_keep = detect_batch(packed, self.n_class, pad, iou_thresh, cls_thresh, self.anchor_table, self.strides,
                                          pre_nms_topk)
        if cand.shape[1] == 0:
            return None
        k = int(n_keep[0])
        if k == 0:  # nothing above the threshold: the reference returns the (empty) thresholded array
            return cand[0, :0]
        return cand[0].index_select(0, keep[0, :k].to(torch.int64))
