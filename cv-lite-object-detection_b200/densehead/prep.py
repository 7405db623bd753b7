"""Label preparation and result formatting on the device: the steps immediately before and after the dense-head path.

Reference counterparts: `swap_xy`, `convert_to_xywh`, `convert_to_corners` (FCOS/utils.py:6-40 and the copies in
RetinaNet/ and CenterNet/), the box half of `random_flip_horizontal` (FCOS/data_preprocess.py:24-41), the label assembly of
`preprocess_data` + `train()` (FCOS/data_preprocess.py:121-131, FCOS/train_fcos.py:131-135) and `detect_bboxes`'
post-processing (RetinaNet/retinanet_module.py:559-569)."""
import numpy as np
import torch

from . import _capi
from ._tensors import current_device, stream_ptr, to_device, uses_stream

SWAP_XY, TO_XYWH, TO_CORNERS, FLIP_HORIZONTAL = 0, 1, 2, 3


@uses_stream
def _convert(boxes, mode, stream=None):
    dev = current_device()
    b = to_device(boxes, torch.float32, dev).contiguous()
    if b.shape[-1] != 4:
        raise ValueError("boxes must have a last axis of 4")
    out = torch.empty_like(b)
    _capi.check(_capi.lib().dh_box_convert(_capi.handle(dev.index), b.data_ptr(), b.numel() // 4, int(mode), out.data_ptr(),
                                           stream_ptr(stream)), "dh_box_convert")
    return out


def swap_xy(boxes):
    """FCOS/utils.py:6 -- (a, b, c, d) -> (b, a, d, c)."""
    return _convert(boxes, SWAP_XY)


def convert_to_xywh(boxes):
    """FCOS/utils.py:16 -- (lo0, lo1, hi0, hi1) -> ((lo + hi) / 2, hi - lo)."""
    return _convert(boxes, TO_XYWH)


def convert_to_corners(boxes):
    """FCOS/utils.py:29 -- (c0, c1, s0, s1) -> (c - s / 2, c + s / 2)."""
    return _convert(boxes, TO_CORNERS)


def flip_boxes_horizontal(boxes):
    """The box half of random_flip_horizontal (FCOS/data_preprocess.py:36-39) on normalised (xmin, ymin, xmax, ymax)."""
    return _convert(boxes, FLIP_HORIZONTAL)


@uses_stream
def prepare_labels(bboxes, classes, flip=None, max_boxes=None, offsets=None, nbox=None, stream=None):
    """Dataset boxes -> (labels [B, max_boxes, 5] = (cy, cx, h, w, class), nbox [B]) on the device.

    Either a list of per-image `[n_i, 4]` (xmin, ymin, xmax, ymax) arrays with a matching list of class-id arrays (packed
    ragged on the host, one copy), or ragged device tensors `bboxes [total, 4]`, `classes [total]` with `offsets [B+1]`, or
    padded `[B, n, 4]` / `[B, n]` tensors with `nbox`.

    `max_boxes` defaults to the longest image, so nothing is dropped.  An explicit `max_boxes` smaller than an image's box
    count keeps that image's first `max_boxes` boxes (the reference keeps all of them, FCOS/train_fcos.py:131-135): with
    host lists this raises ValueError here; with device inputs the kernel sets DH_STATUS_TRUNCATED, which
    `densehead.raise_for_status()` turns into the same ValueError (a synchronising read)."""
    dev = current_device()
    if isinstance(bboxes, (list, tuple)):
        counts = [len(b) for b in bboxes]
        if max_boxes is not None and counts and max(counts) > int(max_boxes):
            raise ValueError("prepare_labels: an image has %d boxes, max_boxes is %d" % (max(counts), int(max_boxes)))
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        bboxes = np.concatenate([np.asarray(b, np.float32).reshape(-1, 4) for b in bboxes] + [np.zeros((0, 4), np.float32)])
        classes = np.concatenate([np.asarray(c, np.float32).reshape(-1) for c in classes] + [np.zeros((0,), np.float32)])
    raw = to_device(bboxes, torch.float32, dev).contiguous()
    cls = to_device(classes, torch.float32, dev).contiguous()
    if offsets is not None:
        off = to_device(offsets, torch.int32, dev)
        batch = int(off.numel()) - 1
        if max_boxes is None:
            max_boxes = int((off[1:] - off[:-1]).max().item()) if batch else 1
        in_max, nb = 0, None
    else:
        if raw.dim() != 3:
            raise ValueError("padded input must be [B, n, 4]")
        off, batch, in_max = None, int(raw.shape[0]), int(raw.shape[1])
        nb = to_device(nbox, torch.int32, dev) if nbox is not None else None
        if max_boxes is None:
            max_boxes = in_max
    max_boxes = max((int(max_boxes) + 3) & ~3, 4)  # keeps every image's rows 16-byte aligned for the TMA bulk load
    fl = to_device(np.asarray(flip, np.int32) if not isinstance(flip, torch.Tensor) else flip, torch.int32, dev) if flip is not None else None
    out = torch.empty((batch, max_boxes, 5), dtype=torch.float32, device=dev)
    out_n = torch.empty((batch,), dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().dh_prepare_labels(
        _capi.handle(dev.index), raw.data_ptr(), cls.data_ptr(), off.data_ptr() if off is not None else None,
        nb.data_ptr() if nb is not None else None, fl.data_ptr() if fl is not None else None, batch, in_max, max_boxes,
        out.data_ptr(), out_n.data_ptr(), stream_ptr(stream)), "dh_prepare_labels")
    return out, out_n


@uses_stream
def format_detections(rows, n_keep, ratios, stream=None):
    """RetinaNet/retinanet_module.py:559-569 for a batch: kept rows [B, n, 6] (y1, x1, y2, x2, score, label) ->
    (boxes [B, n, 4] as (x1, y1, x2, y2) in source-image pixels, scores [B, n], labels int32 [B, n])."""
    dev = current_device()
    r = to_device(rows, torch.float32, dev).contiguous()
    batch, n = int(r.shape[0]), int(r.shape[1])
    nk = to_device(n_keep, torch.int32, dev)
    ra = to_device(np.asarray(ratios, np.float32).reshape(batch, 2) if not isinstance(ratios, torch.Tensor) else ratios, torch.float32, dev).contiguous()
    boxes = torch.empty((batch, n, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((batch, n), dtype=torch.float32, device=dev)
    labels = torch.empty((batch, n), dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().dh_format_detections(_capi.handle(dev.index), r.data_ptr(), nk.data_ptr(), ra.data_ptr(), batch, n,
                                                 boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), stream_ptr(stream)),
                "dh_format_detections")
    return boxes, scores, labels
