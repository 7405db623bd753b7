"""Inference-side device routines: prediction_to_corners, decode front ends, pre-NMS top-k, greedy NMS.

Reference counterparts (paths relative to the reference repository): `prediction_to_corners`
(FCOS/fcos.py:112, fcos_center_v1.py:125, RetinaNet/retinanet_module.py:428, CenterNet/
tf_centernet_resnet_s8.py:210), `image_detections` (FCOS/infer_fcos.py:27, retinanet_module.py:483),
`cpu_nms` (retinanet_module.py:453).
"""
import numpy as np
import torch

from . import _capi
from ._tensors import current_device, stream_ptr, to_device, uses_stream

NMS_AGNOSTIC, NMS_PER_CLASS = 0, 1
NMS_MAX_CANDIDATES = 16384


@uses_stream
def prediction_to_corners(xy_pred, mode, stride, d0=0.0, d1=0.0, scales=None, stream=None):
    """xy_pred: [..., H, W, >=4] (modes 0-2) or [..., H, W, S, >=4] (mode 3) -> same leading shape + [4]."""
    dev = current_device()
    p = to_device(xy_pred, torch.float32, dev).contiguous()
    sub = 1
    if mode == 3:
        sub = int(p.shape[-2])
        h, w = int(p.shape[-4]), int(p.shape[-3])
        lead = p.shape[:-4]
    else:
        h, w = int(p.shape[-3]), int(p.shape[-2])
        lead = p.shape[:-3]
    batch = int(np.prod(lead)) if len(lead) else 1
    out = torch.empty(tuple(p.shape[:-1]) + (4,), dtype=torch.float32, device=dev)
    sc = _capi.float_array([float(v) for v in scales]) if scales is not None else None
    _capi.check(_capi.lib().dh_prediction_to_corners(
        _capi.handle(dev.index), p.data_ptr(), batch, h, w, sub, int(p.shape[-1]), int(mode), float(stride), float(d0),
        float(d1), sc, out.data_ptr(), stream_ptr(stream)), "dh_prediction_to_corners")
    return out


@uses_stream
def nms(dets, iou_thr, mode=NMS_AGNOSTIC, min_score=-float("inf"), score_inclusive=True, n_valid=None, num_classes=0,
        max_per_class=0, max_total=0, max_out=None, stream=None):
    """dets [B, n, >=5(6)] float32 on the device -> (keep int32 [B, max_out], n_keep int32 [B])."""
    dev = current_device()
    d = to_device(dets, torch.float32, dev).contiguous()
    if d.dim() != 3:
        raise ValueError("dets must be [B, n, row]")
    batch, n, row = (int(v) for v in d.shape)
    if n > NMS_MAX_CANDIDATES:
        raise ValueError("%d candidates per image exceed the NMS capacity of %d; use a pre-NMS top-k" % (n, NMS_MAX_CANDIDATES))
    max_out = int(max_out if max_out is not None else (max_total if max_total > 0 else max(n, 1)))
    keep = torch.empty((batch, max_out), dtype=torch.int32, device=dev)
    n_keep = torch.zeros((batch,), dtype=torch.int32, device=dev)
    nv = to_device(n_valid, torch.int32, dev) if n_valid is not None else None
    _capi.check(_capi.lib().dh_nms(
        _capi.handle(dev.index), d.data_ptr(), nv.data_ptr() if nv is not None else None, batch, n, row, int(mode),
        float(iou_thr), float(min_score), 1 if score_inclusive else 0, int(num_classes), int(max_per_class), int(max_total),
        keep.data_ptr(), max_out, n_keep.data_ptr(), stream_ptr(stream)), "dh_nms")
    return keep, n_keep


@uses_stream
def select_topk(dets, seg_offsets, k, min_score, score_inclusive=True, score_col=4, with_source=False, stream=None):
    """Per-segment threshold + exact top-k.  dets [B, n, row] -> [B, n_seg*k, row] (unused slots: score -inf)."""
    dev = current_device()
    d = to_device(dets, torch.float32, dev).contiguous()
    batch, n, row = (int(v) for v in d.shape)
    seg = torch.as_tensor(np.asarray(seg_offsets, dtype=np.int32)).to(dev)
    n_seg = int(seg.numel()) - 1
    out = torch.empty((batch, n_seg * k, row), dtype=torch.float32, device=dev)
    src = torch.empty((batch, n_seg * k), dtype=torch.int32, device=dev) if with_source else None
    _capi.check(_capi.lib().dh_select_topk(
        _capi.handle(dev.index), d.data_ptr(), batch, n, row, int(score_col), seg.data_ptr(), n_seg, int(k), float(min_score),
        1 if score_inclusive else 0, out.data_ptr(), src.data_ptr() if src is not None else None, stream_ptr(stream)),
        "dh_select_topk")
    return (out, src) if with_source else out
