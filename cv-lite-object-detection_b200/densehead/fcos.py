"""FCOS dense-head routines: drop-in for FCOS/fcos.py, FCOS/fcos_center.py and
FCOS/fcos_center_v1.py of the reference (paths relative to the reference repository).

`format_data*` keep the reference signatures and return device tensors (torch CUDA tensors, which
TensorFlow ingests zero-copy through DLPack) instead of host float64 arrays; values equal
`reference_map.astype(float32)`.
"""
import ctypes

import numpy as np
import torch

from . import _capi
from ._batch import image_dims, pack_labels
from ._tensors import as_host, current_device, stream_ptr, to_device

DEFAULT_STRIDES = [8, 16, 32, 64, 128]
DEFAULT_B_DIM = [32, 64, 128, 256]
MODES = {"fcos": 0, "center": 1, "center_only": 2, "center_v1": 3}


def level_shapes(img_pad, strides):
    return [(int(img_pad[0] / s), int(img_pad[1] / s)) for s in strides]


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
