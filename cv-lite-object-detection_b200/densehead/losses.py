"""Dense-head losses on the device: the reference's focal_loss / smooth_l1_loss / iou_loss
(FCOS/fcos.py:380-462; identical copies live in every other module of the reference) and the generic
(prediction, target) map loss that the per-detector `model_loss` wrappers are built on.

Everything returns device tensors; nothing synchronises.  A loss result is a float32 vector
{cls, reg, cen, n_pos}.
"""
import numpy as np
import torch

from . import _capi
from ._tensors import current_device, stream_ptr, to_device, uses_stream

CEN_NONE, CEN_SMOOTH_L1, CEN_FOCAL, CEN_IGNORE = 0, 1, 2, 3
REG_SMOOTH_L1, REG_IOU, REG_GIOU = 0, 1, 2


def reg_mode(reg_type):
    """`reg_type` of the reference's model_loss ("l1" / "iou", FCOS/fcos.py:464) plus the "giou" extension."""
    return {"iou": REG_IOU, "giou": REG_GIOU}.get(str(reg_type).lower(), REG_SMOOTH_L1)
POS_GE1, POS_GT0, POS_MASK = 0, 1, 2
CLS_FOCAL, CLS_SIGMOID_BCE = 0, 1


@uses_stream
def dense_loss(targets, preds, shapes, batch, ch, reg_ch, cen_mode, reg_mode, pos_rule, alpha=0.25, gamma=2.0,
               delta=1.0, masks=None, per_image=True, stream=None, weights=None, cls_mode=CLS_FOCAL):
    """targets/preds: lists of contiguous float32 device tensors, map m holding [B, H*W*sub, ch] rows.
    shapes: list of (H, W, sub).  Returns (per_image [B,4] or None, total [4]); with `weights` = (w_cls, w_reg,
    w_cen) also the gradient maps d(w . {cls, reg, cen}) / d preds (dh_dense_loss_grad, same pass)."""
    dev = current_device()
    for t, p in zip(targets, preds):
        if t.numel() != p.numel():
            raise ValueError("target and prediction maps differ in size: %r vs %r" % (tuple(t.shape), tuple(p.shape)))
    out_pi = torch.empty((batch, 4), dtype=torch.float32, device=dev) if per_image else None
    out_tot = torch.empty((4,), dtype=torch.float32, device=dev)
    n = len(targets)
    common = [_capi.handle(dev.index), n, _capi.ptr_array([t.data_ptr() for t in targets]),
              _capi.ptr_array([p.data_ptr() for p in preds]),
              _capi.ptr_array([m.data_ptr() for m in masks]) if masks is not None else None,
              _capi.int_array([s[0] for s in shapes]), _capi.int_array([s[1] for s in shapes]),
              _capi.int_array([s[2] for s in shapes]), int(batch), int(ch), int(reg_ch), int(cen_mode), int(reg_mode),
              int(pos_rule), int(cls_mode), float(alpha), float(gamma), float(delta)]
    outs = [out_pi.data_ptr() if per_image else None, out_tot.data_ptr(), stream_ptr(stream)]
    if weights is None:
        _capi.check(_capi.lib().dh_dense_loss(*(common + outs)), "dh_dense_loss")
        return out_pi, out_tot
    grads = [torch.empty_like(p) for p in preds]
    _capi.check(_capi.lib().dh_dense_loss_grad(*(common + [float(weights[0]), float(weights[1]), float(weights[2]),
                                                            _capi.ptr_array([g.data_ptr() for g in grads])] + outs)),
                "dh_dense_loss_grad")
    return out_pi, out_tot, grads


class WeightedLoss(torch.autograd.Function):
    """Differentiable scalar `w_cls*cls + w_reg*reg + w_cen*cen` for torch callers (the reference differentiates its
    model_loss with tf.GradientTape, FCOS/train_fcos.py:152-176).  `run(weights)` must return (total [4], grads list)
    computed by one of the *_grad entry points for the prediction tensors passed as `preds`."""

    @staticmethod
    def forward(ctx, run, weights, *preds):
        total, grads = run(weights)
        ctx.grads = grads
        w = torch.tensor([float(weights[0]), float(weights[1]), float(weights[2])], device=total.device)
        return (total[:3] * w).sum()

    @staticmethod
    def backward(ctx, g):
        return (None, None) + tuple(g * x for x in ctx.grads)


def _flat(x, dev):
    return to_device(x, torch.float32, dev).reshape(-1)


def focal_loss(labels, logits, alpha=0.25, gamma=2.0):
    """FCOS/fcos.py:443 -- sum-reduced focal loss (labels may be multi-hot or fractional)."""
    dev = current_device()
    y, x = _flat(labels, dev), _flat(logits, dev)
    if y.numel() != x.numel():
        raise ValueError("labels and logits differ in size")
    n = y.numel()
    if n == 0:
        return torch.zeros((), device=dev)
    _, tot = dense_loss([y], [x], [(n, 1, 1)], 1, 1, 0, CEN_NONE, REG_SMOOTH_L1, POS_GT0, alpha, gamma, 1.0,
                        per_image=False)
    return tot[0]


def smooth_l1_loss(xy_true, xy_pred, mask=1.0, delta=1.0):
    """FCOS/fcos.py:380 -- `sum(mask[..., None] * where(|d| < delta, d*d/2, |d|))`."""
    dev = current_device()
    t = to_device(xy_true, torch.float32, dev)
    p = to_device(xy_pred, torch.float32, dev)
    if t.shape != p.shape:
        raise ValueError("xy_true and xy_pred differ in shape")
    if np.isscalar(mask) or (isinstance(mask, torch.Tensor) and mask.dim() == 0):
        n = t.numel()
        pad = (-n) % 4
        tf, pf = t.reshape(-1), p.reshape(-1)
        if pad:
            z = torch.zeros(pad, device=dev)
            tf, pf = torch.cat([tf, z]), torch.cat([pf, z])
        rows = (n + pad) // 4
        m = torch.full((rows,), float(mask), dtype=torch.float32, device=dev)
    else:
        m = to_device(mask, torch.float32, dev)
        if t.shape[-1] != 4 or tuple(m.shape) != tuple(t.shape[:-1]):
            raise NotImplementedError("mask must be a scalar or have the shape of xy_true without its last (=4) axis")
        tf, pf, m = t.reshape(-1), p.reshape(-1), m.reshape(-1).contiguous()
        rows = m.numel()
    if rows == 0:
        return torch.zeros((), device=dev)
    _, tot = dense_loss([tf], [pf], [(rows, 1, 1)], 1, 4, 4, CEN_NONE, REG_SMOOTH_L1, POS_MASK, 0.25, 2.0, delta,
                        masks=[m], per_image=False)
    return tot[1]


def giou_loss(xy_true, xy_pred, mask):
    """EXTENSION (the reference has no GIoU loss): `sum((1 - GIoU) * mask)` on the box construction of `iou_loss`
    (DH_REG_GIOU; its specification is giou_loss of the CPU oracle)."""
    return iou_loss(xy_true, xy_pred, mask, _reg=REG_GIOU)


def iou_loss(xy_true, xy_pred, mask, _reg=REG_IOU):
    """FCOS/fcos.py:393 -- `sum(-log(IoU + 1e-12) * mask)` for [H, W, 4] tblr maps on the integer grid.
    Rows whose mask is 0 contribute exactly 0 (the reference would propagate a NaN from them)."""
    dev = current_device()
    t = to_device(xy_true, torch.float32, dev)
    p = to_device(xy_pred, torch.float32, dev)
    m = to_device(mask, torch.float32, dev)
    if t.dim() != 3 or t.shape[-1] != 4 or t.shape != p.shape or tuple(m.shape) != tuple(t.shape[:2]):
        raise ValueError("expected [H, W, 4] maps and an [H, W] mask")
    h, w = int(t.shape[0]), int(t.shape[1])
    _, tot = dense_loss([t.reshape(-1)], [p.reshape(-1)], [(h, w, 1)], 1, 4, 4, CEN_NONE, _reg, POS_MASK,
                        masks=[m.reshape(-1).contiguous()], per_image=False)
    return tot[1]
