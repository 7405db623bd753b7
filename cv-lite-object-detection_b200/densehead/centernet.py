"""CenterNet dense-head routines: drop-in for CenterNet/tf_centernet_resnet_s8.py,
CenterNet/tf_centernet_hourglass.py and CenterNet/tf_centernet.py of the reference."""
import numpy as np
import torch

from . import _capi, infer, losses
from ._batch import image_dims, pack_labels, check_classes
from ._tensors import as_host, current_device, stream_ptr, to_device, uses_stream

MODES = {"s8": 0, "hourglass": 1, "falloff": 2, "hourglass4": 3, "gaussian": 4}  # "gaussian": extension (DH_CENTERNET_GAUSSIAN)


@uses_stream
def format_data_batch(boxes, nbox, img_dim, num_classes, img_pad, stride=8, mode="s8", box_scales=None, sigma=0.25,
                      out=None, status=None, stream=None):
    """Encode a padded batch.  Output shape: s8 [B, H, W, S, C+4]; hourglass [B, H, W, C+4];
    falloff and gaussian (extension: DH_CENTERNET_GAUSSIAN) [B, H, W, C+5]; hourglass4 [B, H, W, 4, C+5] (the inline encoder of train_hourglass_voc.py:99-153: `img_dim` is
    the unpadded square side per image, `img_pad` the padded one).  `img_pad` is passed through with the reference's own
    indexing."""
    dev = current_device()
    check_classes(boxes, nbox, num_classes)
    boxes_d = to_device(boxes, torch.float32, dev)
    if boxes_d.dim() != 3 or boxes_d.shape[2] != 5:
        raise ValueError("boxes must be [B, Nmax, 5]")
    batch, nmax = int(boxes_d.shape[0]), int(boxes_d.shape[1])
    nbox_d = to_device(nbox, torch.int32, dev)
    dims_d = to_device(image_dims(img_dim, batch) if not isinstance(img_dim, torch.Tensor) or not img_dim.is_cuda
                       else img_dim, torch.float32, dev)
    pad0, pad1 = int(img_pad[0]), int(img_pad[1])
    m = MODES[mode]
    scales = [float(v) for v in (box_scales if box_scales is not None else [])]
    if m == 0:
        if not scales:
            raise ValueError("mode 's8' needs box_scales")
        shape = (batch, int(pad1 / stride), int(pad0 / stride), len(scales), num_classes + 4)
    elif m == 1:
        shape = (batch, int(pad1 / stride), int(pad0 / stride), num_classes + 4)
    elif m == 3:
        shape = (batch, int(pad0 / stride), int(pad1 / stride), 4, num_classes + 5)
    else:
        shape = (batch, int(pad0 / stride), int(pad1 / stride), num_classes + 5)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dev)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be contiguous float32 %r" % (shape,))
    if status is None:
        status = torch.empty((1,), dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().dh_centernet_encode(
        _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax, pad0, pad1,
        int(stride), len(scales), _capi.float_array(scales) if scales else None, float(sigma), int(num_classes), m,
        out.data_ptr(), status.data_ptr(), stream_ptr(stream)), "dh_centernet_encode")
    return out, status


def _prep(gt_labels, img_dim, img_pad):
    g = as_host(gt_labels, np.float32).reshape(-1, 5)
    dim = as_host(img_dim, np.float32).reshape(2)
    pad = [int(v) for v in (as_host(img_pad, np.float64).reshape(2) if img_pad is not None else dim)]
    boxes, nbox = pack_labels([g])
    return g, dim, pad, boxes, nbox


def format_data_s8(gt_labels, box_scales, img_dim, num_classes, img_pad=None, stride=8):
    """CenterNet/tf_centernet_resnet_s8.py:243 `format_data` -> ([H, W, S, C+4] device tensor, n_targets).
    Raises ValueError when a box is not below the largest scale, like the reference's `min([])`."""
    g, dim, pad, boxes, nbox = _prep(gt_labels, img_dim, img_pad)
    if len(g):
        f = np.float32
        bh = (g[:, 0] + f(.5) * g[:, 2]) * dim[0] - (g[:, 0] - f(.5) * g[:, 2]) * dim[0]
        bw = (g[:, 1] + f(.5) * g[:, 3]) * dim[1] - (g[:, 1] - f(.5) * g[:, 3]) * dim[1]
        if np.any(np.maximum(bh, bw) >= f(max(box_scales))):
            raise ValueError("min() arg is an empty sequence")
    out, _ = format_data_batch(boxes, nbox, dim[None], num_classes, pad, stride, "s8", box_scales)
    return out[0], len(g)


def format_data_hourglass(gt_labels, img_dim, num_classes, img_pad=None, stride=8):
    """CenterNet/tf_centernet_hourglass.py:379 `format_data` -> ([H, W, C+4], n_targets)."""
    g, dim, pad, boxes, nbox = _prep(gt_labels, img_dim, img_pad)
    out, _ = format_data_batch(boxes, nbox, dim[None], num_classes, pad, stride, "hourglass")
    return out[0], len(g)


def format_data(gt_labels, img_dim, num_classes, img_pad=None, stride=8, sigma=0.25):
    """CenterNet/tf_centernet.py:152 `format_data` -> [H, W, C+5] (inverse-power fall-off heat)."""
    g, dim, pad, boxes, nbox = _prep(gt_labels, img_dim, img_pad)
    out, _ = format_data_batch(boxes, nbox, dim[None], num_classes, pad, stride, "falloff", sigma=sigma)
    return out[0]


# ---- losses -------------------------------------------------------------------------------------
def model_loss_s8(y_true, y_pred):
    """CenterNet/tf_centernet_resnet_s8.py:368 `model_loss` on [B, H, W, S, C+4] -> (cls_loss, reg_loss)."""
    dev = current_device()
    yt = to_device(y_true, torch.float32, dev).contiguous()
    yp = to_device(y_pred, torch.float32, dev).contiguous()
    b, h, w, s, ch = (int(v) for v in yp.shape)
    _, tot = losses.dense_loss([yt], [yp], [(h, w, s)], b, ch, 4, losses.CEN_NONE, losses.REG_SMOOTH_L1, losses.POS_GT0)
    return tot[0], tot[1]


def model_loss_hourglass(y_true, y_pred):
    """CenterNet/tf_centernet_hourglass.py:492 `model_loss` on [B, H, W, C+4] -> (cls_loss, reg_loss)."""
    dev = current_device()
    yt = to_device(y_true, torch.float32, dev).contiguous()
    yp = to_device(y_pred, torch.float32, dev).contiguous()
    b, h, w, ch = (int(v) for v in yp.shape)
    _, tot = losses.dense_loss([yt], [yp], [(h, w, 1)], b, ch, 4, losses.CEN_NONE, losses.REG_SMOOTH_L1, losses.POS_GT0)
    return tot[0], tot[1]


def format_data_hourglass4(gt_labels, raw_dims, img_dims, num_classes, stride=8):
    """The 4-scale encoder CenterNet/train_hourglass_voc.py:99-153 keeps inline: one image's (cy, cx, h, w, class) rows
    (normalised by the unpadded side `raw_dims`) -> device tensor [img_dims/8, img_dims/8, 4, C+5]."""
    g = as_host(gt_labels, np.float32).reshape(-1, 5)
    boxes, nbox = pack_labels([g])
    out, _ = format_data_batch(boxes, nbox, [[raw_dims, raw_dims]], num_classes, [int(img_dims), int(img_dims)], stride, "hourglass4")
    return out[0]


def model_loss_hourglass4(bboxes, masks, outputs, loss_type="sigmoid"):
    """CenterNet/tf_hourglass_net.py:372 `model_loss(bboxes, masks, outputs)` on [B, H, W, 4, C+5] maps and a [B, H, W, 4]
    mask -> (total_cls_loss, total_reg_loss): sigmoid cross-entropy (or focal) over channels 4:, masked L1 over :4."""
    dev = current_device()
    yt = to_device(bboxes, torch.float32, dev).contiguous()
    yp = to_device(outputs, torch.float32, dev).contiguous()
    m = to_device(masks, torch.float32, dev).contiguous()
    b, h, w, s, ch = (int(v) for v in yp.shape)
    _, tot = losses.dense_loss([yt], [yp], [(h, w, s)], b, ch, 4, losses.CEN_NONE, losses.REG_SMOOTH_L1, losses.POS_MASK,
                               delta=0.0, masks=[m.reshape(b, -1)],
                               cls_mode=losses.CLS_SIGMOID_BCE if loss_type == "sigmoid" else losses.CLS_FOCAL)
    return tot[0], tot[1]


def model_loss(y_true, y_pred, reg_type="l1", cen_type="l1"):
    """CenterNet/tf_centernet.py:428 `model_loss`: y_true [H, W, C+5], y_pred [1, H, W, C+5] (or batched)."""
    dev = current_device()
    yt = to_device(y_true, torch.float32, dev)
    yp = to_device(y_pred, torch.float32, dev)
    if yt.dim() == 3:
        yt = yt.unsqueeze(0)
    if yp.dim() == 3:
        yp = yp.unsqueeze(0)
    b, h, w, ch = (int(v) for v in yp.shape)
    cen = losses.CEN_SMOOTH_L1 if cen_type.lower() == "l1" else losses.CEN_IGNORE
    reg = losses.reg_mode(reg_type)
    _, tot = losses.dense_loss([yt.contiguous()], [yp.contiguous()], [(h, w, 1)], b, ch, 4, cen, reg, losses.POS_GE1)
    return tot[0], tot[1], tot[2]


@uses_stream
def encode_loss_batch(boxes, nbox, img_dim, num_classes, img_pad, y_pred, stride=8, mode="s8", box_scales=None,
                      sigma=0.25, reg_type="l1", alpha=0.25, gamma=2.0, delta=1.0, stream=None, weights=None, cls_type="focal"):
    """Fused CenterNet encode + loss.  Returns (per_image [B,4], total [4], status [1]); with `weights` = (w_cls,
    w_reg, w_cen) a fourth item, the gradient d(w . {cls, reg, cen}) / d y_pred (dh_centernet_encode_loss_grad)."""
    dev = current_device()
    check_classes(boxes, nbox, num_classes)
    boxes_d = to_device(boxes, torch.float32, dev)
    batch, nmax = int(boxes_d.shape[0]), int(boxes_d.shape[1])
    nbox_d = to_device(nbox, torch.int32, dev)
    dims_d = to_device(image_dims(img_dim, batch) if not isinstance(img_dim, torch.Tensor) or not img_dim.is_cuda
                       else img_dim, torch.float32, dev)
    yp = to_device(y_pred, torch.float32, dev).contiguous()
    scales = [float(v) for v in (box_scales if box_scales is not None else [])]
    out_pi = torch.empty((batch, 4), dtype=torch.float32, device=dev)
    out_tot = torch.empty((4,), dtype=torch.float32, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    if weights is not None:
        grad = torch.empty_like(yp)
        _capi.check(_capi.lib().dh_centernet_encode_loss_grad(
            _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax,
            int(img_pad[0]), int(img_pad[1]), int(stride), len(scales), _capi.float_array(scales) if scales else None,
            float(sigma), int(num_classes), MODES[mode], yp.data_ptr(),
            losses.reg_mode(reg_type),
            losses.CLS_SIGMOID_BCE if cls_type == "sigmoid" else losses.CLS_FOCAL, float(alpha), float(gamma), float(delta),
            float(weights[0]), float(weights[1]), float(weights[2]), grad.data_ptr(), out_pi.data_ptr(), out_tot.data_ptr(),
            status.data_ptr(), stream_ptr(stream)), "dh_centernet_encode_loss_grad")
        return out_pi, out_tot, status, grad
    _capi.check(_capi.lib().dh_centernet_encode_loss(
        _capi.handle(dev.index), boxes_d.data_ptr(), nbox_d.data_ptr(), dims_d.data_ptr(), batch, nmax,
        int(img_pad[0]), int(img_pad[1]), int(stride), len(scales), _capi.float_array(scales) if scales else None,
        float(sigma), int(num_classes), MODES[mode], yp.data_ptr(),
        losses.reg_mode(reg_type),
        losses.CLS_SIGMOID_BCE if cls_type == "sigmoid" else losses.CLS_FOCAL, float(alpha), float(gamma), float(delta),
        out_pi.data_ptr(), out_tot.data_ptr(), status.data_ptr(), stream_ptr(stream)), "dh_centernet_encode_loss")
    return out_pi, out_tot, status


# ---- inference ------------------------------------------------------------------------------------
def prediction_to_corners(xy_pred, stride):
    """CenterNet/tf_centernet.py:128 / tf_centernet_hourglass.py:355 (same as FCOS): tblr -> corners."""
    return infer.prediction_to_corners(xy_pred, 0, stride)


def prediction_to_corners_s8(xy_pred, box_scales, stride=8):
    """CenterNet/tf_centernet_resnet_s8.py:210 -- [H, W, S, >=4] -> [H, W, S, 4]."""
    return infer.prediction_to_corners(xy_pred, 3, stride, scales=box_scales)


def bboxes_iou(boxes1, boxes2):
    """CenterNet/tf_centernet_resnet_s8.py:22 -- float64 IoU of corner boxes ([..., 4]), floored at float32 eps.
    Shapes must match or one side must be a single box."""
    dev = current_device()
    b1 = to_device(np.asarray(as_host(boxes1, np.float64)), torch.float64, dev).reshape(-1, 4).contiguous()
    b2 = to_device(np.asarray(as_host(boxes2, np.float64)), torch.float64, dev).reshape(-1, 4).contiguous()
    n = max(int(b1.shape[0]), int(b2.shape[0])) if min(int(b1.shape[0]), int(b2.shape[0])) else 0
    out = torch.empty((n,), dtype=torch.float64, device=dev)
    _capi.check(_capi.lib().dh_bboxes_iou(_capi.handle(dev.index), b1.data_ptr(), int(b1.shape[0]), b2.data_ptr(),
                                          int(b2.shape[0]), out.data_ptr(), stream_ptr(None)), "dh_bboxes_iou")
    return out


def nms(bboxes, iou_threshold, sigma=0.3, method="nms"):
    """CenterNet/tf_centernet_resnet_s8.py:44 `nms`: rows (xmin, ymin, w, h, score, class) -> list of kept rows
    (x1, y1, x2, y2, score, class) as host float64 arrays, classes visited in ascending order.  Unlike the
    reference the input array is left untouched."""
    assert method in ["nms", "soft-nms"]
    dev = current_device()
    host = as_host(bboxes, np.float64).reshape(-1, 6)
    n = int(host.shape[0])
    if n == 0:
        return []
    classes = np.unique(host[:, 5])
    rows_d = to_device(host, torch.float64, dev)
    cls_d = to_device(classes, torch.float64, dev)
    out = torch.empty((n, 6), dtype=torch.float64, device=dev)
    src = torch.empty((n,), dtype=torch.int32, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().dh_centernet_nms(
        _capi.handle(dev.index), rows_d.data_ptr(), n, cls_d.data_ptr(), int(len(classes)), float(iou_threshold), float(sigma),
        1 if method == "soft-nms" else 0, out.data_ptr(), src.data_ptr(), cnt.data_ptr(), stream_ptr(None)), "dh_centernet_nms")
    k = int(cnt[0])
    return list(out[:k].cpu().numpy())
