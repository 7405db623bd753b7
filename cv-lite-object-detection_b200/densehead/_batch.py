"""Ragged GT lists -> the padded [B, Nmax, 5] + nbox[B] layout the C ABI takes."""
import numpy as np

from ._tensors import as_host


def check_classes(labels, nbox, num_classes):
    """The reference indexes the class channels with the GT class (FCOS/fcos.py:281-283): an id outside
    [0, num_classes) raises IndexError there, so it does here (host inputs; device inputs are flagged by the kernels)."""
    if not isinstance(labels, np.ndarray) or labels.size == 0:
        return
    cls = labels[..., 4]
    if labels.ndim == 3 and nbox is not None and isinstance(nbox, np.ndarray):
        cls = cls[np.arange(labels.shape[1])[None, :] < nbox.reshape(-1, 1)]
    if cls.size and (cls.min() < 0 or cls.max() >= num_classes or not np.all(np.isfinite(cls))):
        bad = cls[(cls < 0) | (cls >= num_classes) | ~np.isfinite(cls)][0]
        raise IndexError("index %d is out of bounds for axis 2 with size %d" % (int(bad) if np.isfinite(bad) else -1, num_classes))


def pack_labels(labels_list, max_boxes=None):
    rows = [as_host(g, np.float32).reshape(-1, 5) for g in labels_list]
    nmax = max([len(r) for r in rows] + [1]) if max_boxes is None else int(max_boxes)
    nmax = (nmax + 3) & ~3  # keeps every image's row block 16-byte aligned for the TMA bulk load
    boxes = np.zeros((len(rows), nmax, 5), dtype=np.float32)
    nbox = np.zeros((len(rows),), dtype=np.int32)
    for b, r in enumerate(rows):
        if len(r) > nmax:
            raise ValueError("image %d has %d boxes > max_boxes %d" % (b, len(r), nmax))
        boxes[b, :len(r)] = r
        nbox[b] = len(r)
    return boxes, nbox


def image_dims(img_dim, batch):
    d = as_host(img_dim, np.float32).reshape(-1, 2)
    if d.shape[0] == 1 and batch != 1:
        d = np.repeat(d, batch, axis=0)
    if d.shape[0] != batch:
        raise ValueError("img_dim has %d rows for a batch of %d" % (d.shape[0], batch))
    return np.ascontiguousarray(d)
