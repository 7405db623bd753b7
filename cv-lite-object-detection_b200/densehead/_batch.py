"""Ragged GT lists -> the padded [B, Nmax, 5] + nbox[B] layout the C ABI takes."""
import numpy as np

from ._tensors import as_host


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
