"""Device-memory plumbing for the Python mirror of the reference interface.

PyTorch is only the allocator / stream / DLPack provider here.  Inputs may be NumPy arrays, objects
with `.numpy()` (TensorFlow eager tensors), torch tensors, or anything speaking DLPack (zero-copy when
already on the device).  Host inputs go through a small ring of pinned staging buffers so that the
host->device copy is asynchronous on the current stream.
"""
import numpy as np
import torch

_RING = 4
_staging = {}  # (device index, nbytes bucket) -> [ (pinned uint8 tensor, event) ... ], next slot


def current_device():
    if not torch.cuda.is_available():
        raise RuntimeError("densehead needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return int(s.cuda_stream)


def _is_dlpack_capsule(x):
    return type(x).__name__ == "PyCapsule"


def as_host(x, dtype=np.float32):
    """Host NumPy view/copy of a small input (GT labels, image size)."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().astype(dtype, copy=False)
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def _stage(host, device):
    nbytes = max(int(host.nbytes), 1)
    bucket = 1 << (nbytes - 1).bit_length()
    key = (device.index, bucket)
    ring = _staging.get(key)
    if ring is None:
        ring = _staging[key] = {"slots": [], "next": 0}
    if len(ring["slots"]) < _RING:
        ring["slots"].append([torch.empty(bucket, dtype=torch.uint8, pin_memory=True), None])
    slot = ring["slots"][ring["next"] % len(ring["slots"])]
    ring["next"] += 1
    if slot[1] is not None:
        slot[1].synchronize()  # the copy that last used this pinned block has completed
    tdtype = torch.from_numpy(np.empty(0, dtype=host.dtype)).dtype
    dev = torch.empty(host.shape, dtype=tdtype, device=device)
    if host.size:
        pinned = slot[0][:nbytes].view(tdtype).reshape(host.shape)
        pinned.copy_(torch.from_numpy(host))
        dev.copy_(pinned, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    slot[1] = ev
    return dev


def to_device(x, dtype, device=None):
    """Contiguous device tensor of `dtype` (torch dtype).  Zero-copy for device-resident inputs."""
    device = device or current_device()
    if _is_dlpack_capsule(x):
        x = torch.utils.dlpack.from_dlpack(x)
    elif not isinstance(x, (torch.Tensor, np.ndarray)) and hasattr(x, "__dlpack__") and not hasattr(x, "numpy"):
        x = torch.from_dlpack(x)
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            if x.device != device:
                raise ValueError("tensor lives on %s but the current device is %s" % (x.device, device))
            return x.to(dtype).contiguous()
        x = x.detach().numpy()
    np_dtype = {torch.float32: np.float32, torch.int32: np.int32, torch.float64: np.float64}[dtype]
    return _stage(as_host(x, np_dtype), device)
