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


_KDL_CUDA = (2, 13)  # DLDeviceType: kDLCUDA, kDLCUDAManaged


def _on_cuda(x):
    """True when `x` speaks DLPack and says it lives in CUDA memory."""
    try:
        return int(x.__dlpack_device__()[0]) in _KDL_CUDA
    except Exception:
        return False


def _tf_gpu_capsule(x):
    """TensorFlow eager tensors before `__dlpack__` existed: tf.experimental.dlpack.to_dlpack."""
    if not type(x).__module__.startswith("tensorflow") or "GPU" not in str(getattr(x, "device", "")):
        return None
    try:
        import tensorflow as tf
        return tf.experimental.dlpack.to_dlpack(x)
    except Exception:
        return None


def uses_stream(fn):
    """Run a wrapper that takes `stream=` with that stream current: its staging copies and output allocations are then
    ordered on (and owned by) the stream the kernels are launched on."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        stream = kwargs.get("stream")
        if stream is None or stream == torch.cuda.current_stream():
            return fn(*args, **kwargs)
        with torch.cuda.stream(stream):
            return fn(*args, **kwargs)
    return wrapped


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
    elif not isinstance(x, (torch.Tensor, np.ndarray)):
        # device-resident objects cross zero-copy even when they also offer .numpy() (a TensorFlow EagerTensor on the
        # GPU does): asking for .numpy() there would be a D2H + H2D round trip
        if hasattr(x, "__dlpack__") and (_on_cuda(x) or not hasattr(x, "numpy")):
            x = torch.from_dlpack(x)
        else:
            cap = _tf_gpu_capsule(x)
            if cap is not None:
                x = torch.utils.dlpack.from_dlpack(cap)
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            if x.device != device:
                raise ValueError("tensor lives on %s but the current device is %s" % (x.device, device))
            return x.to(dtype).contiguous()
        x = x.detach().numpy()
    np_dtype = {torch.float32: np.float32, torch.int32: np.int32, torch.float64: np.float64}[dtype]
    return _stage(as_host(x, np_dtype), device)
