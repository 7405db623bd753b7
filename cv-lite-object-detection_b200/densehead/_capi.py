"""ctypes binding of libdensehead.so (the C ABI in include/densehead.h).

There is NO CPU fallback: if the shared library is missing or a call fails, this raises.
PyTorch is used only as the device-memory / stream provider; tensors cross the boundary as raw
device pointers (`_tensors.to_device` consumes DLPack capsules and `__dlpack__` objects zero-copy; `_dlpack.py` reads a
capsule without any framework, which is what a TensorFlow-side binding uses).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DENSEHEAD_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libdensehead.so"))

DH_OK = 0
DH_ERR_BAD_ARG, DH_ERR_SHAPE, DH_ERR_CUDA, DH_ERR_CAPACITY, DH_ERR_NCCL = -1, -2, -3, -4, -5
DH_OPT_TMA_STORE, DH_OPT_TILE_BYTES, DH_OPT_CTAS_PER_SM = 1, 2, 3
DH_OPT_LOSS_ALLREDUCE, DH_OPT_ALLREDUCE, DH_OPT_FUSED_TAIL, DH_OPT_ENCODE_KERNEL, DH_OPT_FUSED_MAX_CHUNK = 11, 12, 13, 14, 15
DH_OPT_NMS_FILTER, DH_OPT_NMS_CHAIN = 16, 17
DH_STATUS_BAD_SCALE, DH_STATUS_BAD_CLASS, DH_STATUS_COMM_TIMEOUT, DH_STATUS_TRUNCATED = 1, 2, 4, 8
DH_UNIQUE_ID_BYTES, DH_IPC_HANDLE_BYTES = 128, 64
DH_MAX_BOXES_PER_IMAGE = 256

c_fp = ctypes.POINTER(ctypes.c_float)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_vp = ctypes.c_void_p


class DenseHeadError(RuntimeError):
    pass


_lib = None
_lock = threading.RLock()  # re-entrant: handle() creates the handle under the lock and lib() takes it too
_handles = {}


def lib():
    """Load libdensehead.so once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DenseHeadError(
                "libdensehead.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.dh_version.restype = ctypes.c_char_p
        L.dh_last_error.restype = ctypes.c_char_p
        L.dh_create.argtypes = [ctypes.POINTER(c_vp), ctypes.c_int]
        L.dh_destroy.argtypes = [c_vp]
        L.dh_set_option.argtypes = [c_vp, ctypes.c_int, ctypes.c_int]
        L.dh_launch_count.argtypes = [c_vp]
        L.dh_launch_count.restype = ctypes.c_longlong
        L.dh_read_phase_timing.argtypes = [c_vp, ctypes.POINTER(ctypes.c_longlong)]
        I, F, P = ctypes.c_int, ctypes.c_float, c_vp
        L.dh_fcos_encode.argtypes = [P, P, P, P, I, I, I, I, I, c_ip, c_fp, I, I, ctypes.POINTER(c_vp), P, P]
        L.dh_retina_encode.argtypes = [P, P, P, P, I, I, I, I, I, c_ip, I, c_fp, F, I, ctypes.POINTER(c_vp), P, P]
        L.dh_centernet_encode.argtypes = [P, P, P, P, I, I, I, I, I, I, c_fp, F, I, I, P, P, P]
        PP = ctypes.POINTER(c_vp)
        L.dh_dense_loss.argtypes = [P, I, PP, PP, PP, c_ip, c_ip, c_ip, I, I, I, I, I, I, I, F, F, F, P, P, P]
        L.dh_fcos_encode_loss.argtypes = [P, P, P, P, I, I, I, I, I, c_ip, c_fp, I, I, PP, I, I, F, F, F, P, P, P, P]
        L.dh_retina_encode_loss.argtypes = [P, P, P, P, I, I, I, I, I, c_ip, I, c_fp, F, I, PP, F, F, F, P, P, P, P]
        L.dh_centernet_encode_loss.argtypes = [P, P, P, P, I, I, I, I, I, I, c_fp, F, I, I, P, I, I, F, F, F, P, P, P, P]
        L.dh_dense_loss_grad.argtypes = [P, I, PP, PP, PP, c_ip, c_ip, c_ip, I, I, I, I, I, I, I, F, F, F, F, F, F, PP, P, P, P]
        L.dh_fcos_encode_loss_grad.argtypes = [P, P, P, P, I, I, I, I, I, c_ip, c_fp, I, I, PP, I, I, F, F, F, F, F, F, PP, P, P, P, P]
        L.dh_retina_encode_loss_grad.argtypes = [P, P, P, P, I, I, I, I, I, c_ip, I, c_fp, F, I, PP, F, F, F, F, F, PP, P, P, P, P]
        L.dh_centernet_encode_loss_grad.argtypes = [P, P, P, P, I, I, I, I, I, I, c_fp, F, I, I, P, I, I, F, F, F, F, F, F, P, P, P, P, P]
        L.dh_prediction_to_corners.argtypes = [P, P, I, I, I, I, I, I, F, F, F, c_fp, P, P]
        L.dh_fcos_decode.argtypes = [P, PP, I, I, I, I, c_ip, I, I, P, P, P]
        L.dh_retina_decode.argtypes = [P, PP, I, I, I, I, c_ip, I, P, I, P, P]
        L.dh_select_topk.argtypes = [P, P, I, ctypes.c_longlong, I, I, P, I, I, F, I, P, P, P]
        L.dh_nms.argtypes = [P, P, P, I, I, I, I, F, F, I, I, I, I, P, I, P, P]
        D = ctypes.c_double
        L.dh_fcos_detect.argtypes = [P, PP, I, I, I, I, c_ip, I, I, F, F, I, I, I, P, P, P, P, P, P]
        L.dh_retina_detect.argtypes = [P, PP, I, I, I, I, c_ip, I, P, I, F, F, I, P, I, P, P, P, P, P]
        L.dh_box_convert.argtypes = [P, P, ctypes.c_longlong, I, P, P]
        L.dh_prepare_labels.argtypes = [P, P, P, P, P, P, I, I, I, P, P, P]
        L.dh_format_detections.argtypes = [P, P, P, P, I, I, P, P, P, P]
        L.dh_plan_fused_chunks.argtypes = [I, I, I, I, I, I, I, I, P]
        L.dh_fcos_rectangles.argtypes = [P, P, P, P, I, I, P, P]
        L.dh_fcos_sparse_encode.argtypes = [P, P, P, P, I, I, I, I, I, ctypes.c_longlong, P, P, P, P]
        L.dh_compute_iou.argtypes = [P, P, I, P, I, P, P]
        L.dh_bboxes_iou.argtypes = [P, P, I, P, I, P, P]
        L.dh_centernet_nms.argtypes = [P, P, I, P, I, D, D, I, P, P, P, P]
        L.dh_get_status.argtypes = [P, c_ip, I]
        L.dh_set_trace.argtypes = [P, P, ctypes.c_longlong]
        L.dh_comm_get_unique_id.argtypes = [P]
        L.dh_comm_init_rank.argtypes = [P, I, I, P]
        L.dh_comm_init_all.argtypes = [PP, I]
        L.dh_comm_peer_export.argtypes = [P, P]
        L.dh_comm_peer_import.argtypes = [P, I, I, P]
        L.dh_comm_info.argtypes = [P, c_ip, c_ip, c_ip]
        L.dh_allreduce_loss.argtypes = [P, P, I, P]
        L.dh_comm_destroy.argtypes = [P]
        for name, proto in _OPTIONAL.items():
            if hasattr(L, name):
                getattr(L, name).argtypes = proto
        _lib = L
    return _lib


# entry points added after the first milestone; bound when present
_OPTIONAL = {}


def check(rc, what=""):
    if rc == DH_OK:
        return
    msg = lib().dh_last_error().decode("utf-8", "replace")
    exc = {DH_ERR_BAD_ARG: ValueError, DH_ERR_SHAPE: ValueError, DH_ERR_CAPACITY: ValueError}.get(rc, DenseHeadError)
    raise exc("%s failed (%d): %s" % (what or "densehead call", rc, msg))


def handle(device_index):
    """One library handle per CUDA device per process."""
    h = _handles.get(device_index)
    if h is None:
        with _lock:
            h = _handles.get(device_index)
            if h is None:
                out = c_vp()
                check(lib().dh_create(ctypes.byref(out), int(device_index)), "dh_create")
                h = _handles[device_index] = out
    return h


def status(device_index=0, reset=True):
    """Sticky validation bits of the device's handle (DH_STATUS_*); synchronises."""
    out = ctypes.c_int32(0)
    check(lib().dh_get_status(handle(device_index), ctypes.byref(out), 1 if reset else 0), "dh_get_status")
    return int(out.value)


def raise_for_status(device_index=0):
    """Raise what the reference raises for the inputs the kernels flagged: IndexError for a GT class outside
    [0, num_classes) (FCOS/fcos.py:281-283), ValueError for a CenterNet box not below the largest box scale."""
    bits = status(device_index)
    if bits & DH_STATUS_BAD_CLASS:
        raise IndexError("a ground-truth class index is out of bounds for num_classes")
    if bits & DH_STATUS_BAD_SCALE:
        raise ValueError("min() arg is an empty sequence (a box is not below the largest box scale)")
    if bits & DH_STATUS_COMM_TIMEOUT:
        raise DenseHeadError("a peer rank did not arrive at the loss all-reduce")
    if bits & DH_STATUS_TRUNCATED:
        raise ValueError("prepare_labels dropped boxes: an image has more boxes than max_boxes")


def set_option(device_index, option, value):
    check(lib().dh_set_option(handle(device_index), option, int(value)), "dh_set_option")


def launch_count(device_index=0):
    return int(lib().dh_launch_count(handle(device_index)))


def phase_timing(device_index=0):
    out = (ctypes.c_longlong * 8)()
    check(lib().dh_read_phase_timing(handle(device_index), out), "dh_read_phase_timing")
    return list(out)


def int_array(values):
    return (ctypes.c_int32 * len(values))(*[int(v) for v in values])


def float_array(values):
    return (ctypes.c_float * len(values))(*[float(v) for v in values])


def ptr_array(ptrs):
    return (c_vp * len(ptrs))(*[c_vp(int(p)) for p in ptrs])
