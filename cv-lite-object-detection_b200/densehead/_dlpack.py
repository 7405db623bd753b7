"""Read a DLPack capsule without any framework: the raw device pointer, shape, dtype and device of the tensor it manages.

This is the zero-copy hand-over the reference side uses (`tf.experimental.dlpack.to_dlpack(t)`, `torch.utils.dlpack.to_dlpack(t)`,
or any object's `__dlpack__()`): the capsule owns a `DLManagedTensor`; we only look inside it.  The capsule must be kept
alive (and not consumed by another framework) for as long as the pointer is in use; `CapsuleView` holds it for that.
"""
import ctypes

_KNOWN_DEVICES = {1: "cpu", 2: "cuda", 3: "cuda_host", 13: "cuda_managed"}
_CODES = {0: "int", 1: "uint", 2: "float", 4: "bfloat", 6: "bool"}


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("device", DLDevice), ("ndim", ctypes.c_int32), ("dtype", DLDataType),
                ("shape", ctypes.POINTER(ctypes.c_int64)), ("strides", ctypes.POINTER(ctypes.c_int64)),
                ("byte_offset", ctypes.c_uint64)]


class DLManagedTensor(ctypes.Structure):
    _fields_ = [("dl_tensor", DLTensor), ("manager_ctx", ctypes.c_void_p), ("deleter", ctypes.c_void_p)]


_get = ctypes.pythonapi.PyCapsule_GetPointer
_get.restype = ctypes.c_void_p
_get.argtypes = [ctypes.py_object, ctypes.c_char_p]
_name = ctypes.pythonapi.PyCapsule_GetName
_name.restype = ctypes.c_char_p
_name.argtypes = [ctypes.py_object]


class CapsuleView:
    """`ptr` (int, already offset by byte_offset), `shape`, `strides` (in elements, or None = contiguous), `dtype`
    (e.g. "float32"), `device` (("cuda", 0)).  Keeps the capsule alive."""

    def __init__(self, capsule):
        name = _name(capsule)
        if name not in (b"dltensor", b"dltensor_versioned"):
            raise ValueError("not an unconsumed DLPack capsule (name %r)" % (name,))
        addr = _get(capsule, name)
        if name == b"dltensor_versioned":
            addr += 8  # DLManagedTensorVersioned: DLPackVersion {major, minor} precedes manager_ctx/deleter/flags/dl_tensor
            raise NotImplementedError("versioned DLPack capsules: ask the producer for the legacy capsule (max_version=None)")
        self.capsule = capsule
        t = ctypes.cast(addr, ctypes.POINTER(DLManagedTensor)).contents.dl_tensor
        self.ptr = int(t.data or 0) + int(t.byte_offset)
        self.shape = tuple(int(t.shape[i]) for i in range(t.ndim))
        self.strides = tuple(int(t.strides[i]) for i in range(t.ndim)) if t.strides else None
        self.dtype = "%s%d" % (_CODES.get(t.dtype.code, "code%d" % t.dtype.code), t.dtype.bits)
        self.device = (_KNOWN_DEVICES.get(t.device.device_type, str(t.device.device_type)), int(t.device.device_id))

    def is_contiguous(self):
        if self.strides is None:
            return True
        expect = 1
        for n, s in zip(reversed(self.shape), reversed(self.strides)):
            if n != 1 and s != expect:
                return False
            expect *= n
        return True


def view(obj):
    """CapsuleView of a DLPack capsule or of any object with `__dlpack__()`."""
    if type(obj).__name__ != "PyCapsule":
        obj = obj.__dlpack__()
    return CapsuleView(obj)
