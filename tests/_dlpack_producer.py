"""A DLPack producer that is not torch: builds the DLManagedTensor by hand (ctypes) over a raw device
address - in the GPU test the address comes from cuda-python's cudaMalloc - and wraps it in a "dltensor"
PyCapsule exactly as the DLPack protocol prescribes."""
import ctypes as C

from gpflux_b200.dlpack import _DLManagedTensor

_keep = []  # structs / shape arrays / deleter thunks must outlive the capsules

_DELETER = C.CFUNCTYPE(None, C.c_void_p)
C.pythonapi.PyCapsule_New.restype = C.py_object
C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]


def make_capsule(ptr, shape, strides=None, code=2, bits=64, lanes=1, device_type=2, device_id=0, byte_offset=0,
                 on_delete=None):
    m = _DLManagedTensor()
    nd = len(shape)
    shp = (C.c_int64 * nd)(*shape)
    m.dl_tensor.data = ptr
    m.dl_tensor.device.device_type = device_type
    m.dl_tensor.device.device_id = device_id
    m.dl_tensor.ndim = nd
    m.dl_tensor.dtype.code = code
    m.dl_tensor.dtype.bits = bits
    m.dl_tensor.dtype.lanes = lanes
    m.dl_tensor.shape = C.cast(shp, C.POINTER(C.c_int64))
    st = None
    if strides is not None:
        st = (C.c_int64 * nd)(*strides)
        m.dl_tensor.strides = C.cast(st, C.POINTER(C.c_int64))
    m.dl_tensor.byte_offset = byte_offset

    def _del(_p):
        if on_delete is not None:
            on_delete()

    thunk = _DELETER(_del)
    m.deleter = C.cast(thunk, C.c_void_p)
    _keep.append((m, shp, st, thunk))
    return C.pythonapi.PyCapsule_New(C.addressof(m), b"dltensor", None)


class RawArray:
    """Minimal array object over a raw device address with the __dlpack__ protocol."""

    def __init__(self, ptr, shape, **kw):
        self.ptr, self.shape, self.kw = ptr, tuple(shape), kw
        self.deleted = 0

    def _on_delete(self):
        self.deleted += 1

    def __dlpack__(self, stream=None):
        return make_capsule(self.ptr, self.shape, on_delete=self._on_delete, **self.kw)

    def __dlpack_device__(self):
        return (self.kw.get("device_type", 2), self.kw.get("device_id", 0))
