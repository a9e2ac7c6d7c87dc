"""Zero-copy DLPack ingestion with the checks SURVEY §8b asks of the boundary: a producer hands over a
``DLManagedTensor`` (``__dlpack__()`` / a "dltensor" PyCapsule); before any pointer reaches the C ABI
the tensor must be on a CUDA device (``kDLCUDA``), float64 (float32 only for the fp32 opt-in),
C-contiguous (``strides == NULL`` or canonical) and 16-byte aligned.  Violations raise
``DLPackError`` carrying the C-ABI status (``GPFX_E_DTYPE`` / ``GPFX_E_LAYOUT``) - a ``ValueError``
like the reference's dtype check (gpflux/models/deep_gp.py:137-149).

The capsule is parsed here with ctypes (no torch involved in the validation); ``from_dlpack`` then hands
the same capsule to ``torch.from_dlpack`` so the host layer (which uses torch for memory / streams /
autograd) sees the producer's memory without a copy.  ``device_pointer`` returns the raw device pointer
for callers that talk to the C ABI directly (the TF shim's ``tf.experimental.dlpack`` path does the same
in C++: gpflux_b200/tf_shim/)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Tuple

from ._cabi import GpfxError

GPFX_E_DTYPE = -2
GPFX_E_LAYOUT = -3

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
kDLInt, kDLUInt, kDLFloat, kDLBfloat = 0, 1, 2, 4


class DLPackError(GpfxError, ValueError):
    def __init__(self, status: int, msg: str):
        super().__init__(msg)
        self.status = status


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


class _DLPackVersion(C.Structure):
    _fields_ = [("major", C.c_uint32), ("minor", C.c_uint32)]


class _DLManagedTensorVersioned(C.Structure):
    _fields_ = [("version", _DLPackVersion), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p),
                ("flags", C.c_uint64), ("dl_tensor", _DLTensor)]


_api = C.pythonapi
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]


@dataclass(frozen=True)
class DLDescription:
    data: int  # device address of element 0 (data + byte_offset)
    device_type: int
    device_id: int
    code: int
    bits: int
    lanes: int
    shape: Tuple[int, ...]
    strides: Tuple[int, ...] | None  # in elements; None = compact row-major


def capsule_of(obj):
    """The "dltensor" capsule of ``obj`` (a capsule is passed through)."""
    if type(obj).__name__ == "PyCapsule":
        return obj
    if not hasattr(obj, "__dlpack__"):
        raise DLPackError(GPFX_E_LAYOUT, f"{type(obj).__name__} does not export DLPack (__dlpack__)")
    return obj.__dlpack__()


def describe(capsule) -> DLDescription:
    """Reads the DLTensor of an unconsumed capsule without taking ownership."""
    if _api.PyCapsule_IsValid(capsule, b"dltensor"):
        ptr = _api.PyCapsule_GetPointer(capsule, b"dltensor")
        t = C.cast(ptr, C.POINTER(_DLManagedTensor)).contents.dl_tensor
    elif _api.PyCapsule_IsValid(capsule, b"dltensor_versioned"):
        ptr = _api.PyCapsule_GetPointer(capsule, b"dltensor_versioned")
        t = C.cast(ptr, C.POINTER(_DLManagedTensorVersioned)).contents.dl_tensor
    else:
        raise DLPackError(GPFX_E_LAYOUT, "not an unconsumed DLPack capsule (\"dltensor\")")
    nd = int(t.ndim)
    shape = tuple(int(t.shape[i]) for i in range(nd))
    strides = tuple(int(t.strides[i]) for i in range(nd)) if t.strides else None
    return DLDescription(int(t.data or 0) + int(t.byte_offset), int(t.device.device_type), int(t.device.device_id),
                         int(t.dtype.code), int(t.dtype.bits), int(t.dtype.lanes), shape, strides)


def validate(d: DLDescription, *, bits: int = 64, ndim: int | None = None, what: str = "tensor") -> None:
    if d.device_type not in (kDLCUDA, kDLCUDAManaged):
        raise DLPackError(GPFX_E_LAYOUT, f"{what}: DLPack device_type {d.device_type} is not kDLCUDA - gpflux_b200 has "
                                         "no host path, move the tensor to the GPU first")
    if d.code != kDLFloat or d.bits != bits or d.lanes != 1:
        raise DLPackError(GPFX_E_DTYPE, f"{what} needs to have dtype float{bits} (according to gpflow.default_float()), "
                                        f"however got DLPack dtype (code={d.code}, bits={d.bits}, lanes={d.lanes})")
    if ndim is not None and len(d.shape) != ndim:
        raise DLPackError(GPFX_E_LAYOUT, f"{what}: expected {ndim} dimensions, got shape {d.shape}")
    if d.strides is not None:
        want, acc = [], 1
        for n in reversed(d.shape):
            want.append(acc)
            acc *= max(n, 1)
        want = tuple(reversed(want))
        bad = [i for i, (s, w, n) in enumerate(zip(d.strides, want, d.shape)) if n > 1 and s != w]
        if bad:
            raise DLPackError(GPFX_E_LAYOUT, f"{what}: strides {d.strides} of shape {d.shape} are not C-contiguous "
                                             f"(expected {want}); the C ABI takes dense row-major buffers")
    if d.data % 16 != 0 and all(n > 0 for n in d.shape):
        raise DLPackError(GPFX_E_LAYOUT, f"{what}: device address {d.data:#x} is not 16-byte aligned")


def from_dlpack(obj, *, bits: int = 64, ndim: int | None = None, what: str = "tensor"):
    """Validated zero-copy view of a DLPack producer as a torch tensor (the host layer's tensor type)."""
    import torch

    if isinstance(obj, torch.Tensor):
        return obj
    cap = capsule_of(obj)
    d = describe(cap)
    validate(d, bits=bits, ndim=ndim, what=what)
    t = torch.from_dlpack(cap)
    if t.numel() and t.data_ptr() != d.data:
        raise DLPackError(GPFX_E_LAYOUT, f"{what}: DLPack import was not zero-copy")
    return t


def device_pointer(obj, *, bits: int = 64, ndim: int | None = None, what: str = "tensor"):
    """(address, shape, keep_alive) of a validated DLPack producer, for direct C-ABI calls."""
    cap = capsule_of(obj)
    d = describe(cap)
    validate(d, bits=bits, ndim=ndim, what=what)
    return d.data, d.shape, (obj, cap)
