"""DLPack validation of the boundary (SURVEY §8b ownership row): device, dtype, strides and alignment are
checked on the DLManagedTensor itself before any pointer can reach the C ABI.  No GPU needed: the checks
read the descriptor only."""
import numpy as np
import pytest

from _dlpack_producer import RawArray
from gpflux_b200 import dlpack as dl
from gpflux_b200._cabi import GpfxError

ADDR = 0x7F0000000000  # a plausible, 16-byte aligned device address (never dereferenced here)


def test_descriptor_of_a_hand_built_capsule():
    cap = RawArray(ADDR, (5, 3)).__dlpack__()
    d = dl.describe(cap)
    assert d.data == ADDR and d.shape == (5, 3) and d.strides is None
    assert (d.device_type, d.code, d.bits, d.lanes) == (dl.kDLCUDA, dl.kDLFloat, 64, 1)
    dl.validate(d, ndim=2)
    dl.validate(dl.describe(RawArray(ADDR, (5, 3), strides=(3, 1)).__dlpack__()))  # canonical strides are fine
    dl.validate(dl.describe(RawArray(ADDR, (1, 3), strides=(999, 1)).__dlpack__()))  # stride of a length-1 dim is free


@pytest.mark.parametrize("kw,status", [
    (dict(bits=32), dl.GPFX_E_DTYPE),                      # float32 without the fp32 opt-in
    (dict(code=dl.kDLInt), dl.GPFX_E_DTYPE),
    (dict(lanes=2), dl.GPFX_E_DTYPE),
    (dict(strides=(1, 5)), dl.GPFX_E_LAYOUT),              # a transposed view
    (dict(strides=(6, 2)), dl.GPFX_E_LAYOUT),              # every other column
    (dict(byte_offset=8), dl.GPFX_E_LAYOUT),               # not 16-byte aligned
    (dict(device_type=dl.kDLCPU), dl.GPFX_E_LAYOUT),       # host memory: there is no CPU path
    (dict(device_type=dl.kDLCUDAHost), dl.GPFX_E_LAYOUT),
])
def test_violations_carry_the_c_abi_status(kw, status):
    with pytest.raises(dl.DLPackError) as e:
        dl.validate(dl.describe(RawArray(ADDR, (5, 3), **kw).__dlpack__()), what="inputs")
    assert e.value.status == status
    assert isinstance(e.value, ValueError) and isinstance(e.value, GpfxError)  # reference raises ValueError on dtype


def test_fp32_opt_in_and_rank():
    d = dl.describe(RawArray(ADDR, (5, 3), bits=32).__dlpack__())
    dl.validate(d, bits=32)
    with pytest.raises(dl.DLPackError) as e:
        dl.validate(dl.describe(RawArray(ADDR, (5,)).__dlpack__()), ndim=2)
    assert e.value.status == dl.GPFX_E_LAYOUT


def test_numpy_host_array_is_refused():
    x = np.zeros((4, 2))
    with pytest.raises(dl.DLPackError) as e:
        dl.from_dlpack(x, what="inputs")
    assert e.value.status == dl.GPFX_E_LAYOUT and "kDLCUDA" in str(e.value)
    with pytest.raises(dl.DLPackError):
        dl.capsule_of(object())
