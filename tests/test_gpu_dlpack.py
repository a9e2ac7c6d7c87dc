"""Zero-copy DLPack ingestion end to end: device memory allocated and filled by cuda-python (not torch),
exported through a hand-built DLManagedTensor, goes into GPLayer / DeepGP without a copy and gives the
result of the torch-tensor path bit for bit; a float32 or strided producer is refused with the C-ABI
status before anything is launched."""
import numpy as np
import pytest
import torch

from _dlpack_producer import RawArray

pytestmark = pytest.mark.gpu


def _device_array(x: np.ndarray):
    from cuda.bindings import runtime as rt

    err, ptr = rt.cudaMalloc(x.nbytes)
    assert int(err) == 0
    (err,) = rt.cudaMemcpy(ptr, x.ctypes.data, x.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice)
    assert int(err) == 0
    return int(ptr)


def _model(D=3, M=32):
    from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp

    rng = np.random.default_rng(5)
    X = rng.standard_normal((400, D))
    model = build_constant_input_dim_deep_gp(X, 2, Config(M, 1e-2, 1e-2), z_init=X[:M])
    for layer in model.f_layers:
        layer.q_mu.assign(0.1 * rng.standard_normal(tuple(layer.q_mu.shape)))
    return model.to("cuda"), X


def test_raw_cuda_memory_through_dlpack_is_zero_copy_and_exact():
    from cuda.bindings import runtime as rt
    from gpflux_b200 import dlpack as dl

    model, X = _model()
    Xb = np.ascontiguousarray(X[:256])
    ptr = _device_array(Xb)
    arr = RawArray(ptr, Xb.shape)
    t = dl.from_dlpack(arr, ndim=2, what="inputs")
    assert t.data_ptr() == ptr and t.is_cuda and tuple(t.shape) == Xb.shape
    assert torch.equal(t.cpu(), torch.from_numpy(Xb))
    addr, shape, _keep = dl.device_pointer(RawArray(ptr, Xb.shape))
    assert addr == ptr and shape == Xb.shape

    torch.manual_seed(3)  # the sample between the two layers is random: same draw for both calls
    want_m, want_v = model.predict_f(torch.from_numpy(Xb).cuda())
    torch.manual_seed(3)
    got_m, got_v = model.predict_f(RawArray(ptr, Xb.shape))  # the producer object itself, no torch tensor
    assert torch.equal(got_m, want_m) and torch.equal(got_v, want_v)
    layer = model.f_layers[0]
    m1, v1 = layer.predict(RawArray(ptr, Xb.shape))
    m2, v2 = layer.predict(torch.from_numpy(Xb).cuda())
    assert torch.equal(m1, m2) and torch.equal(v1, v2)
    del t, got_m, got_v, m1, v1
    torch.cuda.synchronize()
    rt.cudaFree(ptr)


def test_bad_producers_are_refused_before_any_launch():
    from cuda.bindings import runtime as rt
    from gpflux_b200 import _cabi
    from gpflux_b200 import dlpack as dl

    model, X = _model()
    lib = _cabi.load()
    X32 = np.ascontiguousarray(X[:64].astype(np.float32))
    p32 = _device_array(X32)
    Xt = np.ascontiguousarray(X[:64].T)  # [D, N] memory viewed as [N, D] with strides (1, N)
    pt = _device_array(Xt)
    n0 = lib.gpfx_kernel_launch_count()
    with pytest.raises(ValueError) as e:
        model.predict_f(RawArray(p32, X32.shape, bits=32))
    assert e.value.status == dl.GPFX_E_DTYPE
    with pytest.raises(dl.DLPackError) as e:
        model.predict_f(RawArray(pt, (64, X.shape[1]), strides=(1, 64)))
    assert e.value.status == dl.GPFX_E_LAYOUT
    with pytest.raises(dl.DLPackError) as e:
        model.f_layers[0].predict(RawArray(pt + 8, (32, X.shape[1])))
    assert e.value.status == dl.GPFX_E_LAYOUT
    assert lib.gpfx_kernel_launch_count() == n0
    rt.cudaFree(p32)
    rt.cudaFree(pt)
