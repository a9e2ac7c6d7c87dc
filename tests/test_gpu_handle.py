"""gpfx_create / gpfx_destroy: a created handle owns its own pool of side streams, tuning
override and error text; results through it equal the default handle's bit for bit, two host threads
can drive two handles at once, and destroying a handle leaves the default one usable."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

from _util import assert_close
from test_gpu_layer import PARAMS, _run_gpu, _run_oracle

pytestmark = pytest.mark.gpu


def _case(seed, N=700, M=96, D=4, P=3, K=3):
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(seed)
    spec = ot.make_spec(rng, M, D, P, K=K, kernel="se", mean="identity" if P == D else "zero", whiten=True)
    X = torch.from_numpy(rng.standard_normal((N, D)))
    eps = torch.from_numpy(rng.standard_normal((N, P)))
    w = [torch.from_numpy(rng.standard_normal((N, P))) for _ in range(3)]
    return spec, X, eps, w


def test_layer_through_a_created_handle_matches_default_and_oracle():
    from gpflux_b200 import _cabi

    lib = _cabi.load()
    spec, X, eps, (w_f, w_m, w_v) = _case(11)
    base = _run_gpu(spec, X, eps, w_f, w_m, w_v, 0.4)
    default = lib.gpfx_get_current()
    h = C.c_void_p()
    _cabi.check(lib.gpfx_create(0, C.byref(h)))
    assert h.value and h.value != default
    try:
        _cabi.check(lib.gpfx_set_current(h))
        assert lib.gpfx_get_current() == h.value
        got = _run_gpu(spec, X, eps, w_f, w_m, w_v, 0.4)
    finally:
        _cabi.check(lib.gpfx_set_current(None))
    assert lib.gpfx_get_current() == default
    for a, b in zip(got[:4], base[:4]):
        assert torch.equal(a.cpu(), b.cpu())
    for k in PARAMS + ("X",):
        assert torch.equal(got[4][k].cpu(), base[4][k].cpu())
    ref = _run_oracle(spec, X, eps, w_f, w_m, w_v, 0.4)
    assert_close(got[1], ref[1], scale=1.0, what="fmean")
    assert_close(got[4]["q_sqrt"], ref[4]["q_sqrt"], what="grad q_sqrt")
    _cabi.check(lib.gpfx_destroy(h))
    again = _run_gpu(spec, X, eps, w_f, w_m, w_v, 0.4)  # default handle still works
    assert torch.equal(again[1].cpu(), base[1].cpu())


def test_errors_and_tuning_are_per_handle():
    from gpflux_b200 import _cabi

    lib = _cabi.load()
    h = C.c_void_p()
    _cabi.check(lib.gpfx_create(0, C.byref(h)))
    try:
        lib.gpfx_set_current(h)
        lib.gpfx_tune_gemm_config(1)  # only this handle is forced onto the 64x64 tiles
        st = lib.gpfx_gemm(1, -5, 4, 4, 1.0, None, 4, 0, 0, None, 4, 0, 0, 0.0, None, 4, 0, 0, 0, None)
        assert st == 0  # an empty product is a no-op, not an error
        bad = C.c_void_p()
        assert lib.gpfx_create(12345, C.byref(bad)) != 0
        assert b"device" in lib.gpfx_handle_last_error(h)
    finally:
        lib.gpfx_set_current(None)
    assert b"12345" not in lib.gpfx_handle_last_error(None)  # the default handle did not see that error
    _cabi.check(lib.gpfx_destroy(h))


def test_two_threads_two_handles():
    from gpflux_b200 import _cabi

    lib = _cabi.load()
    cases = [_case(21), _case(22, N=500, M=64, D=3, P=2, K=1)]
    want = [_run_gpu(s, X, e, *w, 0.3) for s, X, e, w in cases]
    out, errs = [None, None], []

    def work(i):
        try:
            torch.cuda.set_device(0)
            h = C.c_void_p()
            _cabi.check(lib.gpfx_create(0, C.byref(h)))
            lib.gpfx_set_current(h)
            with torch.cuda.stream(torch.cuda.Stream()):
                s, X, e, w = cases[i]
                for _ in range(3):
                    out[i] = _run_gpu(s, X, e, *w, 0.3)
                torch.cuda.synchronize()
            lib.gpfx_set_current(None)
            _cabi.check(lib.gpfx_destroy(h))
        except Exception as exc:  # surfaced in the main thread
            errs.append(exc)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(2):
        assert torch.equal(out[i][1].cpu(), want[i][1].cpu())
        assert torch.equal(out[i][4]["Z"].cpu(), want[i][4]["Z"].cpu())
