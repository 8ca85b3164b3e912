"""vgb_math.h (shared by kernels, oracle restatement and the _ref bx shim): accuracy vs float64."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyref import OracleLib


@pytest.fixture(scope="module")
def ev():
    lib = OracleLib().lib
    lib.vgo_math_eval.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]

    def f(op, x, y=None):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y if y is not None else x, np.float32)
        out = np.empty_like(x)
        lib.vgo_math_eval(op, x.ctypes.data, y.ctypes.data, out.ctypes.data, len(x))
        return out
    return f


def ulp_err(got, want64):
    want32 = want64.astype(np.float32)
    ulp = np.spacing(np.abs(want32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - want64) / np.maximum(ulp, 1e-45)


def test_sin_cos(ev):
    x = np.linspace(-20.0, 20.0, 400001).astype(np.float32)
    for op, fn in ((0, np.sin), (1, np.cos)):
        got = ev(op, x)
        err = np.abs(got.astype(np.float64) - fn(x.astype(np.float64)))
        assert err.max() < 1.5e-7, (op, err.max())


def test_acos(ev):
    x = np.linspace(-1.0, 1.0, 400001).astype(np.float32)
    got = ev(2, x)
    assert ulp_err(got, np.arccos(x.astype(np.float64))).max() < 2.5
    assert ev(2, np.array([1.0, -1.0, 2.0, -3.0], np.float32)).tolist() == [0.0, np.float32(np.pi), 0.0, np.float32(np.pi)]


def test_atan2(ev):
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(300000) * 10).astype(np.float32)
    x = (rng.standard_normal(300000) * 10).astype(np.float32)
    got = ev(3, x, y)
    assert np.abs(got.astype(np.float64) - np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() < 4e-7
    # axes and the defined value at the origin
    assert ev(3, np.array([0, 0, 1, -1, 0], np.float32), np.array([0, 1, 0, 0, -1], np.float32)).tolist() == \
        [0.0, np.float32(np.pi / 2), 0.0, np.float32(np.pi), -np.float32(np.pi / 2)]


def test_tan(ev):
    x = np.linspace(0.01, 1.5, 100001).astype(np.float32)
    got = ev(4, x)
    assert (np.abs(got.astype(np.float64) - np.tan(x.astype(np.float64))) / np.tan(x.astype(np.float64))).max() < 1e-6
