"""Accuracy of the library's internal FP64 elementary functions (csrc/cgp_math.cuh) against extended precision."""
import ctypes as C

import numpy as np
import pytest
import torch

from chirpgp_b200 import _native

pytestmark = pytest.mark.gpu


def _eval(kind, x):
    L = _native.lib()
    xd = torch.as_tensor(x, dtype=torch.float64, device='cuda')
    out = torch.empty_like(xd)
    rc = L.cgp_test_math(kind, xd.numel(), C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()),
                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    return out.cpu().numpy()


def _ulp_err(got, want_ld):
    want = want_ld.astype(np.float64)
    ulp = np.spacing(np.abs(want))
    return np.max(np.abs((got.astype(np.longdouble) - want_ld) / ulp.astype(np.longdouble)))


def test_fast_math_ulp():
    rng = np.random.default_rng(0)
    ld = np.longdouble
    # exp on the fast range and beyond (fallback)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-40, 40, 200000), rng.uniform(-1, 1, 100000)])
    assert _ulp_err(_eval(0, x), np.exp(x.astype(ld))) < 2.5      # only used inside softplus (u = exp(-x) << x)
    assert np.isinf(_eval(0, np.array([710., 800.]))).all() and (_eval(0, np.array([-800.])) == 0).all()
    # softplus: series branch (x >= 3) against the true value; general branch against the reference's NAIVE formula
    # log(exp(x) + 1) in float64 (models.py:50 -- inaccurate for very negative x by construction, and we mirror it);
    # overflow like the naive form
    x = np.concatenate([rng.uniform(3, 60, 300000), rng.uniform(60, 700, 50000)])
    want = np.log1p(np.exp(-x.astype(ld))) + x.astype(ld)
    assert _ulp_err(_eval(1, x), want) < 2.0
    assert _ulp_err(_eval(7, x), want) < 2.0
    xn = rng.uniform(-30, 3, 100000)
    assert np.max(np.abs(_eval(1, xn) - np.log(np.exp(xn) + 1.))) < 1e-15
    assert np.isinf(_eval(1, np.array([710., 1000.]))).all()
    assert np.isnan(_eval(1, np.array([np.nan]))).all()
    sg = 1 / (1 + np.exp(-x.astype(ld)))
    assert _ulp_err(_eval(6, x), sg) < 2.5
    # sin / cos: small angles (the chirp models' dt * w), moderate, large (fallback)
    x = np.concatenate([rng.uniform(-0.5, 0.5, 200000), rng.uniform(-100, 100, 200000), rng.uniform(-1e5, 1e5, 100000),
                        rng.uniform(-1e9, 1e9, 1000)])
    # reference in extended precision is only good to ~1e-19 * |x|: restrict the ulp check accordingly
    assert _ulp_err(_eval(2, x[:400000]), np.sin(x[:400000].astype(ld))) < 2.0
    assert _ulp_err(_eval(3, x[:400000]), np.cos(x[:400000].astype(ld))) < 2.0
    big = x[400000:]
    assert np.max(np.abs(_eval(2, big) - np.sin(big))) < 1e-15 and np.max(np.abs(_eval(3, big) - np.cos(big))) < 1e-15
    # rsqrt / rcp
    x = np.exp(rng.uniform(-50, 50, 300000))
    assert _ulp_err(_eval(4, x), 1 / np.sqrt(x.astype(ld))) < 2.0
    assert _ulp_err(_eval(5, x), 1 / x.astype(ld)) < 2.0
    assert np.isnan(_eval(4, np.array([-1.]))).all()
