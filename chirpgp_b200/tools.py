"""Monte-Carlo input generation on the device -- the step *before* the filtering path.

Mirrors the simulators the reference's CRLB jobs and tests are built from (/root/reference/chirpgp/tools.py:81-170
``simulate_lgssm`` / ``simulate_sde``; tetralith/jobs/crlb_ekf.py:41-56; test/test_crlb.py:41-55):

    x_0 = m0 + chol(P0) eps,    x_k = mean(x_{k-1}) + chol(Sigma) eps_k,    y_k = H x_k + sqrt(Xi) eps'_k .

The reference draws ``eps`` with ``jax.random.normal`` from threefry keys; those streams are JAX-specific, so there is no bit
parity on this row.  Here the normals come from the counter-based Philox4x32-10 generator inside the kernel
(csrc/cgp_sim.cu): trajectory ``i`` depends on ``(seed, i)`` only, so sharding a Monte-Carlo batch over GPUs
(``first_trajectory`` = this rank's offset) reproduces the single-GPU samples exactly.  ``key`` arguments of the
reference-named functions are integer seeds."""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import filters_smoothers as fs
from .models import LCDModel, LinearDisc

__all__ = ['simulate', 'simulate_lgssm', 'simulate_sde', 'rmse']

_F64 = torch.float64


def simulate(cond_m_cov, H, Xi, m0, P0, dt, T: int, num_trajectories: int, seed: int, first_trajectory: int = 0,
             states: bool = True):
    """B = ``num_trajectories`` trajectories of a discretised model (tagged ``LCDModel`` / ``LinearDisc``, or a linear Python
    callable) and their measurements.  Returns CUDA tensors ``(x0 (B, d), xs (B, T, d) or None, ys (B, T))``."""
    dev = fs._device()
    dt = float(dt)
    model = fs._disc_model(cond_m_cov, int(np.shape(m0)[-1]), dt)
    if not isinstance(model, (LCDModel, LinearDisc)):
        raise NotImplementedError('simulate: discretised models only')
    model_id, d, nh = fs._model_fields(model)
    consts = fs._consts_on_device(model, dt, dev, *((dt,) if isinstance(model, LCDModel) else ()))
    B = int(num_trajectories)
    bt = fs._Batch()
    consts_t, cs = bt.see(fs._dev(consts, dev), 1, 'model parameters')
    m0_t, m0s = bt.see(fs._dev(m0, dev), 1, 'm0')
    P0_t, P0s = bt.see(fs._dev(P0, dev), 2, 'P0')
    if bt.B not in (None, B):
        raise ValueError('batched parameters (%d) do not match num_trajectories (%d)' % (bt.B, B))
    H_t = fs._dev(H, dev).reshape(-1)
    p = fs._problem(B, int(T), model_id, d, nh, consts_t, cs, m0_t, m0s, P0_t, P0s, H_t, None, 0, None, Xi, dt)
    x0 = torch.empty((B, d), dtype=_F64, device=dev)
    xs = torch.empty((B, int(T), d), dtype=_F64, device=dev) if states else None
    ys = torch.empty((B, int(T)), dtype=_F64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rc = N.lib().cgp_simulate_f64(C.byref(p), C.c_uint64(int(seed)), C.c_uint64(int(first_trajectory)), fs._ptr(x0),
                                  fs._ptr(xs), fs._ptr(ys), stream)
    N.check(rc, 'simulate')
    return x0, xs, ys


def simulate_lgssm(F, Sigma, x0, T: int, key: int):
    """tools.py:81-116: one trajectory (T, d) of x_k = F x_{k-1} + q_k, q_k ~ N(0, Sigma), started at the given x0."""
    d = int(np.shape(x0)[-1])
    _, xs, _ = simulate(LinearDisc(F, Sigma), np.zeros(d), 0., x0, np.zeros((d, d)), 0., T, 1, key)
    return xs[0]


def simulate_sde(m_and_cov, m0, P0, dt, T: int, key: int, const_diag_cov: bool = False):
    """tools.py:119-170: one trajectory (T, d) with Gaussian increments x_k = mean(x_{k-1}) + chol(cov) dw_k, x_0 ~ N(m0, P0).
    ``const_diag_cov`` is accepted for signature parity (the Cholesky factor of a diagonal matrix is its square root)."""
    d = int(np.shape(m0)[-1])
    _, xs, _ = simulate(m_and_cov, np.zeros(d), 0., m0, P0, dt, T, 1, key)
    return xs[0]


def rmse(x1, x2, reduce_sum: bool = True):
    """tools.py:279-293."""
    x1, x2 = torch.as_tensor(x1), torch.as_tensor(x2)
    val = torch.sqrt(torch.mean((x1 - x2) ** 2, dim=0))
    return val.sum() if reduce_sum else val
