"""Maximum-likelihood path: the EKF negative log-likelihood as a differentiable objective.

What the reference does: ``obj_func(theta) = ekf(m_and_cov, H, Xi, m0, P0, dt, ys)[-1][-1]`` differentiated with
``jax.grad`` through ``lax.scan`` and minimised with L-BFGS-B (/root/reference/demos/ekfs_mle.py:42-49,
tetralith/jobs/ekfs_mle.py:41-48).  Here the objective is one CUDA kernel without per-step outputs and the gradient is a
hand-written reverse-mode adjoint kernel (csrc/cgp_nll.cu) wrapped in ``torch.autograd.Function`` -- the torch analogue
of the ``jax.custom_vjp`` wrapper (chirpgp_b200/jax_ffi.py holds the JAX binding of the same C ABI).  The kernels
differentiate w.r.t. the derived model constants, ``m0`` and ``P0``; the small map ``theta -> g(theta) -> constants``
(models.py:50, :56-73, :295-309, :453-456) stays in host autograd, which also reproduces the exact ``lam == 0`` branch.

Batching: ``ys`` (T,) or (B, T); parameters shared or per chirp.  ``candidates=True`` evaluates every chirp against
every parameter set (hyper-parameter grid): ys (B, T) x params (G, ...) -> nll (B, G).
"""
import ctypes as C
from typing import Callable, Optional

import numpy as np
import torch

from . import _native as N
from .filters_smoothers import _device, _problem, _ptr, _h_unit_index
from .models import LCDModel, NC_LCD

__all__ = ['ekf_nll', 'ekf_nll_path', 'filter_nll', 'fit_mle']

_F64 = torch.float64


class _EkfNll(torch.autograd.Function):
    """consts [B, NC_LCD], m0 [B, d], P0 [B, d, d], Xi (0-d tensor) -> nll [B]  (all CUDA float64, contiguous)."""

    @staticmethod
    def forward(ctx, consts, m0, P0, Xi, ys, H, dt, nh, ys_repeat, h_unit, ckpt_every):
        L = N.lib()
        dev = consts.device
        B, d = m0.shape
        T = ys.shape[-1]
        p = _problem(B, T, N.CGP_MODEL_LCD, d, nh, consts, NC_LCD, m0, d, P0, d * d, H, None, 0, None, float(Xi), dt,
                     ys_repeat, h_unit)
        need_grad = any(ctx.needs_input_grad[:4])
        every = int(ckpt_every or L.cgp_ekf_nll_default_ckpt(T))
        ws, nbytes = None, 0
        if need_grad:
            nbytes = L.cgp_ekf_nll_workspace_bytes(C.byref(p), every)
            ws = torch.empty((nbytes // 8,), dtype=_F64, device=dev)
        nll = torch.empty((B,), dtype=_F64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.cgp_ekf_nll_fwd_f64(C.byref(p), _ptr(ys), _ptr(nll), _ptr(ws), C.c_size_t(nbytes), every, stream)
        N.check(rc, 'ekf_nll')
        ctx.save_for_backward(consts, m0, P0, Xi, ys, H)
        ctx.ws, ctx.meta = ws, (dt, nh, ys_repeat, h_unit, every, nbytes)
        return nll

    @staticmethod
    def backward(ctx, nll_bar):
        consts, m0, P0, Xi, ys, H = ctx.saved_tensors
        dt, nh, ys_repeat, h_unit, every, nbytes = ctx.meta
        L = N.lib()
        dev = consts.device
        B, d = m0.shape
        T = ys.shape[-1]
        p = _problem(B, T, N.CGP_MODEL_LCD, d, nh, consts, NC_LCD, m0, d, P0, d * d, H, None, 0, None, float(Xi), dt,
                     ys_repeat, h_unit)
        cb = torch.empty((B, NC_LCD), dtype=_F64, device=dev)
        mb = torch.empty((B, d), dtype=_F64, device=dev)
        Pb = torch.empty((B, d, d), dtype=_F64, device=dev)
        xb = torch.empty((B,), dtype=_F64, device=dev)
        nb = nll_bar.contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.cgp_ekf_nll_bwd_f64(C.byref(p), _ptr(ys), _ptr(nb), _ptr(ctx.ws), C.c_size_t(nbytes), every, _ptr(cb),
                                   _ptr(mb), _ptr(Pb), _ptr(xb), stream)
        N.check(rc, 'ekf_nll backward')
        return cb, mb, Pb, xb.sum(), None, None, None, None, None, None, None


class _EkfNllPath(torch.autograd.Function):
    """As _EkfNll, but the output is the whole cumulative n_ell [B, T] (filters_smoothers.py:180-184) and the backward pass
    takes an arbitrary cotangent on it: the weight of step k's increment is sum_{j >= k} ct_j."""

    @staticmethod
    def forward(ctx, consts, m0, P0, Xi, ys, H, dt, nh, ys_repeat, h_unit, ckpt_every):
        L = N.lib()
        dev = consts.device
        B, d = m0.shape
        T = ys.shape[-1]
        p = _problem(B, T, N.CGP_MODEL_LCD, d, nh, consts, NC_LCD, m0, d, P0, d * d, H, None, 0, None, float(Xi), dt,
                     ys_repeat, h_unit)
        need_grad = any(ctx.needs_input_grad[:4])
        every = int(ckpt_every or L.cgp_ekf_nll_default_ckpt(T))
        ws, nbytes = None, 0
        if need_grad:
            nbytes = L.cgp_ekf_nll_workspace_bytes(C.byref(p), every)
            ws = torch.empty((nbytes // 8,), dtype=_F64, device=dev)
        nell = torch.empty((B, T), dtype=_F64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.cgp_ekf_nll_path_fwd_f64(C.byref(p), _ptr(ys), _ptr(nell), _ptr(ws), C.c_size_t(nbytes), every, stream)
        N.check(rc, 'ekf_nll_path')
        ctx.save_for_backward(consts, m0, P0, Xi, ys, H)
        ctx.ws, ctx.meta = ws, (dt, nh, ys_repeat, h_unit, every, nbytes)
        return nell

    @staticmethod
    def backward(ctx, nell_bar):
        consts, m0, P0, Xi, ys, H = ctx.saved_tensors
        dt, nh, ys_repeat, h_unit, every, nbytes = ctx.meta
        L = N.lib()
        dev = consts.device
        B, d = m0.shape
        T = ys.shape[-1]
        p = _problem(B, T, N.CGP_MODEL_LCD, d, nh, consts, NC_LCD, m0, d, P0, d * d, H, None, 0, None, float(Xi), dt,
                     ys_repeat, h_unit)
        cb = torch.empty((B, NC_LCD), dtype=_F64, device=dev)
        mb = torch.empty((B, d), dtype=_F64, device=dev)
        Pb = torch.empty((B, d, d), dtype=_F64, device=dev)
        xb = torch.empty((B,), dtype=_F64, device=dev)
        w = torch.flip(torch.cumsum(torch.flip(nell_bar, dims=(-1,)), dim=-1), dims=(-1,)).contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.cgp_ekf_nll_path_bwd_f64(C.byref(p), _ptr(ys), _ptr(w), _ptr(ctx.ws), C.c_size_t(nbytes), every, _ptr(cb),
                                        _ptr(mb), _ptr(Pb), _ptr(xb), stream)
        N.check(rc, 'ekf_nll_path backward')
        return cb, mb, Pb, xb.sum(), None, None, None, None, None, None, None


def _as_dev(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=_F64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)


def ekf_nll(cond_m_cov: LCDModel, H, Xi, m0, P0, dt, ys, candidates: bool = False,
            ckpt_every: Optional[int] = None, path: bool = False) -> torch.Tensor:
    """Final cumulative negative log-likelihood of the EKF == ``ekf(cond_m_cov, H, Xi, m0, P0, dt, ys)[-1][-1]``
    (filters_smoothers.py:222-264), differentiable w.r.t. the model hyper-parameters, ``m0``, ``P0`` and ``Xi``.

    Returns a CUDA tensor of shape () for one problem, (B,) for a batch, (B, G) with ``candidates=True``.
    ``path=True`` returns the whole third output of ``ekf`` instead -- the cumulative n_ell at every step, shape (..., T) --
    and accepts any cotangent on it (``ekf_nll_path``)."""
    if not isinstance(cond_m_cov, LCDModel):
        raise NotImplementedError('ekf_nll: the adjoint kernel is compiled for the chirp-family LCD models only')
    dev = _device()
    dt = float(dt)
    d, nh = cond_m_cov.d, cond_m_cov.num_harmonics
    ys_t = _as_dev(ys, dev)
    lead = ys_t.shape[:-1]
    ys2 = ys_t.reshape(-1, ys_t.shape[-1]).contiguous()
    Bc = ys2.shape[0]
    consts = _as_dev(cond_m_cov.consts(dt), dev)
    m0_t, P0_t = _as_dev(m0, dev), _as_dev(P0, dev)
    Xi_t = Xi.to(device=dev, dtype=_F64) if isinstance(Xi, torch.Tensor) else torch.tensor(float(Xi), dtype=_F64, device=dev)
    pb = [int(np.prod(t.shape[:t.dim() - k])) for t, k in ((consts, 1), (m0_t, 1), (P0_t, 2))]
    G = max(pb)
    if any(n not in (1, G) for n in pb):
        raise ValueError('inconsistent parameter batch sizes %s' % pb)
    if candidates:
        B, ys_repeat, out_shape = Bc * G, G, tuple(lead) + (G,)
        expand = lambda t, core: t.reshape((1, -1) + core).expand((Bc, G) + core).reshape((B,) + core)
    else:
        if G > 1 and Bc == 1:
            B, ys_repeat, out_shape = G, G, (G,)
        elif G in (1, Bc):
            B, ys_repeat, out_shape = Bc, 1, tuple(lead)
        else:
            raise ValueError('batched parameters (%d) do not match the %d chirps (use candidates=True for a grid)' % (G, Bc))
        expand = lambda t, core: t.reshape((-1,) + core).expand((B,) + core)
    consts_b = expand(consts, (NC_LCD,)).contiguous()
    m0_b = expand(m0_t, (d,)).contiguous()
    P0_b = expand(P0_t, (d, d)).contiguous()
    h_unit = _h_unit_index(H)
    H_t = _as_dev(H, dev).reshape(-1).contiguous()
    if path:
        nell = _EkfNllPath.apply(consts_b, m0_b, P0_b, Xi_t, ys2, H_t, dt, nh, ys_repeat, h_unit, ckpt_every)
        return nell.reshape(out_shape + (ys2.shape[-1],))
    nll = _EkfNll.apply(consts_b, m0_b, P0_b, Xi_t, ys2, H_t, dt, nh, ys_repeat, h_unit, ckpt_every)
    return nll.reshape(out_shape)


def ekf_nll_path(cond_m_cov: LCDModel, H, Xi, m0, P0, dt, ys, candidates: bool = False,
                 ckpt_every: Optional[int] = None) -> torch.Tensor:
    """The cumulative negative log-likelihood at every step == ``ekf(...)[-1]`` (filters_smoothers.py:180-184, :263-264), shape
    (..., T), differentiable with ANY cotangent (what ``jax.vjp`` of the reference's third output accepts): the adjoint
    kernel weights the increment of step k by the sum of the cotangents of steps >= k."""
    return ekf_nll(cond_m_cov, H, Xi, m0, P0, dt, ys, candidates=candidates, ckpt_every=ckpt_every, path=True)


def filter_nll(method: str, model_args: tuple, H, Xi, m0, P0, dt, ys, sgps=None) -> torch.Tensor:
    """Final cumulative nll of ANY of the five filters, evaluated by its kernel in nll-only mode (nothing but one double
    per problem is stored): ``method`` in {'kf', 'ekf', 'ekf_for_kpt', 'sgp_filter', 'cd_ekf', 'cd_sgp_filter'};
    ``model_args`` are the leading model arguments of that filter (e.g. ``(m_and_cov,)``, ``(drift, dispersion)`` or
    ``(F, Sigma, h)``).  Not differentiable by
    autograd -- ``fit_mle`` differentiates it by fourth-order central differences over a candidate batch (one launch)."""
    from . import filters_smoothers as fs
    dt = float(dt)
    if method in ('kf', 'ekf', 'sgp_filter'):
        model = fs._disc_model(model_args[0], int(m0.shape[-1]), dt) if method != 'kf' else model_args[0]
        consts = fs._consts_on_device(model, dt, fs._device(), *((dt,) if method != 'kf' else ()))
        return fs._run_filter(method, model, consts, H, Xi, m0, P0, dt, ys, sgps=sgps, store=False, last_only=True)
    if method == 'ekf_for_kpt':
        F, Sigma, h = model_args
        lin = fs.LinearDisc(F, Sigma)
        model = fs._KPTModel(lin, h.num_harmonics)
        return fs._run_filter(method, model, fs._dev(lin.consts(), fs._device()), torch.zeros(int(lin.d), dtype=_F64), Xi, m0,
                              P0, dt, ys, store=False, last_only=True)
    if method in ('cd_ekf', 'cd_sgp_filter'):
        model = fs._sde_model(model_args[0], int(m0.shape[-1]))
        b = model_args[1]
        Qc = fs._qc(fs._dispersion_matrix(b, m0) if method == 'cd_ekf' else b)
        return fs._run_filter(method, model, model.consts(), H, Xi, m0, P0, dt, ys, sgps=sgps, Qc=Qc, store=False,
                              last_only=True)
    raise ValueError('unknown filter %r' % method)


def fit_mle(build_model: Callable, init_theta, H, Xi, dt, ys, transform: Optional[Callable] = None, maxiter: int = 200,
            reduce_group=None, method: str = 'ekf', sgps=None, fd_step: Optional[float] = None):
    """L-BFGS-B maximum-likelihood fit driving the nll kernels -- the role of
    ``jaxopt.ScipyMinimize(method='L-BFGS-B', fun=obj_func).run(init_theta)`` (demos/ekfs_mle.py:48-49 and the other
    ``*_mle.py`` demos / tetralith jobs).

    build_model(params) -> (drift, dispersion, m_and_cov, m0, P0, H') as chirpgp_b200.models.build_* (for
    method='ekf_for_kpt': (F, Sigma, m0, P0, h) as build_kpt_chirp_model with fs bound, and ``H`` is ignored); ``transform`` maps
    the unconstrained theta to params (default: the reference's softplus ``g``).  The objective is the SUM of the nll over
    all chirps in ``ys`` (this rank's shard when torch.distributed is initialised: the scalar objective and its 6-vector
    gradient are then summed over ranks with one all-reduce -- the only collective on this path).

    method='ekf' uses the hand-written adjoint kernel.  The other filters ('sgp_filter' with ``sgps``, 'cd_ekf',
    'cd_sgp_filter') have no adjoint kernel: their gradient is taken by fourth-order central differences, with all
    4P + 1 perturbed parameter sets evaluated as ONE candidate batch against the shared signals (nll-only kernels;
    agreement with jax.grad ~1e-8 relative, tests/test_gpu_mle.py).

    Returns (theta_opt (numpy), scipy OptimizeResult); ``result.success`` follows the reference's convention
    (tetralith/jobs/ekfs_mle.py:49, :75-78: a failed fit is reported, not raised)."""
    import scipy.optimize
    from .distributed import allreduce_objective
    from .models import g as _g
    transform = transform or _g
    dev = _device()
    ys_t = _as_dev(ys, dev)
    ys2 = ys_t.reshape(-1, ys_t.shape[-1])

    def fun_adjoint(theta_np):
        theta = torch.tensor(theta_np, dtype=_F64, device=dev, requires_grad=True)
        _, _, m_and_cov, m0, P0, _ = build_model(transform(theta))
        val = ekf_nll(m_and_cov, H, Xi, m0, P0, dt, ys_t).sum()
        grad, = torch.autograd.grad(val, theta)
        v, gr = allreduce_objective(val.detach(), grad, group=reduce_group)
        return float(v.cpu()), gr.cpu().numpy().copy()

    def fun_fd(theta_np):
        # five-point central differences: (-f(+2h) + 8 f(+h) - 8 f(-h) + f(-2h)) / 12h, all 4P + 1 candidates in ONE batch.
        # On this hardware the candidates cost nothing extra (a single chirp leaves the GPU empty and the candidates run
        # side by side), and truncation (h^4) and round-off (eps |f| / h) both sit near 1e-8 relative or below.
        n = theta_np.shape[0]
        # step: the sigma-point objectives carry ~1e-12 relative summation noise (tests/test_noise_floor.py) and want the larger
        # step; the EKF-type objectives are smooth to ~1e-15 and take the smaller one (truncation ~ h^4)
        step = fd_step if fd_step is not None else (2e-3 if method in ('sgp_filter', 'cd_sgp_filter') else 1e-4)
        h = step * np.maximum(1., np.abs(theta_np))
        cand = np.tile(theta_np, (4 * n + 1, 1))
        for i in range(n):
            for j, mult in enumerate((1., -1., 2., -2.)):
                cand[1 + 4 * i + j, i] += mult * h[i]
        built = build_model(transform(torch.as_tensor(cand)))
        if method == 'ekf_for_kpt':                       # build_kpt_chirp_model: (F, Sigma, m0, P0, h)
            F_, Sigma_, m0, P0, h_ = built
            margs = (F_, Sigma_, h_)
        else:
            drift, dispersion, m_and_cov, m0, P0, _ = built
            if method in ('ekf', 'sgp_filter'):
                margs = (m_and_cov,)
            else:
                margs = (drift, dispersion if method == 'cd_ekf' else dispersion.matrix())
        total = torch.zeros(4 * n + 1, dtype=_F64, device=dev)
        for row in ys2:                                   # every chirp against the 4P + 1 candidates
            total = total + filter_nll(method, margs, H, Xi, m0, P0, dt, row, sgps=sgps)
        grad = (8. * (total[1::4] - total[2::4]) - (total[3::4] - total[4::4])) / torch.as_tensor(12 * h, device=dev)
        v, gr = allreduce_objective(total[0], grad, group=reduce_group)
        return float(v.cpu()), gr.cpu().numpy().copy()

    fun = fun_adjoint if method == 'ekf' else fun_fd
    res = scipy.optimize.minimize(fun, np.asarray(init_theta, dtype=np.float64), jac=True, method='L-BFGS-B',
                                  options={'maxiter': maxiter})
    return res.x, res
