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

__all__ = ['ekf_nll', 'ekf_nll_path', 'filter_nll', 'filter_nll_grad', 'fit_mle', 'fit_mle_batched']

_F64 = torch.float64


class _EkfNll(torch.autograd.Function):
    """consts [B, NC_LCD], m0 [B, d], P0 [B, d, d], Xi (0-d tensor) -> nll [B]  (all CUDA float64, contiguous)."""

    @staticmethod
    def forward(ctx, consts, m0, P0, Xi, ys, H, dt, nh, ys_repeat, h_unit, ckpt_every, raw_p0):
        L = N.lib()
        dev = consts.device
        B, d = m0.shape
        T = ys.shape[-1]
        p = _problem(B, T, N.CGP_MODEL_LCD, d, nh, consts, NC_LCD, m0, d, P0, d * d, H, None, 0, None, float(Xi), dt,
                     ys_repeat, h_unit)
        need_grad = any(ctx.needs_input_grad[:4])
        every = int(ckpt_every or _default_ckpt(L, p, T, dev))
        ws, nbytes = None, 0
        if need_grad:
            nbytes = L.cgp_ekf_nll_workspace_bytes(C.byref(p), every)
            ws = torch.empty((nbytes // 8,), dtype=_F64, device=dev)
        nll = torch.empty((B,), dtype=_F64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.cgp_ekf_nll_fwd_f64(C.byref(p), _ptr(ys), _ptr(nll), _ptr(ws), C.c_size_t(nbytes), every, stream)
        N.check(rc, 'ekf_nll')
        ctx.save_for_backward(consts, m0, P0, Xi, ys, H)
        ctx.ws, ctx.meta = ws, (dt, nh, ys_repeat, h_unit, every, nbytes, bool(raw_p0))
        return nll

    @staticmethod
    def backward(ctx, nll_bar):
        consts, m0, P0, Xi, ys, H = ctx.saved_tensors
        dt, nh, ys_repeat, h_unit, every, nbytes, raw_p0 = ctx.meta
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
        fn = L.cgp_ekf_nll_bwd_f64 if raw_p0 else L.cgp_ekf_nll_bwd_sym_f64
        rc = fn(C.byref(p), _ptr(ys), _ptr(nb), _ptr(ctx.ws), C.c_size_t(nbytes), every, _ptr(cb), _ptr(mb), _ptr(Pb),
                _ptr(xb), stream)
        N.check(rc, 'ekf_nll backward')
        return cb, mb, Pb, xb.sum(), None, None, None, None, None, None, None, None


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
            nbytes = L.cgp_ekf_nll_path_workspace_bytes(C.byref(p), every)
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


def _default_ckpt(L, p, T, dev) -> int:
    """Segment length of the checkpointed adjoint: the library default (16 steps: the per-warp scratch slots stay in L2),
    lengthened until the checkpoints (one (m, P, nll) record per problem and segment; config 5 on one GPU: 120 GB at 16
    steps) fit into three quarters of the free memory."""
    every = int(L.cgp_ekf_nll_default_ckpt(T))
    free, _ = torch.cuda.mem_get_info(dev)
    free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)      # blocks the caching allocator can reuse
    while every < T and L.cgp_ekf_nll_workspace_bytes(C.byref(p), every) > 0.75 * free:
        every *= 2
    _default_ckpt.last = {'ckpt_every': every, 'free_bytes': int(free),
                          'workspace_bytes': int(L.cgp_ekf_nll_workspace_bytes(C.byref(p), every))}
    return every


def _as_dev(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=_F64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)


def ekf_nll(cond_m_cov: LCDModel, H, Xi, m0, P0, dt, ys, candidates: bool = False,
            ckpt_every: Optional[int] = None, path: bool = False, raw_p0_cotangent: bool = False) -> torch.Tensor:
    """Final cumulative negative log-likelihood of the EKF == ``ekf(cond_m_cov, H, Xi, m0, P0, dt, ys)[-1][-1]``
    (filters_smoothers.py:222-264), differentiable w.r.t. the model hyper-parameters, ``m0``, ``P0`` and ``Xi``.

    Returns a CUDA tensor of shape () for one problem, (B,) for a batch, (B, G) with ``candidates=True``.
    ``path=True`` returns the whole third output of ``ekf`` instead -- the cumulative n_ell at every step, shape (..., T) --
    and accepts any cotangent on it (``ekf_nll_path``).

    The cotangent of ``P0`` is symmetrised, (X + X^T) / 2 of what ``jax.grad`` of the reference returns for a general
    matrix argument: the same gradient for every parametrisation in which P0 is a symmetric matrix (all callers of the
    reference build it as a diagonal, models.py:56-58).  ``raw_p0_cotangent=True`` makes the adjoint kernel also carry the
    antisymmetric part and return JAX's unsymmetrised matrix."""
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
    nll = _EkfNll.apply(consts_b, m0_b, P0_b, Xi_t, ys2, H_t, dt, nh, ys_repeat, h_unit, ckpt_every, raw_p0_cotangent)
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
        # kf takes (F, Sigma) like the reference (filters_smoothers.py:145-148); the other two a cond_m_cov callable
        model = fs._disc_model(model_args[0], int(m0.shape[-1]), dt) if method != 'kf' else fs.LinearDisc(*model_args[:2])
        consts = fs._consts_on_device(model, dt if method != 'kf' else None, fs._device(), *((dt,) if method != 'kf' else ()))
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


_TANGENT_METHODS = ('ekf', 'sgp_filter', 'cd_ekf', 'cd_sgp_filter')


def filter_nll_grad(method: str, build_model: Callable, theta, H, Xi, dt, ys, sgps=None, transform: Optional[Callable] = None):
    """Final cumulative nll of ``method`` in {'ekf', 'sgp_filter', 'cd_ekf', 'cd_sgp_filter'} AND its exact gradient w.r.t. the
    unconstrained parameters ``theta`` -- what ``jax.value_and_grad(obj_func)(theta)`` gives in demos/ghfs_mle.py:54-61,
    demos/cd_ekfs_mle.py and demos/cd_ghfs_mle.py -- by the forward-mode tangent kernels (csrc/cgp_tangent.cu): one group of
    lanes per (chirp, parameter direction) carries the filter state and its derivative along that direction.

    ``build_model(transform(theta))`` -> (drift, dispersion, m_and_cov, m0, P0, _) as ``chirpgp_b200.models.build_*``;
    ``theta`` is (P,) (shared by all chirps in ``ys``) or (B, P) (one parameter vector per chirp: the lock-step MLE of
    ``fit_mle_batched``).  The tangents of the kernel inputs (d consts / d theta etc.) come from ``torch.func.jacfwd`` of the
    builder on the host.  Returns ``(nll (B,), grad (B, P))`` as CUDA tensors (B = number of chirps)."""
    from . import filters_smoothers as fs
    from .models import g as _g, NC_SDE
    if method not in _TANGENT_METHODS:
        raise ValueError('no tangent kernel for %r' % method)
    transform = transform or _g
    dev = _device()
    dt = float(dt)
    L = N.lib()
    cd = method in ('cd_ekf', 'cd_sgp_filter')
    theta_t = (theta.detach() if isinstance(theta, torch.Tensor) else torch.as_tensor(np.asarray(theta, dtype=np.float64))).to(_F64).cpu()
    ys_t = _as_dev(ys, dev)
    ys2 = ys_t.reshape(-1, ys_t.shape[-1]).contiguous()
    B, T = int(ys2.shape[0]), int(ys2.shape[1])
    meta = {}

    def pack(th):
        drift, dispersion, m_and_cov, m0, P0, _ = build_model(transform(th))
        model = drift if cd else m_and_cov
        meta['d'], meta['nh'] = int(model.d), int(model.num_harmonics)
        parts = [(model.consts() if cd else model.consts(dt)).reshape(-1), m0.reshape(-1), P0.reshape(-1)]
        if cd:
            bm = dispersion.matrix()
            parts.append((bm @ bm.transpose(-1, -2)).reshape(-1))
        return torch.cat(parts)

    per_chirp = theta_t.dim() == 2
    if per_chirp and theta_t.shape[0] != B:
        raise ValueError('theta has %d rows for %d chirps' % (theta_t.shape[0], B))
    if per_chirp:
        vals = torch.func.vmap(pack)(theta_t)                                   # (B, NV)
        jac = torch.func.vmap(torch.func.jacfwd(pack))(theta_t)                 # (B, NV, P)
    else:
        vals = pack(theta_t)[None]
        jac = torch.func.jacfwd(pack)(theta_t)[None]
    d, nh = meta['d'], meta['nh']
    nc = NC_SDE if cd else NC_LCD
    P_ = int(theta_t.shape[-1])
    sizes = [nc, d, d * d] + ([d * d] if cd else [])
    vparts = torch.split(vals.to(dev), sizes, dim=1)
    dparts = torch.split(jac.to(dev).transpose(1, 2).contiguous(), sizes, dim=2)     # (B|1, P, size)
    consts, m0, P0 = [v.contiguous() for v in vparts[:3]]
    consts_d, m0_d, P0_d = [t.contiguous() for t in dparts[:3]]
    Qc = vparts[3].contiguous() if cd else None
    Qc_d = dparts[3].contiguous() if cd else None
    st = (lambda n: n) if per_chirp else (lambda n: 0)
    sig = fs._sigma_tables(sgps, dev) if method in ('sgp_filter', 'cd_sgp_filter') else None
    if method in ('sgp_filter', 'cd_sgp_filter') and sig is None:
        raise ValueError('%s needs sgps' % method)
    H_t = _as_dev(H, dev).reshape(-1).contiguous()
    p = _problem(B, T, N.CGP_MODEL_SDE if cd else N.CGP_MODEL_LCD, d, nh, consts, st(nc), m0, st(d), P0, st(d * d), H_t,
                 Qc, st(d * d) if cd else 0, sig, float(Xi), dt, 1, -1)
    nll = torch.empty((B,), dtype=_F64, device=dev)
    nll_dot = torch.empty((B, P_), dtype=_F64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rc = L.cgp_filter_nll_tangent_f64(method.encode(), C.byref(p), _ptr(ys2), P_, _ptr(consts_d), st(P_ * nc), _ptr(m0_d),
                                      st(P_ * d), _ptr(P0_d), st(P_ * d * d), _ptr(Qc_d), st(P_ * d * d) if cd else 0, None,
                                      _ptr(nll), _ptr(nll_dot), stream)
    N.check(rc, 'filter_nll_grad(%s)' % method)
    return nll, nll_dot


def _objective(build_model, H, Xi, dt, ys_t, method, sgps, transform, fd_step):
    """-> f(theta (n, P) numpy) = (values (n,), grads (n, P)) numpy: row i is the nll of chirp i under theta[i] and its
    gradient.  'ekf': adjoint kernel; 'sgp_filter' / 'cd_ekf' / 'cd_sgp_filter': tangent kernels; 'ekf_for_kpt': five-point
    central differences over a candidate batch."""
    dev = _device()

    def f_adjoint(theta_np, rows):
        theta = torch.tensor(theta_np, dtype=_F64, device=dev, requires_grad=True)
        _, _, m_and_cov, m0, P0, _ = build_model(transform(theta))
        val = ekf_nll(m_and_cov, H, Xi, m0, P0, dt, ys_t[rows])
        grad, = torch.autograd.grad(val.sum(), theta)
        return val.detach(), grad

    def f_tangent(theta_np, rows):
        return filter_nll_grad(method, build_model, torch.as_tensor(theta_np), H, Xi, dt, ys_t[rows], sgps=sgps,
                               transform=transform)

    def f_fd(theta_np, rows):
        # five-point central differences: (-f(+2h) + 8 f(+h) - 8 f(-h) + f(-2h)) / 12h, all 4P + 1 candidates of a chirp in ONE
        # batch against its signal (nll-only kernel)
        vals, grads = [], []
        for th, r in zip(theta_np, rows):
            n = th.shape[0]
            h = (fd_step if fd_step is not None else 1e-4) * np.maximum(1., np.abs(th))
            cand = np.tile(th, (4 * n + 1, 1))
            for i in range(n):
                for j, mult in enumerate((1., -1., 2., -2.)):
                    cand[1 + 4 * i + j, i] += mult * h[i]
            F_, Sigma_, m0, P0, h_ = build_model(transform(torch.as_tensor(cand)))
            total = filter_nll(method, (F_, Sigma_, h_), H, Xi, m0, P0, dt, ys_t[r])
            grads.append((8. * (total[1::4] - total[2::4]) - (total[3::4] - total[4::4])) / torch.as_tensor(12 * h, device=dev))
            vals.append(total[0])
        return torch.stack(vals), torch.stack(grads)

    if method == 'ekf':
        return f_adjoint
    if method in _TANGENT_METHODS:
        return f_tangent
    if method == 'ekf_for_kpt':
        return f_fd
    raise ValueError('unknown filter %r' % method)


def fit_mle(build_model: Callable, init_theta, H, Xi, dt, ys, transform: Optional[Callable] = None, maxiter: int = 200,
            reduce_group=None, method: str = 'ekf', sgps=None, fd_step: Optional[float] = None):
    """L-BFGS-B maximum-likelihood fit driving the nll kernels -- the role of
    ``jaxopt.ScipyMinimize(method='L-BFGS-B', fun=obj_func).run(init_theta)`` (demos/ekfs_mle.py:48-49 and the other
    ``*_mle.py`` demos / tetralith jobs).

    build_model(params) -> (drift, dispersion, m_and_cov, m0, P0, H') as chirpgp_b200.models.build_* (for
    method='ekf_for_kpt': (F, Sigma, m0, P0, h) as build_kpt_chirp_model with fs bound, and ``H`` is ignored); ``transform`` maps
    the unconstrained theta to params (default: the reference's softplus ``g``).  The objective is the SUM of the nll over
    all chirps in ``ys`` under ONE parameter vector (this rank's shard when torch.distributed is initialised: the scalar
    objective and its gradient are then summed over ranks with one all-reduce -- the only collective on this path).  For one
    independent fit per chirp see ``fit_mle_batched``.

    Gradients are exact for every filter of the reference's MLE demos: method='ekf' uses the hand-written reverse-mode adjoint
    kernel (csrc/cgp_nll2.cu), 'sgp_filter' (with ``sgps``), 'cd_ekf' and 'cd_sgp_filter' the forward-mode tangent kernels
    (csrc/cgp_tangent.cu).  Only 'ekf_for_kpt' (KPT model) is differentiated by five-point central differences.

    Returns (theta_opt (numpy), scipy OptimizeResult); ``result.success`` follows the reference's convention
    (tetralith/jobs/ekfs_mle.py:49, :75-78: a failed fit is reported, not raised)."""
    import scipy.optimize
    from .distributed import allreduce_objective
    from .models import g as _g
    transform = transform or _g
    dev = _device()
    ys_t = _as_dev(ys, dev)
    ys2 = ys_t.reshape(-1, ys_t.shape[-1])
    B = int(ys2.shape[0])
    rows = torch.arange(B, device=dev)

    if method == 'ekf_for_kpt':
        per_row = _objective(build_model, H, Xi, dt, ys2, method, sgps, transform, fd_step)

        def shared(theta_np):
            v, gr = per_row(np.tile(theta_np, (B, 1)), rows)
            return v.sum(), gr.sum(0)
    elif method == 'ekf':
        def shared(theta_np):
            theta = torch.tensor(theta_np, dtype=_F64, device=dev, requires_grad=True)
            _, _, m_and_cov, m0, P0, _ = build_model(transform(theta))
            val = ekf_nll(m_and_cov, H, Xi, m0, P0, dt, ys2).sum()
            grad, = torch.autograd.grad(val, theta)
            return val.detach(), grad
    elif method in _TANGENT_METHODS:
        def shared(theta_np):
            v, gr = filter_nll_grad(method, build_model, torch.as_tensor(theta_np), H, Xi, dt, ys2, sgps=sgps, transform=transform)
            return v.sum(), gr.sum(0)
    else:
        raise ValueError('unknown filter %r' % method)

    def fun(theta_np):
        v, gr = shared(theta_np)
        v, gr = allreduce_objective(v, gr, group=reduce_group)
        return float(v.cpu()), gr.cpu().numpy().copy()

    res = scipy.optimize.minimize(fun, np.asarray(init_theta, dtype=np.float64), jac=True, method='L-BFGS-B',
                                  options={'maxiter': maxiter})
    return res.x, res


def fit_mle_batched(build_model: Callable, init_theta, H, Xi, dt, ys, transform: Optional[Callable] = None, maxiter: int = 200,
                    method: str = 'ekf', sgps=None, fd_step: Optional[float] = None, nan_on_failure: bool = True):
    """One INDEPENDENT L-BFGS-B fit per chirp, all fits advanced in lock-step: what tetralith/jobs/ghfs_mle.py:26-86 (and the
    other ``*_mle.py`` jobs) run as 100 Monte-Carlo runs x 3 magnitudes of sequential ``ScipyMinimize(...).run(init_theta)``
    calls.  Every iteration evaluates the objectives and gradients that all still-running optimisers ask for in ONE batched
    kernel launch (per-chirp parameter vectors: theta (n, P) against ys (n, T)).

    Each fit is driven by SciPy's own L-BFGS-B (one optimiser per chirp on its own Python thread; the threads meet at the
    objective, where a coordinator gathers the requested points, launches the batch and hands the results back), so every
    chirp follows exactly the iterates a stand-alone ``fit_mle`` on that chirp would.

    init_theta: (P,) shared start or (B, P).  Returns (thetas (B, P) numpy, results list of scipy OptimizeResult); with
    ``nan_on_failure`` the rows of failed fits are NaN -- the reference's convention
    (tetralith/jobs/ekfs_mle.py:49, :75-78: ``if not opt_state.success: params = nan``)."""
    import threading
    import scipy.optimize
    from .models import g as _g
    transform = transform or _g
    dev = _device()
    ys_t = _as_dev(ys, dev)
    ys2 = ys_t.reshape(-1, ys_t.shape[-1]).contiguous()
    B = int(ys2.shape[0])
    init = np.asarray(init_theta, dtype=np.float64)
    init = np.tile(init, (B, 1)) if init.ndim == 1 else init
    if init.shape[0] != B:
        raise ValueError('init_theta has %d rows for %d chirps' % (init.shape[0], B))
    batch_fun = _objective(build_model, H, Xi, dt, ys2, method, sgps, transform, fd_step)

    cond = threading.Condition()
    pending, answers, state = {}, {}, {'active': B, 'error': None}

    def worker_fun(i):
        def fun(theta_np):
            with cond:
                pending[i] = np.array(theta_np, dtype=np.float64)
                cond.notify_all()
                while i not in answers and state['error'] is None:
                    cond.wait()
                if state['error'] is not None:
                    raise RuntimeError('batched objective failed') from state['error']
                return answers.pop(i)
        return fun

    results = [None] * B

    def worker(i):
        try:
            results[i] = scipy.optimize.minimize(worker_fun(i), init[i], jac=True, method='L-BFGS-B', options={'maxiter': maxiter})
        except Exception as exc:  # noqa: BLE001
            results[i] = exc
        finally:
            with cond:
                state['active'] -= 1
                cond.notify_all()

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(B)]
    for t in threads:
        t.start()
    n_launches = 0
    while True:
        with cond:
            while state['active'] > 0 and len(pending) < state['active']:
                cond.wait()
            if state['active'] == 0:
                break
            idx = sorted(pending)
            thetas = np.stack([pending.pop(i) for i in idx])
        try:
            with torch.cuda.device(dev):
                v, gr = batch_fun(thetas, torch.as_tensor(idx, device=dev))
            v, gr = v.cpu().numpy(), gr.cpu().numpy()
            n_launches += 1
            with cond:
                for j, i in enumerate(idx):
                    answers[i] = (float(v[j]), gr[j].copy())
                cond.notify_all()
        except Exception as exc:  # noqa: BLE001
            with cond:
                state['error'] = exc
                cond.notify_all()
            break
    for t in threads:
        t.join()
    if state['error'] is not None:
        raise state['error']
    for r in results:
        if isinstance(r, Exception):
            raise r
    thetas = np.stack([r.x for r in results])
    if nan_on_failure:
        for i, r in enumerate(results):
            if not r.success:
                thetas[i] = np.nan
    fit_mle_batched.last_launches = n_launches
    return thetas, results
