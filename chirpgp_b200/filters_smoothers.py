"""Gaussian filters and smoothers -- host-side mirror of the reference interface
``chirpgp.filters_smoothers`` (/root/reference/chirpgp/filters_smoothers.py), running on hand-written sm_100a
CUDA kernels through the C ABI of include/chirpgp_b200.h.

Same names, positional signatures and return tuples as the reference:

    kf :145-148, rts :187-188, ekf :222-225, ekf_for_kpt :267-270, eks :317-318, cd_ekf :352-355, cd_eks :400-402,
    sgp_filter :446-450, sgp_smoother :493-496, cd_sgp_filter :534-538, cd_sgp_smoother :585-588

Differences that follow from running compiled kernels instead of traced Python closures:

* model arguments (``cond_m_cov``, ``a``, ``b``) must be the tagged callables that ``chirpgp_b200.models``
  builds (``LCDModel``, ``SDEDrift``, ``Dispersion``, ``LinearDisc``, ``LinearSDE``); a plain Python callable
  is accepted only if probing shows it is linear, anything else raises ``NotImplementedError`` -- there is no
  CPU fallback;
* batching is native instead of ``jax.vmap``: ``ys`` may be ``(T,)`` or ``(B, T)``; model parameters, ``m0`` and
  ``P0`` may carry a matching leading batch axis (or none = shared).  Outputs get the same leading axis;
* arrays may be NumPy arrays or torch tensors (CPU or CUDA); results come back as the kind of ``ys`` /
  ``mfs`` (NumPy in -> NumPy out, CUDA tensor in -> CUDA tensors out, no host round trip).
"""
import ctypes as C
import os
import threading
import weakref
from typing import Tuple

import numpy as np
import torch

from . import _native as N
from .models import LCDModel, SDEDrift, Dispersion, LinearDisc, LinearSDE, KPTMeasurement, MODEL_KPT

__all__ = ['kf', 'rts', 'ekf', 'ekf_for_kpt', 'eks', 'cd_ekf', 'cd_eks', 'sgp_filter', 'sgp_smoother', 'cd_sgp_filter',
           'cd_sgp_smoother', 'sgp_filter_smoother', 'ekf_smoother', 'cd_ekf_smoother', 'cd_sgp_filter_smoother',
           'filter_smoother_batches', 'READOUTS']

_F64 = torch.float64

# sgp_filter on CUDA tensors also produces the smoother gains (see `sgp_filter`); CHIRPGP_B200_FUSE_GAINS=0 turns it off
FUSE_SMOOTHER_GAINS = os.environ.get('CHIRPGP_B200_FUSE_GAINS', '1') != '0'
# sgp_filter_smoother reads PINNED host measurements in place over PCIe (zero-copy) where the filter kernel streams them
# 32 samples per coalesced load, two blocks ahead (cgp_duo.cuh); CHIRPGP_B200_ZERO_COPY=0 makes it upload them first
ZERO_COPY_YS = os.environ.get('CHIRPGP_B200_ZERO_COPY', '1') != '0'


# ------------------------------------------------------------------------------------------------ helpers
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('chirpgp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _dev(x, dev) -> torch.Tensor:
    """-> contiguous float64 tensor on the GPU (no copy if it already is one)."""
    if isinstance(x, torch.Tensor):
        if (x.is_cuda and x.device == dev and x.dtype == _F64 and not x.requires_grad and x.is_contiguous()
                and x.data_ptr() % 32 == 0):
            return x                     # the very object: identity-keyed caches (_h_unit_index) keep hitting across calls
        t = x.detach()
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float64))
    t = t.to(device=dev, dtype=_F64, non_blocking=True).contiguous()
    if t.data_ptr() % 32:
        # the kernels use 16- and 32-byte vector accesses (every ABI pointer is 32-byte aligned); a view with
        # a storage offset (x[1:5]) would raise a misaligned-address fault, which is sticky for the whole CUDA context
        t = t.clone()
    return t


def _kind(x):
    if isinstance(x, torch.Tensor):
        return ('torch', x.device)
    return ('numpy', None)


_PINNED_MIN_BYTES = 1 << 20


def _back(t: torch.Tensor, kind, sync: bool = True):
    """Result tensor -> the kind of the caller's input.  ``sync=False`` (filter_smoother_batches): the device->host copy is
    only queued on the current stream; the caller waits for an event recorded behind it before touching the array."""
    to_host = kind[0] == 'numpy' or (kind[0] == 'torch' and kind[1].type == 'cpu')
    if not to_host:
        return t.to(kind[1])
    # host callers (NumPy arrays / CPU tensors) get results backed by pinned host memory (torch's caching host allocator
    # recycles the blocks): the device->host copy of the (T, d, d) outputs runs at the PCIe rate instead of the pageable rate
    host = None
    # (sync=False: a small result must not take the blocking pageable copy either -- it would stall the batch sequence)
    if t.is_cuda and (not sync or t.numel() * t.element_size() >= _PINNED_MIN_BYTES):
        try:
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=True)
            if sync:
                torch.cuda.current_stream(t.device).synchronize()
        except RuntimeError:
            host = None                  # no pinned memory left: pageable copy below
    if host is None:
        host = t.cpu()
    return host.numpy() if kind[0] == 'numpy' else host


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_sigma_cache = {}
_h_cache = {}


def _h_unit_index(H) -> int:
    """j if H is exactly the unit vector e_j, CGP_H_HARMONIC if it is the measurement row of a harmonic chirp model, else -1
    (a kernel specialisation hint: the H_E1 / H_HARM kernels never read H).
    Device tensors are inspected once per tensor OBJECT and version so that steady-state calls do not synchronise the
    stream.  The cache entry holds a weak reference to the tensor and is honoured only while that very object is alive and
    unmodified -- an address can be recycled by the caching allocator for a different H, an object identity cannot."""
    if isinstance(H, torch.Tensor) and H.is_cuda:
        key = id(H)
        hit = _h_cache.get(key)
        if hit is not None and hit[0]() is H and hit[1] == H._version:
            return hit[2]
        host = H.detach().cpu().numpy().reshape(-1)
    else:
        key = None
        host = (H.detach().numpy() if isinstance(H, torch.Tensor) else np.asarray(H, dtype=np.float64)).reshape(-1)
    ones = np.flatnonzero(host)
    out = int(ones[0]) if ones.size == 1 and host[ones[0]] == 1. else -1
    d = host.shape[0]
    if out < 0 and d >= 6 and d % 2 == 0 and np.array_equal(host, np.array([0., 1.] * ((d - 2) // 2) + [0., 0.])):
        out = N.CGP_H_HARMONIC             # measurement row of the harmonic chirp models (models.py:257)
    if key is not None:
        if len(_h_cache) > 64:
            _h_cache.clear()
        _h_cache[key] = (weakref.ref(H), H._version, out)
    return out


def _consts_on_device(model, dt, dev, *args):
    """model.consts(dt) uploaded once per (model instance, dt, parameter versions): the constants are a handful of
    doubles computed on the host; re-uploading them from pageable memory on every call would serialise the stream."""
    params = [getattr(model, a, None) for a in ('lam', 'b', 'ell', 'sigma', 'F', 'Sigma', 'A')]
    params = [t for t in params if isinstance(t, torch.Tensor)]
    if any(t.requires_grad for t in params) or not hasattr(model, '__dict__'):
        return _dev(model.consts(*args), dev)
    key = (dt, str(dev), tuple((t.data_ptr(), t._version) for t in params))
    cache = model.__dict__.setdefault('_dev_consts', {})
    hit = cache.get(key)
    if hit is None:
        cache.clear()
        hit = _dev(model.consts(*args), dev)
        cache[key] = hit
    return hit


def _sigma_tables(sgps, dev):
    w = np.ascontiguousarray(np.asarray(sgps.w.detach().cpu() if isinstance(sgps.w, torch.Tensor) else sgps.w,
                                        dtype=np.float64))
    xi = np.ascontiguousarray(np.asarray(sgps.xi.detach().cpu() if isinstance(sgps.xi, torch.Tensor) else sgps.xi,
                                         dtype=np.float64))
    key = (w.tobytes(), xi.tobytes(), str(dev))
    hit = _sigma_cache.get(key)
    if hit is None:
        from .quadratures import SigmaPoints
        sp = SigmaPoints(int(xi.shape[1]), int(w.shape[0]), w, None, xi)
        order = sp.gauss_hermite_order()
        if order == 0 and sp.is_cubature():
            order = -1                               # marks the cubature rule
        hit = (torch.as_tensor(w).to(dev), torch.as_tensor(xi).to(dev), order)
        _sigma_cache[key] = hit
    return hit


def _probe_linear(fn, d, with_dt, dt):
    """Accept a plain Python callable only if it is linear: f(u) = F u (and a constant covariance)."""
    def call(u):
        out = fn(u, dt) if with_dt else fn(u)
        if with_dt:
            mean, cov = out
            return np.asarray(mean, dtype=np.float64), np.asarray(cov, dtype=np.float64)
        return np.asarray(out, dtype=np.float64), None
    try:
        c0, S0 = call(np.zeros(d))
        F = np.stack([call(np.eye(d)[i])[0] - c0 for i in range(d)], axis=1)
        x = np.linspace(-1.3, 0.7, d) + 0.1
        fx, Sx = call(x)
        ok = np.all(c0 == 0.) and np.allclose(fx, F @ x, rtol=1e-12, atol=1e-14)
        if with_dt:
            ok = ok and np.array_equal(S0, Sx)
    except Exception as exc:  # noqa: BLE001
        raise NotImplementedError('chirpgp_b200: cannot lower an arbitrary Python callable to a CUDA kernel (%s); '
                                  'use the tagged models from chirpgp_b200.models' % exc)
    if not ok:
        raise NotImplementedError('chirpgp_b200: only linear Python callables can be lowered automatically; use the '
                                  'tagged models from chirpgp_b200.models (no CPU fallback)')
    return (LinearDisc(F, S0) if with_dt else LinearSDE(F))


def _disc_model(cond_m_cov, d, dt):
    if isinstance(cond_m_cov, (LCDModel, LinearDisc)):
        return cond_m_cov
    if callable(cond_m_cov):
        return _probe_linear(cond_m_cov, d, True, dt)
    raise TypeError('cond_m_cov must be a chirpgp_b200.models tagged callable')


def _sde_model(a, d):
    if isinstance(a, (SDEDrift, LinearSDE)):
        return a
    if callable(a):
        return _probe_linear(a, d, False, None)
    raise TypeError('drift must be a chirpgp_b200.models tagged callable')


def _dispersion_matrix(b, m0):
    """cd_ekf / cd_eks take a dispersion *callable* (filters_smoothers.py:362, :409); the kernels need it state
    independent (true for every model of the reference)."""
    if isinstance(b, Dispersion):
        return b.matrix()
    if callable(b):
        probe = np.zeros(int(m0.shape[-1])) if not isinstance(m0, torch.Tensor) else np.zeros(int(m0.shape[-1]))
        b0 = np.asarray(b(probe), dtype=np.float64)
        b1 = np.asarray(b(probe + 1.), dtype=np.float64)
        if not np.array_equal(b0, b1):
            raise NotImplementedError('chirpgp_b200: state-dependent dispersion is not supported by the kernels')
        return torch.as_tensor(b0)
    return b if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b, dtype=np.float64))


def _model_fields(model):
    nh = getattr(model, 'num_harmonics', 0)
    return model.model_id, int(model.d), int(nh)


class _Batch:
    """Resolves the batch size from the leading axes of the per-problem inputs."""

    def __init__(self):
        self.B = None

    def see(self, t: torch.Tensor, core_dims: int, what: str):
        lead = t.shape[:t.dim() - core_dims]
        if len(lead) == 0:
            return t.reshape((1,) + tuple(t.shape)), 0
        n = int(np.prod(lead))
        t = t.reshape((n,) + tuple(t.shape[len(lead):]))
        if n == 1:
            return t, 0
        if self.B is None:
            self.B = n
        elif self.B != n:
            raise ValueError('inconsistent batch sizes: %s has %d, expected %d' % (what, n, self.B))
        return t, int(np.prod(t.shape[1:]))


def _problem(B, T, model, d, nh, consts, cs, m0, m0s, P0, P0s, H, Qc, Qs, sig, Xi, dt, ys_repeat=1, h_unit=-1):
    p = N.CgpProblem()
    p.B, p.T = B, T
    p.model, p.d, p.num_harmonics = model, d, nh
    p.ys_repeat = ys_repeat
    p.consts, p.consts_stride = _ptr(consts), cs
    p.m0, p.m0_stride = _ptr(m0), m0s
    p.P0, p.P0_stride = _ptr(P0), P0s
    p.H = _ptr(H)
    p.h_unit_index = h_unit
    p.in_flight = int(getattr(_in_flight, 'n', 0))
    p.Qc, p.Qc_stride = _ptr(Qc), Qs
    if sig is not None:
        w, xi, order = sig
        p.n_sigma = int(w.shape[0])
        p.sig_w, p.sig_xi = _ptr(w), _ptr(xi)
        p.sigma_kind = (N.CGP_SIGMA_GAUSS_HERMITE if order > 0 else
                        (N.CGP_SIGMA_CUBATURE if order < 0 else N.CGP_SIGMA_GENERIC))
        p.gh_order = max(order, 0)
    p.Xi, p.dt = float(Xi), float(dt)
    return p


class _SmootherGains:
    """Smoother workspace ([G | c | C] per (chirp, step)) that a filter call produced together with (mfs, Pfs).  Rides
    on the returned ``mfs`` tensor; the smoother uses it only if it is called with exactly those tensor OBJECTS (identity,
    through weak references -- not addresses, which the allocator recycles), unmodified according to torch's version
    counters, and the same model constants / sigma points / dt -- otherwise it recomputes the gains.

    Limits: a write into ``mfs`` / ``Pfs`` that bypasses torch's version counter (a raw-pointer kernel, DLPack / cupy view)
    is invisible here -- callers who do that must pass ``smoother_gains=False`` to ``sgp_filter`` or drop the attribute
    (``del mfs._cgp_smoother_gains``).  The record keeps d^2 + d + d (d + 1) / 2 doubles per step alive as long as ``mfs`` lives."""
    __slots__ = ('ws', 'nbytes', 'mfs_ref', 'Pfs_ref', 'mfs_version', 'Pfs_version', 'consts', 'consts_version',
                 'sig', 'dt', 'shape', 'smoother')

    def matches(self, smoother, mfs, Pfs, consts, sig, dt):
        return (self.smoother == smoother and self.mfs_ref() is mfs and self.Pfs_ref() is Pfs
                and mfs._version == self.mfs_version and Pfs._version == self.Pfs_version
                and tuple(mfs.shape) == self.shape
                and consts.data_ptr() == self.consts.data_ptr() and consts._version == self.consts_version
                and sig is self.sig and dt == self.dt)


def _run_filter(name, model, consts, H, Xi, m0, P0, dt, ys, sgps=None, Qc=None, store=True, last_only=False,
                gains_for=None, ys_host=None):
    """ys_host (internal, sgp_filter_smoother): a pinned float64 CPU tensor the kernel reads in place (zero-copy) instead of
    `ys`; honoured only on the fused filter + gains path, whose producer warp streams the measurements in coalesced 256-byte
    blocks two blocks ahead.  Results stay on the device (the caller is responsible for synchronising before `ys_host` dies)."""
    dev = _device()
    L = N.lib()
    kind = _kind(ys) if ys_host is None else ('torch', dev)
    model_id, d, nh = _model_fields(model)
    ys_t = _dev(ys, dev) if ys_host is None else ys_host
    if ys_t.dim() == 0:
        raise ValueError('ys must have at least one axis (T,)')
    out_lead = tuple(ys_t.shape[:-1])
    T = int(ys_t.shape[-1])
    bt = _Batch()
    ys2, _ = bt.see(ys_t, 1, 'ys')
    consts_t, cs = bt.see(_dev(consts, dev), 1, 'model parameters')
    m0_t, m0s = bt.see(_dev(m0, dev), 1, 'm0')
    P0_t, P0s = bt.see(_dev(P0, dev), 2, 'P0')
    Qc_t, Qs = (None, 0)
    if Qc is not None:
        Qc_t, Qs = bt.see(_dev(Qc, dev), 2, 'dispersion')
    B = bt.B or 1
    if ys2.shape[0] == 1 and B > 1:          # one signal, many parameter sets
        ys_repeat = B
        out_lead = (B,)
    else:
        ys_repeat = 1
        if B > 1 and len(out_lead) == 0:
            out_lead = (B,)
    h_unit = _h_unit_index(H)
    H_t = _dev(H, dev).reshape(-1)
    if m0_t.shape[-1] != d or P0_t.shape[-1] != d or H_t.shape[0] != d:
        raise ValueError('state dimension mismatch: model d=%d, m0 %s, P0 %s, H %s'
                         % (d, tuple(m0_t.shape), tuple(P0_t.shape), tuple(H_t.shape)))
    sig = _sigma_tables(sgps, dev) if sgps is not None else None
    if sig is not None and int(sig[1].shape[1]) != d:
        raise ValueError('sigma points have dimension %d, model has %d' % (int(sig[1].shape[1]), d))
    p = _problem(B, T, model_id, d, nh, consts_t, cs, m0_t, m0s, P0_t, P0s, H_t, Qc_t, Qs, sig, Xi, dt, ys_repeat, h_unit)
    mfs = torch.empty((B, T, d), dtype=_F64, device=dev) if store else None
    Pfs = torch.empty((B, T, d, d), dtype=_F64, device=dev) if store else None
    nell = torch.empty((B,) if last_only else (B, T), dtype=_F64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rec = None
    fused = (gains_for is not None and store and kind[0] == 'torch' and kind[1] == dev and T > 1
             and L.cgp_sgp_filter_gains_fused(C.byref(p)) == 1)
    if ys_host is not None and not fused:                # no zero-copy kernel on this path: upload and run as usual
        ys_t = _dev(ys_host, dev)
        ys2 = ys_t.reshape(ys2.shape)
    # only where the filter kernel itself produces the gains: elsewhere the same gain kernel would merely run earlier
    if fused:
        nbytes = L.cgp_workspace_bytes(b'sgp_filter_gains', C.byref(p))
        try:
            ws = torch.empty((max(nbytes, 8) // 8,), dtype=_F64, device=dev)
        except torch.OutOfMemoryError:
            ws = None                    # not enough memory to keep the gains: the smoother will recompute them
        if ws is not None:
            rc = L.cgp_sgp_filter_gains_f64(C.byref(p), _ptr(ys2), _ptr(mfs), _ptr(Pfs), _ptr(nell), int(last_only),
                                            _ptr(ws), C.c_size_t(nbytes), stream)
            N.check(rc, name)
            rec = _SmootherGains()
            rec.ws, rec.nbytes, rec.consts, rec.consts_version = ws, nbytes, consts_t, consts_t._version
            rec.sig, rec.dt, rec.smoother = sig, float(dt), gains_for
    if rec is None:
        rc = getattr(L, 'cgp_%s_f64' % name)(C.byref(p), _ptr(ys2), _ptr(mfs), _ptr(Pfs), _ptr(nell), int(last_only), stream)
        N.check(rc, name)
    if store:
        mfs = mfs.reshape(out_lead + (T, d))
        Pfs = Pfs.reshape(out_lead + (T, d, d))
    nell = nell.reshape(out_lead if last_only else out_lead + (T,))
    if not store:
        return _back(nell, kind)
    mfs, Pfs = _back(mfs, kind), _back(Pfs, kind)
    if rec is not None:
        rec.mfs_ref, rec.Pfs_ref, rec.shape = weakref.ref(mfs), weakref.ref(Pfs), tuple(mfs.shape)
        rec.mfs_version, rec.Pfs_version = mfs._version, Pfs._version
        mfs._cgp_smoother_gains = rec
    return mfs, Pfs, _back(nell, kind)


def _run_smoother(name, model, consts, mfs, Pfs, dt, sgps=None, Qc=None):
    dev = _device()
    L = N.lib()
    kind = _kind(mfs)
    model_id, d, nh = _model_fields(model)
    mfs_t, Pfs_t = _dev(mfs, dev), _dev(Pfs, dev)
    if mfs_t.dim() < 2 or Pfs_t.dim() != mfs_t.dim() + 1:
        raise ValueError('mfs must be (..., T, d) and Pfs (..., T, d, d)')
    out_lead = tuple(mfs_t.shape[:-2])
    T = int(mfs_t.shape[-2])
    if int(mfs_t.shape[-1]) != d:
        raise ValueError('state dimension mismatch: model d=%d, mfs %s' % (d, tuple(mfs_t.shape)))
    bt = _Batch()
    m2, _ = bt.see(mfs_t, 2, 'mfs')
    P2, _ = bt.see(Pfs_t, 3, 'Pfs')
    consts_t, cs = bt.see(_dev(consts, dev), 1, 'model parameters')
    Qc_t, Qs = (None, 0)
    if Qc is not None:
        Qc_t, Qs = bt.see(_dev(Qc, dev), 2, 'dispersion')
    B = int(m2.shape[0])
    if bt.B is not None and bt.B != B:
        raise ValueError('batched model parameters need equally batched mfs / Pfs')
    sig = _sigma_tables(sgps, dev) if sgps is not None else None
    p = _problem(B, T, model_id, d, nh, consts_t, cs, None, 0, None, 0, None, Qc_t, Qs, sig, 0., dt)
    mss = torch.empty_like(m2)
    Pss = torch.empty_like(P2)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rec = getattr(mfs, '_cgp_smoother_gains', None) if isinstance(mfs, torch.Tensor) else None
    if (rec is not None and isinstance(Pfs, torch.Tensor) and mfs_t.data_ptr() == mfs.data_ptr()
            and rec.matches(name, mfs, Pfs, consts_t, sig, float(dt))):
        # the filter call that produced (mfs, Pfs) already evaluated the gains: only the sequential sweep is left
        rc = L.cgp_smoother_sweep_f64(C.byref(p), _ptr(m2), _ptr(P2), _ptr(mss), _ptr(Pss), _ptr(rec.ws),
                                      C.c_size_t(rec.nbytes), stream)
    else:
        nbytes = L.cgp_workspace_bytes(name.encode(), C.byref(p))
        ws = torch.empty((max(nbytes, 8) // 8,), dtype=_F64, device=dev)
        rc = getattr(L, 'cgp_%s_f64' % name)(C.byref(p), _ptr(m2), _ptr(P2), _ptr(mss), _ptr(Pss), _ptr(ws),
                                             C.c_size_t(nbytes), stream)
    N.check(rc, name)
    return _back(mss.reshape(out_lead + (T, d)), kind), _back(Pss.reshape(out_lead + (T, d, d)), kind)


def _state_dim(m0):
    return int(m0.shape[-1])


# ------------------------------------------------------------------------------------------------ public API
def kf(F, Sigma, H, Xi, m0, P0, ys) -> Tuple:
    """Kalman filter for 1-d measurements (filters_smoothers.py:145-184).
    Returns (mfs (T, d), Pfs (T, d, d), n_ell (T,)); n_ell is the cumulative negative log-likelihood."""
    model = LinearDisc(F, Sigma)
    return _run_filter('kf', model, _consts_on_device(model, None, _device()), H, Xi, m0, P0, 0., ys)


def rts(F, Sigma, mfs, Pfs) -> Tuple:
    """RTS smoother (filters_smoothers.py:187-219)."""
    model = LinearDisc(F, Sigma)
    return _run_smoother('rts', model, _consts_on_device(model, None, _device()), mfs, Pfs, 0.)


def ekf(cond_m_cov, H, Xi, m0, P0, dt, ys) -> Tuple:
    """Extended Kalman filter (filters_smoothers.py:222-264)."""
    dt = float(dt)
    model = _disc_model(cond_m_cov, _state_dim(m0), dt)
    return _run_filter('ekf', model, _consts_on_device(model, dt, _device(), dt), H, Xi, m0, P0, dt, ys)


class _KPTModel:
    """(F, Sigma) of the KPT model tagged for cgp_ekf_for_kpt_f64."""
    model_id = MODEL_KPT

    def __init__(self, lin: LinearDisc, num_harmonics: int):
        self.lin, self.d, self.num_harmonics = lin, int(lin.d), int(num_harmonics)


def ekf_for_kpt(F, Sigma, h, Xi, m0, P0, dt, ys) -> Tuple:
    """Ad-hoc extended Kalman filter for the KPT model (filters_smoothers.py:267-314): linear prediction with (F, Sigma),
    nonlinear scalar measurement ``h``.  ``h`` must be the tagged measurement function that
    ``chirpgp_b200.models.build_kpt_chirp_model`` returns (its Jacobian is compiled into the kernel); smooth the result
    with ``rts(F, Sigma, mfs, Pfs)`` as tetralith/jobs/kpt_mle.py:59-62 does."""
    if not isinstance(h, KPTMeasurement):
        raise NotImplementedError('chirpgp_b200.ekf_for_kpt: the measurement function must be the tagged callable of '
                                  'chirpgp_b200.models.build_kpt_chirp_model (no CPU fallback)')
    lin = LinearDisc(F, Sigma)
    if int(lin.d) != h.num_harmonics + 2:
        raise ValueError('KPT state dimension %d does not match %d harmonics' % (int(lin.d), h.num_harmonics))
    model = _KPTModel(lin, h.num_harmonics)
    dev = _device()
    H_unused = torch.zeros(int(lin.d), dtype=_F64)
    return _run_filter('ekf_for_kpt', model, _dev(lin.consts(), dev), H_unused, Xi, m0, P0, float(dt), ys)


def eks(cond_m_cov, mfs, Pfs, dt) -> Tuple:
    """Extended Kalman smoother (filters_smoothers.py:317-349)."""
    dt = float(dt)
    model = _disc_model(cond_m_cov, int(mfs.shape[-1]), dt)
    return _run_smoother('eks', model, _consts_on_device(model, dt, _device(), dt), mfs, Pfs, dt)


def sgp_filter(cond_m_cov, sgps, H, Xi, m0, P0, dt, ys, *, smoother_gains=None, _ys_host=None) -> Tuple:
    """Sigma-point (Gauss--Hermite / cubature) filter (filters_smoothers.py:446-490).

    ``smoother_gains`` (extension; default: on for CUDA-tensor ``ys`` where a fused kernel exists -- chirp LCD model with
    Gauss--Hermite order 3 --, see ``FUSE_SMOOTHER_GAINS``): the filter kernel also evaluates what ``sgp_smoother``'s reverse scan computes from the filtering result alone (:520-527 -- the
    sigma-point prediction from (mf_k, Pf_k) is the one the filter makes for step k + 1) and leaves it in a device
    workspace attached to the returned ``mfs``; ``sgp_smoother(cond_m_cov, sgps, mfs, Pfs, dt)`` on those very tensors
    then only runs the sequential sweep (:83-84).  Any other input to the smoother takes the stand-alone path."""
    dt = float(dt)
    model = _disc_model(cond_m_cov, _state_dim(m0), dt)
    fuse = FUSE_SMOOTHER_GAINS if smoother_gains is None else bool(smoother_gains)
    return _run_filter('sgp_filter', model, _consts_on_device(model, dt, _device(), dt), H, Xi, m0, P0, dt, ys, sgps=sgps,
                       gains_for='sgp_smoother' if fuse else None, ys_host=_ys_host)


def sgp_smoother(cond_m_cov, sgps, mfs, Pfs, dt) -> Tuple:
    """Sigma-point smoother (filters_smoothers.py:493-531)."""
    dt = float(dt)
    model = _disc_model(cond_m_cov, int(mfs.shape[-1]), dt)
    return _run_smoother('sgp_smoother', model, _consts_on_device(model, dt, _device(), dt), mfs, Pfs, dt, sgps=sgps)


READOUTS = ('mfs', 'Pfs', 'n_ell', 'mss', 'Pss', 'n_ell_last', 'freq', 'v_mean', 'v_var')
_gh1_cache = {}


def _frequency(mss: torch.Tensor, Pss: torch.Tensor, order: int = 10) -> torch.Tensor:
    """E[g(V_k)] under the smoothing marginal V_k ~ N(mss[.., k, d-2], Pss[.., k, d-2, d-2]) on the device:
    ``gaussian_expectation(ms=mss[:, 2], chol_Ps=sqrt(Pss[:, 2, 2]), force_shape=True)`` of demos/ghfs_mle.py:87-89
    (quadratures.py:234-274), read straight out of the smoother result with element strides."""
    from .quadratures import SigmaPoints
    d = int(mss.shape[-1])
    v = d - 2
    n = int(mss.numel() // d)
    out = torch.empty(tuple(mss.shape[:-1]), dtype=_F64, device=mss.device)
    if n == 0:
        return out
    tab = _gh1_cache.get(int(order))
    if tab is None:                      # np.roots is an eigen-solve: once per order, not once per call
        sg = SigmaPoints.gauss_hermite(d=1, order=order)
        tab = (np.ascontiguousarray(sg.w, dtype=np.float64), np.ascontiguousarray(np.asarray(sg.xi)[:, 0], dtype=np.float64))
        _gh1_cache[int(order)] = tab
    w, xi = tab
    stream = C.c_void_p(torch.cuda.current_stream(mss.device).cuda_stream)
    rc = N.lib().cgp_gaussian_expectation_softplus_f64(n, C.c_void_p(mss.data_ptr() + 8 * v), d,
                                                       C.c_void_p(Pss.data_ptr() + 8 * (v * d + v)), d * d, 1,
                                                       w.ctypes.data_as(C.c_void_p), xi.ctypes.data_as(C.c_void_p), int(order),
                                                       C.c_void_p(out.data_ptr()), stream)
    N.check(rc, 'frequency readout')
    return out


def _filter_smoother(run_filter, run_smoother, H, m0, P0, ys, readout, order, zero_copy=False, sync=True):
    """filter + smoother in one call on the device; only the requested results travel back to a host caller.
    ``sync=False``: nothing waits for the stream (results and a zero-copy ``ys`` are the caller's to guard with an event)."""
    dev = _device()
    kind = _kind(ys)
    # zero-copy measurements pay off for one blocking call (no separate upload in front of the filter: 5.2 vs 5.7 ms for config 2); in
    # a batch sequence the asynchronous upload hides under the other batches' kernels anyway, and with eight GPUs sharing the host's
    # PCIe path the copy engine's bulk transfer beats the kernel's 256-byte reads (5.7 vs 6.4 ms per batch and rank)
    if (zero_copy and ZERO_COPY_YS and getattr(_in_flight, 'n', 0) <= 1 and isinstance(ys, torch.Tensor) and not ys.is_cuda
            and ys.dtype == _F64 and ys.is_contiguous() and ys.is_pinned() and ys.data_ptr() % 32 == 0):
        f = run_filter(_dev(H, dev), _dev(m0, dev), _dev(P0, dev), ys, True)
    else:
        f = run_filter(_dev(H, dev), _dev(m0, dev), _dev(P0, dev), _dev(ys, dev), False)
    sm = run_smoother(f[0], f[1])
    if readout is None:
        out = tuple(_back(t, kind, sync) for t in f + sm)
        if sync:
            torch.cuda.current_stream(dev).synchronize()      # a zero-copy input must outlive the kernels that read it
        return out
    names = (readout,) if isinstance(readout, str) else tuple(readout)
    d = int(sm[0].shape[-1])
    have = {'mfs': f[0], 'Pfs': f[1], 'n_ell': f[2], 'mss': sm[0], 'Pss': sm[1]}
    out = []
    for nm in names:
        if nm in have:
            t = have[nm]
        elif nm == 'n_ell_last':
            t = f[2][..., -1].contiguous()
        elif nm == 'freq':
            t = _frequency(sm[0], sm[1], order)
        elif nm == 'v_mean':
            t = sm[0][..., d - 2].contiguous()
        elif nm == 'v_var':
            t = sm[1][..., d - 2, d - 2].contiguous()
        else:
            raise ValueError('unknown readout %r (choose from %s)' % (nm, ', '.join(READOUTS)))
        out.append(_back(t, kind, sync))
    if sync and (kind[0] == 'numpy' or (kind[0] == 'torch' and kind[1].type == 'cpu')):
        torch.cuda.current_stream(dev).synchronize()      # a zero-copy input must outlive the kernels that read it
    return tuple(out)


def sgp_filter_smoother(cond_m_cov, sgps, H, Xi, m0, P0, dt, ys, readout=None, order: int = 10, *, _sync: bool = True) -> Tuple:
    """``sgp_filter`` followed by ``sgp_smoother`` in one call (extension; what every demo / job does back to back,
    demos/ghfs_mle.py:69-85).  ``readout=None`` returns ``(mfs, Pfs, n_ell, mss, Pss)`` of the kind of ``ys``.  For host
    (NumPy) callers this is the efficient form of the pair: the measurements are uploaded once, the filtering result never
    travels back up, and the filter kernel hands the smoother its gains; the results come down into pinned host memory.

    ``readout`` -- a name or tuple of names from ``READOUTS`` -- returns just those, in that order, and only those cross PCIe:
    the five arrays above, ``'n_ell_last'`` (the MLE objective), and what the demos compute from the smoothing marginal of
    the frequency state V = x[d-2] right afterwards (demos/ghfs_mle.py:87-89): ``'freq'`` = E[g(V_k)] by Gauss--Hermite of
    ``order`` on the device (``gaussian_expectation``, quadratures.py:234-274), ``'v_mean'``, ``'v_var'``.  Asking for
    ``('freq', 'v_var')`` returns 16 bytes per step instead of 328."""
    dt = float(dt)
    def run_filter(H_, m0_, P0_, ys_, host):
        if host:                                       # pinned host measurements, read in place by the filter kernel
            return sgp_filter(cond_m_cov, sgps, H_, Xi, m0_, P0_, dt, ys_, _ys_host=ys_)
        return sgp_filter(cond_m_cov, sgps, H_, Xi, m0_, P0_, dt, ys_)
    return _filter_smoother(run_filter, lambda mfs, Pfs: sgp_smoother(cond_m_cov, sgps, mfs, Pfs, dt), H, m0, P0, ys, readout,
                            order, zero_copy=True, sync=_sync)


def ekf_smoother(cond_m_cov, H, Xi, m0, P0, dt, ys, readout=None, order: int = 10, *, _sync: bool = True) -> Tuple:
    """``ekf`` + ``eks`` in one call (demos/ekfs_mle.py:68-76); ``readout`` as in ``sgp_filter_smoother``."""
    dt = float(dt)
    return _filter_smoother(lambda H_, m0_, P0_, ys_, host: ekf(cond_m_cov, H_, Xi, m0_, P0_, dt, ys_),
                            lambda mfs, Pfs: eks(cond_m_cov, mfs, Pfs, dt), H, m0, P0, ys, readout, order, sync=_sync)


def cd_ekf_smoother(a, b, H, Xi, m0, P0, dt, ys, readout=None, order: int = 10, *, _sync: bool = True) -> Tuple:
    """``cd_ekf`` + ``cd_eks`` in one call (demos/cd_ekfs_mle.py); ``readout`` as in ``sgp_filter_smoother``."""
    dt = float(dt)
    return _filter_smoother(lambda H_, m0_, P0_, ys_, host: cd_ekf(a, b, H_, Xi, m0_, P0_, dt, ys_),
                            lambda mfs, Pfs: cd_eks(a, b, mfs, Pfs, dt), H, m0, P0, ys, readout, order, sync=_sync)


def cd_sgp_filter_smoother(a, b, sgps, H, Xi, m0, P0, dt, ys, readout=None, order: int = 10, *, _sync: bool = True) -> Tuple:
    """``cd_sgp_filter`` + ``cd_sgp_smoother`` in one call (demos/cd_ghfs_mle.py:61-75); ``readout`` as in ``sgp_filter_smoother``."""
    dt = float(dt)
    return _filter_smoother(lambda H_, m0_, P0_, ys_, host: cd_sgp_filter(a, b, sgps, H_, Xi, m0_, P0_, dt, ys_),
                            lambda mfs, Pfs: cd_sgp_smoother(a, b, sgps, mfs, Pfs, dt), H, m0, P0, ys, readout, order, sync=_sync)


_stream_cache = {}
_batch_stats = {'oom_fallbacks': 0}      # how often a batch sequence had to lower its depth (tests, diagnostics)
_in_flight = threading.local()           # .n = batches filter_smoother_batches keeps in flight (CgpProblem.in_flight hint)


def _batch_streams(dev, depth):
    """The side streams of filter_smoother_batches, kept per device: torch's caching allocator pools freed blocks per
    stream, so fresh streams on every call would cudaMalloc (and synchronise the device for) every batch buffer again."""
    have = _stream_cache.setdefault(str(dev), [])
    while len(have) < depth:
        have.append(torch.cuda.Stream(dev))
    return have[:depth]


_PINNED_SEED_MAX_BYTES = 1 << 30


def _seed_pinned_pool(outs, n):
    """Host results of a batch sequence come from torch's caching host allocator, which hands a freed block out again only once
    an event recorded at free time -- behind everything queued on the batch's stream by then, i.e. up to `depth` later batches --
    has completed.  A steady sequence therefore cycles through ~2 x depth blocks per result; left to grow on demand, the pool takes
    a cudaHostAlloc (~14 ms, synchronising) whenever timing jitter leaves no completed block.  Allocating the blocks up front and
    dropping them (never used on a stream: immediately reusable) moves that cost to the start of the sequence."""
    shapes = [tuple(o.shape) for o in outs if isinstance(o, np.ndarray) or (isinstance(o, torch.Tensor) and not o.is_cuda)]
    shapes = [sh for sh in shapes if 8 * int(np.prod(sh)) >= _PINNED_MIN_BYTES]          # small results: nothing to gain
    total = sum(8 * int(np.prod(sh)) for sh in shapes) * n
    if not shapes or total > _PINNED_SEED_MAX_BYTES:
        return
    try:
        spare = [torch.empty(sh, dtype=_F64, pin_memory=True) for sh in shapes for _ in range(n)]
        del spare
    except RuntimeError:
        pass                              # no pinned memory to spare: the pool grows on demand as before


def filter_smoother_batches(pair, *model_args, batches, readout=None, order: int = 10, depth: int = 4):
    """Run one of the filter + smoother pairs over a SEQUENCE of measurement batches with ``depth`` batches in flight
    (extension).  The Monte-Carlo jobs of the reference call the pair once per run in a Python loop
    (tetralith/jobs/ghfs_mle.py:26-86: ``for mc in range(num_mcs)``); issued one blocking call after the other, every
    batch pays its host->device and device->host transfers and the tail of its latency-bound filter kernel serially.  Here
    batch k + 1 is already queued on a second CUDA stream while batch k runs: its filter kernel fills the SM sub-partitions
    that batch k's early-finishing chirps leave idle, and batch k's readout crosses PCIe underneath it.

        for freq, v_var in cg.filter_smoother_batches(cg.sgp_filter_smoother, m_and_cov, sgps, H, Xi, m0, P0, dt,
                                                      batches=ys_batches, readout=('freq', 'v_var')):
            ...

    ``pair`` is ``sgp_filter_smoother``, ``ekf_smoother``, ``cd_ekf_smoother`` or ``cd_sgp_filter_smoother``; ``model_args`` are
    that function's positional arguments without the trailing ``ys``; ``batches`` is any iterable of ``ys`` arrays (NumPy,
    CPU tensors -- pinned ones are read in place by the filter kernel -- or CUDA tensors).  Yields, in order, exactly what
    ``pair(*model_args, ys, readout=readout, order=order)`` returns for each batch; a yielded result is complete (its stream
    has been waited for).  Host results live in pinned memory that torch's host allocator recycles once they are dropped.
    Memory: ``depth`` batches' device buffers are alive at once (config 2: 1.8 GB each); if that does not fit, the sequence goes on
    with as many batches in flight as did.  The library is told how many batches
    are in flight (``CgpProblem.in_flight``) and picks its kernels for the throughput of the overlapping launches rather than for
    the latency of one: from 4000 chirps in flight the Gauss--Hermite pair runs the 8-lanes-per-chirp kernel (results agree with
    the single-call kernels to rounding, not bit for bit).  Config 2 (1000 chirps per batch), host to host: depth 3: 3.4 ms per
    batch, depth 4 (the default): 3.0 ms, depth 8: 2.3 ms (one blocking call: 5.2 ms)."""
    if pair not in (sgp_filter_smoother, ekf_smoother, cd_ekf_smoother, cd_sgp_filter_smoother):
        raise TypeError('filter_smoother_batches: `pair` must be one of the *_smoother pair functions of chirpgp_b200')
    depth = int(depth)
    if depth < 1:
        raise ValueError('depth must be >= 1')
    from collections import deque
    dev = _device()
    streams = _batch_streams(dev, depth)
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))          # the model arguments may still be on their way to the device
    for st in streams:
        st.wait_event(ready)
    inflight = deque()                                    # (results, event, ys kept alive: a zero-copy input of the kernels)
    limit = depth                                         # batches allowed in flight (lowered if the device runs out of memory)
    try:
        for k, ys in enumerate(batches):
            st = streams[k % depth]
            if isinstance(ys, torch.Tensor) and ys.is_cuda:
                # a device batch may have been produced just now on the caller's stream (a lazy iterable of simulate(...) results)
                st.wait_stream(torch.cuda.current_stream(dev))
            for attempt in (0, 1):
                oom = False
                _in_flight.n = limit
                try:
                    with torch.cuda.stream(st):
                        out = pair(*model_args, ys, readout=readout, order=order, _sync=False)
                        done = torch.cuda.Event()
                        done.record(st)
                except torch.OutOfMemoryError:
                    if attempt or not inflight:
                        raise
                    oom = True
                finally:
                    _in_flight.n = 0
                if not oom:
                    break
                # `depth` batches of this size do not fit: hand out what is in flight, give the streams' pools back to the device
                # and go on with as many batches in flight as did fit (outside the except clause: the traceback of the failed
                # call keeps its partial allocations alive)
                limit = max(1, len(inflight))
                _batch_stats['oom_fallbacks'] += 1
                while inflight:
                    out0, done0, _ = inflight.popleft()
                    done0.synchronize()
                    yield out0
                torch.cuda.synchronize(dev)
                torch.cuda.empty_cache()
            if k == 0:
                _seed_pinned_pool(out, 2 * depth + 2)
            inflight.append((out, done, ys))
            if len(inflight) >= limit:
                out0, done0, _ = inflight.popleft()
                done0.synchronize()
                yield out0
        while inflight:
            out0, done0, _ = inflight.popleft()
            done0.synchronize()
            yield out0
    finally:
        for _, done0, _ in inflight:                      # generator closed early: the queued kernels still read their inputs
            done0.synchronize()


def _qc(bm):
    bm = bm if isinstance(bm, torch.Tensor) else torch.as_tensor(np.asarray(bm, dtype=np.float64))
    bm = bm.to(_F64)
    return bm @ bm.transpose(-1, -2)


def cd_ekf(a, b, H, Xi, m0, P0, dt, ys) -> Tuple:
    """Continuous-discrete EKF, one RK4 step per sample (filters_smoothers.py:352-397).
    ``b`` is the dispersion *callable* as in the reference."""
    model = _sde_model(a, _state_dim(m0))
    return _run_filter('cd_ekf', model, model.consts(), H, Xi, m0, P0, float(dt), ys, Qc=_qc(_dispersion_matrix(b, m0)))


def cd_eks(a, b, mfs, Pfs, dt) -> Tuple:
    """Continuous-discrete EKS (filters_smoothers.py:400-443)."""
    model = _sde_model(a, int(mfs.shape[-1]))
    return _run_smoother('cd_eks', model, model.consts(), mfs, Pfs, float(dt), Qc=_qc(_dispersion_matrix(b, mfs[..., 0, :])))


def cd_sgp_filter(a, b, sgps, H, Xi, m0, P0, dt, ys) -> Tuple:
    """Continuous-discrete sigma-point filter (filters_smoothers.py:534-582).  ``b`` is the dispersion *matrix*
    (d, dw) as in the reference (callers pass ``dispersion(eye(d))``)."""
    model = _sde_model(a, _state_dim(m0))
    return _run_filter('cd_sgp_filter', model, model.consts(), H, Xi, m0, P0, float(dt), ys, sgps=sgps, Qc=_qc(b))


def cd_sgp_smoother(a, b, sgps, mfs, Pfs, dt) -> Tuple:
    """Continuous-discrete sigma-point smoother (filters_smoothers.py:585-632)."""
    model = _sde_model(a, int(mfs.shape[-1]))
    return _run_smoother('cd_sgp_smoother', model, model.consts(), mfs, Pfs, float(dt), sgps=sgps, Qc=_qc(b))
