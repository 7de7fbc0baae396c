"""ctypes front-end of oracle/chirpgp_oracle.c (TEST INFRASTRUCTURE ONLY -- see the C file's header).

Function names follow the reference (/root/reference/chirpgp/filters_smoothers.py); models are described by
`ChirpSpec` (chirp / harmonic / La Scala LCD + SDE, raw hyper-parameters as in models.py:437-494) or
`LinearSpec` (the linear models of test/test_filters_smoothers.py:19-85).  All inputs/outputs are NumPy
float64; a leading batch axis on `ys` (or on `mfs`/`Pfs`) runs chirps in parallel with OpenMP.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libchirpgp_oracle.so')
_lib = None

V = dict(kf=0, rts=1, ekf=2, eks=3, sgp_filter=4, sgp_smoother=5, cd_ekf=6, cd_eks=7, cd_sgp_filter=8,
         cd_sgp_smoother=9)
_FILTERS = ('kf', 'ekf', 'sgp_filter', 'cd_ekf', 'cd_sgp_filter')


class _OrModel(C.Structure):
    _fields_ = [('kind', C.c_int), ('d', C.c_int), ('h', C.c_int), ('lascala', C.c_int),
                ('lam', C.c_double), ('b', C.c_double), ('ell', C.c_double), ('sigma', C.c_double),
                ('freq_scale', C.c_double),
                ('F', C.c_void_p), ('Sigma', C.c_void_p), ('A', C.c_void_p), ('Bm', C.c_void_p), ('dw', C.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'chirpgp_oracle.c')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libchirpgp_oracle.so'], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.or_batch.restype = C.c_int
        _lib.or_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class ChirpSpec:
    """params = (lam, b, ell, sigma) scalars or (B,) arrays; d = 2 * num_harmonics + 2."""

    def __init__(self, lam, b, ell, sigma, num_harmonics=1, freq_scale=1., lascala=False):
        self.hp = np.stack(np.broadcast_arrays(*[_f64(x) for x in (lam, b, ell, sigma, freq_scale)]), axis=-1)
        self.h = int(num_harmonics)
        self.d = 2 * self.h + 2
        self.lascala = bool(lascala)

    def dispersion_matrix(self, i=None):
        """diag(b, b, ..., 0, 2 sigma (sqrt3/ell)^1.5)  (models.py:112-113, :170-171)."""
        hp = self.hp if self.hp.ndim == 1 else self.hp[0 if i is None else i]
        b, ell, sigma = hp[1], hp[2], hp[3]
        bb = 0. if self.lascala else b
        return np.diag([bb, bb] * self.h + [0., 2 * sigma * (math.sqrt(3) / ell) ** 1.5])

    def _proto(self, keep):
        m = _OrModel()
        m.kind, m.d, m.h, m.lascala = 1, self.d, self.h, int(self.lascala)
        hp = self.hp.reshape(-1, 5)[0]
        m.lam, m.b, m.ell, m.sigma, m.freq_scale = [float(x) for x in hp]
        return m


class LinearSpec:
    def __init__(self, F=None, Sigma=None, A=None):
        self.F = None if F is None else _f64(F)
        self.Sigma = None if Sigma is None else _f64(Sigma)
        self.A = None if A is None else _f64(A)
        ref = self.F if self.F is not None else self.A
        self.d = ref.shape[0]
        self.hp = None

    def _proto(self, keep):
        m = _OrModel()
        m.kind, m.d, m.h = 0, self.d, 0
        m.F, m.Sigma, m.A = _p(self.F), _p(self.Sigma), _p(self.A)
        return m


def _run(name, spec, H=None, Xi=0., m0=None, P0=None, dt=0., ys=None, mfs=None, Pfs=None, sgps=None, Bm=None,
         nthreads=0, store=True):
    L = lib()
    d = spec.d
    proto = spec._proto(None)
    keep = []
    if Bm is not None:
        Bm = _f64(Bm)
        proto.Bm, proto.dw = _p(Bm), Bm.shape[1]
    is_filter = name in _FILTERS
    if is_filter:
        ys = _f64(ys)
        batched = ys.ndim == 2
        ys2 = ys.reshape(-1, ys.shape[-1])
        B, T = ys2.shape
    else:
        mfs, Pfs = _f64(mfs), _f64(Pfs)
        batched = mfs.ndim == 3
        mfs = mfs.reshape(-1, mfs.shape[-2], d)
        Pfs = Pfs.reshape(-1, Pfs.shape[-3], d, d)
        B, T = mfs.shape[:2]
    hp = spec.hp
    hp_stride = 0
    if hp is not None:
        hp = _f64(hp.reshape(-1, 5))
        if hp.shape[0] > 1:
            hp_stride = 5
            if is_filter and B == 1:           # one signal, many candidates
                B = hp.shape[0]
                batched = True
                ys2 = np.ascontiguousarray(np.broadcast_to(ys2, (B, T)))
            assert hp.shape[0] == B
    w = xi = None
    n = 0
    if sgps is not None:
        w, xi, n = _f64(sgps.w), _f64(sgps.xi), int(sgps.n_points)
    if is_filter:
        H = _f64(H)
        m0 = _f64(m0).reshape(-1, d)
        P0 = _f64(P0).reshape(-1, d, d)
        m0s = d if m0.shape[0] > 1 else 0
        P0s = d * d if P0.shape[0] > 1 else 0
        out_m = np.empty((B, T, d)) if store else None
        out_P = np.empty((B, T, d, d)) if store else None
        out_n = np.empty((B, T))
        rc = L.or_batch(V[name], C.c_int64(B), C.c_int64(T), C.byref(proto), _p(hp), C.c_int64(hp_stride),
                        _p(w), _p(xi), C.c_int(n), _p(H), C.c_double(Xi), _p(m0), C.c_int64(m0s), _p(P0),
                        C.c_int64(P0s), C.c_double(dt), _p(ys2), C.c_int64(T), None, None,
                        _p(out_m), _p(out_P), _p(out_n), C.c_int(nthreads))
        assert rc == 0
        if not store:
            return None, None, (out_n if batched else out_n[0])
        if not batched:
            return out_m[0], out_P[0], out_n[0]
        return out_m, out_P, out_n
    out_m = np.empty((B, T, d))
    out_P = np.empty((B, T, d, d))
    rc = L.or_batch(V[name], C.c_int64(B), C.c_int64(T), C.byref(proto), _p(hp), C.c_int64(hp_stride),
                    _p(w), _p(xi), C.c_int(n), None, C.c_double(0.), None, C.c_int64(0), None, C.c_int64(0),
                    C.c_double(dt), None, C.c_int64(0), _p(mfs), _p(Pfs), _p(out_m), _p(out_P), None,
                    C.c_int(nthreads))
    assert rc == 0
    if not batched:
        return out_m[0], out_P[0]
    return out_m, out_P


# ---- reference-named entry points (filters_smoothers.py:145, :187, :222, :317, :352, :400, :446, :493, :534, :585)
def kf(F, Sigma, H, Xi, m0, P0, ys, **kw):
    return _run('kf', LinearSpec(F=F, Sigma=Sigma), H, Xi, m0, P0, 0., ys, **kw)


def rts(F, Sigma, mfs, Pfs, **kw):
    return _run('rts', LinearSpec(F=F, Sigma=Sigma), mfs=mfs, Pfs=Pfs, **kw)


def ekf(spec, H, Xi, m0, P0, dt, ys, **kw):
    return _run('ekf', spec, H, Xi, m0, P0, dt, ys, **kw)


def eks(spec, mfs, Pfs, dt, **kw):
    return _run('eks', spec, mfs=mfs, Pfs=Pfs, dt=dt, **kw)


def sgp_filter(spec, sgps, H, Xi, m0, P0, dt, ys, **kw):
    return _run('sgp_filter', spec, H, Xi, m0, P0, dt, ys, sgps=sgps, **kw)


def sgp_smoother(spec, sgps, mfs, Pfs, dt, **kw):
    return _run('sgp_smoother', spec, mfs=mfs, Pfs=Pfs, dt=dt, sgps=sgps, **kw)


def cd_ekf(spec, Bm, H, Xi, m0, P0, dt, ys, **kw):
    return _run('cd_ekf', spec, H, Xi, m0, P0, dt, ys, Bm=Bm, **kw)


def cd_eks(spec, Bm, mfs, Pfs, dt, **kw):
    return _run('cd_eks', spec, mfs=mfs, Pfs=Pfs, dt=dt, Bm=Bm, **kw)


def cd_sgp_filter(spec, Bm, sgps, H, Xi, m0, P0, dt, ys, **kw):
    return _run('cd_sgp_filter', spec, H, Xi, m0, P0, dt, ys, sgps=sgps, Bm=Bm, **kw)


def cd_sgp_smoother(spec, Bm, sgps, mfs, Pfs, dt, **kw):
    return _run('cd_sgp_smoother', spec, mfs=mfs, Pfs=Pfs, dt=dt, sgps=sgps, Bm=Bm, **kw)


def ekf_for_kpt(F, Sigma, num_harmonics, Xi, m0, P0, ys, nthreads=0):
    """filters_smoothers.py:267-314 with the measurement function of build_kpt_chirp_model (models.py:572-578).
    ys (T,) or (B, T); F / Sigma / m0 / P0 shared or with a leading batch axis."""
    L = lib()
    F, Sigma, m0, P0, ys = _f64(F), _f64(Sigma), _f64(m0), _f64(P0), _f64(ys)
    batched = ys.ndim == 2
    ys2 = ys.reshape(-1, ys.shape[-1])
    B, T = ys2.shape
    d = m0.shape[-1]
    fs_stride = d * d if F.ndim == 3 else 0
    assert F.shape == Sigma.shape
    out_m, out_P, out_n = np.empty((B, T, d)), np.empty((B, T, d, d)), np.empty((B, T))
    rc = L.or_ekf_kpt_batch(C.c_int64(B), C.c_int64(T), C.c_int(d), C.c_int(int(num_harmonics)), _p(F), _p(Sigma),
                            C.c_int64(fs_stride), C.c_double(float(Xi)), _p(m0), C.c_int64(d if m0.ndim == 2 else 0), _p(P0),
                            C.c_int64(d * d if P0.ndim == 3 else 0), _p(ys2), C.c_int64(T), _p(out_m), _p(out_P), _p(out_n),
                            C.c_int(nthreads))
    assert rc == 0
    if not batched:
        return out_m[0], out_P[0], out_n[0]
    return out_m, out_P, out_n


def kpt_model(params, fs, num_harmonics=1):
    """build_kpt_chirp_model (models.py:522-580) without the callable: F, Sigma, m0, P0 as NumPy arrays."""
    q1, q2, p0, f0, a0 = [float(x) for x in params]
    d = num_harmonics + 2
    F = np.eye(d); F[-1, 0] = 1.
    Sigma = np.zeros((d, d))
    Sigma[0, 0] = (2 * math.pi * q1 / fs) ** 2
    for k in range(1, num_harmonics + 1):
        Sigma[k, k] = q2
    m0 = np.array([2 * math.pi * f0 / fs] + [a0] * num_harmonics + [0.])
    return F, Sigma, m0, p0 * np.eye(d)


def chirp_m0_P0_H(delta, ell, sigma, m0_v, num_harmonics=1, kind='chirp'):
    """m0, P0, H of build_chirp_model (models.py:453-459) / build_harmonic_chirp_model (:484-494)."""
    h = num_harmonics
    if kind == 'chirp':
        m0 = np.array([0., 0., m0_v, 0.])
    else:
        m0 = np.array([0., 1.] * h + [m0_v, 0.])
    P0 = np.diag([delta] * (2 * h) + [sigma ** 2, (math.sqrt(3) / ell) ** 2 * sigma ** 2])
    H = np.array([0., 1.] * h + [0., 0.])
    return m0, P0, H


def disc_mean_cov(spec, dt, u, want_jac=True):
    L = lib()
    proto = spec._proto(None)
    d = spec.d
    u = _f64(u)
    mean, J, Sig = np.empty(d), np.empty((d, d)), np.empty((d, d))
    L.or_disc_mean_cov(C.byref(proto), C.c_double(dt), _p(u), _p(mean), _p(J), _p(Sig))
    return mean, J, Sig


def sde_drift(spec, u):
    L = lib()
    proto = spec._proto(None)
    d = spec.d
    u = _f64(u)
    a, J = np.empty(d), np.empty((d, d))
    L.or_sde_drift(C.byref(proto), _p(u), _p(a), _p(J))
    return a, J


def m32_solution(ell, sigma, dt):
    L = lib()
    Ft, St = np.empty((2, 2)), np.empty((2, 2))
    L.or_m32_solution(C.c_double(ell), C.c_double(sigma), C.c_double(dt), _p(Ft), _p(St))
    return Ft, St


def max_threads():
    return lib().or_max_threads()
