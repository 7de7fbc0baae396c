"""Chirp-IF estimation models, host side (torch float64).

Mirrors the reference interface ``chirpgp.models`` (/root/reference/chirpgp/models.py):
``g`` :50, ``g_inv`` :53, ``model_chirp`` :76-119, ``model_harmonic_chirp`` :122-178, ``model_lascala`` :181-261,
``disc_chirp_lcd`` :264-311, ``disc_harmonic_chirp_lcd`` :332-386, ``disc_model_lascala_lcd`` :419-434,
``disc_m32`` :408-416, ``build_chirp_model`` :437-459, ``build_harmonic_chirp_model`` :462-494,
``build_lascala_model`` :497-519, ``build_kpt_chirp_model`` :522-580.

The reference hands the filters *Python closures* and differentiates them with ``jax.jacfwd`` inside the scan.
Here the drift / dispersion / conditional-mean functions are compiled into the CUDA kernels as device
functions, so the builders return **tagged callables**: objects that still evaluate on the host when user code
calls them (``m_and_cov(zeros(4), dt)``, ``dispersion(eye(4))``: tetralith/jobs/crlb_ekf.py:35,
demos/cd_ghfs_mle.py:48) but that the filters recognise and lower to ``(model id, constants tensor)``.

Hyper-parameters may be Python floats, NumPy arrays or torch tensors, may carry ``requires_grad`` (MLE) and may
have a leading batch axis (one model per chirp / per hyper-parameter candidate).  The map
``theta -> g(theta) -> derived constants`` stays in torch autograd; the kernels (and the hand-written adjoint
kernel) work on the derived constants only (SURVEY Appendix A "recommended factoring").
"""
import math
from typing import Tuple

import numpy as np
import torch

__all__ = ['g', 'g_inv', 'model_chirp', 'model_harmonic_chirp', 'model_lascala', 'disc_chirp_lcd',
           'disc_harmonic_chirp_lcd', 'disc_model_lascala_lcd', 'disc_m32', 'build_chirp_model',
           'build_harmonic_chirp_model', 'build_lascala_model', 'build_kpt_chirp_model', 'posterior_cramer_rao', 'LinearDisc',
           'LinearSDE',
           'LCDModel', 'SDEDrift', 'Dispersion', 'KPTMeasurement',
           'MODEL_LINEAR_DISC', 'MODEL_LCD', 'MODEL_LINEAR_SDE', 'MODEL_SDE', 'MODEL_KPT', 'NC_LCD', 'NC_SDE']

# model ids shared with include/chirpgp_b200.h
MODEL_LINEAR_DISC = 0   # x_k = F x_{k-1} + q               consts = [F (d*d), Sigma (d*d)]
MODEL_LCD = 1           # chirp / harmonic / La Scala LCD    consts = NC_LCD doubles, see LCDModel.consts
MODEL_LINEAR_SDE = 2    # dx = A x dt + B dW                 consts = [A (d*d)]
MODEL_SDE = 3           # chirp / harmonic / La Scala SDE    consts = NC_SDE doubles, see SDEDrift.consts
MODEL_KPT = 4           # KPT model (ekf_for_kpt only)       consts = [F (d*d), Sigma (d*d)], d = num_harmonics + 2
NC_LCD = 10
NC_SDE = 4

_F64 = torch.float64


def _t(x, like=None):
    """-> float64 torch tensor (keeps device / grad of tensors, wraps python / numpy numbers)."""
    if isinstance(x, torch.Tensor):
        return x if x.dtype == _F64 else x.to(_F64)
    dev = like.device if isinstance(like, torch.Tensor) else None
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)


def g(x):
    """Naive softplus ``log(exp(x) + 1)`` exactly as the reference writes it (overflows for x > 709 alike)."""
    if isinstance(x, torch.Tensor):
        return torch.log(torch.exp(x) + 1.)
    return np.log(np.exp(np.asarray(x, dtype=np.float64)) + 1.)


def g_inv(x):
    if isinstance(x, torch.Tensor):
        return torch.log(torch.exp(x) - 1.)
    return np.log(np.exp(np.asarray(x, dtype=np.float64)) - 1.)


def _m32_solution(ell, sigma, dt):
    """models.py:61-73; returns (F00, F01, F10, F11), (S00, S01, S11) as tensors broadcast over batch."""
    gamma = math.sqrt(3) / ell
    eta = dt * gamma
    beta = sigma ** 2 * torch.exp(-2 * eta)
    ee = torch.exp(-eta)
    f = ((1 + eta) * ee, (dt + 0 * eta) * ee, (-dt * gamma ** 2) * ee, (1 - eta) * ee)
    s = (sigma ** 2 - beta * (2 * eta + 2 * eta ** 2 + 1),
         2 * dt ** 2 * gamma ** 3 * beta,
         gamma ** 2 * (sigma ** 2 + beta * (2 * eta - 2 * eta ** 2 - 1)))
    return f, s


def _stationary_cov_m32_diag(ell, sigma):
    return sigma ** 2, (math.sqrt(3) / ell) ** 2 * sigma ** 2


class LCDModel:
    """Locally-conditional discretisation ``u, dt -> (cond mean, cond cov)`` of the (harmonic) chirp SDE.

    num_harmonics = h gives state dimension d = 2h + 2 (chirp model: h = 1, freq_scale = 1);
    ``lascala=True`` drops damping and chirp noise (models.py:419-434).
    """
    model_id = MODEL_LCD

    def __init__(self, lam, b, ell, sigma, num_harmonics: int = 1, freq_scale: float = 1., lascala: bool = False):
        ell = _t(ell)
        self.ell, self.sigma = ell, _t(sigma, ell)
        self.lam = None if lascala else _t(lam, ell)
        self.b = None if lascala else _t(b, ell)
        self.num_harmonics = int(num_harmonics)
        self.freq_scale = float(freq_scale)
        self.lascala = bool(lascala)
        self.d = 2 * self.num_harmonics + 2

    def consts(self, dt) -> torch.Tensor:
        """(..., NC_LCD) = [e, F00, F01, F10, F11, q, S00, S01, S11, freq_scale]   (differentiable)

        e = exp(-lam dt) (1 for La Scala); F, S = Matern-3/2 transition / covariance (models.py:61-73);
        q = chirp-block process variance with the exact ``lam == 0`` branch of models.py:302-308."""
        dt = float(dt)
        f, s = _m32_solution(self.ell, self.sigma, dt)
        if self.lascala:
            e = torch.ones_like(f[0])
            q = torch.zeros_like(f[0])
        else:
            lam, b = self.lam, self.b
            e = torch.exp(-lam * dt)
            is0 = lam == 0.
            lam_safe = torch.where(is0, torch.ones_like(lam), lam)
            q = torch.where(is0, b ** 2 * dt + 0 * lam, b ** 2 / (2 * lam_safe) * (1 - torch.exp(-2 * lam_safe * dt)))
        parts = torch.broadcast_tensors(e, *f, q, *s)
        fs = torch.full_like(parts[0], self.freq_scale)
        return torch.stack(list(parts) + [fs], dim=-1)

    def __call__(self, u, dt) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host evaluation of (cond mean, cond cov) at one state u (d,) -- convenience for user code, never
        used on the filtering path."""
        u = _t(u, self.ell)
        c = self.consts(dt)
        if c.dim() != 1:
            raise ValueError('host evaluation of a batched model is not supported')
        h, d = self.num_harmonics, self.d
        e, f00, f01, f10, f11, q, s00, s01, s11, fsc = c.unbind(-1)
        w = 2 * math.pi * g(u[d - 2]) * self.freq_scale
        rows = []
        for k in range(1, h + 1):
            cth, sth = torch.cos(dt * k * w), torch.sin(dt * k * w)
            rows.append((cth * e) * u[2 * k - 2] + (-sth * e) * u[2 * k - 1])
            rows.append((sth * e) * u[2 * k - 2] + (cth * e) * u[2 * k - 1])
        rows.append(f00 * u[d - 2] + f01 * u[d - 1])
        rows.append(f10 * u[d - 2] + f11 * u[d - 1])
        cov = torch.zeros((d, d), dtype=_F64, device=c.device)
        idx = torch.arange(2 * h)
        cov[idx, idx] = q
        cov[d - 2, d - 2], cov[d - 2, d - 1], cov[d - 1, d - 2], cov[d - 1, d - 1] = s00, s01, s01, s11
        return torch.stack(rows), cov


class SDEDrift:
    """Drift ``a(u)`` of the (harmonic) chirp SDE (models.py:104-110, :164-168, :246-252)."""
    model_id = MODEL_SDE

    def __init__(self, lam, ell, num_harmonics: int = 1, freq_scale: float = 1., lascala: bool = False):
        ell = _t(ell)
        self.ell = ell
        self.lam = torch.zeros_like(ell) if lascala else _t(lam, ell)
        self.num_harmonics = int(num_harmonics)
        self.freq_scale = float(freq_scale)
        self.d = 2 * self.num_harmonics + 2

    def consts(self) -> torch.Tensor:
        """(..., NC_SDE) = [lam, gamma^2, 2 gamma, freq_scale], gamma = sqrt(3) / ell."""
        gamma = math.sqrt(3) / self.ell
        parts = torch.broadcast_tensors(self.lam, gamma ** 2, 2 * gamma)
        return torch.stack(list(parts) + [torch.full_like(parts[0], self.freq_scale)], dim=-1)

    def __call__(self, u) -> torch.Tensor:
        u = _t(u, self.ell)
        h, d = self.num_harmonics, self.d
        lam, g2, tg, _ = self.consts().unbind(-1)
        w = 2 * math.pi * g(u[d - 2]) * self.freq_scale
        rows = []
        for k in range(1, h + 1):
            rows.append(-lam * u[2 * k - 2] + (-(w * k)) * u[2 * k - 1])
            rows.append((w * k) * u[2 * k - 2] + (-lam) * u[2 * k - 1])
        rows.append(u[d - 1] + 0 * lam)
        rows.append(-g2 * u[d - 2] + (-tg) * u[d - 1])
        return torch.stack(rows)


class Dispersion:
    """State-independent dispersion ``diag(b, b, ..., 0, 2 sigma (sqrt3/ell)^1.5)`` (models.py:112-113, :170-171).
    Calling it with anything returns the (d, d) matrix, as the reference's ``dispersion(_)`` does."""

    def __init__(self, b, ell, sigma, num_harmonics: int = 1, lascala: bool = False):
        ell = _t(ell)
        self.ell, self.sigma = ell, _t(sigma, ell)
        self.b = torch.zeros_like(ell) if lascala else _t(b, ell)
        self.num_harmonics = int(num_harmonics)
        self.d = 2 * self.num_harmonics + 2

    def matrix(self) -> torch.Tensor:
        last = 2 * self.sigma * (math.sqrt(3) / self.ell) ** 1.5
        b, last = torch.broadcast_tensors(self.b, last)
        diag = torch.stack([b] * (2 * self.num_harmonics) + [torch.zeros_like(b), last], dim=-1)
        return torch.diag_embed(diag)

    def __call__(self, _=None) -> torch.Tensor:
        return self.matrix()


class LinearDisc:
    """Tagged linear discrete model ``(u, dt) -> (F u, Sigma)`` -- what the reference's tests pass as the lambda
    ``m_and_cov`` (test/test_filters_smoothers.py:68-70); also what ``kf`` / ``rts`` run on."""
    model_id = MODEL_LINEAR_DISC

    def __init__(self, F, Sigma):
        self.F = _t(F)
        self.Sigma = _t(Sigma, self.F)
        self.d = self.F.shape[-1]

    def consts(self, dt=None) -> torch.Tensor:
        F, S = torch.broadcast_tensors(self.F, self.Sigma)
        return torch.cat([F.reshape(*F.shape[:-2], -1), S.reshape(*S.shape[:-2], -1)], dim=-1)

    def __call__(self, u, dt=None):
        return self.F @ _t(u, self.F), self.Sigma


class LinearSDE:
    """Tagged linear drift ``u -> A u`` (test/test_filters_smoothers.py:30)."""
    model_id = MODEL_LINEAR_SDE

    def __init__(self, A):
        self.A = _t(A)
        self.d = self.A.shape[-1]

    def consts(self) -> torch.Tensor:
        return self.A.reshape(*self.A.shape[:-2], -1)

    def __call__(self, u):
        return self.A @ _t(u, self.A)


def _chirp_P0(delta, ell, sigma, h):
    s0, s1 = _stationary_cov_m32_diag(ell, sigma)
    parts = torch.broadcast_tensors(delta, s0, s1)
    diag = torch.stack([parts[0]] * (2 * h) + [parts[1], parts[2]], dim=-1)
    return torch.diag_embed(diag)


def model_chirp(lam, b, ell, sigma, delta):
    """models.py:76-119 -> (drift, dispersion, m0, P0, H)."""
    ell = _t(ell)
    drift = SDEDrift(lam, ell)
    dispersion = Dispersion(b, ell, sigma)
    m0 = torch.tensor([0., 1., 0., 0.], dtype=_F64, device=ell.device)
    P0 = _chirp_P0(_t(delta, ell), ell, _t(sigma, ell), 1)
    H = torch.tensor([0., 1., 0., 0.], dtype=_F64, device=ell.device)
    return drift, dispersion, m0, P0, H


def model_harmonic_chirp(lam, b, ell, sigma, delta, num_harmonics: int = 1, freq_scale: float = 1.):
    """models.py:122-178."""
    ell = _t(ell)
    h = num_harmonics
    drift = SDEDrift(lam, ell, h, freq_scale)
    dispersion = Dispersion(b, ell, sigma, h)
    m0 = torch.tensor([0., 1.] * h + [0., 0.], dtype=_F64, device=ell.device)
    P0 = _chirp_P0(_t(delta, ell), ell, _t(sigma, ell), h)
    H = torch.tensor([0., 1.] * h + [0., 0.], dtype=_F64, device=ell.device)
    return drift, dispersion, m0, P0, H


def model_lascala(ell, sigma, delta):
    """models.py:181-261."""
    ell = _t(ell)
    drift = SDEDrift(0., ell, lascala=True)
    dispersion = Dispersion(0., ell, sigma, lascala=True)
    m0 = torch.tensor([0., 1., 0., 0.], dtype=_F64, device=ell.device)
    P0 = _chirp_P0(_t(delta, ell), ell, _t(sigma, ell), 1)
    H = torch.tensor([0., 1., 0., 0.], dtype=_F64, device=ell.device)
    return drift, dispersion, m0, P0, H


def disc_chirp_lcd(lam, b, ell, sigma):
    return LCDModel(lam, b, ell, sigma)


def disc_harmonic_chirp_lcd(lam, b, ell, sigma, num_harmonics: int = 1, freq_scale: float = 1.):
    return LCDModel(lam, b, ell, sigma, num_harmonics, freq_scale)


def disc_model_lascala_lcd(ell, sigma):
    return LCDModel(None, None, ell, sigma, lascala=True)


def disc_m32(ell, sigma):
    """Exact discretisation of the Matern-3/2 SDE (models.py:408-416) as a tagged linear model (dt bound late)."""
    ell_t, sigma_t = _t(ell), _t(sigma)

    class _M32(LinearDisc):
        def __init__(self):
            self.d = 2

        def _mats(self, dt):
            f, s = _m32_solution(ell_t, sigma_t, float(dt))
            F = torch.stack([torch.stack([f[0], f[1]], -1), torch.stack([f[2], f[3]], -1)], -2)
            S = torch.stack([torch.stack([s[0], s[1]], -1), torch.stack([s[1], s[2]], -1)], -2)
            return F, S

        def consts(self, dt=None):
            F, S = self._mats(dt)
            return torch.cat([F.reshape(*F.shape[:-2], -1), S.reshape(*S.shape[:-2], -1)], dim=-1)

        def __call__(self, u, dt=None):
            F, S = self._mats(dt)
            return F @ _t(u, F), S

    return _M32()


def _split_params(params, n):
    p = _t(params)
    if p.shape[-1] != n:
        raise ValueError('expected %d parameters in the last axis, got shape %s' % (n, tuple(p.shape)))
    return p.unbind(-1)


def build_chirp_model(params):
    """models.py:437-459.  params (..., 6) = lam, b, delta, ell, sigma, m0_1 ->
    (drift, dispersion, m_and_cov, m0, P0, H)."""
    lam, b, delta, ell, sigma, m0_v = _split_params(params, 6)
    drift, dispersion, _, P0, H = model_chirp(lam, b, ell, sigma, delta)
    z = torch.zeros_like(m0_v)
    m0 = torch.stack([z, z, m0_v, z], dim=-1)
    m_and_cov = disc_chirp_lcd(lam, b, ell, sigma)
    return drift, dispersion, m_and_cov, m0, P0, H


def build_harmonic_chirp_model(params, num_harmonics: int = 1, freq_scale: float = 1.):
    """models.py:462-494."""
    lam, b, delta, ell, sigma, m0_v = _split_params(params, 6)
    drift, dispersion, _, P0, H = model_harmonic_chirp(lam, b, ell, sigma, delta, num_harmonics=num_harmonics,
                                                       freq_scale=freq_scale)
    z, o = torch.zeros_like(m0_v), torch.ones_like(m0_v)
    m0 = torch.stack([z, o] * num_harmonics + [m0_v, z], dim=-1)
    m_and_cov = disc_harmonic_chirp_lcd(lam, b, ell, sigma, num_harmonics=num_harmonics, freq_scale=freq_scale)
    return drift, dispersion, m_and_cov, m0, P0, H


def build_lascala_model(params):
    """models.py:497-519.  params (..., 4) = delta, ell, sigma, m0_1."""
    delta, ell, sigma, m0_v = _split_params(params, 4)
    drift, dispersion, _, P0, H = model_lascala(ell, sigma, delta)
    z = torch.zeros_like(m0_v)
    m0 = torch.stack([z, z, m0_v, z], dim=-1)
    m_and_cov = disc_model_lascala_lcd(ell, sigma)
    return drift, dispersion, m_and_cov, m0, P0, H


class KPTMeasurement:
    """Tagged measurement function of the KPT model (models.py:572-578): ``h(x) = sum_k x[k] sin(k g(x[0] + x[-1]))``,
    k = 1 .. num_harmonics.  Callable on the host like the reference's closure; ``ekf_for_kpt`` recognises it and runs the
    kernel that has ``h`` and its Jacobian compiled in."""

    def __init__(self, num_harmonics: int):
        self.num_harmonics = int(num_harmonics)

    def __call__(self, x):
        x = _t(x)
        k = torch.arange(1, self.num_harmonics + 1, dtype=_F64, device=x.device)
        return (x[..., 1:-1] * torch.sin(g(x[..., 0] + x[..., -1])[..., None] * k)).sum(-1)


def build_kpt_chirp_model(params, fs: float, num_harmonics: int = 1):
    """models.py:522-580 (Shi et al. 2017 Kalman pitch tracking).  params (..., 5) = q1, q2, p0, f0, a0 ->
    (F, Sigma, m0, P0, h) with the state x = [omega, a_1 .. a_h, phase]."""
    q1, q2, p0, f0, a0 = _split_params(params, 5)
    d = num_harmonics + 2
    eye = torch.eye(d, dtype=_F64, device=q1.device)
    P0 = p0[..., None, None] * eye
    m0 = torch.stack([2 * math.pi * f0 / fs] + [a0] * num_harmonics + [torch.zeros_like(a0)], dim=-1)
    F = eye.clone()
    F[-1, 0] = 1.
    diag = torch.stack([(2 * math.pi * q1 / fs) ** 2] + [q2] * num_harmonics + [torch.zeros_like(q2)], dim=-1)
    Sigma = torch.diag_embed(diag)
    return F, Sigma, m0, P0, KPTMeasurement(num_harmonics)


def posterior_cramer_rao(xss, yss, j0, logpdf_transition, logpdf_likelihood):
    """Posterior Cramer--Rao lower bound by Monte Carlo (models.py:583-644; Tichavsky et al. 1998), for 1-d measurements.

    xss (T + 1, N, d) state trajectories (initial samples first), yss (T, N) measurements, j0 (d, d) = -E[Hess log p(x0)],
    ``logpdf_transition(xt, xs)`` and ``logpdf_likelihood(yt, xt)`` scalar-valued callables written with torch operations.
    Returns J (T, d, d), the inverses of the bound matrices.

    Analysis tool next to the hot path, not a kernel: the reference takes the Hessians with jax.hessian / jacfwd / jacrev under
    vmap (:631-634); this is the same recursion (:636-646) with ``torch.func`` doing the differentiation on whatever device
    the samples live on (``chirpgp_b200.tools.simulate`` leaves them on the GPU)."""
    from torch.func import hessian, jacfwd, jacrev, vmap
    xss, yss, j = _t(xss), _t(yss), _t(j0)
    htt_tr = vmap(hessian(logpdf_transition, argnums=0))
    hts_tr = vmap(jacfwd(jacrev(logpdf_transition, argnums=1), argnums=0))
    hss_tr = vmap(hessian(logpdf_transition, argnums=1))
    htt_li = vmap(hessian(logpdf_likelihood, argnums=1))
    js = []
    for k in range(yss.shape[0]):
        yt, xt, xs = yss[k], xss[k + 1], xss[k]
        d11 = -hss_tr(xt, xs).mean(0)
        d12 = -hts_tr(xt, xs).mean(0)
        d22 = -(htt_tr(xt, xs) + htt_li(yt, xt)).mean(0)
        j = d22 - d12.transpose(-1, -2) @ torch.linalg.solve(j + d11, d12)
        js.append(j)
    return torch.stack(js)
