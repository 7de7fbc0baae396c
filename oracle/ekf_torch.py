"""Autodiff twin of the EKF negative log-likelihood on the chirp-family LCD models (TEST INFRASTRUCTURE ONLY).

Plain torch float64 on the CPU, Python loop over time; the Jacobian comes from torch.func.jacfwd exactly as the
reference takes it from jax.jacfwd (filters_smoothers.py:255), so it is independent of the hand-derived closed forms
in the CUDA kernels.  Used by tests/ to check the adjoint kernel's cotangents (consts, m0, P0 incl. the unsymmetrised
P0 convention, Xi) and d nll / d theta."""
import math

import torch

_F64 = torch.float64


def lcd_mean(consts, u, dt, nh):
    """Conditional mean of LCDModel from its derived constants (models.py:295-301, :369-376)."""
    e, f00, f01, f10, f11 = consts[0], consts[1], consts[2], consts[3], consts[4]
    fs = consts[9]
    d = 2 * nh + 2
    w = 2 * math.pi * torch.log(torch.exp(u[d - 2]) + 1.) * fs
    rows = []
    for k in range(1, nh + 1):
        c, s = torch.cos(dt * k * w), torch.sin(dt * k * w)
        rows.append((c * e) * u[2 * k - 2] + (-s * e) * u[2 * k - 1])
        rows.append((s * e) * u[2 * k - 2] + (c * e) * u[2 * k - 1])
    rows.append(f00 * u[d - 2] + f01 * u[d - 1])
    rows.append(f10 * u[d - 2] + f11 * u[d - 1])
    return torch.stack(rows)


def lcd_cov(consts, nh):
    d = 2 * nh + 2
    q, s00, s01, s11 = consts[5], consts[6], consts[7], consts[8]
    S = torch.zeros((d, d), dtype=_F64)
    idx = torch.arange(2 * nh)
    S = S.index_put((idx, idx), q.expand(2 * nh))
    S = S.index_put((torch.tensor([d - 2, d - 2, d - 1, d - 1]), torch.tensor([d - 2, d - 1, d - 2, d - 1])),
                    torch.stack([s00, s01, s01, s11]))
    return S


def ekf_nll(consts, H, Xi, m0, P0, dt, ys, nh):
    """Final cumulative nll of filters_smoothers.py:222-264 with _linear_update :55-68."""
    mf, Pf = m0, P0
    Sigma = lcd_cov(consts, nh)
    nll = torch.zeros((), dtype=_F64)
    for y in ys:
        J = torch.func.jacfwd(lambda u: lcd_mean(consts, u, dt, nh))(mf)
        mp = lcd_mean(consts, mf, dt, nh)
        Pp = J @ Pf @ J.T + Sigma
        S = H @ Pp @ H + Xi
        K = Pp @ H / S
        pred = H @ mp
        mf = mp + K * (y - pred)
        Pf = Pp - torch.outer(K, K) * S
        sc = torch.sqrt(S)
        nll = nll + (torch.log(2 * math.pi * sc * sc) + (y - pred) ** 2 / (sc * sc)) / 2
    return nll
