"""Synthetic chirp generators (host side, NumPy) -- the input side of the BASELINE configs.

Mirrors ``chirpgp.toymodels`` (/root/reference/chirpgp/toymodels.py): ``gen_chirp`` :37-70,
``gen_harmonic_chirp`` :73-104, ``constant_mag`` :122, ``damped_exp_mag`` :133, ``random_ou_mag`` :144-167,
``affine_freq`` :170, ``polynomial_freq`` :194, ``meow_freq`` :226-268.

The reference draws noise from JAX threefry keys; JAX is not available here, so ``random_ou_mag`` takes a
``numpy.random.Generator`` (or an int seed) instead of a key.  The OU recursion is the one
``simulate_sde(..., const_diag_cov=True)`` performs (tools.py:119-170):
``x0 ~ N(0, sigma^2)``, ``x_{k+1} = exp(-dt/ell) x_k + sqrt(sigma^2 (1 - exp(-2 dt/ell))) eps_k``.
"""
import math
from typing import Callable, List, Sequence, Tuple, Union

import numpy as np

__all__ = ['gen_chirp', 'gen_harmonic_chirp', 'gen_chirp_envelope', 'constant_mag', 'damped_exp_mag',
           'random_ou_mag', 'affine_freq', 'polynomial_freq', 'meow_freq', 'synthetic_batch']


def gen_chirp(ts, magnitude_func, phase_func, base_phase: float = 0.):
    ts = np.asarray(ts, dtype=np.float64)
    return magnitude_func(ts) * np.sin(base_phase + 2 * math.pi * phase_func(ts))


def gen_harmonic_chirp(ts, magnitude_funcs: Sequence[Callable], fundamental_phase_func, base_phase: float = 0.):
    ts = np.asarray(ts, dtype=np.float64)
    ys = np.zeros_like(ts)
    for i, mag_func in enumerate(magnitude_funcs):
        ys = ys + mag_func(ts) * np.sin(base_phase + (i + 1) * 2 * math.pi * fundamental_phase_func(ts))
    return ys


def gen_chirp_envelope(ts, magnitude_func, phase_func, base_phase: float = 0.):
    ts = np.asarray(ts, dtype=np.float64)
    return magnitude_func(ts) * np.exp((base_phase + 2 * math.pi * phase_func(ts)) * 1.j)


def constant_mag(b: float):
    return lambda ts: np.ones_like(np.asarray(ts, dtype=np.float64)) * b


def damped_exp_mag(damp_rate: float):
    return lambda ts: np.exp(-damp_rate * np.asarray(ts, dtype=np.float64))


def random_ou_mag(ell: float, sigma: float, key: Union[int, np.random.Generator]):
    rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)

    def generate_ou(ts):
        ts = np.asarray(ts, dtype=np.float64)
        dt = np.diff(ts)[0]
        n = ts.size
        decay = math.exp(-dt / ell)
        std = math.sqrt(sigma ** 2 * (1 - math.exp(-2 * dt / ell)))
        x = sigma * rng.standard_normal()
        eps = rng.standard_normal(n)
        out = np.empty(n)
        for k in range(n):
            x = decay * x + std * eps[k]
            out[k] = x
        return out

    return generate_ou


def affine_freq(a: float, b: float):
    return (lambda ts: a * np.asarray(ts) + b), (lambda ts: 0.5 * a * np.asarray(ts) ** 2 + b * np.asarray(ts))


def polynomial_freq(coeffs: List[float]):
    # NB the reference accumulates into jnp.empty_like (zeros under XLA); zeros here.
    def freq_func(ts):
        ts = np.asarray(ts, dtype=np.float64)
        f = np.zeros_like(ts)
        for k, c in enumerate(coeffs):
            f = f + c * ts ** k
        return f

    def phase_func(ts):
        ts = np.asarray(ts, dtype=np.float64)
        p = np.zeros_like(ts)
        for k, c in enumerate(coeffs):
            p = p + c / (k + 1) * ts ** (k + 1)
        return p

    return freq_func, phase_func


def meow_freq(mag: float = 500, scale: float = 5, offset: float = 5.5) -> Tuple[Callable, Callable]:
    """Valid on (0, pi): phase a exp(-b / sin t) + c t and its derivative."""

    def freq_func(ts):
        ts = np.asarray(ts, dtype=np.float64)
        return mag * scale * np.cos(ts) / (np.sin(ts) ** 2) * np.exp(-scale / np.sin(ts)) + offset

    def phase_func(ts):
        ts = np.asarray(ts, dtype=np.float64)
        return mag * np.exp(-scale / np.sin(ts)) + offset * ts

    return freq_func, phase_func


def synthetic_batch(B: int, T: int, dt: float, Xi: float = 0.1, num_harmonics: int = 1, seed: int = 2,
                    offset: float = 8.):
    """The synthetic (B, T) measurement batch of SURVEY 8(d): chirp i uses magnitude i % 3 of
    (constant 1, damped exp 0.3, OU(1, 1)), the `meow` phase with offset 8, and noise from
    ``default_rng([seed, i])``.  Returns (ts (T,), ys (B, T), true_freq (T,))."""
    ts = np.linspace(dt, dt * T, T)
    freq_func, phase_func = meow_freq(offset=offset)
    ys = np.empty((B, T))
    for i in range(B):
        rng = np.random.default_rng([seed, i])
        kind = i % 3
        if kind == 0:
            mag = constant_mag(1.)
        elif kind == 1:
            mag = damped_exp_mag(0.3)
        else:
            mag = random_ou_mag(1., 1., rng)
        if num_harmonics == 1:
            clean = gen_chirp(ts, mag, phase_func)
        else:
            m = mag(ts)
            clean = gen_harmonic_chirp(ts, [lambda _t, _m=m: _m] * num_harmonics, phase_func)
        ys[i] = clean + math.sqrt(Xi) * rng.standard_normal(T)
    return ts, ys, freq_func(ts)
