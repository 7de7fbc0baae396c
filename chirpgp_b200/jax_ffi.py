"""JAX binding of the C ABI: XLA FFI custom calls wrapped in ``jax.custom_vjp`` so that ``jax.grad`` of the MLE objective
(``demos/ekfs_mle.py:42-49``) keeps working, and ``jax.vmap`` maps to the kernels' native batch axis
(``vmap_method="broadcast_all"``).

EXPERIMENTAL -- IMPORT-GUARDED AND NEVER EXECUTED: jax / jaxlib are not installed in the image this repository was developed in and cannot
be installed (no wheel, no network), so neither this module nor ``csrc/xla_ffi_shim.cc`` has ever been executed.  The
tested binding of the same entry points is ``chirpgp_b200._native`` (ctypes) with ``torch.autograd.Function`` playing the
role of ``custom_vjp`` (``chirpgp_b200.mle``).  Build the shim first (command at the top of csrc/xla_ffi_shim.cc).
"""
import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
    _HAVE_JAX = hasattr(jax, 'ffi')
except ImportError:   # the normal case in this image
    jax = None
    _HAVE_JAX = False

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, 'libchirpgp_b200_xla.so')
_registered = False


def available() -> bool:
    return _HAVE_JAX and os.path.exists(_SHIM)


def _register():
    global _registered
    if _registered:
        return
    if not available():
        raise RuntimeError('chirpgp_b200.jax_ffi needs jax >= 0.4.38 and the XLA shim %s (see csrc/xla_ffi_shim.cc)' % _SHIM)
    lib = ctypes.CDLL(_SHIM)
    for name in ('CgpFilter', 'CgpFilterGains', 'CgpSmoother', 'CgpSmootherSweep', 'CgpEkfNllFwd', 'CgpEkfNllBwd'):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform='CUDA')
    _registered = True


def _f64(shape):
    return jax.ShapeDtypeStruct(shape, jnp.float64)


def _sigma_kind(sig, gh_order) -> int:
    """CGP_SIGMA_* hint of include/chirpgp_b200.h: 1 = Gauss-Hermite table with gh_order nodes per dimension, 2 = cubature
    table (2 d points), 0 = generic.  The hint only selects a kernel specialisation; results do not depend on it."""
    if sig is None:
        return 0
    if gh_order:
        return 1
    n, d = int(sig.w.shape[0]), int(sig.xi.shape[1])
    return 2 if n == 2 * d else 0


def filter_call(fn, model, consts, H, Xi, m0, P0, dt, ys, Qc=None, sig=None, num_harmonics=0, gh_order=0, h_unit_index=-1):
    """fn in {'kf','ekf','sgp_filter','cd_ekf','cd_sgp_filter'}: ys (T,) -> (mfs (T,d), Pfs (T,d,d), n_ell (T,)).
    Wrap in jax.vmap for batches: the custom call receives the broadcast batch axis natively."""
    _register()
    d = m0.shape[-1]
    T = ys.shape[-1]
    lead = ys.shape[:-1]
    empty = jnp.zeros((0,), jnp.float64)
    w, xi = (empty, empty) if sig is None else (jnp.asarray(sig.w), jnp.asarray(sig.xi))
    call = jax.ffi.ffi_call('CgpFilter', (_f64(lead + (T, d)), _f64(lead + (T, d, d)), _f64(lead + (T,))),
                            vmap_method='broadcast_all')
    return call(ys, consts, m0, P0, H, empty if Qc is None else Qc, w, xi, fn=fn, model=int(model),
                num_harmonics=int(num_harmonics), sigma_kind=_sigma_kind(sig, gh_order), gh_order=int(gh_order), ys_repeat=1,
                h_unit_index=int(h_unit_index), Xi=float(Xi), dt=float(dt))


def sgp_filter_smoother_call(model, consts, H, Xi, m0, P0, dt, ys, sig, num_harmonics=0, gh_order=0, h_unit_index=-1):
    """sgp_filter + sgp_smoother as two custom calls that share the smoother workspace (cgp_sgp_filter_gains_f64 +
    cgp_smoother_sweep_f64): ys (T,) -> (mfs, Pfs, n_ell, mss, Pss).  What `filtering(ys)` followed by
    `smoothing(mfs, Pfs)` computes in the demos (demos/ghfs_mle.py:69-85), without evaluating the sigma points twice."""
    _register()
    d = m0.shape[-1]
    T = ys.shape[-1]
    lead = ys.shape[:-1]
    w, xi = jnp.asarray(sig.w), jnp.asarray(sig.xi)
    f = jax.ffi.ffi_call('CgpFilterGains', (_f64(lead + (T, d)), _f64(lead + (T, d, d)), _f64(lead + (T,)),
                                            _f64(lead + (T, 2 * d * d + d))), vmap_method='broadcast_all')
    mfs, Pfs, nell, ws = f(ys, consts, m0, P0, H, w, xi, model=int(model), num_harmonics=int(num_harmonics),
                           sigma_kind=_sigma_kind(sig, gh_order), gh_order=int(gh_order), ys_repeat=1,
                           h_unit_index=int(h_unit_index), Xi=float(Xi), dt=float(dt))
    mss, Pss = jax.ffi.ffi_call('CgpSmootherSweep', (_f64(mfs.shape), _f64(Pfs.shape)), vmap_method='broadcast_all')(mfs, Pfs, ws)
    return mfs, Pfs, nell, mss, Pss


def smoother_call(fn, model, consts, mfs, Pfs, dt, Qc=None, sig=None, num_harmonics=0, gh_order=0):
    _register()
    d = mfs.shape[-1]
    empty = jnp.zeros((0,), jnp.float64)
    w, xi = (empty, empty) if sig is None else (jnp.asarray(sig.w), jnp.asarray(sig.xi))
    ws_shape = mfs.shape[:-1] + (2 * d * d + d,) if fn in ('rts', 'eks', 'sgp_smoother') else (1,)
    call = jax.ffi.ffi_call('CgpSmoother', (_f64(mfs.shape), _f64(Pfs.shape), _f64(ws_shape)), vmap_method='broadcast_all')
    mss, Pss, _ = call(mfs, Pfs, consts, empty if Qc is None else Qc, w, xi, fn=fn, model=int(model),
                       num_harmonics=int(num_harmonics), sigma_kind=_sigma_kind(sig, gh_order), gh_order=int(gh_order), dt=float(dt))
    return mss, Pss


def make_ekf_nll(num_harmonics: int, Xi: float, dt: float, B: int, T: int, h_unit_index: int = 1, ckpt_every: int = 16):
    """Returns nll(consts (B,NC), m0 (B,d), P0 (B,d,d), H (d,), ys (B,T)) -> (B,) with a custom VJP that launches the
    adjoint kernel.  (For a cotangent on the whole n_ell (T,) output bind cgp_ekf_nll_path_{fwd,bwd}_f64 the same way: the
    backward rule passes the reversed cumulative sum of the cotangent as step weights; chirpgp_b200.mle._EkfNllPath is the
    tested torch version.)  Cotangents for (consts, m0, P0); zeros for H, ys (constants in every caller of the reference).

    The workspace (checkpoints, scheduling words, per-warp scratch) is WRITTEN by both kernels: the forward call returns it
    as a result, the backward call takes it as an operand that is aliased to a result (input_output_aliases), so XLA never
    sees an input buffer being mutated behind its back."""
    _register()
    from . import _native as N
    from .filters_smoothers import _problem
    d = 2 * num_harmonics + 2
    p = _problem(B, T, N.CGP_MODEL_LCD, d, num_harmonics, None, 0, None, 0, None, 0, None, None, 0, None, Xi, dt, 1, h_unit_index)
    ws_len = N.lib().cgp_ekf_nll_workspace_bytes(ctypes.byref(p), int(ckpt_every)) // 8
    attrs = dict(num_harmonics=int(num_harmonics), ys_repeat=1, h_unit_index=int(h_unit_index), ckpt_every=int(ckpt_every),
                 Xi=float(Xi), dt=float(dt))

    def _fwd_call(consts, m0, P0, H, ys):
        return jax.ffi.ffi_call('CgpEkfNllFwd', (_f64((B,)), _f64((ws_len,))))(ys, consts, m0, P0, H, **attrs)

    @jax.custom_vjp
    def nll(consts, m0, P0, H, ys):
        return _fwd_call(consts, m0, P0, H, ys)[0]

    def fwd(consts, m0, P0, H, ys):
        out, ws = _fwd_call(consts, m0, P0, H, ys)
        return out, (consts, m0, P0, H, ys, ws)

    def bwd(res, nll_bar):
        consts, m0, P0, H, ys, ws = res
        cb, mb, Pb, _, _ = jax.ffi.ffi_call(
            'CgpEkfNllBwd', (_f64(consts.shape), _f64(m0.shape), _f64(P0.shape), _f64((B,)), _f64((ws_len,))),
            input_output_aliases={6: 4})(ys, consts, m0, P0, H, nll_bar, ws, **attrs)
        return cb, mb, Pb, jnp.zeros_like(H), jnp.zeros_like(ys)

    nll.defvjp(fwd, bwd)
    return nll
