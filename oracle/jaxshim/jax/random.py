"""jax.random stand-in. NOT threefry: keys are numpy SeedSequence entropy lists; streams differ from real JAX."""
import numpy as _np
import torch as _torch


def PRNGKey(seed):
    return _np.array([0, int(seed)], dtype=_np.uint32)


def split(key, num=2):
    ss = _np.random.SeedSequence([int(k) for k in _np.asarray(key).ravel()])
    return [_np.array(c.generate_state(2), dtype=_np.uint32) for c in ss.spawn(num)]


def normal(key, shape=()):
    rng = _np.random.default_rng([int(k) for k in _np.asarray(key).ravel()])
    return _torch.as_tensor(rng.standard_normal(shape))
