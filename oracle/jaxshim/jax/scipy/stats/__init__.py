import math as _math
import torch as _torch
from ... import numpy as _jnp


class norm:
    @staticmethod
    def logpdf(x, loc=0., scale=1.):
        """Same operation order as jax.scipy.stats.norm.logpdf:
        -(log(2*pi*scale^2) + (x-loc)^2/scale^2) / 2."""
        x, loc, scale = _jnp.asarray(x), _jnp.asarray(loc), _jnp.asarray(scale)
        scale_sqrd = scale * scale
        log_normalizer = _torch.log(2 * _math.pi * scale_sqrd)
        quadratic = (x - loc) * (x - loc) / scale_sqrd
        return (log_normalizer + quadratic) / (-2.)
