"""Import the UNMODIFIED reference hot-path modules (filters_smoothers, quadratures, models) from
/root/reference on top of oracle/jaxshim (torch-float64 stand-in for the absent jax/jaxlib).

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no /root/reference); used by
tests/golden/make_golden.py to generate the committed golden vectors and by `-m "not gpu"` tests (skipped
when the reference is absent) to validate the C restatement in oracle/.

The reference's package __init__ pulls in out-of-scope modules (classical_methods -> jax.scipy.optimize ...),
so a bare namespace package `chirpgp` is registered first and only the hot-path submodules are imported.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('CHIRPGP_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'jaxshim')


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'chirpgp', 'filters_smoothers.py'))


def load():
    """Returns (filters_smoothers, quadratures, models) modules of the reference, run over the shim."""
    if not available():
        raise RuntimeError('reference sources not present at %s' % REFERENCE_ROOT)
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    name = '_chirpgp_reference'
    if name not in sys.modules:
        pkg = types.ModuleType('chirpgp')
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'chirpgp')]
        saved = sys.modules.get('chirpgp')
        sys.modules['chirpgp'] = pkg
        try:
            fs = importlib.import_module('chirpgp.filters_smoothers')
            qd = importlib.import_module('chirpgp.quadratures')
            md = importlib.import_module('chirpgp.models')
        finally:
            # do not leave the reference registered under the product-adjacent name
            for k in [k for k in sys.modules if k == 'chirpgp' or k.startswith('chirpgp.')]:
                del sys.modules[k]
            if saved is not None:
                sys.modules['chirpgp'] = saved
        holder = types.ModuleType(name)
        holder.fs, holder.qd, holder.md = fs, qd, md
        sys.modules[name] = holder
    h = sys.modules[name]
    return h.fs, h.qd, h.md
