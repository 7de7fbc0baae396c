"""ctypes binding of the C ABI in include/chirpgp_b200.h (the in-tree ``libchirpgp_b200.so``).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libchirpgp_b200.so')
CSRC = os.path.join(_HERE, 'csrc')

CGP_MODEL_LINEAR_DISC, CGP_MODEL_LCD, CGP_MODEL_LINEAR_SDE, CGP_MODEL_SDE, CGP_MODEL_KPT = 0, 1, 2, 3, 4
CGP_SIGMA_GENERIC, CGP_SIGMA_GAUSS_HERMITE, CGP_SIGMA_CUBATURE = 0, 1, 2
CGP_H_HARMONIC = -2            # CgpProblem.h_unit_index: H = [0 1 0 1 ... 0 0] (include/chirpgp_b200.h)
ABI_VERSION = 3

_ERRORS = {-1: 'CGP_ERR_BAD_ARG', -2: 'CGP_ERR_UNSUPPORTED (no kernel compiled for this model / state dimension)',
           -3: 'CGP_ERR_WORKSPACE'}


class CgpProblem(C.Structure):
    _fields_ = [
        ('B', C.c_int64), ('T', C.c_int64),
        ('model', C.c_int32), ('d', C.c_int32), ('num_harmonics', C.c_int32), ('n_sigma', C.c_int32),
        ('sigma_kind', C.c_int32), ('gh_order', C.c_int32),
        ('ys_repeat', C.c_int64),
        ('h_unit_index', C.c_int32), ('in_flight', C.c_int32),
        ('consts', C.c_void_p), ('consts_stride', C.c_int64),
        ('m0', C.c_void_p), ('m0_stride', C.c_int64),
        ('P0', C.c_void_p), ('P0_stride', C.c_int64),
        ('H', C.c_void_p),
        ('Qc', C.c_void_p), ('Qc_stride', C.c_int64),
        ('sig_w', C.c_void_p), ('sig_xi', C.c_void_p),
        ('Xi', C.c_double), ('dt', C.c_double),
    ]


FILTER_FUNCS = ('kf', 'ekf', 'ekf_for_kpt', 'sgp_filter', 'cd_ekf', 'cd_sgp_filter')
SMOOTHER_FUNCS = ('rts', 'eks', 'sgp_smoother', 'cd_eks', 'cd_sgp_smoother')
EXPORTED = (['cgp_abi_version', 'cgp_workspace_bytes'] + ['cgp_%s_f64' % f for f in FILTER_FUNCS + SMOOTHER_FUNCS]
            + ['cgp_ekf_nll_default_ckpt', 'cgp_ekf_nll_workspace_bytes', 'cgp_ekf_nll_fwd_f64', 'cgp_ekf_nll_bwd_f64', 'cgp_ekf_nll_bwd_sym_f64',
               'cgp_ekf_nll_path_workspace_bytes', 'cgp_ekf_nll_path_fwd_f64', 'cgp_ekf_nll_path_bwd_f64', 'cgp_filter_nll_tangent_f64',
               'cgp_sgp_filter_gains_fused', 'cgp_sgp_filter_gains_f64', 'cgp_smoother_sweep_f64',
               'cgp_gaussian_expectation_softplus_f64', 'cgp_simulate_f64', 'cgp_test_philox', 'cgp_test_normals',
               'cgp_bench_dfma', 'cgp_test_math'])

_lib = None


def build(force: bool = False, jobs: int = 8) -> str:
    """Compile every CUDA source for sm_100a into chirpgp_b200/libchirpgp_b200.so (nvcc cross-compiles
    without a GPU)."""
    args = ['make', '-C', CSRC, '-j%d' % jobs]
    if force:
        args.append('-B')
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('chirpgp_b200: %s is missing -- run `python -c "import __graft_entry__ as g; '
                               'g.build()"` (there is no CPU fallback)' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.cgp_abi_version.restype = C.c_int
        if L.cgp_abi_version() != ABI_VERSION:
            raise RuntimeError('chirpgp_b200: stale libchirpgp_b200.so (ABI %d, expected %d); rebuild'
                               % (L.cgp_abi_version(), ABI_VERSION))
        L.cgp_workspace_bytes.restype = C.c_size_t
        L.cgp_workspace_bytes.argtypes = [C.c_char_p, C.POINTER(CgpProblem)]
        for f in FILTER_FUNCS:
            fn = getattr(L, 'cgp_%s_f64' % f)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(CgpProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        for f in SMOOTHER_FUNCS:
            fn = getattr(L, 'cgp_%s_f64' % f)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(CgpProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_size_t, C.c_void_p]
        L.cgp_sgp_filter_gains_fused.restype = C.c_int
        L.cgp_sgp_filter_gains_fused.argtypes = [C.POINTER(CgpProblem)]
        L.cgp_sgp_filter_gains_f64.restype = C.c_int
        L.cgp_sgp_filter_gains_f64.argtypes = [C.POINTER(CgpProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_size_t, C.c_void_p]
        L.cgp_smoother_sweep_f64.restype = C.c_int
        L.cgp_smoother_sweep_f64.argtypes = [C.POINTER(CgpProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_void_p]
        L.cgp_ekf_nll_default_ckpt.restype = C.c_int64
        L.cgp_ekf_nll_default_ckpt.argtypes = [C.c_int64]
        L.cgp_ekf_nll_workspace_bytes.restype = C.c_size_t
        L.cgp_ekf_nll_workspace_bytes.argtypes = [C.POINTER(CgpProblem), C.c_int64]
        L.cgp_ekf_nll_path_workspace_bytes.restype = C.c_size_t
        L.cgp_ekf_nll_path_workspace_bytes.argtypes = [C.POINTER(CgpProblem), C.c_int64]
        L.cgp_ekf_nll_fwd_f64.restype = C.c_int
        L.cgp_ekf_nll_fwd_f64.argtypes = [C.POINTER(CgpProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64,
                                          C.c_void_p]
        L.cgp_ekf_nll_bwd_f64.restype = C.c_int
        L.cgp_ekf_nll_bwd_f64.argtypes = [C.POINTER(CgpProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cgp_ekf_nll_bwd_sym_f64.restype = C.c_int
        L.cgp_ekf_nll_bwd_sym_f64.argtypes = L.cgp_ekf_nll_bwd_f64.argtypes
        L.cgp_gaussian_expectation_softplus_f64.restype = C.c_int
        L.cgp_gaussian_expectation_softplus_f64.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                                            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.cgp_simulate_f64.restype = C.c_int
        L.cgp_simulate_f64.argtypes = [C.POINTER(CgpProblem), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        L.cgp_test_philox.restype = C.c_int
        L.cgp_test_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p, C.c_void_p]
        L.cgp_test_normals.restype = C.c_int
        L.cgp_test_normals.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p]
        L.cgp_ekf_nll_path_fwd_f64.restype = C.c_int
        L.cgp_ekf_nll_path_fwd_f64.argtypes = L.cgp_ekf_nll_fwd_f64.argtypes
        L.cgp_ekf_nll_path_bwd_f64.restype = C.c_int
        L.cgp_ekf_nll_path_bwd_f64.argtypes = L.cgp_ekf_nll_bwd_f64.argtypes
        L.cgp_filter_nll_tangent_f64.restype = C.c_int
        L.cgp_filter_nll_tangent_f64.argtypes = [C.c_char_p, C.POINTER(CgpProblem), C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cgp_bench_dfma.restype = C.c_double
        L.cgp_bench_dfma.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cgp_test_math.restype = C.c_int
        L.cgp_test_math.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        if rc == -2:
            raise NotImplementedError('chirpgp_b200.%s: %s' % (what, _ERRORS[rc]))
        raise ValueError('chirpgp_b200.%s: %s' % (what, _ERRORS.get(rc, 'error %d' % rc)))
    raise RuntimeError('chirpgp_b200.%s: CUDA error %d' % (what, rc))
