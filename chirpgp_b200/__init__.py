"""chirpgp_b200 -- B200-native (sm_100a) batched Gaussian filters and smoothers for chirp / instantaneous-
frequency estimation, behind the reference's ``chirpgp`` Python interface.

Re-exports follow /root/reference/chirpgp/__init__.py:1-6 (hot-path names only)."""
from .filters_smoothers import (kf, rts, ekf, eks, cd_ekf, cd_eks, sgp_filter, sgp_smoother, cd_sgp_filter,  # noqa
                                cd_sgp_smoother, ekf_for_kpt, sgp_filter_smoother, ekf_smoother, cd_ekf_smoother,
                                cd_sgp_filter_smoother, filter_smoother_batches)
from .models import (g, g_inv, model_chirp, model_harmonic_chirp, model_lascala, disc_chirp_lcd,  # noqa: F401
                     disc_harmonic_chirp_lcd, disc_model_lascala_lcd, disc_m32, build_chirp_model,
                     build_harmonic_chirp_model, build_lascala_model, build_kpt_chirp_model, posterior_cramer_rao,
                     LinearDisc, LinearSDE)
from .quadratures import SigmaPoints, gaussian_expectation  # noqa: F401
from .toymodels import (gen_chirp, gen_harmonic_chirp, constant_mag, damped_exp_mag, random_ou_mag,  # noqa: F401
                        affine_freq, polynomial_freq, meow_freq)

__version__ = '0.1.0'
