// cgp_dispatch.cuh -- runtime (model, d, group size) -> compiled kernel instance.
#pragma once
#include "cgp_cubduo.cuh"
#include "cgp_oct.cuh"

namespace cgp {

template <class T> struct Tag { using type = T; };

// Compiled state dimensions.  Linear models: d in 1..5 (5: rts after ekf_for_kpt with three harmonics).  Chirp family: discrete (LCD) models num_harmonics in 1..5 (d = 4 .. 12), SDE drifts 1..3.
template <class F> int dispatch_disc(const CgpProblem &p, F &&f) {
    if (p.model == CGP_MODEL_LINEAR_DISC) {
        switch (p.d) {
            case 1: return f(Tag<ModelLinearDisc<1>>{});
            case 2: return f(Tag<ModelLinearDisc<2>>{});
            case 3: return f(Tag<ModelLinearDisc<3>>{});
            case 4: return f(Tag<ModelLinearDisc<4>>{});
            case 5: return f(Tag<ModelLinearDisc<5>>{});
            default: return CGP_ERR_UNSUPPORTED;
        }
    }
    if (p.model == CGP_MODEL_LCD) {
        if (p.d != 2 * p.num_harmonics + 2) return CGP_ERR_BAD_ARG;
        switch (p.num_harmonics) {
            case 1: return f(Tag<ModelLCD<1>>{});
            case 2: return f(Tag<ModelLCD<2>>{});
            case 3: return f(Tag<ModelLCD<3>>{});
            case 4: return f(Tag<ModelLCD<4>>{});       // d = 10: real_applications/bats/myotis_myotis_analysis.py:50
            case 5: return f(Tag<ModelLCD<5>>{});       // d = 12
            default: return CGP_ERR_UNSUPPORTED;
        }
    }
    return CGP_ERR_BAD_ARG;
}
template <class F> int dispatch_sde(const CgpProblem &p, F &&f) {
    if (p.model == CGP_MODEL_LINEAR_SDE) {
        switch (p.d) {
            case 1: return f(Tag<ModelLinearSDE<1>>{});
            case 2: return f(Tag<ModelLinearSDE<2>>{});
            case 3: return f(Tag<ModelLinearSDE<3>>{});
            case 4: return f(Tag<ModelLinearSDE<4>>{});
            default: return CGP_ERR_UNSUPPORTED;
        }
    }
    if (p.model == CGP_MODEL_SDE) {
        if (p.d != 2 * p.num_harmonics + 2) return CGP_ERR_BAD_ARG;
        switch (p.num_harmonics) {
            case 1: return f(Tag<ModelSDE<1>>{});
            case 2: return f(Tag<ModelSDE<2>>{});
            case 3: return f(Tag<ModelSDE<3>>{});
            default: return CGP_ERR_UNSUPPORTED;
        }
    }
    return CGP_ERR_BAD_ARG;
}

// Lanes per chirp for the sigma-point kernels: enough lanes for one "work item" each (a point, or with the
// Gauss-Hermite sharing a base index), capped at a warp.
// The sharing specialisation is compiled for Gauss-Hermite order 3 (the reference's default, quadratures.py:157).
inline bool use_share(const CgpProblem &p) {
    if (p.sigma_kind != CGP_SIGMA_GAUSS_HERMITE || p.gh_order != 3) return false;
    if (p.model == CGP_MODEL_LINEAR_DISC || p.model == CGP_MODEL_LINEAR_SDE) return false;
    int64_t n = 1;
    for (int i = 0; i < p.d; i++) n *= p.gh_order;
    return n == p.n_sigma;
}
inline int group_size(const CgpProblem &p, bool share) {
    const int work = share ? p.n_sigma / p.gh_order : p.n_sigma;
    if (work <= 8) return 8;
    if (work <= 16) return 16;
    return 32;
}

inline bool aligned16(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }
inline bool aligned32(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// entry points implemented in the individual translation units
int launch_ekf(const CgpProblem &p, const FilterIO &io, cudaStream_t s);
int launch_ekf_kpt(const CgpProblem &p, const FilterIO &io, cudaStream_t s);
int launch_sgp_filter(const CgpProblem &p, const FilterIO &io, cudaStream_t s);
int launch_cd_ekf(const CgpProblem &p, const FilterIO &io, cudaStream_t s);
int launch_cd_sgp_filter(const CgpProblem &p, const FilterIO &io, cudaStream_t s);
int launch_eks(const CgpProblem &p, const SmootherIO &io, cudaStream_t s);
int launch_sgp_smoother(const CgpProblem &p, const SmootherIO &io, cudaStream_t s);
int launch_sgp_gains(const CgpProblem &p, const SmootherIO &io, cudaStream_t s);
int launch_smoother_sweep(const CgpProblem &p, const SmootherIO &io, cudaStream_t s);
bool sgp_filter_fuses_gains(const CgpProblem &p);
int launch_cd_eks(const CgpProblem &p, const SmootherIO &io, cudaStream_t s);
int launch_cd_sgp_smoother(const CgpProblem &p, const SmootherIO &io, cudaStream_t s);

}  // namespace cgp
