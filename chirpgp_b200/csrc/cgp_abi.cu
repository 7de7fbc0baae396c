// cgp_abi.cu -- extern "C" entry points declared in include/chirpgp_b200.h: argument validation + dispatch.
#include <string.h>
#include "cgp_dispatch.cuh"

using namespace cgp;

namespace {

int check_common(const CgpProblem *p) {
    if (!p) return CGP_ERR_BAD_ARG;
    if (p->B < 1 || p->T < 1 || p->d < 1) return CGP_ERR_BAD_ARG;
    if (!p->consts) return CGP_ERR_BAD_ARG;
    return 0;
}
int check_filter(const CgpProblem *p, const double *ys, double *mfs, double *Pfs) {
    int rc = check_common(p);
    if (rc) return rc;
    if (!ys || !p->m0 || !p->P0 || !p->H) return CGP_ERR_BAD_ARG;
    if ((mfs == nullptr) != (Pfs == nullptr)) return CGP_ERR_BAD_ARG;
    if (p->ys_repeat < 1) return CGP_ERR_BAD_ARG;
    return 0;
}
int check_sigma(const CgpProblem *p) {
    if (p->n_sigma < 1 || !p->sig_w || !p->sig_xi) return CGP_ERR_BAD_ARG;
    return 0;
}
int check_smoother(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss) {
    int rc = check_common(p);
    if (rc) return rc;
    if (!mfs || !Pfs || !mss || !Pss) return CGP_ERR_BAD_ARG;
    if (mss == mfs || Pss == Pfs) return CGP_ERR_BAD_ARG;
    return 0;
}
size_t disc_ws_bytes(const CgpProblem *p) {
    return (size_t)p->B * (size_t)p->T * (size_t)ws_record_doubles(p->d) * sizeof(double);
}

}  // namespace

extern "C" {

int cgp_abi_version(void) { return CGP_ABI_VERSION; }

int cgp_kf_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell, int last_only, void *stream) {
    int rc = check_filter(p, ys, mfs, Pfs);
    if (rc) return rc;
    if (p->model != CGP_MODEL_LINEAR_DISC) return CGP_ERR_BAD_ARG;
    return launch_ekf(*p, FilterIO{ys, mfs, Pfs, nell, last_only}, (cudaStream_t)stream);
}
int cgp_ekf_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell, int last_only, void *stream) {
    int rc = check_filter(p, ys, mfs, Pfs);
    if (rc) return rc;
    return launch_ekf(*p, FilterIO{ys, mfs, Pfs, nell, last_only}, (cudaStream_t)stream);
}
int cgp_ekf_for_kpt_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell, int last_only,
                        void *stream) {
    if (!p || p->B < 1 || p->T < 1 || !p->consts || !ys || !p->m0 || !p->P0 || p->ys_repeat < 1) return CGP_ERR_BAD_ARG;
    if ((mfs == nullptr) != (Pfs == nullptr)) return CGP_ERR_BAD_ARG;
    if (p->model != CGP_MODEL_KPT || p->d != p->num_harmonics + 2) return CGP_ERR_BAD_ARG;
    return launch_ekf_kpt(*p, FilterIO{ys, mfs, Pfs, nell, last_only}, (cudaStream_t)stream);
}
int cgp_sgp_filter_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell, int last_only,
                       void *stream) {
    int rc = check_filter(p, ys, mfs, Pfs);
    if (rc) return rc;
    if ((rc = check_sigma(p))) return rc;
    return launch_sgp_filter(*p, FilterIO{ys, mfs, Pfs, nell, last_only}, (cudaStream_t)stream);
}
int cgp_cd_ekf_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell, int last_only,
                   void *stream) {
    int rc = check_filter(p, ys, mfs, Pfs);
    if (rc) return rc;
    if (!p->Qc) return CGP_ERR_BAD_ARG;
    return launch_cd_ekf(*p, FilterIO{ys, mfs, Pfs, nell, last_only}, (cudaStream_t)stream);
}
int cgp_cd_sgp_filter_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell, int last_only,
                          void *stream) {
    int rc = check_filter(p, ys, mfs, Pfs);
    if (rc) return rc;
    if ((rc = check_sigma(p))) return rc;
    if (!p->Qc) return CGP_ERR_BAD_ARG;
    return launch_cd_sgp_filter(*p, FilterIO{ys, mfs, Pfs, nell, last_only}, (cudaStream_t)stream);
}

size_t cgp_workspace_bytes(const char *fn, const CgpProblem *p) {
    if (!fn || !p) return 0;
    if (!strcmp(fn, "rts") || !strcmp(fn, "eks") || !strcmp(fn, "sgp_smoother") || !strcmp(fn, "sgp_filter_gains") ||
        !strcmp(fn, "smoother_sweep"))
        return disc_ws_bytes(p);
    return 0;
}

int cgp_sgp_filter_gains_fused(const CgpProblem *p) {
    if (!p) return 0;
    return sgp_filter_fuses_gains(*p) ? 1 : 0;
}
int cgp_sgp_filter_gains_f64(const CgpProblem *p, const double *ys, double *mfs, double *Pfs, double *nell, int last_only,
                             void *ws, size_t ws_bytes, void *stream) {
    int rc = check_filter(p, ys, mfs, Pfs);
    if (rc) return rc;
    if ((rc = check_sigma(p))) return rc;
    if (!mfs || !Pfs) return CGP_ERR_BAD_ARG;
    if (!ws || ws_bytes < disc_ws_bytes(p)) return CGP_ERR_WORKSPACE;
    if (sgp_filter_fuses_gains(*p) && aligned16(ws))
        return launch_sgp_filter(*p, FilterIO{ys, mfs, Pfs, nell, last_only, (double *)ws}, (cudaStream_t)stream);
    rc = launch_sgp_filter(*p, FilterIO{ys, mfs, Pfs, nell, last_only}, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_sgp_gains(*p, SmootherIO{mfs, Pfs, nullptr, nullptr, (double *)ws}, (cudaStream_t)stream);
}
int cgp_smoother_sweep_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss, void *ws,
                           size_t ws_bytes, void *stream) {
    if (!p || p->B < 1 || p->T < 1 || p->d < 1) return CGP_ERR_BAD_ARG;
    if (!mfs || !Pfs || !mss || !Pss || mss == mfs || Pss == Pfs) return CGP_ERR_BAD_ARG;
    if (!ws || ws_bytes < disc_ws_bytes(p)) return CGP_ERR_WORKSPACE;
    return launch_smoother_sweep(*p, SmootherIO{mfs, Pfs, mss, Pss, (double *)ws}, (cudaStream_t)stream);
}

int cgp_rts_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss, void *ws,
                size_t ws_bytes, void *stream) {
    int rc = check_smoother(p, mfs, Pfs, mss, Pss);
    if (rc) return rc;
    if (p->model != CGP_MODEL_LINEAR_DISC) return CGP_ERR_BAD_ARG;
    if (!ws || ws_bytes < disc_ws_bytes(p)) return CGP_ERR_WORKSPACE;
    return launch_eks(*p, SmootherIO{mfs, Pfs, mss, Pss, (double *)ws}, (cudaStream_t)stream);
}
int cgp_eks_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss, void *ws,
                size_t ws_bytes, void *stream) {
    int rc = check_smoother(p, mfs, Pfs, mss, Pss);
    if (rc) return rc;
    if (!ws || ws_bytes < disc_ws_bytes(p)) return CGP_ERR_WORKSPACE;
    return launch_eks(*p, SmootherIO{mfs, Pfs, mss, Pss, (double *)ws}, (cudaStream_t)stream);
}
int cgp_sgp_smoother_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss, void *ws,
                         size_t ws_bytes, void *stream) {
    int rc = check_smoother(p, mfs, Pfs, mss, Pss);
    if (rc) return rc;
    if ((rc = check_sigma(p))) return rc;
    if (!ws || ws_bytes < disc_ws_bytes(p)) return CGP_ERR_WORKSPACE;
    return launch_sgp_smoother(*p, SmootherIO{mfs, Pfs, mss, Pss, (double *)ws}, (cudaStream_t)stream);
}
int cgp_cd_eks_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss, void *ws,
                   size_t ws_bytes, void *stream) {
    int rc = check_smoother(p, mfs, Pfs, mss, Pss);
    if (rc) return rc;
    if (!p->Qc) return CGP_ERR_BAD_ARG;
    return launch_cd_eks(*p, SmootherIO{mfs, Pfs, mss, Pss, nullptr}, (cudaStream_t)stream);
}
int cgp_cd_sgp_smoother_f64(const CgpProblem *p, const double *mfs, const double *Pfs, double *mss, double *Pss, void *ws,
                            size_t ws_bytes, void *stream) {
    int rc = check_smoother(p, mfs, Pfs, mss, Pss);
    if (rc) return rc;
    if ((rc = check_sigma(p))) return rc;
    if (!p->Qc) return CGP_ERR_BAD_ARG;
    return launch_cd_sgp_smoother(*p, SmootherIO{mfs, Pfs, mss, Pss, nullptr}, (cudaStream_t)stream);
}

}  // extern "C"
