// xla_ffi_shim.cc -- XLA FFI (jax.ffi) handlers over the C ABI of include/chirpgp_b200.h.
//
// NOT part of the default build: it needs the jaxlib headers (`jax.ffi.include_dir()`), and JAX cannot be installed
// in the image this repository was developed in (no wheel, no network), so this translation unit has never been
// compiled or run.  It documents -- in code -- the binding a maintainer adds to call the kernels from JAX:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include xla_ffi_shim.cc -L.. -lchirpgp_b200 -o ../libchirpgp_b200_xla.so
//
// One handler per family; the reference function is selected by the `fn` attribute.  XLA hands device buffers and
// the stream; shapes carry B, T, d (vmap with vmap_method="broadcast_all" simply adds the leading batch axis).
#include <cstdint>
#include <string_view>

#include "xla/ffi/api/ffi.h"
#include "../../include/chirpgp_b200.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using RF64 = ffi::ResultBuffer<ffi::F64>;

namespace {

struct Attrs {
    int64_t model, num_harmonics, sigma_kind, gh_order, ys_repeat, h_unit_index;
    double Xi, dt;
};

CgpProblem make_problem(const Attrs &a, int64_t B, int64_t T, int64_t d, const F64 &consts, const double *m0, int64_t m0_rows,
                        const double *P0, int64_t P0_rows, const double *H, const double *Qc, int64_t Qc_rows, const double *w,
                        const double *xi, int64_t n_sigma) {
    CgpProblem p{};
    p.B = B; p.T = T; p.model = (int32_t)a.model; p.d = (int32_t)d; p.num_harmonics = (int32_t)a.num_harmonics;
    p.n_sigma = (int32_t)n_sigma; p.sigma_kind = (int32_t)a.sigma_kind; p.gh_order = (int32_t)a.gh_order;
    p.ys_repeat = a.ys_repeat; p.h_unit_index = (int32_t)a.h_unit_index;
    const auto cd = consts.dimensions();
    const int64_t nc = cd.back();
    p.consts = consts.typed_data();
    p.consts_stride = (consts.element_count() / nc) > 1 ? nc : 0;
    p.m0 = m0; p.m0_stride = m0_rows > 1 ? d : 0;
    p.P0 = P0; p.P0_stride = P0_rows > 1 ? d * d : 0;
    p.H = H;
    p.Qc = Qc; p.Qc_stride = Qc_rows > 1 ? d * d : 0;
    p.sig_w = w; p.sig_xi = xi;
    p.Xi = a.Xi; p.dt = a.dt;
    return p;
}

ffi::Error status(int rc, const char *what) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, what);
}

// filters: ys [B,T] -> mfs [B,T,d], Pfs [B,T,d,d], nell [B,T]
ffi::Error FilterImpl(cudaStream_t stream, std::string_view fn, int64_t model, int64_t num_harmonics, int64_t sigma_kind,
                      int64_t gh_order, int64_t ys_repeat, int64_t h_unit_index, double Xi, double dt, F64 ys, F64 consts, F64 m0,
                      F64 P0, F64 H, F64 Qc, F64 sig_w, F64 sig_xi, RF64 mfs, RF64 Pfs, RF64 nell) {
    const auto od = mfs->dimensions();
    const int64_t d = od.back(), T = od[od.size() - 2], B = mfs->element_count() / (T * d);
    const Attrs a{model, num_harmonics, sigma_kind, gh_order, ys_repeat, h_unit_index, Xi, dt};
    const int64_t n = sig_w.element_count();
    CgpProblem p = make_problem(a, B, T, d, consts, m0.typed_data(), m0.element_count() / d, P0.typed_data(),
                                P0.element_count() / (d * d), H.typed_data(), Qc.typed_data(), Qc.element_count() / (d * d),
                                n ? sig_w.typed_data() : nullptr, n ? sig_xi.typed_data() : nullptr, n);
    int rc = CGP_ERR_UNSUPPORTED;
    double *o0 = mfs->typed_data(), *o1 = Pfs->typed_data(), *o2 = nell->typed_data();
    if (fn == "kf") rc = cgp_kf_f64(&p, ys.typed_data(), o0, o1, o2, 0, stream);
    else if (fn == "ekf") rc = cgp_ekf_f64(&p, ys.typed_data(), o0, o1, o2, 0, stream);
    else if (fn == "sgp_filter") rc = cgp_sgp_filter_f64(&p, ys.typed_data(), o0, o1, o2, 0, stream);
    else if (fn == "cd_ekf") rc = cgp_cd_ekf_f64(&p, ys.typed_data(), o0, o1, o2, 0, stream);
    else if (fn == "cd_sgp_filter") rc = cgp_cd_sgp_filter_f64(&p, ys.typed_data(), o0, o1, o2, 0, stream);
    return status(rc, "chirpgp_b200 filter");
}

// sgp_filter that also fills the smoother workspace: ys [B,T] -> mfs, Pfs, nell, ws [B,T,2d^2+d]
ffi::Error FilterGainsImpl(cudaStream_t stream, int64_t model, int64_t num_harmonics, int64_t sigma_kind, int64_t gh_order,
                           int64_t ys_repeat, int64_t h_unit_index, double Xi, double dt, F64 ys, F64 consts, F64 m0, F64 P0,
                           F64 H, F64 sig_w, F64 sig_xi, RF64 mfs, RF64 Pfs, RF64 nell, RF64 ws) {
    const auto od = mfs->dimensions();
    const int64_t d = od.back(), T = od[od.size() - 2], B = mfs->element_count() / (T * d);
    const Attrs a{model, num_harmonics, sigma_kind, gh_order, ys_repeat, h_unit_index, Xi, dt};
    CgpProblem p = make_problem(a, B, T, d, consts, m0.typed_data(), m0.element_count() / d, P0.typed_data(),
                                P0.element_count() / (d * d), H.typed_data(), nullptr, 1, sig_w.typed_data(), sig_xi.typed_data(),
                                sig_w.element_count());
    return status(cgp_sgp_filter_gains_f64(&p, ys.typed_data(), mfs->typed_data(), Pfs->typed_data(), nell->typed_data(), 0,
                                           ws->typed_data(), ws->element_count() * sizeof(double), stream),
                  "chirpgp_b200 sgp_filter_gains");
}
// sequential half of rts / eks / sgp_smoother on a filled workspace: mfs, Pfs, ws -> mss, Pss
ffi::Error SweepImpl(cudaStream_t stream, F64 mfs, F64 Pfs, F64 ws, RF64 mss, RF64 Pss) {
    const auto od = mfs.dimensions();
    const int64_t d = od.back(), T = od[od.size() - 2], B = mfs.element_count() / (T * d);
    CgpProblem p{};
    p.B = B; p.T = T; p.d = (int32_t)d;
    return status(cgp_smoother_sweep_f64(&p, mfs.typed_data(), Pfs.typed_data(), mss->typed_data(), Pss->typed_data(),
                                         const_cast<double *>(ws.typed_data()), ws.element_count() * sizeof(double), stream),
                  "chirpgp_b200 smoother_sweep");
}

// smoothers: mfs, Pfs -> mss, Pss; `ws` is an extra result buffer XLA allocates as scratch ([B,T,2d^2+d] or [1])
ffi::Error SmootherImpl(cudaStream_t stream, std::string_view fn, int64_t model, int64_t num_harmonics, int64_t sigma_kind,
                        int64_t gh_order, double dt, F64 mfs, F64 Pfs, F64 consts, F64 Qc, F64 sig_w, F64 sig_xi, RF64 mss,
                        RF64 Pss, RF64 ws) {
    const auto od = mfs.dimensions();
    const int64_t d = od.back(), T = od[od.size() - 2], B = mfs.element_count() / (T * d);
    const Attrs a{model, num_harmonics, sigma_kind, gh_order, 1, -1, 0., dt};
    const int64_t n = sig_w.element_count();
    CgpProblem p = make_problem(a, B, T, d, consts, nullptr, 1, nullptr, 1, nullptr, Qc.typed_data(), Qc.element_count() / (d * d),
                                n ? sig_w.typed_data() : nullptr, n ? sig_xi.typed_data() : nullptr, n);
    const size_t wb = ws->element_count() * sizeof(double);
    int rc = CGP_ERR_UNSUPPORTED;
    if (fn == "rts") rc = cgp_rts_f64(&p, mfs.typed_data(), Pfs.typed_data(), mss->typed_data(), Pss->typed_data(), ws->typed_data(), wb, stream);
    else if (fn == "eks") rc = cgp_eks_f64(&p, mfs.typed_data(), Pfs.typed_data(), mss->typed_data(), Pss->typed_data(), ws->typed_data(), wb, stream);
    else if (fn == "sgp_smoother") rc = cgp_sgp_smoother_f64(&p, mfs.typed_data(), Pfs.typed_data(), mss->typed_data(), Pss->typed_data(), ws->typed_data(), wb, stream);
    else if (fn == "cd_eks") rc = cgp_cd_eks_f64(&p, mfs.typed_data(), Pfs.typed_data(), mss->typed_data(), Pss->typed_data(), nullptr, 0, stream);
    else if (fn == "cd_sgp_smoother") rc = cgp_cd_sgp_smoother_f64(&p, mfs.typed_data(), Pfs.typed_data(), mss->typed_data(), Pss->typed_data(), nullptr, 0, stream);
    return status(rc, "chirpgp_b200 smoother");
}

// MLE: forward (nll + checkpoints as a residual) and the adjoint
ffi::Error NllFwdImpl(cudaStream_t stream, int64_t num_harmonics, int64_t ys_repeat, int64_t h_unit_index, int64_t ckpt_every,
                      double Xi, double dt, F64 ys, F64 consts, F64 m0, F64 P0, F64 H, RF64 nll, RF64 ws) {
    const int64_t d = 2 * num_harmonics + 2, B = nll->element_count(), T = ys.dimensions().back();
    const Attrs a{CGP_MODEL_LCD, num_harmonics, 0, 0, ys_repeat, h_unit_index, Xi, dt};
    CgpProblem p = make_problem(a, B, T, d, consts, m0.typed_data(), m0.element_count() / d, P0.typed_data(),
                                P0.element_count() / (d * d), H.typed_data(), nullptr, 1, nullptr, nullptr, 0);
    return status(cgp_ekf_nll_fwd_f64(&p, ys.typed_data(), nll->typed_data(), ws->typed_data(), ws->element_count() * sizeof(double),
                                      ckpt_every, stream), "chirpgp_b200 ekf_nll fwd");
}
ffi::Error NllBwdImpl(cudaStream_t stream, int64_t num_harmonics, int64_t ys_repeat, int64_t h_unit_index, int64_t ckpt_every,
                      double Xi, double dt, F64 ys, F64 consts, F64 m0, F64 P0, F64 H, F64 nll_bar, F64 ws, RF64 consts_bar,
                      RF64 m0_bar, RF64 P0_bar, RF64 Xi_bar, RF64 ws_out) {
    // the workspace is read AND written by the adjoint kernel: the caller aliases the operand `ws` to the result `ws_out`
    // (input_output_aliases={6: 4} in chirpgp_b200/jax_ffi.py), so this is the same buffer and XLA knows it is mutated
    if (ws_out->typed_data() != ws.typed_data()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "ws must be aliased to ws_out");
    const int64_t d = 2 * num_harmonics + 2, B = nll_bar.element_count(), T = ys.dimensions().back();
    const Attrs a{CGP_MODEL_LCD, num_harmonics, 0, 0, ys_repeat, h_unit_index, Xi, dt};
    CgpProblem p = make_problem(a, B, T, d, consts, m0.typed_data(), m0.element_count() / d, P0.typed_data(),
                                P0.element_count() / (d * d), H.typed_data(), nullptr, 1, nullptr, nullptr, 0);
    return status(cgp_ekf_nll_bwd_f64(&p, ys.typed_data(), nll_bar.typed_data(), ws_out->typed_data(),
                                      ws_out->element_count() * sizeof(double), ckpt_every, consts_bar->typed_data(), m0_bar->typed_data(),
                                      P0_bar->typed_data(), Xi_bar->typed_data(), stream), "chirpgp_b200 ekf_nll bwd");
}

}  // namespace

#define CGP_COMMON_FILTER_ATTRS                                                                                          \
    .Attr<std::string_view>("fn").Attr<int64_t>("model").Attr<int64_t>("num_harmonics").Attr<int64_t>("sigma_kind")      \
    .Attr<int64_t>("gh_order")

XLA_FFI_DEFINE_HANDLER_SYMBOL(CgpFilter, FilterImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() CGP_COMMON_FILTER_ATTRS
                                  .Attr<int64_t>("ys_repeat").Attr<int64_t>("h_unit_index").Attr<double>("Xi").Attr<double>("dt")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgpFilterGains, FilterGainsImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model").Attr<int64_t>("num_harmonics").Attr<int64_t>("sigma_kind")
                                  .Attr<int64_t>("gh_order").Attr<int64_t>("ys_repeat").Attr<int64_t>("h_unit_index")
                                  .Attr<double>("Xi").Attr<double>("dt")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgpSmootherSweep, SweepImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgpSmoother, SmootherImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>() CGP_COMMON_FILTER_ATTRS
                                  .Attr<double>("dt")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgpEkfNllFwd, NllFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("num_harmonics").Attr<int64_t>("ys_repeat").Attr<int64_t>("h_unit_index")
                                  .Attr<int64_t>("ckpt_every").Attr<double>("Xi").Attr<double>("dt")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgpEkfNllBwd, NllBwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("num_harmonics").Attr<int64_t>("ys_repeat").Attr<int64_t>("h_unit_index")
                                  .Attr<int64_t>("ckpt_every").Attr<double>("Xi").Attr<double>("dt")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
