// kf / ekf / sgp_filter launchers (discrete-time models).
#include <stdlib.h>
#include "cgp_dispatch.cuh"
#include "cgp_multi.cuh"

namespace cgp {

// Below this many chirps one thread per chirp leaves the GPU to a handful of warps that each walk ~600 instructions per step;
// 16 lanes per chirp (cgp_fast.cuh: ekf_lane_kernel) halve the instructions on the chain.  CGP_EKF_LANE=0 / 1 forces the choice
// (tests, measurements).
static const int64_t kEkfLaneMaxB = 2048;     // measured (profiles/r2_ekf_lane.txt): 2.45 -> 1.87 ms for one chirp, 3.35 -> 2.34 ms at 1000, 3.38 -> 4.12 ms at 4000
static bool use_ekf_lane(const CgpProblem &p) {
    if (!(p.model == CGP_MODEL_LCD && p.num_harmonics == 1 && p.d == 4)) return false;
    const char *v = getenv("CGP_EKF_LANE");
    if (v && *v) return atoi(v) != 0;
    return p.B * (int64_t)(p.in_flight > 1 ? p.in_flight : 1) <= kEkfLaneMaxB;     // chirps in flight over all overlapping launches
}

int launch_ekf(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    if (use_ekf_lane(p)) {
        if (p.h_unit_index == 1) ekf_lane_kernel<1, true><<<(unsigned)ceil_div(p.B, 2), 32, 0, s>>>(p, io);
        else ekf_lane_kernel<1, false><<<(unsigned)ceil_div(p.B, 2), 32, 0, s>>>(p, io);
        return check_launch();
    }
    return dispatch_disc(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        const int block = 128;
        if constexpr (Model::D % 4 == 0) {
            if (io.mfs != nullptr && aligned32(io.mfs) && aligned32(io.Pfs)) {
                ekf_thread_kernel<Model, true><<<(unsigned)ceil_div(p.B, block), block, 0, s>>>(p, io);
                return check_launch();
            }
        }
        ekf_thread_kernel<Model><<<(unsigned)ceil_div(p.B, block), block, 0, s>>>(p, io);
        return check_launch();
    });
}

int launch_ekf_kpt(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    const int block = 128;
    const unsigned grid = (unsigned)ceil_div(p.B, block);
    switch (p.num_harmonics) {
        case 1: ekf_kpt_thread_kernel<1><<<grid, block, 0, s>>>(p, io); break;
        case 2: ekf_kpt_thread_kernel<2><<<grid, block, 0, s>>>(p, io); break;
        case 3: ekf_kpt_thread_kernel<3><<<grid, block, 0, s>>>(p, io); break;
        default: return CGP_ERR_UNSUPPORTED;
    }
    return check_launch();
}

static const int64_t kThreadPerChirpMinB = 30000;   // measured: 1.76 G vs 1.48 G filter steps/s at 64 000 chirps

template <class Model, int G, int P>
static int launch_sgp_one(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    const int block = GroupCfg<Model, G, false>::kBlock;
    sgp_filter_kernel<Model, G, P, false><<<(unsigned)ceil_div(p.B * G, block), block, 0, s>>>(p, io);
    return check_launch();
}

// The filter kernel that also fills the smoother workspace exists for the headline configuration (chirp LCD model,
// Gauss-Hermite order 3, warp per chirp); everything else runs the filter and then the time-parallel gain kernel.
// Harmonic chirp models (d = 6, 8) with the cubature rule: cub_duo_filter_kernel (cgp_cubduo.cuh), two chirps per CTA.
// 8-lanes-per-chirp Gauss-Hermite kernel (cgp_oct.cuh): batches from kOctMinB chirps; CGP_GH_OCT=0 / 1 forces it off / on
// (tests, measurements).
static const int64_t kOctMinB = 2000;     // measured (profiles/r2_oct_kernel.txt): pair +10 % at 2000 chirps, +34 % at 4000, +61 % at 32 000
// A caller that keeps several batches in flight (filter_smoother_batches) says so in CgpProblem::in_flight: what counts then is
// the throughput of the overlapping launches, and the 8-lane kernel issues 204 instead of 335 FP64 instructions per chirp and step
// (profiles/r2_batches.txt: 1000 chirps per batch, 3 in flight: 3.00 vs 3.09 ms per batch; 4: 3.00 vs 2.57; 8: 3.00 vs 1.91).
static const int64_t kOctMinInFlightChirps = 4000;
static bool use_oct(const CgpProblem &p) {
    const char *v = getenv("CGP_GH_OCT");
    if (v && *v) return atoi(v) != 0;
    return p.B >= kOctMinB || p.B * (int64_t)(p.in_flight > 1 ? p.in_flight : 1) >= kOctMinInFlightChirps;
}
static bool use_cub_duo(const CgpProblem &p) {
    return p.model == CGP_MODEL_LCD && (p.num_harmonics == 2 || p.num_harmonics == 3) && p.d == 2 * p.num_harmonics + 2 &&
           p.sigma_kind == CGP_SIGMA_CUBATURE && p.n_sigma == 2 * p.d && p.B < kThreadPerChirpMinB;
}
bool sgp_filter_fuses_gains(const CgpProblem &p) {
    return (p.model == CGP_MODEL_LCD && p.num_harmonics == 1 && p.d == 4 && use_share(p)) || use_cub_duo(p);
}

template <int NH, bool H_HARM> static int launch_cub_duo_h(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    static bool configured = false;
    const int smem = (int)sizeof(CubDuoSmem<NH>);
    if (!configured) {
        cudaFuncSetAttribute(cub_duo_filter_kernel<NH, H_HARM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(cub_duo_filter_kernel<NH, H_HARM>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    cub_duo_filter_kernel<NH, H_HARM><<<(unsigned)ceil_div(p.B, 2), 64, smem, s>>>(p, io);
    return check_launch();
}
template <int NH> static int launch_cub_duo(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    if (p.h_unit_index == CGP_H_HARMONIC) return launch_cub_duo_h<NH, true>(p, io, s);
    return launch_cub_duo_h<NH, false>(p, io, s);
}

int launch_sgp_filter(const CgpProblem &p, const FilterIO &io_in, cudaStream_t s) {
    FilterIO io = io_in;
    if (io.ws != nullptr && (io.mfs == nullptr || !sgp_filter_fuses_gains(p) || !aligned16(io.ws))) io.ws = nullptr;
    const bool share = use_share(p);
    const int g = group_size(p, share);
    return dispatch_disc(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        if constexpr (Model::kLinear) {
            return launch_sgp_one<Model, 32, 0>(p, io, s);
        } else {
            // Very large batches are FP64-throughput-bound: one thread per chirp (all sigma points serially, no
            // replicated work, no shuffles) issues ~2.4x fewer FP64 warp-instructions per step than a warp per chirp.
            if (p.B >= kThreadPerChirpMinB && io.ws == nullptr && !(Model::NH == 1 && share && use_oct(p))) {
                if (share) return launch_sgp_one<Model, 1, 3>(p, io, s);
                return launch_sgp_one<Model, 1, 0>(p, io, s);
            }
            if constexpr (Model::NH == 2 || Model::NH == 3) {
                // the 16-byte stores of the consumer warp need aligned outputs; everything else takes the generic kernel
                if (use_cub_duo(p) && (io.mfs == nullptr || (aligned16(io.mfs) && aligned16(io.Pfs))))
                    return launch_cub_duo<Model::NH>(p, io, s);
            }
            if constexpr (Model::NH == 1) {
                // headline path: chirp model, Gauss-Hermite order 3 -> 27 base indices, one per lane
                if (share) {
                    using Pred = GhPredictLCD<1, 3>;
                    // large batches are bound by the FP64 pipe, not by one warp's dependency chain: 8 lanes per chirp (cgp_oct.cuh)
                    if (use_oct(p)) {
                        const unsigned grid = (unsigned)ceil_div(p.B, OctCfg::CH);
                        if (io.ws != nullptr) {
                            if (p.h_unit_index == 1) gh_oct_filter_kernel<true, true><<<grid, 32, 0, s>>>(p, io);
                            else gh_oct_filter_kernel<false, true><<<grid, 32, 0, s>>>(p, io);
                        } else {
                            if (p.h_unit_index == 1) gh_oct_filter_kernel<true, false><<<grid, 32, 0, s>>>(p, io);
                            else gh_oct_filter_kernel<false, false><<<grid, 32, 0, s>>>(p, io);
                        }
                        return check_launch();
                    }
                    if (io.ws != nullptr) {
                        // filter + smoother gains: producer / consumer warp pair per chirp (cgp_duo.cuh)
                        if (p.h_unit_index == 1) gh_duo_filter_kernel<true><<<(unsigned)p.B, 64, 0, s>>>(p, io);
                        else gh_duo_filter_kernel<false><<<(unsigned)p.B, 64, 0, s>>>(p, io);
                        return check_launch();
                    }
                    // experiment (profiles/r2_multi_chirp.txt): NCH chirps interleaved per warp, nll-only
                    const char *mv = getenv("CGP_GH_MULTI");
                    const int nch = (mv && *mv) ? atoi(mv) : 0;
                    if (nch == 16 && io.mfs == nullptr && io.nell && io.nell_last_only && p.h_unit_index == 1) {
                        gh_half_nll_kernel<true><<<(unsigned)ceil_div(p.B, 2), 32, 0, s>>>(p, io.ys, io.nell);
                        return check_launch();
                    }
                    if (nch >= 1 && nch <= 3 && io.mfs == nullptr && io.nell && io.nell_last_only && p.h_unit_index == 1) {
                        const unsigned grid = (unsigned)ceil_div(p.B, nch);
                        if (nch == 1) gh_warp_multi_nll_kernel<1, true><<<grid, 32, 0, s>>>(p, io.ys, io.nell);
                        else if (nch == 2) gh_warp_multi_nll_kernel<2, true><<<grid, 32, 0, s>>>(p, io.ys, io.nell);
                        else gh_warp_multi_nll_kernel<3, true><<<grid, 32, 0, s>>>(p, io.ys, io.nell);
                        return check_launch();
                    }
                    if (p.h_unit_index == 1) gh_warp_filter_kernel<Pred, false, true><<<(unsigned)p.B, 32, 0, s>>>(p, io);
                    else gh_warp_filter_kernel<Pred, false, false><<<(unsigned)p.B, 32, 0, s>>>(p, io);
                    return check_launch();
                }
            }
            if (share) return launch_sgp_one<Model, 32, 3>(p, io, s);
            if (g == 8) return launch_sgp_one<Model, 8, 0>(p, io, s);
            if (g == 16) return launch_sgp_one<Model, 16, 0>(p, io, s);
            return launch_sgp_one<Model, 32, 0>(p, io, s);
        }
    });
}

}  // namespace cgp
