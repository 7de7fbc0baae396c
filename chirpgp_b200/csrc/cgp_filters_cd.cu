// cd_ekf / cd_sgp_filter launchers (continuous-discrete models, RK4).
#include "cgp_dispatch.cuh"

namespace cgp {

// Below this many chirps one thread per chirp cannot fill the GPU (148 SMs x 4 sub-partitions x a few warps); the
// 16-lanes-per-chirp kernels give each chirp 16x the issue slots and a 16x shorter matrix-product chain.
static const int64_t kLaneKernelMaxB = 40000;

int launch_cd_ekf(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    if (p.model == CGP_MODEL_SDE && p.num_harmonics == 1 && p.d == 4 && p.B <= kLaneKernelMaxB) {
        if (p.h_unit_index == 1) cd_ekf_lane_kernel<1, true><<<(unsigned)ceil_div(p.B, 2), 32, 0, s>>>(p, io);
        else cd_ekf_lane_kernel<1, false><<<(unsigned)ceil_div(p.B, 2), 32, 0, s>>>(p, io);
        return check_launch();
    }
    return dispatch_sde(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        const int block = 128;
        cd_ekf_thread_kernel<Model><<<(unsigned)ceil_div(p.B, block), block, 0, s>>>(p, io);
        return check_launch();
    });
}

template <class Model, int G, int P>
static int launch_cd_sgp_one(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    const int block = GroupCfg<Model, G, true>::kBlock;
    sgp_filter_kernel<Model, G, P, true><<<(unsigned)ceil_div(p.B * G, block), block, 0, s>>>(p, io);
    return check_launch();
}

int launch_cd_sgp_filter(const CgpProblem &p, const FilterIO &io, cudaStream_t s) {
    const bool share = use_share(p);
    const int g = group_size(p, share);
    return dispatch_sde(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        if constexpr (Model::kLinear) {
            return launch_cd_sgp_one<Model, 32, 0>(p, io, s);
        } else {
            if constexpr (Model::NH == 1) {
                // chirp SDE, Gauss-Hermite order 3: tuned warp-per-chirp kernel (27 base indices, one per lane)
                if (share) {
                    using Rhs = GhRhsSDE<1, 3>;
                    if (p.h_unit_index == 1) gh_warp_filter_kernel<Rhs, true, true><<<(unsigned)p.B, 32, 0, s>>>(p, io);
                    else gh_warp_filter_kernel<Rhs, true, false><<<(unsigned)p.B, 32, 0, s>>>(p, io);
                    return check_launch();
                }
            }
            if (share) return launch_cd_sgp_one<Model, 32, 3>(p, io, s);
            if (g == 8) return launch_cd_sgp_one<Model, 8, 0>(p, io, s);
            if (g == 16) return launch_cd_sgp_one<Model, 16, 0>(p, io, s);
            return launch_cd_sgp_one<Model, 32, 0>(p, io, s);
        }
    });
}

}  // namespace cgp
