// cd_eks / cd_sgp_smoother launchers (sequential RK4 backwards in time).
#include "cgp_dispatch.cuh"

namespace cgp {

int launch_cd_eks(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    if (p.model == CGP_MODEL_SDE && p.num_harmonics == 1 && p.d == 4 && p.B <= 40000) {
        cd_eks_lane_kernel<1><<<(unsigned)ceil_div(p.B, 2), 32, 0, s>>>(p, io);
        return check_launch();
    }
    return dispatch_sde(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        const int block = 64;
        cd_eks_thread_kernel<Model><<<(unsigned)ceil_div(p.B, block), block, 0, s>>>(p, io);
        return check_launch();
    });
}

template <class Model, int G, int P>
static int launch_one(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    const int block = GroupCfg<Model, G, true>::kBlock;
    cd_sgp_smoother_kernel<Model, G, P><<<(unsigned)ceil_div(p.B * G, block), block, 0, s>>>(p, io);
    return check_launch();
}

int launch_cd_sgp_smoother(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    const bool share = use_share(p);
    const int g = group_size(p, share);
    return dispatch_sde(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        if constexpr (Model::kLinear) {
            return launch_one<Model, 32, 0>(p, io, s);
        } else {
            if constexpr (Model::NH == 1) {
                if (share && aligned16(io.mss) && aligned16(io.Pss)) {
                    cd_ghs_warp_kernel<1, 3><<<(unsigned)p.B, 32, 0, s>>>(p, io);
                    return check_launch();
                }
            }
            if (share) return launch_one<Model, 32, 3>(p, io, s);
            if (g == 8) return launch_one<Model, 8, 0>(p, io, s);
            if (g == 16) return launch_one<Model, 16, 0>(p, io, s);
            return launch_one<Model, 32, 0>(p, io, s);
        }
    });
}

}  // namespace cgp
