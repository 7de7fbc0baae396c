// rts / eks / sgp_smoother launchers: time-parallel gain kernel + sequential sweep.
#include "cgp_dispatch.cuh"

namespace cgp {

// Sweep: warp-per-chirp kernel (shared-memory tiles, one matrix entry per lane) whenever the 16-byte async
// copies are legal (even d, 16-byte aligned buffers); otherwise the thread-per-chirp fallback.
template <int D> static int launch_sweep(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    if constexpr (D % 2 == 0) {
        if (aligned16(io.ws)) {
            if constexpr (D == 4) {
                // tile of 16 steps, double buffered (measured at B = 1000 / 10 000 / 64 000: 0.66 / 4.01 / 24.5 ms; 8 steps x 4 stages:
                // 0.71 / 4.11 / 25.4; 16 x 4 needs 57 KB per warp and loses occupancy at large B)
                constexpr int TS = 16, NSTAGE = 2;
                const size_t smem = sizeof(double) * 2 * NSTAGE * TS * ws_record<4>();
                smoother_sweep_lane4_kernel<TS, NSTAGE><<<(unsigned)ceil_div(p.B, 2), 32, smem, s>>>(p, io);
                return check_launch();
            }
            if constexpr (D == 8) {
                constexpr int TS = 4, NSTAGE = 4;
                const size_t smem = sizeof(double) * NSTAGE * TS * ws_record<8>();
                smoother_sweep_lane8_kernel<TS, NSTAGE><<<(unsigned)p.B, 32, smem, s>>>(p, io);
                return check_launch();
            }
            using Cfg = SweepCfg<D>;
            if (Cfg::smem_bytes() > 48 * 1024)             // d >= 10: opt in to the large dynamic shared-memory carve-out
                cudaFuncSetAttribute(smoother_sweep_warp_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes());
            smoother_sweep_warp_kernel<D><<<(unsigned)p.B, 32, Cfg::smem_bytes(), s>>>(p, io);
            return check_launch();
        }
    }
    const int block = 64;
    smoother_sweep_kernel<D><<<(unsigned)ceil_div(p.B, block), block, 0, s>>>(p, io);
    return check_launch();
}

// From this many chirps on one thread per chirp fills the GPU (148 SMs x 16 resident warps x 32) and the one-pass kernel's
// 3.3x lower DRAM traffic wins over the time-parallel gain kernel + sweep (profiles/r1_ekf_eks.txt).
static const int64_t kOnePassMinB = 32768;

int launch_eks(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    return dispatch_disc(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        const int block = 128;
        if constexpr (Model::D == 2 || Model::D == 4) {          // everything of a step fits the register file up to d = 4
            if (p.B >= kOnePassMinB && aligned16(io.mfs) && aligned16(io.Pfs) && aligned16(io.mss) && aligned16(io.Pss)) {
                const unsigned grid = (unsigned)ceil_div(p.B, block);
                if (Model::D == 4 && aligned32(io.mfs) && aligned32(io.Pfs) && aligned32(io.mss) && aligned32(io.Pss))
                    eks_onepass_thread_kernel<Model, true><<<grid, block, 0, s>>>(p, io);
                else
                    eks_onepass_thread_kernel<Model, false><<<grid, block, 0, s>>>(p, io);
                return check_launch();
            }
        }
        const int64_t items = p.B * (p.T - 1);
        if (items > 0) {
            eks_gain_kernel<Model><<<(unsigned)ceil_div(items, block), block, 0, s>>>(p, io);
            int rc = check_launch();
            if (rc) return rc;
        }
        return launch_sweep<Model::D>(p, io, s);
    });
}

// Time-parallel half of sgp_smoother (filters_smoothers.py:520-527): fills io.ws for steps 0 .. T-2.
int launch_sgp_gains(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    const bool share = use_share(p);
    return dispatch_disc(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        const int block = 128;
        const int64_t items = p.B * (p.T - 1);
        if (items > 0) {
            const unsigned grid = (unsigned)ceil_div(items, block);
            if constexpr (Model::kLinear) {
                sgp_gain_kernel<Model, 0><<<grid, block, 0, s>>>(p, io);
            } else if (p.sigma_kind == CGP_SIGMA_CUBATURE && p.n_sigma == 2 * Model::D && Model::D >= 6) {
                constexpr int D = Model::D;
                const size_t smem = sizeof(double) * 64 * (D * (D + 1) / 2 + D * D);
                if (smem > 48 * 1024)
                    cudaFuncSetAttribute(cubature_gain_kernel<Model::NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cubature_gain_kernel<Model::NH><<<(unsigned)ceil_div(items, 64), 64, smem, s>>>(p, io);
            } else {
                if (share) sgp_gain_kernel<Model, 3><<<grid, block, 0, s>>>(p, io);
                else sgp_gain_kernel<Model, 0><<<grid, block, 0, s>>>(p, io);
            }
            return check_launch();
        }
        return 0;
    });
}

// Sequential half (filters_smoothers.py:83-84) on a filled workspace; needs only B, T, d of the problem.
int launch_smoother_sweep(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    switch (p.d) {
        case 1: return launch_sweep<1>(p, io, s);
        case 2: return launch_sweep<2>(p, io, s);
        case 3: return launch_sweep<3>(p, io, s);
        case 4: return launch_sweep<4>(p, io, s);
        case 5: return launch_sweep<5>(p, io, s);
        case 6: return launch_sweep<6>(p, io, s);
        case 8: return launch_sweep<8>(p, io, s);
        case 10: return launch_sweep<10>(p, io, s);
        case 12: return launch_sweep<12>(p, io, s);
        default: return CGP_ERR_UNSUPPORTED;
    }
}

int launch_sgp_smoother(const CgpProblem &p, const SmootherIO &io, cudaStream_t s) {
    int rc = launch_sgp_gains(p, io, s);
    if (rc) return rc;
    return launch_smoother_sweep(p, io, s);
}

}  // namespace cgp
