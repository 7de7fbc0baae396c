// cgp_sim.cu -- Monte-Carlo input side of the batched filters: trajectories and measurements of the
// discretised models, one thread per trajectory, random numbers generated in the kernel.
//
// What the reference does (tetralith/jobs/crlb_ekf.py:41-56, test/test_crlb.py:41-55, tools.py:81-170 simulate_lgssm /
// simulate_sde):   x_0 = m0 + chol(P0) eps,   x_k = mean(x_{k-1}) + chol(Sigma) eps_k,   y_k = H x_k + sqrt(Xi) eps'_k
// with jax.random.normal from threefry keys.  Those streams cannot be reproduced without JAX; here the normals come from
// the counter-based Philox4x32-10 generator (Salmon et al., SC'11; key = seed, counter = (draw, step, trajectory)) and
// Box-Muller in float64, so that a trajectory depends on (seed, index) only -- any sharding of the batch over GPUs gives
// the same samples -- and the NumPy restatement in oracle/sim_oracle.py reproduces them to rounding.
#include "cgp_dispatch.cuh"

namespace cgp {

struct Philox {
    uint32_t k0, k1;
    CGP_DEV void round(uint32_t (&c)[4], uint32_t ka, uint32_t kb) const {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ ka, n1 = lo1, n2 = hi0 ^ c[3] ^ kb, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    CGP_DEV void block(uint32_t (&c)[4]) const {                      // philox4x32-10
        uint32_t ka = k0, kb = k1;
        CGP_UNROLL for (int r = 0; r < 10; r++) {
            round(c, ka, kb);
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
    }
    // two standard normals from one block: 53-bit uniforms u1 in (0, 1], u2 in [0, 1), Box-Muller
    CGP_DEV void normal2(uint32_t draw, uint32_t step, uint64_t traj, double &z0, double &z1) const {
        uint32_t c[4] = {draw, step, (uint32_t)traj, (uint32_t)(traj >> 32)};
        block(c);
        const uint64_t a = (((uint64_t)c[0] << 32) | c[1]) >> 11, b = (((uint64_t)c[2] << 32) | c[3]) >> 11;
        const double u1 = ((double)a + 1.) * 0x1.0p-53, u2 = (double)b * 0x1.0p-53;
        const double r = sqrt(-2. * log(u1));
        double sn, cs;
        sincospi(2. * u2, &sn, &cs);
        z0 = r * cs; z1 = r * sn;
    }
};

// eps[0..N): draws 0, 1, ... of (step, trajectory)
template <int N> CGP_DEV void normals(const Philox &g, uint32_t first_draw, uint32_t step, uint64_t traj, double (&eps)[N]) {
    CGP_UNROLL for (int i = 0; i < N; i += 2) {
        double z0, z1;
        g.normal2(first_draw + i / 2, step, traj, z0, z1);
        eps[i] = z0;
        if (i + 1 < N) eps[i + 1] = z1;
    }
}

struct SimIO {
    double *__restrict__ xs;      // [B, T, d] or NULL
    double *__restrict__ ys;      // [B, T] or NULL
    double *__restrict__ x0;      // [B, d] or NULL
    uint64_t seed, first_traj;
};

// Model = ModelLinearDisc<D> | ModelLCD<NH>.  chol(P0) and chol(Sigma) are formed once per trajectory (Sigma is state independent
// for every model of the reference; a singular Sigma -- La Scala's zero chirp noise -- is handled by the semi-definite
// convention chol column = 0 when the pivot is 0).
template <class Model>
__global__ void __launch_bounds__(128) simulate_kernel(const CgpProblem p, const SimIO io) {
    constexpr int D = Model::D;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int64_t T = p.T;
    Model mdl;
    mdl.load(p.consts + b * p.consts_stride, p.dt);
    const Philox g{(uint32_t)io.seed, (uint32_t)(io.seed >> 32)};
    const uint64_t traj = io.first_traj + (uint64_t)b;
    auto chol_psd = [](const double (&A)[D][D], double (&L)[D][D]) {
        CGP_UNROLL for (int j = 0; j < D; j++) {
            double s = A[j][j];
            CGP_UNROLL for (int k = 0; k < j; k++) s = fma(-L[j][k], L[j][k], s);
            const double ljj = s > 0. ? sqrt(s) : 0.;
            L[j][j] = ljj;
            CGP_UNROLL for (int i = j + 1; i < D; i++) {
                double t = A[i][j];
                CGP_UNROLL for (int k = 0; k < j; k++) t = fma(-L[i][k], L[j][k], t);
                L[i][j] = ljj > 0. ? t / ljj : 0.;
            }
        }
    };
    double x[D], H[D], Ls[D][D];
    {
        double P0[D][D], L0[D][D], m0[D], eps[D];
        load_vec<D>(p.m0 + b * p.m0_stride, m0);
        load_mat<D>(p.P0 + b * p.P0_stride, P0);
        chol_psd(P0, L0);
        normals<D>(g, 0, 0, traj, eps);
        CGP_UNROLL for (int r = 0; r < D; r++) {
            double s = m0[r];
            CGP_UNROLL for (int c = 0; c <= r; c++) s = fma(L0[r][c], eps[c], s);
            x[r] = s;
        }
        if (io.x0) { CGP_UNROLL for (int r = 0; r < D; r++) io.x0[b * D + r] = x[r]; }
        double Sg[D][D];
        CGP_UNROLL for (int r = 0; r < D; r++) CGP_UNROLL for (int c = 0; c < D; c++) Sg[r][c] = Model::has_sig(r, c) ? mdl.sig(r, c) : 0.;
        chol_psd(Sg, Ls);
    }
    CGP_UNROLL for (int i = 0; i < D; i++) H[i] = p.H[i];
    const double sxi = sqrt(p.Xi);
    for (int64_t t = 0; t < T; t++) {
        double mean[D], eps[D + 1];
        mdl.mean(x, mean);
        normals<D + 1>(g, 0, (uint32_t)(t + 1), traj, eps);
        CGP_UNROLL for (int r = 0; r < D; r++) {
            double s = mean[r];
            CGP_UNROLL for (int c = 0; c <= r; c++) s = fma(Ls[r][c], eps[c], s);
            x[r] = s;
        }
        double y = H[0] * x[0];
        CGP_UNROLL for (int i = 1; i < D; i++) y = fma(H[i], x[i], y);
        y = fma(sxi, eps[D], y);
        if (io.xs) gstore_vec_auto<D>(io.xs + (b * T + t) * D, x);
        if (io.ys) io.ys[b * T + t] = y;
    }
}

int launch_simulate(const CgpProblem &p, const SimIO &io, cudaStream_t s) {
    return dispatch_disc(p, [&](auto tag) {
        using Model = typename decltype(tag)::type;
        const int block = 128;
        simulate_kernel<Model><<<(unsigned)ceil_div(p.B, block), block, 0, s>>>(p, io);
        return check_launch();
    });
}

__global__ void philox_probe_kernel(uint64_t seed, int64_t n, double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox g{(uint32_t)seed, (uint32_t)(seed >> 32)};
    double z0, z1;
    g.normal2((uint32_t)(i & 3), (uint32_t)(i >> 2), (uint64_t)i * 7919u, z0, z1);
    out[2 * i] = z0; out[2 * i + 1] = z1;
}
__global__ void philox_raw_kernel(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) {
    const Philox g{k0, k1};
    uint32_t c[4] = {c0, c1, c2, c3};
    g.block(c);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

}  // namespace cgp

using namespace cgp;

extern "C" {

int cgp_simulate_f64(const CgpProblem *p, uint64_t seed, uint64_t first_trajectory, double *x0, double *xs, double *ys,
                     void *stream) {
    if (!p || p->B < 1 || p->T < 1 || !p->consts || !p->m0 || !p->P0 || !p->H) return CGP_ERR_BAD_ARG;
    if (p->model != CGP_MODEL_LINEAR_DISC && p->model != CGP_MODEL_LCD) return CGP_ERR_BAD_ARG;
    return launch_simulate(*p, SimIO{xs, ys, x0, seed, first_trajectory}, (cudaStream_t)stream);
}

/* test hooks: the raw Philox4x32-10 block (known-answer vectors) and 2 n normals of a fixed counter pattern */
int cgp_test_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out_dev, void *stream) {
    if (!out_dev) return CGP_ERR_BAD_ARG;
    philox_raw_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(k0, k1, c0, c1, c2, c3, out_dev);
    return check_launch();
}
int cgp_test_normals(uint64_t seed, int64_t n, double *out_dev, void *stream) {
    if (n < 1 || !out_dev) return CGP_ERR_BAD_ARG;
    philox_probe_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(seed, n, out_dev);
    return check_launch();
}

}  // extern "C"
