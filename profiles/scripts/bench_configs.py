"""Timings of the other BASELINE.json configs (bench.py is configs[1]).  One JSON line per config.

  python profiles/scripts/bench_configs.py [1 3 4 5] [--small]

config 1: single chirp EKF+EKS (latency: us/step);  config 3: CD-EKF+CD-EKS and CD-GHF+CD-GHS on 1000 chirps;
config 4: harmonic d=8 cubature CKF+CKS on 1000 chirps;  config 5: EKF nll + adjoint, chirps x 16 candidates
(sharded over ranks under torchrun, one all-reduce of the 16 x (1+6) objective/gradient block)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels, mle
from chirpgp_b200.distributed import init_from_env, shard, allreduce_objective
from chirpgp_b200.models import g as gfun

PARAMS = np.array([0.1, 0.1, 0.1, 1., 1., 7.])


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    small = '--small' in sys.argv
    which = [int(a) for a in args] or [1, 3, 4, 5]
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dt, T = 1e-3, 3141
    out = []
    if 1 in which and rank == 0:
        _, ys, _ = toymodels.synthetic_batch(1, T, dt, seed=1)
        ys = torch.as_tensor(ys[0]).cuda()
        _, _, mc, m0, P0, H = cg.build_chirp_model(PARAMS)
        f = cg.ekf(mc, H, 0.1, m0, P0, dt, ys)
        tf = timed(lambda: cg.ekf(mc, H, 0.1, m0, P0, dt, ys))
        ts = timed(lambda: cg.eks(mc, f[0], f[1], dt))
        out.append(dict(config=1, what='single chirp EKF+EKS, T=3141, d=4', ekf_ms=tf, eks_ms=ts,
                        us_per_step=(tf + ts) * 1e3 / T, steps_per_s=T / ((tf + ts) * 1e-3)))
    if 3 in which and rank == 0:
        B = 1000
        _, ys, _ = toymodels.synthetic_batch(B, T, dt, seed=2)
        ys = torch.as_tensor(ys).cuda()
        drift, disp, mc, m0, P0, H = cg.build_chirp_model(PARAMS)
        sg = cg.SigmaPoints.gauss_hermite(4, 3)
        f = cg.cd_ekf(drift, disp, H, 0.1, m0, P0, dt, ys)
        tf = timed(lambda: cg.cd_ekf(drift, disp, H, 0.1, m0, P0, dt, ys))
        ts = timed(lambda: cg.cd_eks(drift, disp, f[0], f[1], dt))
        out.append(dict(config=3, what='CD-EKF+CD-EKS, 1000 chirps x 3141', filter_ms=tf, smoother_ms=ts,
                        steps_per_s=B * T / ((tf + ts) * 1e-3)))
        bm = disp(None)
        f = cg.cd_sgp_filter(drift, bm, sg, H, 0.1, m0, P0, dt, ys)
        tf = timed(lambda: cg.cd_sgp_filter(drift, bm, sg, H, 0.1, m0, P0, dt, ys))
        ts = timed(lambda: cg.cd_sgp_smoother(drift, bm, sg, f[0], f[1], dt))
        out.append(dict(config=3, what='CD-GHF+CD-GHS (81 points), 1000 chirps x 3141', filter_ms=tf, smoother_ms=ts,
                        steps_per_s=B * T / ((tf + ts) * 1e-3)))
    if 4 in which and rank == 0:
        B = 1000
        _, ys, _ = toymodels.synthetic_batch(B, T, dt, num_harmonics=3, seed=4)
        ys = torch.as_tensor(ys).cuda()
        _, _, mc, m0, P0, H = cg.build_harmonic_chirp_model(PARAMS, num_harmonics=3)
        sg = cg.SigmaPoints.cubature(8)
        f = cg.sgp_filter(mc, sg, H, 0.1, m0, P0, dt, ys)
        tf = timed(lambda: cg.sgp_filter(mc, sg, H, 0.1, m0, P0, dt, ys))
        ts = timed(lambda: cg.sgp_smoother(mc, sg, f[0], f[1], dt))
        out.append(dict(config=4, what='harmonic d=8 cubature CKF+CKS, 1000 chirps x 3141', filter_ms=tf, smoother_ms=ts,
                        steps_per_s=B * T / ((tf + ts) * 1e-3)))
    if 5 in which:
        Bc, T5, G = (2000, 10000, 16) if small else (10000, 100000, 16)
        dt5 = 3.141 / T5
        lo_hi = (Bc * rank // world, Bc * (rank + 1) // world)
        nloc = lo_hi[1] - lo_hi[0]
        # synthetic chirps generated on the device (the host generator is a Python loop): same model as toymodels
        gen = torch.Generator(device='cuda').manual_seed(1234 + rank)
        ts_ = torch.linspace(dt5, dt5 * T5, T5, dtype=torch.float64, device='cuda')
        phase = 500 * torch.exp(-5 / torch.sin(ts_)) + 8 * ts_
        ys = torch.sin(2 * np.pi * phase)[None, :] + np.sqrt(0.1) * torch.randn((nloc, T5), dtype=torch.float64, device='cuda', generator=gen)
        lam = np.array([0.1, 0.4, 0.7, 1.0]); bb = np.array([0.05, 0.1, 0.2, 0.4])
        grid = np.array([[l, b_, 0.1, 1., 1., 7.] for l in lam for b_ in bb])
        theta = torch.tensor(np.log(np.exp(grid) - 1.), dtype=torch.float64, device='cuda', requires_grad=True)
        H = np.array([0., 1., 0., 0.])

        def objective():
            _, _, mc, m0, P0, _ = cg.build_chirp_model(gfun(theta))
            nll = mle.ekf_nll(mc, H, 0.1, m0, P0, dt5, ys, candidates=True)     # (nloc, G)
            val = nll.sum(dim=0)                                                # per candidate
            grad, = torch.autograd.grad(val.sum(), theta)                       # (G, 6): rows independent
            return allreduce_objective(val.detach(), grad)

        if world > 1:
            torch.distributed.barrier()
        ms = timed(objective, reps=2)
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        if rank == 0:
            out.append(dict(config=5, what='EKF nll + adjoint gradient, %d chirps x %d candidates x T=%d, %d GPU(s)' % (Bc, G, T5, world),
                            ms=float(t), steps_per_s=Bc * G * T5 / (float(t) * 1e-3), n_gpus=world))
    if rank == 0:
        for o in out:
            print(json.dumps(o), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
