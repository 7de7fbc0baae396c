#!/usr/bin/env python
"""EKF nll + adjoint throughput vs the number of problems on ONE GPU (config 5 shards 160 000 problems over N GPUs:
160 000 / N problems per GPU -- the strong-scaling curve can be read off a single GPU).

    python profiles/scripts/time_nll.py [--T 10000] [--G 16] [P ...]        P = problems (chirps x candidates) per launch

Prints one JSON line per P: fwd ms, bwd ms, nll+gradient steps/s, and the strong-scaling efficiency relative to the
largest P (time(Pmax) / (Pmax / P) / time(P))."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg  # noqa: E402
from chirpgp_b200 import mle  # noqa: E402
from chirpgp_b200.models import g as gfun  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--T', type=int, default=10000)
    ap.add_argument('--G', type=int, default=16)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('P', type=int, nargs='*')
    a = ap.parse_args()
    Ps = a.P or [160000, 80000, 40000, 20000]
    dev = torch.device('cuda', 0)
    T, G = a.T, a.G
    dt = 3.141 / T
    lam = np.array([0.1, 0.4, 0.7, 1.0]); bb = np.array([0.05, 0.1, 0.2, 0.4])
    grid = np.array([[l, b_, 0.1, 1., 1., 7.] for l in lam for b_ in bb])[:G]
    H = np.array([0., 1., 0., 0.])
    rows = []
    for P in Ps:
        nch = P // G
        gen = torch.Generator(device='cuda').manual_seed(1234)
        ts_ = torch.linspace(dt, dt * T, T, dtype=torch.float64, device=dev)
        phase = 500 * torch.exp(-5 / torch.sin(ts_)) + 8 * ts_
        ys = torch.sin(2 * np.pi * phase)[None, :] + np.sqrt(0.1) * torch.randn((nch, T), dtype=torch.float64, device=dev, generator=gen)
        theta = torch.tensor(np.log(np.exp(grid) - 1.), dtype=torch.float64, device=dev, requires_grad=True)
        best = (1e30, 0., 0.)
        for it in range(a.reps + 1):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            _, _, mc, m0, P0, _ = cg.build_chirp_model(gfun(theta))
            nll = mle.ekf_nll(mc, H, 0.1, m0, P0, dt, ys, candidates=True)
            val = nll.sum(dim=0)
            e[1].record()
            grad, = torch.autograd.grad(val.sum(), theta)
            e[2].record()
            torch.cuda.synchronize()
            tot = e[0].elapsed_time(e[2])
            if it > 0 and tot < best[0]:
                best = (tot, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]))
        rows.append(dict(problems=P, T=T, ms=best[0], fwd_ms=best[1], bwd_ms=best[2], steps_per_s=P * T / (best[0] * 1e-3),
                         val0=float(val[0]), grad0=[float(x) for x in grad[0]]))
        del ys
    ref = max(rows, key=lambda r: r['problems'])
    for r in rows:
        r['strong_eff_vs_%d' % ref['problems']] = ref['ms'] / (ref['problems'] / r['problems']) / r['ms']
        print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
