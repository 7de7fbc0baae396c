#!/usr/bin/env python
"""Times the headline pair (sgp_filter, sgp_smoother; chirp model, Gauss-Hermite order 3) on device-resident inputs:
    python profiles/scripts/time_ghf.py [B ...]
Prints filter / smoother / total ms (CUDA events, L2 flushed between passes, median of 7) with and without the fused
smoother gains."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg  # noqa: E402
from chirpgp_b200 import toymodels  # noqa: E402

T, DT, XI = 3141, 1e-3, 0.1


def main():
    Bs = [int(a) for a in sys.argv[1:]] or [1000]
    dev = torch.device('cuda', 0)
    _, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
    sg = cg.SigmaPoints.gauss_hermite(4, 3)
    m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    for B in Bs:
        _, ys, _ = toymodels.synthetic_batch(min(B, 1000), T, DT, Xi=XI, seed=2)
        ys = torch.as_tensor(np.tile(ys, (-(-B // ys.shape[0]), 1))[:B]).to(dev)
        for gains in (False, True):
            tf, ts = [], []
            for it in range(10):
                flush.fill_(1.)
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                f = cg.sgp_filter(mc, sg, H, XI, m0, P0, DT, ys, smoother_gains=gains)
                e[1].record()
                s = cg.sgp_smoother(mc, sg, f[0], f[1], DT)
                e[2].record()
                torch.cuda.synchronize()
                if it >= 3:
                    tf.append(e[0].elapsed_time(e[1])); ts.append(e[1].elapsed_time(e[2]))
                del f, s
            a, b = statistics.median(tf), statistics.median(ts)
            print('B=%6d gains=%d  filter %7.3f ms  smoother %7.3f ms  total %7.3f ms  %7.1f Msteps/s  [CGP_DUO=%s SEL=%s]'
                  % (B, gains, a, b, a + b, B * T / (a + b) / 1e3, os.environ.get('CGP_DUO', '-'),
                     os.environ.get('CGP_DUO_SEL', '-')), flush=True)


if __name__ == '__main__':
    main()
