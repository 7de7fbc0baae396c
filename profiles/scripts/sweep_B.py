"""Times cg.sgp_filter / cg.sgp_smoother (GH order 3, chirp model) for a range of batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels

T, dt = 3141, 1e-3
Bs = [int(a) for a in sys.argv[1:]] or [148, 296, 592, 1000, 2000, 4000, 8000]
_, ys_all, _ = toymodels.synthetic_batch(64, T, dt, Xi=0.1, seed=2)
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
sg = cg.SigmaPoints.gauss_hermite(4, 3)
for B in Bs:
    ys = torch.as_tensor(np.tile(ys_all, (B // 64 + 1, 1))[:B]).cuda()
    ts = []
    for it in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        f = cg.sgp_filter(mc, sg, H, 0.1, m0, P0, dt, ys)
        e[1].record()
        s = cg.sgp_smoother(mc, sg, f[0], f[1], dt)
        e[2].record()
        torch.cuda.synchronize()
        ts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
        del f, s
    tf, tsm = min(t[0] for t in ts[1:]), min(t[1] for t in ts[1:])
    print('B=%6d  filter %8.3f ms (%7.1f Msteps/s, %6.0f cycles/step/warp)  smoother %8.3f ms  total %7.1f Msteps/s'
          % (B, tf, B * T / tf / 1e3, tf * 1e-3 * 1.965e9 / T, tsm, B * T / (tf + tsm) / 1e3))
