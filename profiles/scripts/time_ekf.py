#!/usr/bin/env python
"""EKF / EKS throughput at the CRLB-job shape (tetralith/jobs/crlb_ekf.py: T = 500, dt = 0.01, many trajectories):
    python profiles/scripts/time_ekf.py [B ...]
Reports filter / smoother ms, steps/s and the HBM roofline fraction (algorithmic bytes: filter 176 B, smoother 320 B per step)."""
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg  # noqa: E402

T, DT, XI = 500, 0.01, 0.1


def main():
    Bs = [int(a) for a in sys.argv[1:]] or [100000]
    dev = torch.device('cuda', 0)
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:  # noqa: BLE001
        peak = 6554.2
    _, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
    m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
    for B in Bs:
        ys = torch.randn((B, T), dtype=torch.float64, device=dev)
        tf, ts = [], []
        for it in range(5):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            f = cg.ekf(mc, H, XI, m0, P0, DT, ys)
            e[1].record()
            s = cg.eks(mc, f[0], f[1], DT)
            e[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                tf.append(e[0].elapsed_time(e[1])); ts.append(e[1].elapsed_time(e[2]))
            del f, s
        a, b = statistics.median(tf), statistics.median(ts)
        n = B * T
        print('B=%8d T=%d  ekf %8.3f ms (%6.2f G steps/s, %5.1f%% of HBM)  eks %8.3f ms (%6.2f G steps/s, %5.1f%% of HBM)  '
              'pair %6.2f G steps/s' % (B, T, a, n / a / 1e6, 100 * n * 176 / (a * 1e-3) / 1e9 / peak, b, n / b / 1e6,
                                        100 * n * 320 / (b * 1e-3) / 1e9 / peak, n / (a + b) / 1e6), flush=True)


if __name__ == '__main__':
    main()
