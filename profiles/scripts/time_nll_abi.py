#!/usr/bin/env python
"""Kernel-level timing of the EKF nll forward / adjoint pair through the C ABI (no torch autograd around it), over the
number of problems per GPU and the tuning knobs of csrc/cgp_nll2.cu.

    python profiles/scripts/time_nll_abi.py [--T 10000] [--every 16] [--sweep name=v1,v2 ...] [P ...]

Knobs (environment, read by the library at every call): CGP_NLL_LANES (8|16|32 problems per chain), CGP_NLL_KF / CGP_NLL_KB
(persistent 4-warp blocks per SM, forward / adjoint), CGP_NLL_UNIT_F / CGP_NLL_UNIT_B (time steps per scheduling unit).
One JSON line per (P, knob setting): fwd ms, bwd ms, nll+gradient steps/s."""
import argparse
import ctypes as C
import itertools
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg  # noqa: E402
from chirpgp_b200 import _native as N  # noqa: E402
from chirpgp_b200.filters_smoothers import _problem, _ptr  # noqa: E402
from chirpgp_b200.models import NC_LCD  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--T', type=int, default=10000)
    ap.add_argument('--G', type=int, default=16)
    ap.add_argument('--every', type=int, default=0)
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--raw', action='store_true')
    ap.add_argument('--sweep', action='append', default=[], help='ENVNAME=v1,v2,...')
    ap.add_argument('P', type=int, nargs='*')
    a = ap.parse_args()
    Ps = a.P or [160000, 20000]
    dev = torch.device('cuda', 0)
    L = N.lib()
    T, G = a.T, a.G
    dt = 3.141 / T
    lam = np.array([0.1, 0.4, 0.7, 1.0]); bb = np.array([0.05, 0.1, 0.2, 0.4])
    grid = np.array([[l, b_, 0.1, 1., 1., 7.] for l in lam for b_ in bb])[:G]
    _, _, mc, m0, P0, H = cg.build_chirp_model(grid)
    consts = mc.consts(dt).to(dev)
    H = H.to(dev)
    names = [s.split('=')[0] for s in a.sweep]
    values = [s.split('=')[1].split(',') for s in a.sweep]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for P in Ps:
        nch = P // G
        gen = torch.Generator(device='cuda').manual_seed(1234)
        ts_ = torch.linspace(dt, dt * T, T, dtype=torch.float64, device=dev)
        phase = 500 * torch.exp(-5 / torch.sin(ts_)) + 8 * ts_
        ys = torch.sin(2 * np.pi * phase)[None, :] + np.sqrt(0.1) * torch.randn((nch, T), dtype=torch.float64, device=dev, generator=gen)
        cb_ = consts.repeat(nch, 1).contiguous()
        m0_ = m0.to(dev).repeat(nch, 1).contiguous()
        P0_ = P0.to(dev).repeat(nch, 1, 1).contiguous()
        B = nch * G
        p = _problem(B, T, N.CGP_MODEL_LCD, 4, 1, cb_, NC_LCD, m0_, 4, P0_, 16, H, None, 0, None, 0.1, dt, G, 1)
        nll = torch.empty(B, dtype=torch.float64, device=dev)
        cbar = torch.empty((B, NC_LCD), dtype=torch.float64, device=dev)
        mbar = torch.empty((B, 4), dtype=torch.float64, device=dev)
        Pbar = torch.empty((B, 4, 4), dtype=torch.float64, device=dev)
        for combo in itertools.product(*values) if values else [()]:
            for k, v in zip(names, combo):
                os.environ[k] = v
            every = a.every or int(L.cgp_ekf_nll_default_ckpt(T))
            nbytes = L.cgp_ekf_nll_workspace_bytes(C.byref(p), every)
            ws = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
            bwd = L.cgp_ekf_nll_bwd_f64 if a.raw else L.cgp_ekf_nll_bwd_sym_f64
            best = (1e30, 0., 0.)
            for it in range(a.reps + 1):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                N.check(L.cgp_ekf_nll_fwd_f64(C.byref(p), _ptr(ys), _ptr(nll), _ptr(ws), C.c_size_t(nbytes), every, stream), 'fwd')
                e[1].record()
                N.check(bwd(C.byref(p), _ptr(ys), None, _ptr(ws), C.c_size_t(nbytes), every, _ptr(cbar), _ptr(mbar), _ptr(Pbar),
                            None, stream), 'bwd')
                e[2].record()
                torch.cuda.synchronize()
                tot = e[0].elapsed_time(e[2])
                if it > 0 and tot < best[0]:
                    best = (tot, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]))
            ok = bool(torch.isfinite(nll).all() and torch.isfinite(cbar).all())
            print(json.dumps(dict(problems=B, T=T, every=every, knobs=dict(zip(names, combo)), ms=round(best[0], 3),
                                  fwd_ms=round(best[1], 3), bwd_ms=round(best[2], 3), gsteps_per_s=round(B * T / best[0] / 1e6, 3),
                                  ws_gb=round(nbytes / 1e9, 2), finite=ok, nll_sum=float(nll.sum()),
                                  cbar_sum=float(cbar.sum()))), flush=True)
            del ws
        del ys


if __name__ == '__main__':
    main()
