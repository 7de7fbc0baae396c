#!/usr/bin/env python
"""bench.py -- headline benchmark of the chirpgp_b200 hot path (contract: see the task statement / DESIGN.md).

Workload (BASELINE.json configs[1]): per GPU B = 1000 synthetic toymodel chirps, T = 3141,
dt = 1e-3, chirp model d = 4, Gauss-Hermite order 3 (81 sigma points), sgp_filter + sgp_smoother, float64.
One "step" = one filter + smoother pass over the whole batch.  metric = smoothed chirp time-steps / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); independent chirps are sharded across ranks (weak scaling:
every rank owns B = 1000 chirps), no data-path collective.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU = 1000
T = 3141
DT = 1e-3
XI = 0.1
PARAMS = [0.1, 0.1, 0.1, 1., 1., 7.]
D = 4
N_SIGMA = 81
METRIC = 'smoothed chirp time-steps/sec (batch x T), GHF+GHS'
WORKLOAD = ('configs[1]: 1000 toymodel chirps per GPU x T=3141, dt=1e-3, chirp model d=4, sgp_filter + sgp_smoother with '
            'gauss_hermite(d=4, order=3) (81 points)')
UNIT = 'steps/s'
N_DISTINCT = 6           # distinct input batches rotating through the batch-sequence legs (6 x 25 MB > L2)
E2E_DEPTH = int(os.environ.get('CGP_E2E_DEPTH', '8'))            # batches in flight in the end-to-end leg (cg.filter_smoother_batches); profiles/r2_batches.txt

# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel: read from the committed ncu capture of
# this very command (`ncu --set full`, raw page as CSV; see profiles/README.md), never a literal
NCU_RAW_CSV = os.path.join(ROOT, 'profiles', 'r2_ncu_headline_raw.csv')


def ncu_traffic_bytes(kernel_substr: str):
    """(bytes per launch, source) of the first kernel whose name contains `kernel_substr` in the committed ncu raw CSV."""
    import csv
    try:
        rows = list(csv.reader(open(NCU_RAW_CSV)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        scale = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        for r in rows[2:]:
            if kernel_substr in r[ik]:
                return (float(r[ir]) * scale.get(units[ir], 1.) + float(r[iw]) * scale.get(units[iw], 1.),
                        os.path.relpath(NCU_RAW_CSV, ROOT))
    except Exception:  # noqa: BLE001
        pass
    return None, None

# algorithmic bytes / flops per chirp time-step (DESIGN.md "Roofline accounting")
BYTES_FILTER = 8 + 8 * (D + D * D + 1)          # ys in, mf + Pf + nell out                      = 176
BYTES_SMOOTHER = 2 * 8 * (D + D * D)            # mf, Pf in; ms, Ps out                          = 320
BYTES_WORKSPACE = 8 * (D * D + D + D * (D + 1) // 2)   # [G | c | C packed] record the filter leaves for the sweep = 240 (not algorithmic)
FLOPS_FILTER = 9403                             # sgp_filter, d=4, n=81, counting convention v1
FLOPS_SMOOTHER = 13668                          # sgp_smoother


def flops_per_step(variant: str, d: int = 4, n: int = 81) -> int:
    """Counting convention v1 of SURVEY 8(d) (dense d x d algebra, no symmetry / sparsity credit, FMA = 2 flops, every
    exp / log / sin / cos / sqrt / div = 1 flop); model cost constants of the chirp family as tabulated there."""
    U = 7 * d * d + 7 * d + 9
    CH = d ** 3 / 3 + d * d
    harmonic = d > 4
    c = 80 if harmonic else 35           # LCD mean + Jacobian
    c_pt = 53 if harmonic else 21        # LCD mean per sigma point
    c_a, c_a_pt = 25, 12                 # drift + Jacobian, drift per sigma point
    SP = CH + n * (2 * d * d + d) + n * c_pt + 2 * n * d + 3 * n * d * d + 4 * d * d
    ST = CH + n * (2 * d * d + d) + n * c_a_pt + 2 * n * d + n * d + 3 * n * d * d + 2 * d * d
    table = {
        'ekf': c + 4 * d ** 3 + d * d + U,
        'eks': c + 4 * d ** 3 + d * d + 2 * d ** 3 + CH + 2 * d ** 3 + (2 * d * d + 2 * d) + 4 * d ** 3 + 2 * d * d,
        'sgp_filter': SP + U,
        'sgp_smoother': SP + 3 * n * d * d + 2 * d * d + CH + 2 * d ** 3 + (2 * d * d + 2 * d) + 4 * d ** 3 + 2 * d * d,
        'cd_ekf': 4 * (c_a + 4 * d ** 3 + 2 * d * d + 2 * (d + d * d)) + 6 * (d + d * d) + U,
        'cd_eks': CH + 2 * d ** 3 + 4 * (c_a + d * d + 4 * d * d + 4 * d ** 3 + 2 * d * d + 2 * (d + d * d)) + 6 * (d + d * d),
        'cd_sgp_filter': 4 * ST + 6 * (d + d * d) + U,
        'cd_sgp_smoother': 4 * ST + 6 * (d + d * d) + 4 * (4 * d ** 3 + 4 * d * d) + CH + 2 * d ** 3,
    }
    if variant not in table:
        raise ValueError(variant)
    return int(round(table[variant]))


def bytes_per_smoothed_step(d: int) -> int:
    """Algorithmic HBM bytes of one filter + smoother step: ys in, (mf, Pf, nell) out; (mf, Pf) in, (ms, Ps) out."""
    return 8 + 8 * (d + d * d + 1) + 2 * 8 * (d + d * d)


def synthetic_inputs(rank: int):
    from chirpgp_b200 import toymodels
    _, ys, _ = toymodels.synthetic_batch(B_PER_GPU, T, DT, Xi=XI, seed=2 + 1000 * rank)
    return ys


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_sample(n_chirps: int, nthreads: int = 0):
    """Times the CPU restatement of the reference path (oracle/, OpenMP over chirps) on `n_chirps` chirps of the
    same workload (more than 1000: the 1000 synthetic chirps repeated); returns (steps/s, seconds, threads)."""
    from oracle import oracle as orc
    from chirpgp_b200.quadratures import SigmaPoints
    from chirpgp_b200 import toymodels
    orc.build()
    _, ys, _ = toymodels.synthetic_batch(min(n_chirps, B_PER_GPU), T, DT, Xi=XI, seed=2)
    if n_chirps > ys.shape[0]:
        ys = np.tile(ys, (-(-n_chirps // ys.shape[0]), 1))[:n_chirps]
    spec = orc.ChirpSpec(PARAMS[0], PARAMS[1], PARAMS[3], PARAMS[4])
    m0, P0, H = orc.chirp_m0_P0_H(PARAMS[2], PARAMS[3], PARAMS[4], PARAMS[5])
    sg = SigmaPoints.gauss_hermite(D, 3)
    # all the host threads this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    threads = nthreads or len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    f = orc.sgp_filter(spec, sg, H, XI, m0, P0, DT, ys, nthreads=threads)
    orc.sgp_smoother(spec, sg, f[0], f[1], DT, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_chirps * T / dt, dt, threads


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path.  JAX cannot be installed in this image
    (no wheel, no network), so the arm runs the C restatement in oracle/ ('port') on all host threads, each step a
    bounded sample of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpu_sample(max(2 * cores, 16))                              # spin up the OpenMP team
    rate, _, _ = cpu_sample(max(4 * cores, 32))                 # calibrate, then ~8 s of CPU work per timed step
    n_chirps = int(min(8000, max(2 * cores, rate * 8. / T)))
    for _ in range(args.warmup):
        cpu_sample(max(cores, 4))
    vals, secs = [], []
    for _ in range(args.steps):
        v, s, threads = cpu_sample(n_chirps)
        vals.append(v); secs.append(s)
    value = statistics.mean(vals)
    sample = '%d chirps x %d steps per timed step (GHF+GHS, d=4, 81 points), C restatement, OpenMP' % (n_chirps, T)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * statistics.mean(secs), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        # the same `config` dict as the GPU arm prints (the driver compares them); what one timed CPU step covers is in `sample`
        'config': {'workload': WORKLOAD, 'batch_per_gpu': B_PER_GPU, 'T': T,
                   'parallelism': 'chirps sharded x%d, no collective' % int(os.environ.get('WORLD_SIZE', '1')),
                   'l2': 'flushed between timed steps (256 MiB write)'},
        'sample': sample,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'JAX not installable here: CPU arm = oracle/ C restatement of the reference algorithm (proxy)',
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ the other BASELINE configs
def _cpu_rate(fn, n_small, target_s, unit_steps):
    """Times fn(n) (oracle on n chirps) on a bounded sample sized for ~target_s of CPU work; returns (steps/s, n, seconds)."""
    fn(max(2, n_small // 4))                                     # spin up the OpenMP team
    t0 = time.perf_counter(); fn(n_small); dt = time.perf_counter() - t0
    n = int(max(n_small, min(50 * n_small, n_small * target_s / max(dt, 1e-6))))
    t0 = time.perf_counter(); fn(n); dt = time.perf_counter() - t0
    return n * unit_steps / dt, n, dt


def run_other_configs(dev, fp64_peak, hbm_peak, with_cpu, flush):
    """BASELINE.json configs[0], [2], [3] on this GPU: filter + smoother time (CUDA events, L2
    flushed, mean of 3 after warm-up), the two roofline fractions of SURVEY 8d, and the oracle's CPU rate beside each."""
    import torch
    import chirpgp_b200 as cg
    from chirpgp_b200 import toymodels
    out = []
    ev = lambda: torch.cuda.Event(enable_timing=True)
    threads = len(os.sched_getaffinity(0))

    def timed(filt, smooth, reps=3):
        for _ in range(2):
            f = filt(); sm = smooth(f); del f, sm
        tf, ts = [], []
        for _ in range(reps):
            flush.fill_(1.)
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(); f = filt(); e1.record(); sm = smooth(f); e2.record()
            torch.cuda.synchronize(dev)
            tf.append(e0.elapsed_time(e1)); ts.append(e1.elapsed_time(e2))
            del f, sm
        return statistics.mean(tf), statistics.mean(ts)

    def sequenced(pair, pargs, ys_dev, depth=8, n=24):
        """ms per batch of the same pair as a device-resident batch sequence (cg.filter_smoother_batches, `depth` in flight)."""
        def run(k):
            for _ in cg.filter_smoother_batches(pair, *pargs, batches=(ys_dev for _ in range(k)), depth=depth):
                pass
        run(2 * depth)
        torch.cuda.synchronize(dev)
        e0, e1 = ev(), ev()
        e0.record(); run(n); e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    def with_sequence(e, ms_seq, depth=8):
        e['batch_sequence'] = {'value': e['value'] * (e['filter_ms'] + e['smoother_ms']) / ms_seq, 'unit': UNIT, 'ms_per_batch': ms_seq,
                               'batches_in_flight': depth,
                               'what': 'same pair, device-resident, as a sequence of batches through cg.filter_smoother_batches'}
        return e

    def entry(config, what, B, d, n, variants, tf, ts, cpu):
        steps = B * T
        fl = sum(flops_per_step(v, d, n) for v in variants)
        sec = (tf + ts) * 1e-3
        e = {'config': config, 'workload': what, 'filter_ms': tf, 'smoother_ms': ts, 'value': steps / sec, 'unit': UNIT,
             'flops_per_step': fl, 'bytes_per_step': bytes_per_smoothed_step(d),
             'fp64_tflops': steps * fl / sec / 1e12, 'fp64_frac': steps * fl / sec / 1e12 / fp64_peak if fp64_peak else None,
             'hbm_gbs': steps * bytes_per_smoothed_step(d) / sec / 1e9,
             'hbm_frac': steps * bytes_per_smoothed_step(d) / sec / 1e9 / hbm_peak, 'cpu_baseline': cpu}
        if B == 1:
            e['us_per_step'] = sec * 1e6 / T
            e['note'] = 'single chirp: latency-bound, report us/step (SURVEY 8d)'
        return e

    from oracle import oracle as orc
    params = np.array(PARAMS)
    spec = orc.ChirpSpec(PARAMS[0], PARAMS[1], PARAMS[3], PARAMS[4])
    m0o, P0o, Ho = orc.chirp_m0_P0_H(PARAMS[2], PARAMS[3], PARAMS[4], PARAMS[5])
    drift, disp, mc, m0, P0, H = cg.build_chirp_model(params)
    m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
    sg = cg.SigmaPoints.gauss_hermite(4, 3)
    _, ys_all, _ = toymodels.synthetic_batch(B_PER_GPU, T, DT, Xi=XI, seed=2)

    def cpu_of(fn, n_small, target=3.):
        if not with_cpu:
            return None
        rate, n, sec = _cpu_rate(fn, n_small, target, T)
        return {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                'sample': '%d chirps x %d steps, oracle/ C restatement with OpenMP (%.1f s)' % (n, T, sec)}

    def tile(n):
        return ys_all if n <= ys_all.shape[0] else np.tile(ys_all, (-(-n // ys_all.shape[0]), 1))

    # ---- config 1: single chirp, EKF + EKS
    _, y1, _ = toymodels.synthetic_batch(1, T, DT, Xi=XI, seed=1)
    y1d = torch.as_tensor(y1[0]).to(dev)
    tf, ts = timed(lambda: cg.ekf(mc, H, XI, m0, P0, DT, y1d), lambda f: cg.eks(mc, f[0], f[1], DT))

    def cpu1(n):
        f = orc.ekf(spec, Ho, XI, m0o, P0o, DT, tile(n)[:n], nthreads=threads)
        orc.eks(spec, f[0], f[1], DT, nthreads=threads)
    out.append(entry(1, 'configs[0]: single toymodel chirp x T=%d, ekf + eks, d=4' % T, 1, 4, 0, ('ekf', 'eks'), tf, ts,
                     cpu_of(cpu1, 256)))
    # ---- config 3: continuous-discrete, 1000 chirps
    ysd = torch.as_tensor(ys_all).to(dev)
    tf, ts = timed(lambda: cg.cd_ekf(drift, disp, H, XI, m0, P0, DT, ysd), lambda f: cg.cd_eks(drift, disp, f[0], f[1], DT))
    Bm = spec.dispersion_matrix()

    def cpu3a(n):
        f = orc.cd_ekf(spec, Bm, Ho, XI, m0o, P0o, DT, tile(n)[:n], nthreads=threads)
        orc.cd_eks(spec, Bm, f[0], f[1], DT, nthreads=threads)
    out.append(with_sequence(entry(3, 'configs[2]: %d chirps x T=%d, cd_ekf + cd_eks (one RK4 step per sample)' % (B_PER_GPU, T),
                                   B_PER_GPU, 4, 0, ('cd_ekf', 'cd_eks'), tf, ts, cpu_of(cpu3a, 128)),
                             sequenced(cg.cd_ekf_smoother, (drift, disp, H, XI, m0, P0, DT), ysd)))
    bm = disp(None)
    tf, ts = timed(lambda: cg.cd_sgp_filter(drift, bm, sg, H, XI, m0, P0, DT, ysd),
                   lambda f: cg.cd_sgp_smoother(drift, bm, sg, f[0], f[1], DT))

    def cpu3b(n):
        f = orc.cd_sgp_filter(spec, Bm, sg, Ho, XI, m0o, P0o, DT, tile(n)[:n], nthreads=threads)
        orc.cd_sgp_smoother(spec, Bm, sg, f[0], f[1], DT, nthreads=threads)
    out.append(with_sequence(entry(3, 'configs[2]: %d chirps x T=%d, cd_sgp_filter + cd_sgp_smoother, gauss_hermite(4, 3)' % (B_PER_GPU, T),
                                   B_PER_GPU, 4, 81, ('cd_sgp_filter', 'cd_sgp_smoother'), tf, ts, cpu_of(cpu3b, 16)),
                             sequenced(cg.cd_sgp_filter_smoother, (drift, bm, sg, H, XI, m0, P0, DT), ysd, n=16)))
    # ---- config 4: harmonic model d = 8, cubature
    _, ys4, _ = toymodels.synthetic_batch(B_PER_GPU, T, DT, num_harmonics=3, seed=4)
    ys4d = torch.as_tensor(ys4).to(dev)
    _, _, mc4, m04, P04, H4 = cg.build_harmonic_chirp_model(params, num_harmonics=3)
    m04, P04, H4 = m04.to(dev), P04.to(dev), H4.to(dev)
    sg4 = cg.SigmaPoints.cubature(8)
    tf, ts = timed(lambda: cg.sgp_filter(mc4, sg4, H4, XI, m04, P04, DT, ys4d), lambda f: cg.sgp_smoother(mc4, sg4, f[0], f[1], DT))
    spec4 = orc.ChirpSpec(PARAMS[0], PARAMS[1], PARAMS[3], PARAMS[4], num_harmonics=3)
    m0o4, P0o4, Ho4 = orc.chirp_m0_P0_H(PARAMS[2], PARAMS[3], PARAMS[4], PARAMS[5], num_harmonics=3, kind='harmonic')

    def cpu4(n):
        yy = ys4 if n <= ys4.shape[0] else np.tile(ys4, (-(-n // ys4.shape[0]), 1))
        f = orc.sgp_filter(spec4, sg4, Ho4, XI, m0o4, P0o4, DT, yy[:n], nthreads=threads)
        orc.sgp_smoother(spec4, sg4, f[0], f[1], DT, nthreads=threads)
    out.append(with_sequence(entry(4, 'configs[3]: %d harmonic chirps (3 harmonics, d=8) x T=%d, sgp_filter + sgp_smoother, cubature(8)'
                                   % (B_PER_GPU, T), B_PER_GPU, 8, 16, ('sgp_filter', 'sgp_smoother'), tf, ts, cpu_of(cpu4, 32)),
                             sequenced(cg.sgp_filter_smoother, (mc4, sg4, H4, XI, m04, P04, DT), ys4d, depth=4, n=16), depth=4))
    torch.cuda.empty_cache()
    # ---- config 2 at the batch size north_star names for the CPU comparison: 10 000 chirps (large-batch kernel, cgp_oct.cuh)
    del ysd, ys4d
    torch.cuda.empty_cache()
    B10 = 10000
    y10 = torch.as_tensor(np.ascontiguousarray(tile(B10)[:B10])).to(dev)
    tf, ts = timed(lambda: cg.sgp_filter(mc, sg, H, XI, m0, P0, DT, y10), lambda f: cg.sgp_smoother(mc, sg, f[0], f[1], DT))
    e = entry('2 @ 10 000 chirps', 'configs[1] at 10 000 chirps x T=%d (north_star: "10k-chirp batch"), sgp_filter + sgp_smoother, '
              'gauss_hermite(4, 3): gh_oct_filter_kernel (8 lanes per chirp, smoother records inline) + sweep' % T, B10, 4, 81,
              ('sgp_filter', 'sgp_smoother'), tf, ts, None)
    e['filter_fp64_frac'] = (B10 * T * flops_per_step('sgp_filter') / (tf * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None
    # the pair's flop count includes the smoother's own sigma-point prediction, which the fused filter does not execute
    e['fp64_tflops'] = e['fp64_frac'] = None
    e['note'] = ('cpu_baseline: the headline cpu_baseline of this line (same workload per chirp); filter_fp64_frac counts the '
                 'algorithmic flops of sgp_filter only against the filter kernel, as `roofline` does')
    out.append(e)
    del y10
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ config 5: the MLE sweep
MLE_CHIRPS, MLE_T, MLE_G = 10000, 100000, 16


def run_mle_sweep(world, rank, dev, fp64_peak, chirps=MLE_CHIRPS, T5=MLE_T, reps=2):
    """BASELINE.json configs[4]: EKF nll + adjoint gradient of 10 000 chirps x 16 hyper-parameter candidates
    x T = 1e5, chirps sharded over the ranks (STRONG scaling: the total is fixed), one NCCL all-reduce of the 16 x (1 + 6)
    objective / gradient block per evaluation -- the only collective on the path (demos/ekfs_mle.py:42-49,
    tetralith/run_crlbs.sh:1-2)."""
    import torch
    import torch.distributed as dist
    import chirpgp_b200 as cg
    from chirpgp_b200 import mle
    from chirpgp_b200.distributed import shard_range, allreduce_objective
    from chirpgp_b200.models import g as gfun
    dt5 = 3.141 / T5                                        # total duration < pi keeps toymodels.meow_freq valid
    lo, hi = shard_range(chirps, rank, world)
    nloc = hi - lo
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    ts_ = torch.linspace(dt5, dt5 * T5, T5, dtype=torch.float64, device=dev)
    phase = 500 * torch.exp(-5 / torch.sin(ts_)) + 8 * ts_
    ys = torch.sin(2 * np.pi * phase)[None, :].repeat(nloc, 1)
    for i in range(0, nloc, 256):                           # noise in slabs: no second 8 GB temporary
        ys[i:i + 256] += np.sqrt(XI) * torch.randn((min(256, nloc - i), T5), dtype=torch.float64, device=dev, generator=gen)
    lam = np.array([0.1, 0.4, 0.7, 1.0]); bb = np.array([0.05, 0.1, 0.2, 0.4])
    grid = np.array([[l, b_, 0.1, 1., 1., 7.] for l in lam for b_ in bb])
    theta = torch.tensor(np.log(np.exp(grid) - 1.), dtype=torch.float64, device=dev, requires_grad=True)
    H = np.array([0., 1., 0., 0.])
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def objective(marks=None):
        _, _, mc, m0, P0, _ = cg.build_chirp_model(gfun(theta))
        nll = mle.ekf_nll(mc, H, XI, m0, P0, dt5, ys, candidates=True)              # (nloc, G): forward kernel
        val = nll.sum(dim=0)
        grad, = torch.autograd.grad(val.sum(), theta)                               # adjoint kernel; rows independent
        if marks is not None:
            marks[0].record()
        out = allreduce_objective(val.detach(), grad)                               # G x (1 + 6) doubles, one NCCL call
        if marks is not None:
            marks[1].record()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    objective()
    barrier()
    ms, ms_ar = [], []
    for _ in range(reps):
        e0, e1, ea, eb = ev(), ev(), ev(), ev()
        barrier()
        e0.record()
        val, grad = objective((ea, eb))
        e1.record()
        torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1)); ms_ar.append(ea.elapsed_time(eb))
    t = torch.tensor([statistics.mean(ms), statistics.mean(ms_ar)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_eval, ms_allreduce = float(t[0]), float(t[1])
    steps = chirps * MLE_G * T5
    fl = 3 * flops_per_step('ekf')
    free, total = torch.cuda.mem_get_info(dev)
    res = {'workload': 'configs[4]: EKF nll + adjoint gradient, %d chirps x %d candidates x T=%d, chirps sharded x%d'
                       % (chirps, MLE_G, T5, world),
           'metric': 'nll+gradient chirp time-steps/sec (B x G x T)', 'value': steps / (ms_eval * 1e-3), 'unit': UNIT,
           'n_gpus': world, 'scaling': 'strong', 'ms_per_evaluation': ms_eval, 'ms_in_allreduce': ms_allreduce,
           'evaluations_timed': reps, 'problems_per_gpu': nloc * MLE_G,
           'collective': 'one all-reduce of %d x (1 + 6) doubles per evaluation (NCCL)' % MLE_G,
           'flops_per_step': fl, 'flops_convention': '3 x flops_per_step(ekf): forward + recomputation + adjoint (SURVEY 8d v1)',
           'fp64_tflops_per_gpu': steps * fl / (ms_eval * 1e-3) / 1e12 / world,
           'fp64_frac': steps * fl / (ms_eval * 1e-3) / 1e12 / world / fp64_peak if fp64_peak else None,
           'nll_sum': float(val.sum()), 'grad_norm': float(grad.norm()), 'hbm_in_use_gb': (total - free) / 1e9,
           'checkpointing': getattr(mle._default_ckpt, 'last', None)}
    del ys
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import chirpgp_b200 as cg
    from chirpgp_b200 import _native

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    ys_host = torch.as_tensor(synthetic_inputs(rank)).pin_memory()
    ys = ys_host.to(dev)
    _, _, m_and_cov, m0, P0, H = cg.build_chirp_model(np.array(PARAMS))
    sgps = cg.SigmaPoints.gauss_hermite(d=D, order=3)
    m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys)
        s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)
        del f, s
    # ---- device-resident timing: K steps, L2 flushed between steps (flush outside the per-step events)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    marks = []
    for _ in range(args.steps):
        flush.fill_(1.)
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys)
        e1.record()
        s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)
        e2.record()
        marks.append((e0, e1, e2))
        del f, s
    barrier()
    t_filter = [a.elapsed_time(b) for a, b, _ in marks]
    t_step = [a.elapsed_time(c) for a, _, c in marks]
    ms_step = statistics.mean(t_step)
    ms_filter = statistics.mean(t_filter)

    # ---- device-resident throughput of a SEQUENCE of batches through the product's streaming call (E2E_DEPTH batches in flight
    # on alternating streams): the filter of batch i+1 (latency-bound, leaves most issue slots idle) overlaps the tail of batch
    # i's filter and its sweep.  All five outputs (mfs, Pfs, n_ell, mss, Pss) are produced on the device for every batch.
    # Reported as an extra key -- it is the device-resident counterpart of `e2e`; `value` stays the serial, L2-flushed figure.
    n_pipe = max(3 * args.steps, 10 * E2E_DEPTH)
    more_inputs = [torch.as_tensor(synthetic_inputs(rank + 1000 * k)).pin_memory() for k in range(1, N_DISTINCT)]
    dev_inputs = [ys] + [t.to(dev) for t in more_inputs]

    def dev_sequence(n):
        for _ in cg.filter_smoother_batches(cg.sgp_filter_smoother, m_and_cov, sgps, H, XI, m0, P0, DT,
                                            batches=(dev_inputs[i % N_DISTINCT] for i in range(n)), depth=E2E_DEPTH):
            pass
    dev_sequence(4 * E2E_DEPTH); dev_sequence(2 * E2E_DEPTH)   # warm-up: every stream's device pool gets allocated (cudaMalloc synchronises)
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    dev_sequence(n_pipe)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_pipe = e0.elapsed_time(e1) / n_pipe

    # ---- end-to-end through the PRODUCT API, host buffers in, host buffers out, every copy inside the timed region:
    #   e2e            cg.sgp_filter_smoother(..., ys_host, readout=('freq', 'v_var')): the posterior frequency estimate
    #                  E[g(V_k)] and the marginal variance -- what the demos / jobs compute from the smoother output right
    #                  afterwards (demos/ghfs_mle.py:87-89, quadratures.py:234-274) -- 16 bytes per step come back;
    #   e2e_full       the same call returning all of (mss, Pss): 160 bytes per step come back (PCIe-bound);
    #   e2e_numpy_api  the literal drop-in: NumPy in, sgp_filter then sgp_smoother as two calls, NumPy out.
    # Blocking calls issued back to back on the default stream (what `for batch in batches: f(batch)` does).
    n_e2e = max(4, min(args.steps, 10))
    H_h, m0_h, P0_h = H.cpu(), m0.cpu(), P0.cpu()

    def timed_calls(fn, n):
        out = fn(); out = fn(); out = fn()      # steady state: the previous result is alive while the next call allocates
        barrier()
        flush.fill_(1.)
        torch.cuda.synchronize(dev)
        e0, e1 = ev(), ev()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) * 1e3 / n
        return e0.elapsed_time(e1) / n, wall, out

    ms_e2e_blk, ms_e2e_blk_wall, out_r = timed_calls(
        lambda: cg.sgp_filter_smoother(m_and_cov, sgps, H, XI, m0, P0, DT, ys_host, readout=('freq', 'v_var')), n_e2e)
    # the same work as a SEQUENCE of batches through cg.filter_smoother_batches (E2E_DEPTH batches in flight on alternating
    # streams inside the product): what a Monte-Carlo job does (tetralith/jobs/ghfs_mle.py:26-86: one call per run).  Distinct
    # pinned input batches in rotation; every batch's measurements cross PCIe (asynchronous upload on the batch's stream) and
    # its 16 B/step readout comes back into pinned host memory; the clock stops when the last result has landed.  N_DISTINCT
    # input batches (150 MB > the 126 MB L2) rotate, and every batch writes 1.8 GB between two uses of the same input.
    hosts = [ys_host] + more_inputs
    n_seq = max(3 * args.steps, 10 * E2E_DEPTH)

    def batch_sequence(n):
        last = None
        for last in cg.filter_smoother_batches(cg.sgp_filter_smoother, m_and_cov, sgps, H, XI, m0, P0, DT,
                                               batches=(hosts[i % N_DISTINCT] for i in range(n)), readout=('freq', 'v_var'),
                                               depth=E2E_DEPTH):
            pass
        return last
    batch_sequence(4 * E2E_DEPTH); batch_sequence(4 * E2E_DEPTH)      # warm-up: the per-stream device pools and the pinned result blocks get allocated
    # three timed sequences of n_seq batches each; the median is reported (all three are in the line): the host path of a
    # sequence shows occasional slow episodes from one run to the next on the same box (2.4 ... 6 ms per batch) that the
    # device-resident sequence does not
    e2e_samples, e2e_walls = [], []
    for _ in range(3):
        barrier()
        flush.fill_(1.)
        torch.cuda.synchronize(dev)
        e0, e1 = ev(), ev()
        t0 = time.perf_counter()
        e0.record()
        batch_sequence(n_seq)
        e1.record()
        torch.cuda.synchronize(dev)
        e2e_walls.append((time.perf_counter() - t0) * 1e3 / n_seq)
        e2e_samples.append(e0.elapsed_time(e1) / n_seq)
    ms_e2e = statistics.median(e2e_samples)
    ms_e2e_wall = statistics.median(e2e_walls)
    del hosts[1:], more_inputs, dev_inputs[1:]
    ms_e2e_full, ms_e2e_full_wall, _ = timed_calls(
        lambda: cg.sgp_filter_smoother(m_and_cov, sgps, H, XI, m0, P0, DT, ys_host, readout=('mss', 'Pss')), n_e2e)
    ys_np = ys_host.numpy()

    def numpy_api():
        f = cg.sgp_filter(m_and_cov, sgps, H_h.numpy(), XI, m0_h.numpy(), P0_h.numpy(), DT, ys_np)
        return cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)
    ms_e2e_np, _, _ = timed_calls(numpy_api, max(2, n_e2e // 2))
    clocks = sampler.stop()              # sampled through the timed regions (serial, two-stream, end-to-end)
    freq_mean = float(out_r[0].mean())
    # bare device->host ceiling with every rank copying at once: the 402 MB of Pss into an existing pinned buffer (what bounds
    # e2e_full_outputs; the ranks of one box share the host's PCIe / DRAM path)
    src = torch.empty((B_PER_GPU, T, D, D), dtype=torch.float64, device=dev)
    dst = torch.empty((B_PER_GPU, T, D, D), dtype=torch.float64).pin_memory()
    dst.copy_(src, non_blocking=True)
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_d2h = e0.elapsed_time(e1) / 3
    del src, dst
    try:
        import jax  # noqa: F401
        jax_note = 'importable: version %s' % getattr(jax, '__version__', '?')
    except Exception:  # noqa: BLE001
        jax_note = 'not installed (reference arm = C restatement; chirpgp_b200/jax_ffi.py never executed)'

    # ---- FP64 peak (DFMA-only kernel) on this GPU
    L = _native.lib()
    blocks, iters = 148 * 8, 20000
    out = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    best = 0.
    for i in range(4):
        e0, e1 = ev(), ev()
        e0.record()
        fl = L.cgp_bench_dfma(C.c_void_p(out.data_ptr()), blocks, iters, stream)
        e1.record()
        torch.cuda.synchronize(dev)
        if i > 0:
            best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    fp64_peak = best

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.))
    hbm_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'

    # ---- the other BASELINE configs (rank 0 of a single-GPU run) and the MLE sweep (every N: sharded, with the all-reduce)
    other = None
    if world == 1 and not args.no_configs:
        other = run_other_configs(dev, fp64_peak, hbm_peak, not args.no_cpu, flush)
    torch.cuda.empty_cache()
    mle_line = None
    if not args.no_mle:
        mle_line = run_mle_sweep(world, rank, dev, fp64_peak, chirps=args.mle_chirps, T5=args.mle_T)

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_step, ms_filter, ms_e2e, ms_pipe, ms_e2e_full, ms_e2e_np, ms_d2h, ms_e2e_blk], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_filter, ms_e2e, ms_pipe, ms_e2e_full, ms_e2e_np, ms_d2h, ms_e2e_blk = [float(x) for x in t.tolist()]
    n_steps_total = world * B_PER_GPU * T
    value = n_steps_total / (ms_step * 1e-3)

    line = None
    if rank == 0:
        filt_bytes = B_PER_GPU * T * BYTES_FILTER
        filt_flops = B_PER_GPU * T * flops_per_step('sgp_filter')
        ach_gbs = filt_bytes / (ms_filter * 1e-3) / 1e9
        ach_tf = filt_flops / (ms_filter * 1e-3) / 1e12
        traffic, traffic_src = ncu_traffic_bytes('gh_duo_filter_kernel')
        # CPU baseline on a bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            cpu_sample(max(2 * cores, 16))                      # spin up the OpenMP team
            rate, _, _ = cpu_sample(max(4 * cores, 32))         # calibrate, then ~12 s of CPU work
            n = int(min(8000, max(2 * cores, rate * 12. / T)))
            v, sec, threads = cpu_sample(n)
            cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                   'sample': '%d chirps x %d steps, GHF+GHS, oracle/ C restatement with OpenMP (%.1f s)' % (n, T, sec)}
        if cpu and other:
            for e in other:
                if isinstance(e.get('config'), str) and e['config'].startswith('2 @'):
                    e['x_cpu_baseline'] = e['value'] / cpu['value']      # device-resident rate / the headline cpu_baseline
        cyc = ms_filter * 1e-3 * (clocks.get('sm_mhz') or 1965.) * 1e6 / T
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_gpu': B_PER_GPU, 'T': T,
                       'parallelism': 'chirps sharded x%d, no collective' % world,
                       'l2': 'flushed between timed steps (256 MiB write)'},
            'clocks': clocks,
            'value_batch_sequence': {'value': n_steps_total / (ms_pipe * 1e-3), 'unit': UNIT, 'ms_per_step': ms_pipe,
                                  'batches_in_flight': E2E_DEPTH, 'steps': n_pipe,
                                  'what': 'same passes, device-resident, as a sequence of batches through cg.filter_smoother_batches '
                                          '(all five outputs on the device per batch; %d distinct input batches in rotation, 1.8 GB written per batch)' % N_DISTINCT},
            'e2e': {'value': n_steps_total / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': B_PER_GPU * T * 8, 'd2h_bytes_per_step': B_PER_GPU * T * 8 * 2, 'steps': n_seq,
                    'ms_per_step_wall_clock': ms_e2e_wall, 'batches_in_flight': E2E_DEPTH,
                    'ms_per_step_samples': e2e_samples, 'statistic': 'median of 3 sequences of %d batches (rank 0 samples shown; max over ranks of the medians)' % n_seq,
                    'what': "a sequence of batches through the product's streaming call: for freq, v_var in "
                            "cg.filter_smoother_batches(cg.sgp_filter_smoother, m_and_cov, sgps, H, Xi, m0, P0, dt, batches=<pinned "
                            "host ys, %d distinct batches in rotation>, readout=('freq', 'v_var'), depth=%d) -- every batch: pinned host "
                            "ys uploaded on the batch's stream (asynchronous H2D copy inside the product call), "
                            "posterior frequency estimate E[g(V_k)] (gaussian_expectation on the device) and marginal variance out into "
                            "pinned host memory, 16 B/step (demos/ghfs_mle.py:87-89); the batches in flight overlap on the device (kernels of "
                            "different batches share the SMs, sweep / readout / D2H run under the next filters); the library is told "
                            "how many batches are in flight (CgpProblem.in_flight) and runs gh_oct_filter_kernel (8 lanes per chirp) "
                            "from 4000 chirps in flight; timed from the first call to the last result on the host" % (N_DISTINCT, E2E_DEPTH),
                    'kernels_per_step': ['gh_oct_filter_kernel (sgp_filter + smoother records)', 'smoother_sweep_lane4_kernel',
                                         'expect_softplus_kernel (freq readout)']},
            'e2e_blocking': {'value': n_steps_total / (ms_e2e_blk * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e_blk,
                             'h2d_bytes_per_step': B_PER_GPU * T * 8, 'd2h_bytes_per_step': B_PER_GPU * T * 8 * 2, 'steps': n_e2e,
                             'ms_per_step_wall_clock': ms_e2e_blk_wall, 'check_mean_frequency_hz': freq_mean,
                             'what': "one blocking product call per step, nothing overlapped: cg.sgp_filter_smoother(m_and_cov, sgps, H, "
                                     "Xi, m0, P0, dt, ys_host, readout=('freq', 'v_var')); the pinned host ys are read in place over PCIe "
                                     "by the filter kernel (the h2d bytes cross the bus inside the kernel)"},
            'e2e_full_outputs': {'value': n_steps_total / (ms_e2e_full * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e_full,
                                 'h2d_bytes_per_step': B_PER_GPU * T * 8, 'd2h_bytes_per_step': B_PER_GPU * T * 8 * (D + D * D),
                                 'd2h_gb_per_s': B_PER_GPU * T * 8 * (D + D * D) / (ms_e2e_full * 1e-3) / 1e9,
                                 'bare_d2h_gb_per_s_per_rank': B_PER_GPU * T * 8 * D * D / (ms_d2h * 1e-3) / 1e9,
                                 'what': "same call with readout=('mss', 'Pss'): 160 B/step come back, PCIe-bound; bare_d2h = a plain "
                                         "402 MB device->pinned-host copy issued by all ranks at once (max over ranks)"},
            'e2e_numpy_api': {'value': n_steps_total / (ms_e2e_np * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e_np,
                              'what': 'literal drop-in: NumPy ys -> sgp_filter -> (mfs, Pfs, n_ell) NumPy -> sgp_smoother -> '
                                      '(mss, Pss) NumPy; blocking, filtering result goes down and up again'},
            'gpu_launches': args.steps * 2,
            'kernels_per_step': ['gh_duo_filter_kernel (sgp_filter + smoother gains)', 'smoother_sweep_lane4_kernel (sgp_smoother)'],
            # the binding bound of GHF+GHS is the FP64 pipe; HBM is the extra
            'roofline': {'bound': 'fp64', 'kernel': 'gh_duo_filter_kernel (sgp_filter + smoother gains, producer/consumer warp pair per chirp)',
                         'achieved': ach_tf, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': ach_tf / fp64_peak if fp64_peak else None,
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'flops_per_step': flops_per_step('sgp_filter'), 'kernel_ms': ms_filter,
                         'peak_source': 'DFMA-only kernel measured in this run (MEASURED_PEAKS.json holds no FP64 figure; nominal 37)',
                         'note': 'algorithmic flops of sgp_filter only (convention v1); the kernel also evaluates the smoother '
                                 'gains, which are not counted'},
            'roofline_hbm': {'bound': 'hbm', 'kernel': 'gh_duo_filter_kernel', 'achieved': ach_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                             'frac': ach_gbs / hbm_peak, 'peak_source': hbm_src, 'algorithmic_bytes_per_step': BYTES_FILTER,
                             'workspace_bytes_per_step': BYTES_WORKSPACE},
            # at 1000 chirps per GPU neither throughput roofline is reachable: the kernel time is T x (cycles one producer
            # warp needs per step); floor = the pure dependency latency of one step
            'roofline_chain': {'bound': 'dependency-chain latency of one filter step', 'kernel': 'gh_duo_filter_kernel',
                               'achieved_cycles_per_step': cyc, 'floor_cycles_per_step': 840,
                               'alone_cycles_per_step': 1448, 'frac': 840. / cyc,
                               'source': 'floor: dependent-issue latencies of profiles/microbench/fp64_latency.cu; alone: one '
                                         'chain warp per SM sub-partition, profiles/r2_chain_timeline_plain592.txt'},
            # the batch-sequence legs run the 8-lane kernel with ~3 launches overlapping: no per-kernel time exists, so the whole
            # per-batch time (filter + records + sweep of a batch) is charged against the filter's algorithmic flops -- a lower bound
            'roofline_batch_sequence': {'bound': 'fp64', 'kernel': 'gh_oct_filter_kernel (+ smoother_sweep_lane4_kernel), %d batches in flight' % E2E_DEPTH,
                                        'achieved': filt_flops / (ms_pipe * 1e-3) / 1e12, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                                        'frac': filt_flops / (ms_pipe * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                                        'ms_per_batch': ms_pipe, 'flops_per_step': flops_per_step('sgp_filter'),
                                        'note': 'algorithmic flops of sgp_filter only over the whole per-batch time of the '
                                                'device-resident sequence; ncu of the kernel alone at 10 000 chirps: FP64 pipe 61 % busy '
                                                '(profiles/r2_oct_kernel.txt)'},
            'cpu_baseline': cpu,
            'jax': jax_note,
            'configs': other,
            'mle_sweep': mle_line,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline legs')
    ap.add_argument('--no-configs', action='store_true', help='skip the `configs` key (BASELINE configs 1, 3, 4)')
    ap.add_argument('--no-mle', action='store_true', help='skip the `mle_sweep` key (BASELINE config 5)')
    ap.add_argument('--mle-chirps', type=int, default=MLE_CHIRPS)
    ap.add_argument('--mle-T', type=int, default=MLE_T)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
