"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding of independent chirps and the single all-reduce of
the MLE objective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chirpgp_b200.distributed import shard_range, shard, allreduce_objective


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1000, 10007):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _partial_objective(ys, theta):
    # stand-in for the per-chirp nll and its gradient: any function that is a SUM over independent chirps
    val = (torch.sin(ys * theta[0]) ** 2).sum(dim=1) * theta[1]
    grad = torch.stack([(2 * torch.sin(ys * theta[0]) * torch.cos(ys * theta[0]) * ys).sum(dim=1) * theta[1],
                        (torch.sin(ys * theta[0]) ** 2).sum(dim=1)], dim=-1)
    return val.sum(), grad.sum(dim=0)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    ys = torch.randn(11, 50, dtype=torch.float64)            # identical on every rank; each takes its shard
    theta = torch.tensor([0.7, 1.3], dtype=torch.float64)
    mine = shard(ys, rank, world)
    val, grad = _partial_objective(mine, theta)
    val, grad = allreduce_objective(val, grad)
    full_val, full_grad = _partial_objective(ys, theta)
    ok = torch.allclose(val, full_val, rtol=1e-13) and torch.allclose(grad, full_grad, rtol=1e-13)
    # a grid of candidates at once: value (G,), grad (G, P)
    vals = torch.stack([_partial_objective(mine, theta * s)[0] for s in (1., 2., 3.)])
    grads = torch.stack([_partial_objective(mine, theta * s)[1] for s in (1., 2., 3.)])
    v2, g2 = allreduce_objective(vals, grads)
    fv = torch.stack([_partial_objective(ys, theta * s)[0] for s in (1., 2., 3.)])
    ok = ok and torch.allclose(v2, fv, rtol=1e-13) and g2.shape == (3, 2)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_objective_gloo_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))
