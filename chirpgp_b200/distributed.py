"""Multi-GPU plumbing: independent chirps (and hyper-parameter candidates) are sharded across ranks -- one process per
GPU, launched by torchrun -- and never interact, so filtering / smoothing needs no data-path collective.  The single
exchange on this path is the scalar log-likelihood (and its small gradient vector) during MLE, which is summed with one
all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no distributed layer at all: its only
parallelism is vmap over chirps and independent OS processes (tetralith/run_filters_smoothers.sh:23-31)."""
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ['shard_range', 'shard', 'init_from_env', 'allreduce_objective']


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n independent units: ranks < n % world get one extra unit."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('bad rank / world: %d / %d' % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(x, rank: int, world: int, dim: int = 0):
    """This rank's slice of a batch-major array / tensor."""
    lo, hi = shard_range(x.shape[dim], rank, world)
    index = [slice(None)] * x.ndim
    index[dim] = slice(lo, hi)
    return x[tuple(index)]


def init_from_env(backend: Optional[str] = None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).  Returns
    (rank, world, local_rank).  Single-process runs return (0, 1, 0) without creating a group."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def allreduce_objective(value: torch.Tensor, grad: torch.Tensor, group=None):
    """Sum the per-rank partial objective(s) and gradient(s) in ONE collective: value (...,), grad (..., P) are packed
    into a (..., 1 + P) buffer (message size G x (1 + P) x 8 bytes: latency-bound, a few hundred bytes)."""
    packed = torch.cat([value.reshape(value.shape + (1,)), grad], dim=-1).contiguous()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[..., 0], packed[..., 1:]
