"""Walker sharding and the per-step statistics exchange (one process per GPU).

reference: src/deepqmc/parallel.py (pmap axis 'device_axis', pmean / all_gather helpers,
scatter of walkers :296-317, per-device RNG :126-133); loss/energy.py:63-74 (mean energy =
all-device mean); loss/loss_function.py:187-188 (all_gather of E_loc).  Here: contiguous blocks
of B/world walkers per rank, parameters replicated, ONE fused all-reduce of the packed
statistics vector per step and (optionally) one all-gather of E_loc, over torch.distributed
(NCCL on GPUs, gloo in the CPU tests).  The walker update itself has no collective
(electron_samplers.py:121-126: acceptance and tau are per device).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's RANK/WORLD_SIZE/MASTER_* if WORLD_SIZE > 1."""
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend=backend)
    return world()


def shard_bounds(n_walkers: int, rank: int | None = None, world_size: int | None = None):
    """Contiguous walker block of this rank; electron_batch_size % device_count == 0 is required
    (reference: validate_kwargs.py:45-48)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if n_walkers % world_size:
        raise ValueError('electron batch size must be divisible by the number of devices')
    per = n_walkers // world_size
    return rank * per, (rank + 1) * per


def rank_seed(seed: int, rank: int | None = None) -> int:
    """Per-rank RNG stream (reference: train.py:134 seed + process_index)."""
    r, _ = world()
    return int(seed) + (r if rank is None else rank)


def energy_statistics(E_loc: torch.Tensor, stats: dict | None = None):
    """Global mean / variance / min / max of E_loc and means of the 6 hamil stats with ONE
    all-reduce(sum) of a packed fp64 vector plus one all-reduce(max) of a 2-vector
    (reference: loss/energy.py:63-74, observable.py:474-479)."""
    E = E_loc.double()
    keys = sorted(stats) if stats else []
    packed = torch.stack([E.sum(), (E * E).sum(), torch.tensor(float(E.numel()), device=E.device, dtype=torch.float64)]
                         + [stats[k].double().sum() for k in keys])
    mm = torch.stack([E.max(), -E.min()])
    if world()[1] > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        dist.all_reduce(mm, op=dist.ReduceOp.MAX)
    n = packed[2]
    mean = packed[0] / n
    out = {'energy/mean': mean, 'energy/var': packed[1] / n - mean * mean, 'energy/max': mm[0], 'energy/min': -mm[1],
           'energy/count': n}
    for i, k in enumerate(keys):
        out[k] = packed[3 + i] / n
    return out


def all_gather_walkers(x: torch.Tensor) -> torch.Tensor:
    """all_gather along the walker axis (reference: parallel.py:239-245; needed by the global
    median clipping of the training loss, loss/clip.py:93)."""
    _, w = world()
    if w == 1:
        return x
    out = [torch.empty_like(x) for _ in range(w)]
    dist.all_gather(out, x.contiguous())
    return torch.cat(out, 0)
