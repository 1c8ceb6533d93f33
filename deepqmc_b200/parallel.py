"""Walker sharding and the per-step statistics exchange (one process per GPU).

reference: src/deepqmc/parallel.py (pmap axis 'device_axis', pmean / all_gather helpers,
scatter of walkers :296-317, per-device RNG :126-133); loss/energy.py:63-74 (mean energy =
all-device mean); loss/loss_function.py:187-188 (all_gather of E_loc).  Here: contiguous blocks
of B/world walkers per rank, parameters replicated, ONE all-gather of an 11-double packed
statistics vector per step (sums, count, max, -min; reduced locally) and (optionally) one all-gather of E_loc, over torch.distributed
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
    """Global mean / variance / min / max of E_loc and means of the 6 hamil stats with ONE collective per step: every
    rank packs {sum E, sum E^2, n, 6 stat sums, max, -min} into 11 doubles (one launch of the engine's stats kernel
    when the statistics come from it), one all_gather moves them, the reduction over ranks is local.
    (reference: loss/energy.py:63-74, observable.py:474-479)."""
    keys = sorted(stats) if stats else []
    raw, eng = getattr(stats, 'raw', None), getattr(stats, 'engine', None)
    if E_loc.is_cuda and eng is not None and raw is not None and E_loc.dim() == 1 and E_loc.is_contiguous():
        packed = eng.stats_pack(E_loc, raw)  # hamil.STAT_KEYS order
        from .hamil import STAT_KEYS

        order, i_max = [3 + STAT_KEYS.index(k) for k in keys], 9
    else:  # CPU tensors (gloo tests) or statistics that did not come from the engine
        E = E_loc.double()
        packed = torch.stack([E.sum(), (E * E).sum(), torch.tensor(float(E.numel()), device=E.device, dtype=torch.float64)]
                             + [stats[k].double().sum() for k in keys]
                             + [E.max(), -E.min()])
        order, i_max = [3 + i for i in range(len(keys))], 3 + len(keys)
    if world()[1] > 1:
        flat = torch.empty(world()[1] * packed.numel(), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(flat, packed)
        allp = flat.view(world()[1], packed.numel())
        packed = torch.cat([allp[:, :i_max].sum(0), allp[:, i_max:].max(0).values])
    n = packed[2]
    mean = packed[0] / n
    out = {'energy/mean': mean, 'energy/var': packed[1] / n - mean * mean, 'energy/max': packed[i_max], 'energy/min': -packed[i_max + 1],
           'energy/count': n}
    for i, k in zip(order, keys):
        out[k] = packed[i] / n
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
