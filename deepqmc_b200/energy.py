"""Caller-side helpers of the boundary (reference: src/deepqmc/loss/energy.py:19-74)."""
from __future__ import annotations

import torch

from . import parallel
from .types import PhysicalConfiguration


def compute_local_energy(rng, hamil, ansatz_apply, params, phys_conf: PhysicalConfiguration, batch_size=None):
    """reference: loss/energy.py:19-60.  phys_conf batch shape [mol, state, walker] (or [walker]);
    params: one tree per state (list) or a single tree.  Returns (E_loc[batch_shape], stats{key: mean over walkers})."""
    loc = hamil.local_energy(ansatz_apply)
    r, R = phys_conf.r, phys_conf.R
    if r.dim() == 3:
        E, stats = loc(rng, params if not isinstance(params, (list, tuple)) else params[0], phys_conf)
        return E, {k: v.mean(-1) for k, v in stats.items()}
    Mb, S, B = r.shape[:3]
    E = torch.empty(Mb, S, B, dtype=r.dtype, device=r.device)
    acc: dict = {}
    for m in range(Mb):
        for s in range(S):
            p = params[s] if isinstance(params, (list, tuple)) else params
            seed = None if rng is None else int(rng) * 1000003 + m * S + s
            e, st = loc(seed, p, PhysicalConfiguration(R[m, s, 0] if R.dim() == 5 else R, r[m, s], phys_conf.mol_idx[m, s]))
            E[m, s] = e
            for k, v in st.items():
                acc.setdefault(k, torch.empty(Mb, S, dtype=r.dtype, device=r.device))[m, s] = v.mean()
    return E, acc


def compute_mean_energy(local_energy, weight=None):
    """all-device mean (reference: loss/energy.py:63-74 -> parallel.all_device_mean)."""
    x = local_energy if weight is None else local_energy * weight
    return parallel.energy_statistics(x.reshape(-1))['energy/mean']
