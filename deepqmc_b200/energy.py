"""Caller-side helpers of the boundary (reference: src/deepqmc/loss/energy.py:19-74)."""
from __future__ import annotations

import torch

from . import parallel
from .types import PhysicalConfiguration


def compute_local_energy(rng, hamil, ansatz_apply, params, phys_conf: PhysicalConfiguration, batch_size=None,
                         jax_compatible_rng=False):
    """reference: loss/energy.py:19-60.  phys_conf batch shape [mol, state, walker] (or [walker]);
    params: one tree per state (list) or a single tree.  Returns (E_loc[batch_shape], stats{key: mean over walkers}).
    ``jax_compatible_rng``: the ECP quadrature twists follow the reference's streams for PRNGKey(rng): the key is split over
    the batch shape (:43), every walker folds in its nucleus slot and electron index."""
    loc = hamil.local_energy(ansatz_apply)
    r, R = phys_conf.r, phys_conf.R
    if r.dim() == 3:
        E, stats = loc(rng, params if not isinstance(params, (list, tuple)) else params[0], phys_conf,
                       jax_compatible_rng=jax_compatible_rng)
        return E, {k: v.mean(-1) for k, v in stats.items()}
    Mb, S, B = r.shape[:3]
    E = torch.empty(Mb, S, B, dtype=r.dtype, device=r.device)
    acc: dict = {}
    keys = None
    if jax_compatible_rng and rng is not None:
        from . import jaxrand

        keys = jaxrand.split(jaxrand.prng_key(int(rng)), Mb * S * B).reshape(Mb, S, B, 2)  # split(rng, batch_shape)
    for m in range(Mb):
        for s in range(S):
            p = params[s] if isinstance(params, (list, tuple)) else params
            seed = None if rng is None else (keys[m, s] if keys is not None else int(rng) * 1000003 + m * S + s)
            kw = {'jax_compatible_rng': True} if keys is not None else {}
            e, st = loc(seed, p, PhysicalConfiguration(R[m, s, 0] if R.dim() == 5 else R, r[m, s], phys_conf.mol_idx[m, s]), **kw)
            E[m, s] = e
            for k, v in st.items():
                acc.setdefault(k, torch.empty(Mb, S, dtype=r.dtype, device=r.device))[m, s] = v.mean()
    return E, acc


def compute_mean_energy(local_energy, weight=None):
    """all-device mean (reference: loss/energy.py:63-74 -> parallel.all_device_mean)."""
    x = local_energy if weight is None else local_energy * weight
    return parallel.energy_statistics(x.reshape(-1))['energy/mean']


def median_clip_and_mask(x, clip_width: float, median_center: bool = True, exclude_width: float = float('inf')):
    """Hard-clip local energies to `clip_width` mean absolute deviations around the (all-device) median / mean and
    flag outliers beyond `exclude_width` MADs (reference: loss/clip.py:73-98; the median needs the all-gather of
    E_loc, parallel.py:185-192)."""
    allx = parallel.all_gather_walkers(x.reshape(-1))
    center = all_walker_median(allx) if median_center else allx.mean()
    abs_all = (allx - center).abs()
    mad = abs_all.mean()
    x_clip = torch.clamp(x, center - clip_width * mad, center + clip_width * mad)
    return x_clip, (x - center).abs() < exclude_width


def all_walker_median(allx):
    """jnp.median convention (mean of the two middle values for an even count; torch.median would return the lower
    one) on the already gathered walkers (reference parallel.py:185-192)."""
    return torch.quantile(allx.reshape(-1), 0.5)


def log_squeeze(x):
    """reference utils.py:186-188"""
    sgn, a = torch.sign(x), x.abs()
    return sgn * torch.log1p((a + 0.5 * a**2 + a**3) / (1 + a**2))


def median_log_squeeze_and_mask(x, clip_width: float = 1.0, quantile: float = 0.95, exclude_width: float = float('inf')):
    """Soft squeeze toward the (all-device) median: x_med + 2w log_squeeze((x - x_med) / 2w), w = clip_width times the
    `quantile`-th quantile of |x - x_med| over all walkers (reference: loss/clip.py:101-141)."""
    allx = parallel.all_gather_walkers(x.reshape(-1))
    med = all_walker_median(allx)
    q = torch.quantile((allx - med).abs(), quantile)
    width = clip_width * q
    diff = x - med
    return med + 2 * width * log_squeeze(diff / (2 * width)), diff.abs() / q < exclude_width


def clip_local_energy(clip_mask_fn, local_energy):
    """Apply `clip_mask_fn` to every (molecule, electronic state) electron batch of local_energy[..., B]
    (reference: loss/clip.py:32-48; vmap over the two leading axes)."""
    lead = local_energy.shape[:-1]
    flat = local_energy.reshape(-1, local_energy.shape[-1])
    out = [clip_mask_fn(row) for row in flat]
    return (torch.stack([o[0] for o in out]).reshape(*lead, -1), torch.stack([o[1] for o in out]).reshape(*lead, -1))


def compute_mean_energy_tangent(local_energy, weight, gradient_mask, ansatz, params, phys_conf):
    """Gradient of the mean energy w.r.t. the ansatz parameters: the reference contracts
    (E_loc - <E_loc>) * weight * mask / n_mask with the parameter tangent of log|psi|
    (loss/energy.py:77-102 with loss/loss_function.py:53-82); here the same per-walker factors are the cotangent of
    ONE reverse pass through the CUDA engine (dqmc_wf_vjp_params).  Returns {haiku name: gradient} summed over all
    ranks (all-reduce of the packed gradient, reference optimizer.py:142 pmean)."""
    if local_energy.dim() != 1:
        # the reference subtracts the mean per (molecule, state) row (loss/energy.py:88, axis=-1); one call = one row here
        raise ValueError('compute_mean_energy_tangent takes the local energies of ONE (molecule, electronic state) batch [B]; '
                         'loop over the leading axes with that state\'s parameters (as overlap.compute_mean_overlap_tangent does)')
    E = local_energy
    w = torch.ones_like(E) if weight is None else weight.reshape(-1)
    mask = torch.ones_like(E, dtype=torch.bool) if gradient_mask is None else gradient_mask.reshape(-1)
    stats = parallel.energy_statistics(E * w)
    n_mask = mask.sum().double()
    if parallel.world()[1] > 1:
        torch.distributed.all_reduce(n_mask)
    cot = ((E - stats['energy/mean'].to(E.dtype)) * w * mask.to(E.dtype) / n_mask.to(E.dtype)).contiguous()
    _, grads = ansatz.log_psi_vjp(params, phys_conf, cot)
    if parallel.world()[1] > 1:
        keys = sorted(grads)
        flat = torch.cat([grads[k].reshape(-1) for k in keys])
        torch.distributed.all_reduce(flat)
        o = 0
        for k in keys:
            n = grads[k].numel()
            grads[k] = flat[o:o + n].reshape(grads[k].shape)
            o += n
    return grads
