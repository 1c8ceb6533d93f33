"""Excited-state overlap estimates on top of the CUDA wave-function forward.

Mirror of the value-level part of the reference's ``loss/overlap.py``: all wave functions on the
samples of all wave functions (``compute_wave_function_values`` :19-49), sample-wise ratios
(:52-74), their batch version (:77-99), the clipped-geometric-mean symmetrisation (:102-121) and
the mean overlap / penalty (:124-150).  ``params`` is a sequence with one parameter tree per
electronic state, ``phys_conf`` carries a leading state axis ``[n_wfs, B, ...]``.  Each
``Psi_i(r ~ Psi_j^2)`` block is one ``dqmc_wf_forward`` call (plain-forward kernels); the small
[n_wfs, n_wfs, B] algebra that follows is elementwise torch on the device.  The parameter
tangent of the overlap (:182-229) is one reverse pass (``dqmc_wf_vjp_params``, row N1) per state;
ratio clipping follows loss/clip.py:51-70,144-174.
"""
from __future__ import annotations

import torch

from . import parallel
from .types import PhysicalConfiguration, Psi


def compute_wave_function_values(ansatz, params, phys_conf: PhysicalConfiguration):
    """-> Psi with sign/log [n_wfs (i), n_wfs (j), B]:  Psi_i(r ~ Psi_j^2)"""
    n = len(params)
    assert phys_conf.r.shape[0] == n, 'leading axis of phys_conf must run over the electronic states'
    signs, logs = [], []
    for i in range(n):
        R = phys_conf.R[0] if phys_conf.R.dim() == 4 else phys_conf.R
        R = R[0] if R.dim() == 3 else R
        B = phys_conf.r.shape[1]
        flat = PhysicalConfiguration(R, phys_conf.r.reshape(n * B, *phys_conf.r.shape[2:]),
                                     torch.zeros(n * B, device=phys_conf.r.device))
        psi = ansatz.apply(params[i], flat)  # one forward over the samples of ALL states
        signs.append(psi.sign.reshape(n, B))
        logs.append(psi.log.reshape(n, B))
    return Psi(torch.stack(signs), torch.stack(logs)), {}


def compute_psi_ratio(ansatz, params, phys_conf):
    """ratio[i, j, b] = Psi_i(r_b ~ Psi_j^2) / Psi_j(r_b ~ Psi_j^2)  (reference :52-99; the mean
    log magnitude of each wave function is subtracted before exponentiating)."""
    psi, stats = compute_wave_function_values(ansatz, params, phys_conf)
    mean_log = psi.log.mean(dim=(-1, -2))  # [n_wfs]
    shifted = psi.log - mean_log[:, None, None]
    diag_log = torch.diagonal(shifted, dim1=0, dim2=1).transpose(0, 1)    # [j, b] -> log|Psi_j(r~j)| shifted
    diag_sign = torch.diagonal(psi.sign, dim1=0, dim2=1).transpose(0, 1)
    ratio = psi.sign * diag_sign[None] * torch.exp(shifted - diag_log[None])
    return ratio, stats


def symmetrize_overlap_with_clipped_geometric_mean(x):
    """y_ij = sign(x_ij) sqrt(max(0, x_ij x_ji))  (reference :102-121)"""
    return torch.sign(x) * torch.sqrt(torch.clamp(x * x.transpose(-1, -2), min=0.0))


def compute_mean_overlap(psi_ratio, weight=None):
    """-> (overlap penalty = sum_{i<j} S_ij^2, {'overlap/pairwise/mean': S[n_wfs, n_wfs]}); the sample mean
    is the all-device mean (one all-reduce, reference parallel.py all_device_mean)."""
    w = torch.ones_like(psi_ratio[0]) if weight is None else weight
    local = (w[None] * psi_ratio).double()
    packed = torch.cat([local.sum(-1).reshape(-1), torch.tensor([float(local.shape[-1])], device=local.device,
                                                                dtype=torch.float64)])
    if parallel.world()[1] > 1:
        torch.distributed.all_reduce(packed, op=torch.distributed.ReduceOp.SUM)
    n = psi_ratio.shape[0]
    mean = (packed[:-1] / packed[-1]).reshape(n, n)
    symm = symmetrize_overlap_with_clipped_geometric_mean(mean)
    iu = torch.triu_indices(n, n, 1)
    return (symm[iu[0], iu[1]] ** 2).sum(), {'overlap/pairwise/mean': symm}


def compute_mean_overlap_tangent(psi_ratio, weight, ratio_gradient_mask, ansatz, params, phys_conf, scale=None,
                                 ordering=None):
    """Parameter gradient of the overlap penalty sum_{i<j} S_ij^2 (reference loss/overlap.py:182-229, Eq. 16 of
    Entwistle et al. 2023 restricted -- like the reference -- to the parameters of the SECOND state of every ordered
    pair).  For state j the reference contracts

        2 <ratio_ji> (ratio_ij[b] - <ratio_ij>) w_b mask_b / n_mask * scale_ij      summed over the pairs i < j

    with the parameter tangent of log|psi_j| on state j's own samples; here these factors are the cotangent of one
    reverse pass per state (dqmc_wf_vjp_params).  psi_ratio[i, j, b] as returned by compute_psi_ratio;
    `ordering`: permutation of the states that defines "i < j" (reference permute_matrix with data['ordering']).
    Returns a list with one {haiku name: gradient} dict per state (None for a state that is never the second one)."""
    n, _, B = psi_ratio.shape
    w = torch.ones(n, B, dtype=psi_ratio.dtype, device=psi_ratio.device) if weight is None else weight
    mask = torch.ones(n, n, B, dtype=torch.bool, device=psi_ratio.device) if ratio_gradient_mask is None else ratio_gradient_mask
    local = (w[None] * psi_ratio).double()
    packed = torch.cat([local.sum(-1).reshape(-1), mask.double().sum(-1).reshape(-1),
                        torch.tensor([float(B)], device=local.device, dtype=torch.float64)])
    if parallel.world()[1] > 1:
        torch.distributed.all_reduce(packed)
    mean_ratio = (packed[:n * n] / packed[-1]).reshape(n, n).to(psi_ratio.dtype)
    n_mask = packed[n * n:2 * n * n].reshape(n, n).to(psi_ratio.dtype)
    sc = torch.ones(n, n, dtype=psi_ratio.dtype, device=psi_ratio.device) if scale is None else scale
    order = list(range(n)) if ordering is None else [int(x) for x in ordering]
    pos = {s: k for k, s in enumerate(order)}  # position of every state in the ordering
    grads = []
    for j in range(n):
        cot = torch.zeros(B, dtype=psi_ratio.dtype, device=psi_ratio.device)
        used = False
        for i in range(n):
            if pos[i] < pos[j]:  # (i, j) is in the upper triangle of the permuted matrix
                cot = cot + (2 * mean_ratio[j, i] * (psi_ratio[i, j] - mean_ratio[i, j]) * w[j] * mask[i, j].to(cot.dtype)
                             / n_mask[i, j] * sc[i, j])
                used = True
        if not used:
            grads.append(None)
            continue
        R = phys_conf.R[0] if phys_conf.R.dim() == 4 else phys_conf.R
        R = R[0] if R.dim() == 3 else R
        pc_j = PhysicalConfiguration(R, phys_conf.r[j], torch.zeros(B, device=cot.device))
        _, g = ansatz.log_psi_vjp(params[j], pc_j, cot.contiguous())
        if parallel.world()[1] > 1:
            keys = sorted(g)
            flat = torch.cat([g[k].reshape(-1) for k in keys])
            torch.distributed.all_reduce(flat)
            o = 0
            for k in keys:
                m = g[k].numel()
                g[k] = flat[o:o + m].reshape(g[k].shape)
                o += m
        grads.append(g)
    return grads


def psi_ratio_clip_and_mask(psi_ratio, *, clip_width: float = 10.0, exclude_width: float = float('inf')):
    """Clip the wave-function ratios of ONE electron batch to `clip_width` median absolute deviations around the
    (all-device) median and flag outliers beyond `exclude_width` (reference: loss/clip.py:144-174)."""
    from . import parallel
    from .energy import all_walker_median

    allr = parallel.all_gather_walkers(psi_ratio.reshape(-1))
    center = all_walker_median(allr)
    sigma = all_walker_median((allr - center).abs())
    clipped = torch.clamp(psi_ratio, center - clip_width * sigma, center + clip_width * sigma)
    return clipped, (psi_ratio - center).abs() < exclude_width


def clip_psi_ratio(clip_mask_fn, psi_ratio):
    """Apply `clip_mask_fn` to every [molecule, state i, state j] electron batch of psi_ratio[..., B]
    (reference: loss/clip.py:51-70; vmap over the three leading axes)."""
    lead = psi_ratio.shape[:-1]
    flat = psi_ratio.reshape(-1, psi_ratio.shape[-1])
    out = [clip_mask_fn(row) for row in flat]
    return (torch.stack([o[0] for o in out]).reshape(*lead, -1), torch.stack([o[1] for o in out]).reshape(*lead, -1))

