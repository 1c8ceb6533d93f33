"""Restatement of the Metropolis walker update (TEST INFRASTRUCTURE).

reference: src/deepqmc/sampling/electron_samplers.py:102-163 (proposal, acceptance, age,
tau adaptation) and :333-357 (DecorrSampler).  Random numbers are INJECTED (normal[B,N,3],
uniform[B]) because JAX threefry streams cannot be reproduced (RNG).
"""
import torch


def metropolis_step(wf_batch, state, normal, uniform, target_acceptance=0.57, max_age=None):
    """state = dict(r[B,N,3], sign[B], log[B], age[B] int32, tau scalar)."""
    r_prop = state['r'] + state['tau'] * normal
    s_p, l_p = wf_batch(r_prop)
    log_prob = 2 * (l_p - state['log'])
    accepted = log_prob > torch.log(uniform)
    if max_age is not None:
        accepted = accepted | (state['age'] >= max_age)
    acceptance = accepted.to(torch.float64).sum() / accepted.shape[0]
    tau = state['tau']
    if target_acceptance is not None:
        tau = tau / (target_acceptance / torch.clamp(acceptance, min=0.05))
    age = torch.where(accepted, torch.zeros_like(state['age']), state['age'] + 1)
    new = {
        'r': torch.where(accepted[:, None, None], r_prop, state['r']),
        'sign': torch.where(accepted, s_p, state['sign']),
        'log': torch.where(accepted, l_p, state['log']),
        'age': age,
        'tau': tau,
    }
    return new, acceptance


def clean_force(force, r, R, charges, tau):
    """reference: src/deepqmc/sampling/sampling_utils.py:71-101 (diffs_to_nearest_nuc, crossover_parameter,
    clean_force).  force, r: [B, N, 3]; R: [M, 3]; charges: [M]."""
    d = r[:, :, None, :] - R[None, None]  # [B, N, M, 3]
    d2 = (d * d).sum(-1)
    idx = d2.argmin(-1)  # [B, N]
    z = torch.gather(d, 2, idx[..., None, None].expand(-1, -1, 1, 3))[:, :, 0]
    z2 = torch.gather(d2, 2, idx[..., None])[..., 0]
    ch = torch.as_tensor(charges, dtype=r.dtype)[idx]
    eps = torch.finfo(force.dtype).eps
    z_unit = z / torch.linalg.norm(z, dim=-1, keepdim=True)
    f_unit = force / torch.clamp(torch.linalg.norm(force, dim=-1, keepdim=True), min=eps)
    Z2z2 = ch**2 * z2
    a = (1 + (f_unit * z_unit).sum(-1)) / 2 + Z2z2 / (10 * (4 + Z2z2))
    av2tau = a * (force**2).sum(-1) * tau
    force = (2 / (torch.sqrt(1 + 2 * av2tau) + 1))[..., None] * force
    norm_factor = torch.clamp(torch.sqrt(z2) / (tau * torch.clamp(torch.linalg.norm(force, dim=-1), min=eps)), max=1.0)
    return force * norm_factor[..., None]


def langevin_step(wf_and_grad_batch, R, charges, state, normal, uniform, target_acceptance=0.57, max_age=None):
    """reference: electron_samplers.py:176-232 (LangevinSampler inside MetropolisSampler.sample).
    state = dict(r, sign, log, force, age, tau); wf_and_grad_batch(r) -> (sign, log, grad[B, N, 3])."""
    tau = state['tau']
    r_prop = state['r'] + tau * state['force'] + torch.sqrt(tau) * normal
    s_p, l_p, g_p = wf_and_grad_batch(r_prop)
    f_p = clean_force(g_p, r_prop, R, charges, tau)
    log_g = ((state['force'] + f_p) * ((state['r'] - r_prop) + tau / 2 * (state['force'] - f_p))).sum((1, 2))
    log_prob = log_g + 2 * (l_p - state['log'])
    accepted = log_prob > torch.log(uniform)
    if max_age is not None:
        accepted = accepted | (state['age'] >= max_age)
    acceptance = accepted.to(torch.float64).sum() / accepted.shape[0]
    new_tau = tau
    if target_acceptance is not None:
        new_tau = tau / (target_acceptance / torch.clamp(acceptance, min=0.05))
    sel = accepted[:, None, None]
    new = {
        'r': torch.where(sel, r_prop, state['r']), 'force': torch.where(sel, f_p, state['force']),
        'sign': torch.where(accepted, s_p, state['sign']), 'log': torch.where(accepted, l_p, state['log']),
        'age': torch.where(accepted, torch.zeros_like(state['age']), state['age'] + 1), 'tau': new_tau,
    }
    return new, acceptance


def spin_exchange_step(wf_batch, state, up_idx, down_idx, uniform, n_up):
    """reference: electron_samplers.py:235-330 OppositeSpinExchangeSampler exchange step: swap the positions of electron
    up_idx[b] (among the spin-up ones) and down_idx[b] (among the spin-down ones), accept with 2 dlog|psi| > log u; no max_age
    override, step size untouched."""
    r = state['r']
    b = torch.arange(len(r))
    r_prop = r.clone()
    r_prop[b, up_idx] = r[b, n_up + down_idx]
    r_prop[b, n_up + down_idx] = r[b, up_idx]
    s_p, l_p = wf_batch(r_prop)
    accepted = 2 * (l_p - state['log']) > torch.log(uniform)
    acceptance = accepted.to(torch.float64).sum() / accepted.shape[0]
    new = {
        'r': torch.where(accepted[:, None, None], r_prop, r),
        'sign': torch.where(accepted, s_p, state['sign']),
        'log': torch.where(accepted, l_p, state['log']),
        'age': torch.where(accepted, torch.zeros_like(state['age']), state['age'] + 1),
        'tau': state['tau'],
    }
    return new, acceptance

