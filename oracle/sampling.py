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
