"""Accuracy study (GPU): log|psi| of the fp32 engine's plain forward under the tensor-core path switches against the fp64
CUDA-core engine, on walkers of the full-size LiH / benzene Psiformer.  Prints the error distribution per configuration.
Usage: python tools/acc_study.py [LiH|benzene] [walkers]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepqmc_b200 import params as PN
from deepqmc_b200.ansatz import B200Ansatz
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule

name = sys.argv[1] if len(sys.argv) > 1 else 'benzene'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
mol = Molecule.from_name(name)
hamil = MolecularHamiltonian(mol=mol, ecp_type='ccECP' if name == 'benzene' else None)
N = hamil.n_up + hamil.n_down
rng = np.random.default_rng(5)
p = hamil.ns_valence / hamil.ns_valence.sum()
r = torch.as_tensor(mol.coords[rng.choice(len(mol.coords), size=(B, N), p=p)] + rng.normal(size=(B, N, 3)) * 0.7, device='cuda')
R = torch.as_tensor(mol.coords, device='cuda')
a64 = B200Ansatz(hamil, 'psiformer', dtype='float64')
params = PN.perturb_params(a64.init(0))
e64 = a64.engine_for(hamil, params)
s64, l64 = e64.wf_forward(r, R)
E64 = None
nE = 8
if name == 'benzene':
    E64 = e64.local_energy(r[:nE], R, seed=7)[0]


def run(tag, backend, env):
    for k in ('DQMC_TC_F16', 'DQMC_TC_FUSE_MLP', 'DQMC_ATTN_MMA'):
        os.environ.pop(k, None)
    os.environ.update(env)
    a = B200Ansatz(hamil, 'psiformer', dtype='float32', gemm_backend=backend)
    e = a.engine_for(hamil, params)
    s, l = e.wf_forward(r.float(), R.float())
    d = (l.double() - l64).abs()
    q = torch.quantile(d, torch.tensor([0.5, 0.9, 0.99], device=d.device, dtype=d.dtype))
    msg = f'{tag:34s} |dlog|: median {q[0].item():.2e} p90 {q[1].item():.2e} p99 {q[2].item():.2e} max {d.max().item():.2e} sign mismatches {(s.double() != s64).sum().item()}'
    if E64 is not None:
        E = e.local_energy(r[:nE].float(), R.float(), seed=7)[0]
        rel = ((E.double() - E64).abs() / E64.abs().clamp(min=1.0))
        msg += f' | E_loc rel err max {rel.max().item():.2e} median {rel.median().item():.2e}'
    print(msg, flush=True)
    e.close()


run('cuda-core fp32', 0, {'DQMC_ATTN_MMA': '0'})
run('cuda-core fp32 + attn mma', 0, {'DQMC_ATTN_MMA': '1'})
run('3xTF32', 1, {'DQMC_TC_F16': '0', 'DQMC_ATTN_MMA': '0'})
run('3xTF32 + attn mma', 1, {'DQMC_TC_F16': '0', 'DQMC_ATTN_MMA': '1'})
run('3xF16 gemm', 1, {'DQMC_TC_F16': '1', 'DQMC_TC_FUSE_MLP': '0', 'DQMC_ATTN_MMA': '0'})
run('3xF16 gemm + fused mlp', 1, {'DQMC_TC_F16': '1', 'DQMC_TC_FUSE_MLP': '1', 'DQMC_ATTN_MMA': '0'})
run('3xF16 gemm + fused mlp + attn mma', 1, {'DQMC_TC_F16': '1', 'DQMC_TC_FUSE_MLP': '1', 'DQMC_ATTN_MMA': '1'})
