"""Profiling driver (run under ncu on the GPU box): plain forwards of the benzene Psiformer on one non-local-ECP-sized batch
(86 400 walkers = 40 walkers x 2160 quadrature points, 2.6 M rows per dense layer), i.e. the launches that make up 96 % of the
headline step.  Usage: [ncu ...] python tools/prof_fwd.py [n_forwards] [walkers]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepqmc_b200 import params as PN
from deepqmc_b200.ansatz import B200Ansatz
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule

n_fwd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 86400
mol = Molecule.from_name('benzene')
hamil = MolecularHamiltonian(mol=mol, ecp_type='ccECP')
ansatz = B200Ansatz(hamil, 'psiformer', dtype='float32', gemm_backend=1)
params = PN.perturb_params(ansatz.init(0))
eng = ansatz.engine_for(hamil, params)
rng = np.random.default_rng(0)
N = hamil.n_up + hamil.n_down
p = hamil.ns_valence / hamil.ns_valence.sum()
r = torch.as_tensor(mol.coords[rng.choice(len(mol.coords), size=(B, N), p=p)] + rng.normal(size=(B, N, 3)) * 0.7,
                    dtype=torch.float32, device='cuda')
R = torch.as_tensor(mol.coords, dtype=torch.float32, device='cuda')
for i in range(n_fwd):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    s, l = eng.wf_forward(r, R)
    t1.record()
    torch.cuda.synchronize()
    print(f'forward {i}: {t0.elapsed_time(t1):.2f} ms for {B} walkers, launches so far {eng.launch_count}, mean log|psi| {l.mean().item():.4f}')
