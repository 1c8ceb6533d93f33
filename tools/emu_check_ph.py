"""Development tool: pseudo-Hamiltonian local energy on the CPU emulator (tools/cuda_emu) against the oracle.
    python tools/emu_check_ph.py [--kind psiformer|transpsiformer] [--dtype float64] [--nobuild]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--kind', default='psiformer')
    ap.add_argument('--dtype', default='float64')
    ap.add_argument('--charges', default='17,1')
    ap.add_argument('--B', type=int, default=3)
    ap.add_argument('--nobuild', action='store_true')
    a = ap.parse_args()
    from emu_check import build_emu
    lib = '/tmp/libdqmc_emu.so' if a.nobuild else build_emu()
    from ph_fixture import write_synthetic_ph
    from deepqmc_b200 import params as PN
    from deepqmc_b200.engine import Engine
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.spec import ferminet_spec, paulinet_default_spec, paulinet_spec, psiformer_spec, transpsiformer_spec
    from oracle import wf
    from oracle.hamil import OracleHamiltonian

    d = write_synthetic_ph('/tmp/ph_syn')
    ch = [int(c) for c in a.charges.split(',')]
    coords = [[0.0, 0.0, 0.0], [2.4, 0.0, 0.0], [0.0, 2.6, 0.3]][:len(ch)]
    mol = Molecule(coords=coords, charges=ch, charge=0, spin=0)
    hamil = MolecularHamiltonian(mol=mol, ecp_type='PH', ph_data_dir=d)
    oh = OracleHamiltonian(mol, ecp_type='PH', ph_dir=d)
    assert np.allclose(hamil.ns_valence, oh.ns_valence)
    if a.kind == 'paulinet':
        spec = paulinet_spec(oh, n_layers=2)
    elif a.kind == 'paulinet_default':
        spec = paulinet_default_spec(oh, n_layers=2, embedding_dim=16, n_determinants=3, edge_dim=8)
    else:
        mk = {'psiformer': psiformer_spec, 'transpsiformer': transpsiformer_spec, 'ferminet': ferminet_spec}[a.kind]
        spec = mk(oh, embedding_dim=16, n_layers=2, n_heads=2, n_determinants=3)
    params = PN.perturb_params(PN.init_params(spec, 0))
    pt = wf.to_torch(params)
    rng = np.random.default_rng(0)
    N = spec.n_elec
    R = torch.as_tensor(mol.coords)
    r = torch.as_tensor(mol.coords[rng.integers(0, len(mol.coords), size=(a.B, N))] + 0.8 * rng.normal(size=(a.B, N, 3)))
    eng = Engine(spec, hamil, dtype=a.dtype, _lib_path=lib)
    eng.set_params(params)
    E, stats, s2, l2, grad = eng.local_energy(r, R, want_grad=True)
    for b in range(a.B):
        f = lambda x: wf.log_psi(spec, pt, x, R)
        eo, st = oh.local_energy(f, r[b], R)
        x = r[b].clone().requires_grad_(True)
        g, = torch.autograd.grad(f(x)[1], x)
        print(f'b={b} dlog {abs(l2[b].item()-f(r[b])[1].item()):.2e} dE {abs(E[b].item()-eo.item()):.2e} (E={eo.item():.6f}) '
              + ' '.join(f"{k.split('/')[1]}:{abs(stats[i, b].item()-v.item()):.1e}({v.item():.3f})" for i, (k, v) in enumerate(st.items()))
              + f' dgrad {(grad[b].reshape(-1, 3) - g).abs().max().item():.1e}')


if __name__ == '__main__':
    main()
