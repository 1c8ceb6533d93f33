"""Development tool (not part of the product): config-space fuzz: random chains / clusters with many nuclei, narrow networks, odd spin counts, charged systems, ECPs.

Runs against the CPU emulation of the kernels (build it with `python tools/emu_run_tests.py test_single_sample_signature`,
which leaves /tmp/libdqmc_emu.so; EMU_LIB selects another build; DQMC_EMU_REVERSE / DQMC_EMU_REVERSE_BLOCKS as there).
Usage: python tools/emu_fuzz_vjp.py SEED N_TRIALS
"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepqmc_b200 import params as PN
from deepqmc_b200.engine import Engine
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.spec import ferminet_spec, paulinet_default_spec, paulinet_spec, psiformer_spec, transpsiformer_spec
from oracle import wf
from oracle.hamil import OracleHamiltonian
lib=os.environ.get('EMU_LIB','/tmp/libdqmc_emu.so')
rng = np.random.default_rng(int(sys.argv[1]))
kinds = ['psiformer', 'ferminet', 'transpsiformer', 'paulinet', 'paulinet_default']
worst = 0; worst32 = 0
for trial in range(int(sys.argv[2])):
    M = int(rng.integers(1, 14))
    charges = rng.choice([1, 1, 1, 2, 3, 4], size=M).astype(float)
    while charges.sum() > 13: charges[rng.integers(M)] = 1
    if charges.sum() > 13: continue
    charge = int(rng.choice([0, 0, 1, -1])) if charges.sum() > 2 else 0
    N = int(charges.sum()) - charge
    if N < 2: continue
    spin = int(rng.choice([N % 2, N % 2 + 2])) if N >= 3 else N % 2
    coords = rng.normal(size=(M, 3)) * 1.5 + np.arange(M)[:, None] * np.array([1.5, 0, 0])
    mol = Molecule(coords=coords, charges=charges, charge=charge, spin=spin)
    kind = kinds[rng.integers(len(kinds))]
    hamil = MolecularHamiltonian(mol=mol); oh = OracleHamiltonian(mol)
    d = int(rng.choice([4, 8, 16, 32])); L = int(rng.integers(1, 3)); K = int(rng.integers(1, 4)); H = int(rng.choice([1, 2, 4]))
    de = int(rng.choice([2, 4, 8]))
    tag = f'M={M} N={N} up={hamil.n_up} dn={hamil.n_down} q={charge} {kind:16s} d={d} L={L} K={K} H={H} de={de}'
    try:
        if kind == 'psiformer': spec = psiformer_spec(hamil, embedding_dim=d, n_layers=L, n_heads=H, n_determinants=K)
        elif kind == 'transpsiformer': spec = transpsiformer_spec(hamil, embedding_dim=d, n_layers=L, n_heads=H, n_determinants=K)
        elif kind == 'ferminet': spec = ferminet_spec(hamil, embedding_dim=d, n_layers=L, n_determinants=K, edge_dim=de)
        elif kind == 'paulinet': spec = paulinet_spec(hamil, n_layers=L, n_determinants=K)
        else: spec = paulinet_default_spec(hamil, embedding_dim=max(d, 8), n_layers=L + 1, n_determinants=K, edge_dim=max(de, 4))
    except (AssertionError, ValueError) as e:
        print(tag, 'spec rejected:', str(e)[:80]); continue
    if kind == 'transpsiformer' and d < max(hamil.n_up, hamil.n_down): continue  # rank-deficient determinant by construction
    try:
        params = PN.perturb_params(PN.init_params(spec, trial))
        B = 3
        R = torch.as_tensor(mol.coords)
        r = torch.as_tensor(mol.coords[rng.integers(0, M, size=(B, N))] + rng.normal(size=(B, N, 3)))
        w = torch.as_tensor(rng.normal(size=B))
        t = time.time()
        eng = Engine(spec, hamil, dtype='float64', _lib_path=lib); eng.set_params(params)
        sign, log, grads = eng.vjp_params(r, R, w)
        pt = {k: torch.as_tensor(v).requires_grad_(True) for k, v in params.items()}
        tot = sum(w[b] * wf.log_psi(spec, pt, r[b], R)[1] for b in range(B)); tot.backward()
        assert set(grads) == set(pt), set(grads) ^ set(pt)
        err = max((grads[k].detach().cpu().double().reshape(pt[k].shape) - pt[k].grad).abs().max().item() / max(1e-9, pt[k].grad.abs().max().item()) for k in pt)
        worst = max(worst, err)
        print(f'{tag}: vjp rel err {err:.1e} ({time.time()-t:.1f}s)', flush=True)
        if err > 1e-7: print('   *** FAIL')
    except Exception as e:
        print(f'{tag}: EXC {type(e).__name__}: {str(e)[:200]}', flush=True)
print('worst', worst)
