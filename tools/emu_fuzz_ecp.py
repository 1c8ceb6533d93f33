"""Development tool (not part of the product): ECP fuzz: random clusters of C / Li (ccECP, bfd for C) and H, every ansatz kind, E_loc with injected quadrature twists vs oracle.

Runs against the CPU emulation of the kernels (build it with `python tools/emu_run_tests.py test_single_sample_signature`,
which leaves /tmp/libdqmc_emu.so; EMU_LIB selects another build; DQMC_EMU_REVERSE / DQMC_EMU_REVERSE_BLOCKS as there).
Usage: python tools/emu_fuzz_ecp.py SEED N_TRIALS
"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import deepqmc_b200.engine as E
orig = E.Engine.__init__
def init(self,*a,**k):
    k['_lib_path']=os.environ.get('EMU_LIB','/tmp/libdqmc_emu.so'); k['gemm_backend']=0; orig(self,*a,**k)
E.Engine.__init__=init
import test_gpu_parity as P
P.DEV='cpu'
from deepqmc_b200 import params as PN
from deepqmc_b200.ansatz import B200Ansatz
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.types import PhysicalConfiguration
from oracle.hamil import OracleHamiltonian
rng = np.random.default_rng(int(sys.argv[1]))
kinds = ['psiformer', 'ferminet', 'transpsiformer', 'paulinet', 'paulinet_default']
worst = 0
for trial in range(int(sys.argv[2])):
    M = int(rng.integers(1, 5))
    ecp = str(rng.choice(['ccECP', 'ccECP', 'bfd']))
    pool = [1, 6] if ecp == 'bfd' else [1, 3, 6]
    charges = rng.choice(pool, size=M).astype(float)
    if not (charges > 1).any(): charges[0] = 6.0
    val = sum({1: 1, 3: 1, 6: 4}[int(z)] for z in charges)
    if val < 2 or val > 10: continue
    spin = val % 2
    coords = rng.normal(size=(M, 3)) * 0.7 + np.arange(M)[:, None] * np.array([2.0, 0, 0])
    mol = Molecule(coords=coords, charges=charges, charge=0, spin=spin)
    kind = kinds[rng.integers(len(kinds))]
    hyper = dict(psiformer=dict(embedding_dim=16, n_layers=1, n_heads=2, n_determinants=2), transpsiformer=dict(embedding_dim=16, n_layers=1, n_heads=2, n_determinants=2),
                 ferminet=dict(embedding_dim=16, n_layers=2, n_determinants=2, edge_dim=4), paulinet={}, paulinet_default=dict(embedding_dim=16, n_determinants=2, edge_dim=8))[kind]
    tag = f'{ecp} Z={charges.astype(int).tolist()} {kind}'
    try:
        hamil = MolecularHamiltonian(mol=mol, ecp_type=ecp); oh = OracleHamiltonian(mol, ecp_type=ecp)
        N = hamil.n_up + hamil.n_down
        J = len(hamil.pot.nuc_with_nl_pot)
        a = B200Ansatz(hamil, kind, dtype='float64', **hyper)
        p = PN.perturb_params(a.init(trial))
        B = 2
        r = torch.as_tensor(mol.coords[rng.integers(0, M, size=(B, N))] + 0.8 * rng.normal(size=(B, N, 3)))
        R = torch.as_tensor(mol.coords)
        tw = torch.as_tensor(rng.uniform(0, np.pi / 5, size=(B, max(J, 1), N)))[:, :J]
        t = time.time()
        pc = PhysicalConfiguration(R, r, torch.zeros(B))
        El, st = hamil.local_energy(a.apply)(None, p, pc, ecp_twist=tw if J else None)
        ref = P.oracle_eval(a, oh, p, r, R, twist=tw if J else None)
        err = 0
        for b, (s, l, e, sto) in enumerate(ref):
            err = max(err, abs(El[b].item() - e) / max(1, abs(e), 0.5 * abs(sto['hamil/lap'])), abs(st['hamil/V_nl'][b].item() - sto['hamil/V_nl']) / max(1, abs(sto['hamil/V_nl'])),
                      abs(st['hamil/V_loc'][b].item() - sto['hamil/V_loc']) / max(1, abs(sto['hamil/V_loc'])))
        worst = max(worst, err)
        print(f'{tag} N={N} J={J}: err {err:.1e} ({time.time()-t:.1f}s)', flush=True)
        if err > 1e-8: print('   *** FAIL')
    except Exception as e:
        print(f'{tag}: EXC {type(e).__name__}: {str(e)[:200]}', flush=True)
print('worst', worst)
