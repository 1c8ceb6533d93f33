"""Development tool (not part of the product): sampler fuzz: Metropolis sweeps with injected noise on random clusters / walker counts / kinds vs the oracle step.

Runs against the CPU emulation of the kernels (build it with `python tools/emu_run_tests.py test_single_sample_signature`,
which leaves /tmp/libdqmc_emu.so; EMU_LIB selects another build; DQMC_EMU_REVERSE / DQMC_EMU_REVERSE_BLOCKS as there).
Usage: python tools/emu_fuzz_sampler.py SEED N_TRIALS
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
from deepqmc_b200 import params as PN
from deepqmc_b200.ansatz import B200Ansatz
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from oracle import wf
from oracle.sampling import metropolis_step
rng = np.random.default_rng(int(sys.argv[1]))
kinds = ['psiformer', 'ferminet', 'transpsiformer', 'paulinet', 'paulinet_default']
for trial in range(int(sys.argv[2])):
    M = int(rng.integers(1, 8))
    charges = rng.choice([1, 1, 2, 3], size=M).astype(float)
    N = int(charges.sum())
    if N < 2 or N > 10: continue
    coords = rng.normal(size=(M, 3)) + np.arange(M)[:, None] * np.array([1.6, 0, 0])
    mol = Molecule(coords=coords, charges=charges, charge=0, spin=N % 2)
    kind = kinds[rng.integers(len(kinds))]
    hyper = dict(psiformer=dict(embedding_dim=16, n_layers=1, n_heads=2, n_determinants=2), transpsiformer=dict(embedding_dim=16, n_layers=1, n_heads=2, n_determinants=2),
                 ferminet=dict(embedding_dim=8, n_layers=2, n_determinants=2, edge_dim=4), paulinet={}, paulinet_default=dict(embedding_dim=16, n_determinants=2, edge_dim=8))[kind]
    B = int(rng.choice([1, 3, 7, 130, 257])); nsub = int(rng.integers(1, 4)); max_age = int(rng.choice([-1, 0, 2]))
    if B > 100 and N > 6: B = 33
    tag = f'M={M} N={N} {kind} B={B} nsub={nsub} max_age={max_age}'
    try:
        hamil = MolecularHamiltonian(mol=mol)
        a = B200Ansatz(hamil, kind, dtype='float64', **hyper)
        p = PN.perturb_params(a.init(trial)); pt = wf.to_torch(p)
        eng = a.engine_for(hamil, p)
        r = torch.as_tensor(mol.coords[rng.integers(0, M, size=(B, N))] + rng.normal(size=(B, N, 3)))
        R = torch.as_tensor(mol.coords)
        sign, log = eng.wf_forward(r, R)
        nn = torch.as_tensor(rng.normal(size=(nsub, B, N, 3))); nu = torch.as_tensor(rng.uniform(size=(nsub, B)))
        state = dict(r=r.clone(), sign=sign.clone(), log=log.clone(), age=torch.zeros(B, dtype=torch.int32), tau=torch.tensor([0.4], dtype=torch.float64))
        t = time.time()
        stats = eng.mcmc_sweep(state, R, nsub, target_acceptance=0.57, max_age=None if max_age < 0 else max_age, noise_normal=nn, noise_uniform=nu)
        wfb = lambda rr: tuple(torch.stack(x) for x in zip(*[wf.log_psi(a.spec, pt, rr[b], R) for b in range(B)]))
        ost = dict(r=r.clone(), sign=sign.clone(), log=log.clone(), age=torch.zeros(B, dtype=torch.int32), tau=torch.tensor(0.4, dtype=torch.float64))
        for s in range(nsub):
            ost, acc = metropolis_step(wfb, ost, nn[s], nu[s], 0.57, None if max_age < 0 else max_age)
        ok = (torch.allclose(state['r'], ost['r'], atol=1e-12) and torch.allclose(state['log'], ost['log'], atol=1e-9) and torch.equal(state['age'], ost['age'])
              and abs(state['tau'].item() - ost['tau'].item()) < 1e-12 and abs(stats[0].item() - acc.item()) < 1e-12)
        print(f'{tag}: {"ok" if ok else "*** FAIL"} acc={acc.item():.2f} ({time.time()-t:.1f}s)', flush=True)
    except Exception as e:
        print(f'{tag}: EXC {type(e).__name__}: {str(e)[:200]}', flush=True)
