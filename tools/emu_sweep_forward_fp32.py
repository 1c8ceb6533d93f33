"""Development tool (not part of the product): fp32 production fast paths of the plain forward (dh = 64 attention instances, two-phase / register / thread-per-determinant Slater kernels, register-tiled embedding) over every electron count 2..34 and many nuclei counts, against the fp64 oracle.

Runs against the CPU emulation of the kernels (/tmp/libdqmc_emu.so from tools/emu_run_tests.py, or EMU_LIB).
Usage: python tools/emu_sweep_forward_fp32.py p|t   (Psiformer | TransPsiformer)
"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepqmc_b200 import params as PN
from deepqmc_b200.engine import Engine
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.spec import psiformer_spec, transpsiformer_spec
from oracle import wf
lib=os.environ.get('EMU_LIB','/tmp/libdqmc_emu.so')
which = sys.argv[1]
rng = np.random.default_rng(7)
cases = []
for N in range(2, 35):
    cases.append(('atom', N, 1))
for n in (3, 5, 6, 9, 12, 14, 16, 17):
    cases.append(('chain', n, n))
worst = 0
for shape, N, M in cases:
    if shape == 'atom':
        mol = Molecule(coords=np.zeros((1, 3)), charges=np.array([float(N)]), charge=0, spin=N % 2)
    else:
        mol = Molecule(coords=np.stack([1.8 * np.arange(M), np.zeros(M), np.zeros(M)], 1), charges=np.ones(M), charge=0, spin=M % 2)
    hamil = MolecularHamiltonian(mol=mol)
    mk = psiformer_spec if which == 'p' else transpsiformer_spec
    if which == 't' and N + M > 60: continue
    spec = mk(hamil, embedding_dim=64, n_layers=2, n_heads=1, n_determinants=2)
    params = PN.perturb_params(PN.init_params(spec, N)); pt = wf.to_torch(params)
    B = 3
    R = torch.as_tensor(mol.coords)
    r = torch.as_tensor(mol.coords[rng.integers(0, M, size=(B, N))] + rng.normal(size=(B, N, 3)) * (1.0 if shape == 'chain' else 1.5))
    t = time.time()
    try:
        e32 = Engine(spec, hamil, dtype='float32', _lib_path=lib); e32.set_params(params)
        s32, l32 = e32.wf_forward(r.float(), R.float())
        e64 = Engine(spec, hamil, dtype='float64', _lib_path=lib); e64.set_params(params)
        s64, l64 = e64.wf_forward(r, R)
        err32 = err64 = 0
        for b in range(B):
            so, lo = wf.log_psi(spec, pt, r[b], R)
            err64 = max(err64, abs(l64[b].item() - lo.item()) / max(1, abs(lo.item())))
            err32 = max(err32, abs(l32[b].item() - lo.item()) / max(1, abs(lo.item())))
            assert s64[b].item() == so.item(), 'sign64'
        worst = max(worst, err32)
        flag = '' if (err32 < 2e-4 and err64 < 1e-10) else '   *** CHECK'
        print(f'{which} {shape} N={N} M={M}: fp64 {err64:.1e} fp32 {err32:.1e} ({time.time()-t:.1f}s){flag}', flush=True)
    except Exception as e:
        print(f'{which} {shape} N={N} M={M}: EXC {type(e).__name__}: {str(e)[:200]}', flush=True)
print('worst32', worst)
