"""Development tool (not part of the product): fp32 forward-Laplacian path (attn_fl_f32 specialisations, dh = 64 and 16) over electron counts: fp32 engine vs fp64 engine.

Runs against the CPU emulation of the kernels (/tmp/libdqmc_emu.so from tools/emu_run_tests.py, or EMU_LIB).
Usage: python tools/emu_sweep_eloc_fp32.py
"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepqmc_b200 import params as PN
from deepqmc_b200.engine import Engine
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.spec import psiformer_spec
lib=os.environ.get('EMU_LIB','/tmp/libdqmc_emu.so')
rng = np.random.default_rng(8)
worst = 0
for dh, H in ((64, 1), (16, 2), (32, 1)):
  for N in list(range(2, 19)) + [22, 26, 30]:
    if N > 18 and dh != 64: continue
    mol = Molecule(coords=np.zeros((1, 3)), charges=np.array([float(N)]), charge=0, spin=N % 2)
    hamil = MolecularHamiltonian(mol=mol)
    spec = psiformer_spec(hamil, embedding_dim=dh * H, n_layers=1, n_heads=H, n_determinants=2)
    params = PN.perturb_params(PN.init_params(spec, N))
    B = 2
    R = torch.as_tensor(mol.coords)
    r = torch.as_tensor(rng.normal(size=(B, N, 3)) * 1.5)
    t = time.time()
    try:
        e32 = Engine(spec, hamil, dtype='float32', _lib_path=lib); e32.set_params(params)
        E32, st32, s32, l32, g32 = e32.local_energy(r.float(), R.float(), want_grad=True)
        e64 = Engine(spec, hamil, dtype='float64', _lib_path=lib); e64.set_params(params)
        E64, st64, s64, l64, g64 = e64.local_energy(r, R, want_grad=True)
        scale = torch.maximum(torch.maximum(E64.abs(), 0.5 * (st64[4] if st64.shape[0] == 6 else st64[:, 4]).abs()), torch.ones_like(E64))
        err = ((E32.double() - E64).abs() / scale).max().item()
        errl = ((l32.double() - l64).abs() / l64.abs().clamp(min=1)).max().item()
        worst = max(worst, err)
        flag = '' if err < 2e-4 and errl < 2e-4 else '   *** CHECK'
        print(f'dh={dh} H={H} N={N}: E rel {err:.1e} log {errl:.1e} ({time.time()-t:.1f}s){flag}', flush=True)
    except Exception as e:
        print(f'dh={dh} H={H} N={N}: EXC {type(e).__name__}: {str(e)[:200]}', flush=True)
print('worst', worst)
