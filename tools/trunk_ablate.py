"""Development aid: whole-trunk kernel timed with parts switched off (DQMC_TRUNK_ABLATE bits: 1 no weight loads, 2 no MMAs,
4 no Q/K/V image stores, 8 no image loads; results are garbage, only the time is meaningful).
Usage: python tools/trunk_ablate.py [walkers]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepqmc_b200 import params as PN
from deepqmc_b200.ansatz import B200Ansatz
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 148 * 60
hamil = MolecularHamiltonian(mol=Molecule.from_name('benzene'), ecp_type='ccECP')
a = B200Ansatz(hamil, 'psiformer', dtype='float32', gemm_backend=1)
params = PN.perturb_params(a.init(0))
eng = a.engine_for(hamil, params)
X0 = torch.randn(B * 30, 256, device='cuda')
tiles = B / 4
for ab in [0, 1, 2, 4, 8, 3, 12, 13, 14, 15, 0]:
    os.environ['DQMC_TRUNK_ABLATE'] = str(ab)
    eng.debug_trunk(X0)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        eng.debug_trunk(X0)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    print(f'ablate {ab:2d}: {ms:8.3f} ms  = {ms * 1e-3 * 1.92e9 / (tiles / 148 * 4):9.0f} clk per tile-layer (at 1.92 GHz)')
