"""Development aid: one launch of the whole-trunk kernel with DQMC_TRUNK_TRACE=1 (clock64 stamps of one steady-state tile of
block 0, layer 1, printed by the library to stderr).  Usage: DQMC_TRUNK_TRACE=1 python tools/trunk_trace.py [walkers]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepqmc_b200 import params as PN
from deepqmc_b200.ansatz import B200Ansatz
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 148 * 4
hamil = MolecularHamiltonian(mol=Molecule.from_name('benzene'), ecp_type='ccECP')
a = B200Ansatz(hamil, 'psiformer', dtype='float32', gemm_backend=1)
params = PN.perturb_params(a.init(0))
eng = a.engine_for(hamil, params)
X0 = torch.randn(B * 30, 256, device='cuda')
for _ in range(2):
    out = eng.debug_trunk(X0)
torch.cuda.synchronize()
print('ok', float(out.abs().max()))
