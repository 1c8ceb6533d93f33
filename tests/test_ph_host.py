"""Pseudo-Hamiltonian (reference ecp/pseudo_hamiltonian.py): host table loader and the oracle
restatement, CPU only.  The CUDA path is compared with this oracle in tests/test_gpu_z_next_rows.py."""
import numpy as np
import pytest
import torch

from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.ph import read_ph_xml
from ph_fixture import channels, write_synthetic_ph


def test_table_loader_matches_positional_restatement(tmp_path):
    """The product's tag-based reader and the oracle's positional reader (the reference's access pattern,
    pseudo_hamiltonian.py:49-68) agree on the same file; r V_loc = r V_s + Z_val, r V_L2 = (r V_d - r V_s) / 6."""
    from oracle.ph import parse_xml

    d = write_synthetic_ph(str(tmp_path))
    for sym, zval in (('P', 5.0), ('S', 6.0), ('Cl', 7.0)):
        loc, l2, nv, r_max = read_ph_xml(f'{d}/{sym}.cc.xml')
        loc_o, l2_o, nv_o = parse_xml(f'{d}/{sym}.cc.xml')
        _, ch = channels(sym)
        assert nv == nv_o == zval and r_max == 10.0 and len(loc) == 10001
        assert np.abs(loc - loc_o).max() < 1e-12 and np.abs(l2 - l2_o).max() < 1e-12
        assert np.abs(loc - (ch['s'] + zval)).max() < 1e-12 and np.abs(l2 - (ch['d'] - ch['s']) / 6).max() < 1e-12


def test_hamiltonian_attributes_with_pseudo_hamiltonian(tmp_path):
    d = write_synthetic_ph(str(tmp_path))
    mol = Molecule(coords=[[0, 0, 0], [2.4, 0, 0], [0, 2.6, 0.3]], charges=[17, 15, 1], charge=0, spin=1)
    h = MolecularHamiltonian(mol=mol, ecp_type='PH', ph_data_dir=d)
    assert h.ns_valence.tolist() == [7.0, 5.0, 1.0] and (h.n_up, h.n_down) == (7, 6)
    assert h.ph.tab_of_nuc.tolist() == [0, 1, -1] and h.ph.tables.shape == (2, 2, 10001)
    assert h.loc_params is None and h.nl_params is None and len(h.pot.nuc_with_nl_pot) == 0
    with pytest.raises(ValueError):
        MolecularHamiltonian(mol=Molecule(coords=[[0, 0, 0], [2, 0, 0]], charges=[6, 6], charge=0, spin=0), ecp_type='PH',
                             ph_data_dir=d)  # no pseudo-Hamiltonian for carbon (pseudo_hamiltonian.py:84-87)


def test_oracle_kinetic_term_reduces_to_laplacian_without_l2_channel(tmp_path):
    """With V_L2 = 0 the mass tensor is 1/2: the PH kinetic term must equal -1/2 (lap + |grad|^2) of the plain path, and
    its statistics are those of the coordinates v = sqrt(2) r."""
    from oracle.hamil import OracleHamiltonian
    from oracle.laplacian import laplacian_hessian

    d = write_synthetic_ph(str(tmp_path), l2_scale=0.0)
    mol = Molecule(coords=[[0, 0, 0], [2.4, 0, 0]], charges=[17, 1], charge=0, spin=0)
    oh = OracleHamiltonian(mol, ecp_type='PH', ph_dir=d)
    R = torch.as_tensor(mol.coords)
    r = torch.as_tensor(np.random.default_rng(0).normal(size=(8, 3)))
    c = torch.as_tensor(np.random.default_rng(1).normal(size=(8, 3)))
    log_psi = lambda x: -(x * x).sum() * 0.3 + torch.sin((x * c).sum()) + torch.log1p((x[0] - x[1]).norm())
    e_kin, lap, qf = oh.ph.kinetic_term(log_psi, r, R)
    lap_r, grad = laplacian_hessian(lambda x: log_psi(x.reshape(-1, 3)), r.reshape(-1))
    assert abs(e_kin.item() + 0.5 * (lap_r + (grad**2).sum()).item()) < 1e-10
    assert abs(lap.item() - 0.5 * lap_r.item()) < 1e-10 and abs(qf.item() - 0.5 * (grad**2).sum().item()) < 1e-10
