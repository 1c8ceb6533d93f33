"""CPU: the workspace planner of libdqmc_b200.so, through the C ABI on PLAN-ONLY handles (dqmc_create with device = -1).

Round 1 shipped a plan (`per_walker_elems`) that forgot one buffer `carve()` took, with an emulator-only slack that hid it; the
hardware runs died with "carved workspace exceeds the planned size".  The plan is now a dry pass of the carving code itself
(`dqmc_workspace_bytes`), and `dqmc_debug_plan` walks the real entry point (forward / local energy / reverse pass / Metropolis /
Langevin) with every CUDA call skipped and reports the highest workspace offset it takes.  This file sweeps ansatz kind x mode x
batch size x dtype x Hamiltonian and asserts carved <= given for the planned size, for capped workspaces (the engine then
chunks) and for the documented minimum.
"""
import numpy as np
import pytest

from deepqmc_b200.engine import (MODE_FORWARD, MODE_LANGEVIN, MODE_LOCAL_ENERGY, MODE_MCMC, MODE_VJP, Engine)
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.spec import ferminet_spec, paulinet_default_spec, paulinet_spec, psiformer_spec, transpsiformer_spec

MODES = {'forward': MODE_FORWARD, 'local_energy': MODE_LOCAL_ENERGY, 'vjp': MODE_VJP, 'mcmc': MODE_MCMC, 'langevin': MODE_LANGEVIN}
BATCHES = [1, 257, 4096, 69120]


def _specs(h, small):
    kw = dict(embedding_dim=32, n_layers=2, n_heads=2, n_determinants=3) if small else {}
    yield 'psiformer', psiformer_spec(h, **kw)
    yield 'transpsiformer', transpsiformer_spec(h, **kw)
    yield 'ferminet', ferminet_spec(h, **(dict(kw, edge_dim=8) if small else {}))
    if small:
        yield 'ferminet_narrow', ferminet_spec(h, embedding_dim=4, n_layers=2, n_determinants=2, edge_dim=2)
    yield 'paulinet', paulinet_spec(h)
    yield 'paulinet_default', paulinet_default_spec(h)
    yield 'psiformer_add', psiformer_spec(h, backflow_transform='both', **kw)


def _engine(spec, h, dtype, built_lib):
    return Engine(spec, h, dtype=dtype, plan_only=True, gemm_backend=1 if dtype == 'float32' and spec.embedding_dim % 32 == 0 else 0)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('mol,ecp', [('LiH', None), ('LiH', 'ccECP'), ('benzene', 'ccECP'), ('N2', None)])
def test_carved_never_exceeds_planned(built_lib, mol, ecp, dtype):
    h = MolecularHamiltonian(mol=Molecule.from_name(mol), ecp_type=ecp)
    n_checked = 0
    for name, spec in _specs(h, small=mol != 'benzene'):
        if mol == 'benzene' and name not in ('psiformer', 'transpsiformer', 'ferminet'):
            continue
        eng = _engine(spec, h, dtype, built_lib)
        for mname, mode in MODES.items():
            if mode == MODE_VJP and spec.backflow_transform != 'mult':
                continue  # no reverse pass for the additive branch (status 2)
            for B in BATCHES:
                planned, carved = eng.debug_plan(B, mode)
                assert planned == eng.workspace_bytes(B, mode)
                assert 0 < carved <= planned, (name, mname, B, dtype, planned, carved)
                floor = eng.workspace_bytes_min(B, mode)
                assert floor <= planned
                # capped workspaces (the engine chunks the walkers), down to the documented minimum
                for cap in {max(floor, planned // 3), max(floor, planned // 50), floor}:
                    _, c2 = eng.debug_plan(B, mode, cap)
                    assert 0 < c2 <= cap, (name, mname, B, dtype, cap, c2)
                n_checked += 1
        eng.close()
    assert n_checked >= 3 * 4 * len(BATCHES)


def test_too_small_workspace_is_refused(built_lib):
    h = MolecularHamiltonian(mol=Molecule.from_name('LiH'), ecp_type='ccECP')
    eng = _engine(psiformer_spec(h, embedding_dim=32, n_layers=2, n_heads=2, n_determinants=3), h, 'float32', built_lib)
    for mode in MODES.values():
        floor = eng.workspace_bytes_min(64, mode)
        with pytest.raises(RuntimeError, match='workspace'):
            eng.debug_plan(64, mode, floor // 2)
    eng.close()


def test_plan_only_engine_has_no_compute_entry_points(built_lib):
    import ctypes as C

    h = MolecularHamiltonian(mol=Molecule.from_name('LiH'))
    eng = _engine(psiformer_spec(h, embedding_dim=32, n_layers=1, n_heads=2, n_determinants=2), h, 'float64', built_lib)
    flat = np.zeros(eng.n_packed)
    rc = eng.lib.dqmc_set_params(eng.h, flat.ctypes.data_as(C.POINTER(C.c_double)), eng.n_packed, None)
    assert rc == 2 and b'plan-only' in eng.lib.dqmc_last_error(eng.h)
    rc = eng.lib.dqmc_wf_forward(eng.h, None, None, 0, 4, None, None, None, 0, None)
    assert rc == 2
    eng.close()


def test_pseudo_hamiltonian_and_headline_sizes(built_lib):
    """The headline configuration's numbers: the plain-forward chunk of the benzene ECP pass (86 400 virtual walkers per group)
    and a pseudo-Hamiltonian engine (extra per-electron metric records in the forward-Laplacian chunk)."""
    import ctypes as C

    h = MolecularHamiltonian(mol=Molecule.from_name('benzene'), ecp_type='ccECP')
    eng = _engine(psiformer_spec(h), h, 'float32', built_lib)
    planned, carved = eng.debug_plan(4096, MODE_LOCAL_ENERGY)
    assert carved <= planned
    for cap in (8 << 30, 60 << 30, 100 << 30):
        _, c = eng.debug_plan(4096, MODE_LOCAL_ENERGY, cap)
        assert c <= cap
    eng.close()
    h2 = MolecularHamiltonian(mol=Molecule.from_name('LiH'))
    eng = _engine(psiformer_spec(h2, embedding_dim=32, n_layers=2, n_heads=2, n_determinants=3), h2, 'float64', built_lib)
    before = eng.workspace_bytes(100, MODE_LOCAL_ENERGY)
    tabs = np.zeros((1, 2, 8))
    ton = np.zeros(2, dtype=np.int32)
    rc = eng.lib.dqmc_set_pseudo_hamiltonian(eng.h, 1, 8, 4.0, tabs.ctypes.data_as(C.POINTER(C.c_double)),
                                             ton.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    after, carved = eng.debug_plan(100, MODE_LOCAL_ENERGY)
    assert after > before and carved <= after
    eng.close()
