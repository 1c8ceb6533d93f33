"""Fixtures recorded with the reference itself by tools/export_reference_fixture.py (on a machine that has JAX) for the ansatz
kinds the reference's own tests do not cover.  Every tests/golden/external/*.npz is checked against the CPU oracle here (and
against the CUDA engine in test_gpu_z_next_rows.py::test_engine_external_fixtures).  The directory is empty in this repository
-- the build container has no JAX -- so these tests skip; they are the hook for pinning the Psiformer / FermiNet /
TransPsiformer trunks on a box with the reference's dependencies."""
import glob
import os

import numpy as np
import pytest
import torch

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'external', '*.npz')))
KIND = {'psiformer': 'psiformer', 'ferminet': 'ferminet', 'transpsiformer': 'transpsiformer', 'default': 'paulinet_default'}


def load_external(path):
    d = np.load(path)
    params = {k[len('param/'):]: np.asarray(d[k], dtype=np.float64) for k in d.files if k.startswith('param/')}
    hyper = {}
    for o in [str(x) for x in d['overrides']]:  # e.g. omni_factory.embedding_dim=32 n_determinants=4
        key, val = o.split('=')
        key = key.split('.')[-1]
        hyper[{'n_interactions': 'n_layers', 'num_heads': 'n_heads', 'two_particle_stream_dim': 'edge_dim'}.get(key, key)] = int(val)
    return d, params, KIND[str(d['ansatz'])], hyper


@pytest.mark.skipif(not FILES, reason='no externally recorded reference fixtures (tests/golden/external is empty)')
@pytest.mark.parametrize('path', FILES or [None])
def test_oracle_against_externally_recorded_fixture(path):
    from deepqmc_b200 import spec as S
    from deepqmc_b200.molecule import Molecule
    from oracle import wf
    from oracle.hamil import OracleHamiltonian
    from oracle.laplacian import laplacian_hessian

    d, params, kind, hyper = load_external(path)
    mol = Molecule.from_name(str(d['molecule']))
    oh = OracleHamiltonian(mol)
    spec = getattr(S, kind + '_spec')(oh, **hyper)
    from deepqmc_b200 import params as PN

    assert {k: tuple(v.shape) for k, v in params.items()} == {k: tuple(v) for k, v in PN.param_shapes(spec).items()}
    pt, r, R = wf.to_torch(params), torch.as_tensor(d['r']), torch.as_tensor(d['R'])
    f = lambda x: wf.log_psi(spec, pt, x, R)
    s, l = f(r)
    assert s.item() == float(d['sign']) and abs(l.item() - float(d['log'])) < 1e-6
    lap, grad = laplacian_hessian(lambda x: f(x.reshape(-1, 3))[1], r.reshape(-1))
    assert abs(lap.item() - float(d['lap'])) < 1e-5 * max(1.0, abs(float(d['lap'])))
    assert np.allclose(grad.numpy(), d['grad'], rtol=1e-5, atol=1e-6)
    e, _ = oh.local_energy(f, r, R)
    assert abs(e.item() - float(d['e_loc'])) < 1e-5 * max(1.0, abs(float(d['e_loc'])))
