"""The reference's parameter- and random-stream-dependent regression fixtures, reproduced WITHOUT JAX (CPU): oracle +
oracle/jaxrand.py (numpy restatement of jax.random / haiku initialisation).  Everything here is checked against numbers
the reference recorded itself (tests/golden/reference_goldens.json, extracted by tools/extract_reference_goldens.py):
the electron initialiser, the carbon ccECP potentials (which validates the restated ccECP table), and the Metropolis /
Decorr / Langevin samplers.  The wave-function fixtures (psi, gradient, Laplacian, E_loc) are in test_oracle_goldens.py.
Agreement beyond ~1e-6 is not expected where the wave function enters: the reference evaluates part of its first GNN
layer in float32."""
import json
import os

import numpy as np
import pytest
import torch

from deepqmc_b200.molecule import Molecule
from deepqmc_b200.spec import paulinet_spec
from deepqmc_b200 import jaxrand as PJ  # the product's JAX-compatible walker initialiser / quadrature twists
from oracle import jaxrand as J
from oracle import wf
from oracle.hamil import OracleHamiltonian
from oracle.sampling import clean_force, langevin_step, metropolis_step


@pytest.fixture(scope='module')
def g():
    return json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_goldens.json')))


@pytest.fixture(scope='module')
def lih():
    mol = Molecule.from_name('LiH')
    oh = OracleHamiltonian(mol)
    spec = paulinet_spec(oh)
    pt = wf.to_torch(J.haiku_init_conv_gnn_ansatz(spec, seed=0))
    R = torch.as_tensor(mol.coords)
    r0 = np.stack([PJ.atom_centered_initializer(k, mol.charges, mol.charges, mol.coords, 2, 2) for k in J.split(J.prng_key(0), 10)])
    return mol, spec, pt, R, r0


def test_electron_initializer_reproduces_reference_walkers(g):
    """AtomCenteredElectronInitializer(ShellBasedDistribution()) (sampling/electron_sample_initializers.py:43-288) driven by
    the jax.random streams: split, categorical tie-break, exponential radii, Haar-orthogonal directions."""
    mol = Molecule.from_name('LiH')
    r = PJ.atom_centered_initializer(J.prng_key(0), mol.charges, mol.charges, mol.coords, 2, 2)
    assert np.abs(r - np.asarray(g['edge_builder_LiH']['ne'])[0]).max() < 1e-14  # the walker of every n = 1 fixture
    rs = np.stack([PJ.atom_centered_initializer(k, mol.charges, mol.charges, mol.coords, 2, 2) for k in J.split(J.prng_key(0), 5)])
    assert np.abs(rs - np.asarray(g['init_sample_Molecular']['rs'])).max() < 1e-14


def test_carbon_ccecp_potentials_match_reference_fixture(g):
    """tests/test_potential.py for C: the ccECP and bfd tables restated in oracle/hamil.py (the reference reads them from
    pyscf) reproduce the recorded local potential, and -- with the Haiku-initialised ansatz and the fold_in-derived quadrature
    twists -- the non-local potential (the reference notes that term is 'not particularly numerically stable')."""
    mol = Molecule.from_name('C')
    oh0 = OracleHamiltonian(mol)
    r = PJ.atom_centered_initializer(J.prng_key(0), mol.charges, oh0.ns_valence, mol.coords, oh0.n_up, oh0.n_down)
    v = oh0.local_potential(torch.as_tensor(r), torch.as_tensor(mol.coords))
    assert abs(v.item() - g['potential_C_None']['local_potential']) < 1e-9 * abs(v.item())
    for ecp in ('ccECP', 'bfd'):
        oh = OracleHamiltonian(mol, ecp_type=ecp)
        assert (oh.n_up, oh.n_down) == (3, 1)
        r = torch.as_tensor(PJ.atom_centered_initializer(J.prng_key(0), mol.charges, oh.ns_valence, mol.coords, oh.n_up, oh.n_down))
        R = torch.as_tensor(mol.coords)
        assert abs(oh.local_potential(r, R).item() - g[f'potential_C_{ecp}']['local_potential']) < 1e-11 * 99.0
        spec = paulinet_spec(oh)
        pt = wf.to_torch(J.haiku_init_conv_gnn_ansatz(spec, seed=0))
        twists = torch.as_tensor(PJ.ecp_quadrature_twists(J.prng_key(0), 1, spec.n_elec))
        vnl = oh.nonloc_potential(r, R, lambda x: wf.log_psi(spec, pt, x, R), twists)
        assert abs(vnl.item() / g[f'potential_C_{ecp}']['nonlocal_potential'] - 1) < 1e-5


def _wf_batch(spec, pt, R):
    return lambda rr: tuple(torch.stack(x) for x in zip(*[wf.log_psi(spec, pt, rr[b], R) for b in range(len(rr))]))


@pytest.mark.parametrize('kind', ['Metropolis', 'DecorrMetropolis'])
def test_metropolis_sampler_reproduces_reference_fixture(g, lih, kind):
    """tests/test_sampling.py TestSampling: init(PRNGKey(0)) then sample(PRNGKey(step)) for step < 4; Decorr: 20 sub-steps per
    sample from split(rng, 20), max_age 20.  Proposal / acceptance noise = the reference's jax.random streams."""
    mol, spec, pt, R, r0 = lih
    wfb = _wf_batch(spec, pt, R)
    s0, l0 = wfb(torch.as_tensor(r0))
    gi = g['sampling']['init_Metropolis']
    assert np.abs(r0 - np.asarray(gi['r'])).max() < 1e-14 and np.abs(l0.numpy() - np.asarray(gi['psi:log'])).max() < 2e-5
    st = dict(r=torch.as_tensor(r0), sign=s0, log=l0, age=torch.zeros(10, dtype=torch.int32), tau=torch.tensor(0.1, dtype=torch.float64))
    decorr = kind == 'DecorrMetropolis'
    for step in range(4):
        key = J.prng_key(step)
        for k in (J.split(key, 20) if decorr else [key]):
            kp, ka = J.split(k, 2)
            st, acc = metropolis_step(wfb, st, torch.as_tensor(J.normal(kp, (10, 4, 3))), torch.as_tensor(J.uniform(ka, (10,))),
                                      0.57, 20 if decorr else None)
    gs = g['sampling'][f'sample_{kind}']
    assert np.abs(st['r'].numpy() - np.asarray(gs['smpl_state:r'])).max() < 1e-12  # the same accept / reject history
    assert st['age'].tolist() == gs['smpl_state:age'] and abs(st['tau'].item() - gs['smpl_state:tau']) < 1e-12
    assert np.abs(st['log'].numpy() - np.asarray(gs['smpl_state:psi:log'])).max() < 2e-5
    assert acc.item() == gs['stats:sampling/acceptance']
    assert abs(st['log'].mean().item() - gs['stats:sampling/log_psi/mean']) < 2e-5
    assert abs(st['log'].std(unbiased=False).item() - gs['stats:sampling/log_psi/std']) < 2e-5
    i, j = torch.triu_indices(4, 4, 1)
    dm = torch.sqrt(torch.finfo(torch.float64).eps + ((st['r'][:, i] - st['r'][:, j]) ** 2).sum(-1)).mean().item()
    assert abs(dm - gs['stats:sampling/dists/mean']) < 1e-9


def test_langevin_sampler_reproduces_reference_fixture(g, lih):
    mol, spec, pt, R, r0 = lih

    def wfg(rr):
        out = []
        for b in range(len(rr)):
            x = rr[b].clone().requires_grad_(True)
            s, l = wf.log_psi(spec, pt, x, R)
            out.append((s.detach(), l.detach(), torch.autograd.grad(l, x)[0]))
        return tuple(torch.stack(x) for x in zip(*out))

    tau0 = torch.tensor(0.1, dtype=torch.float64)
    r = torch.as_tensor(r0)
    s, l, gr = wfg(r)
    st = dict(r=r, sign=s, log=l, force=clean_force(gr, r, R, mol.charges, tau0), age=torch.zeros(10, dtype=torch.int32), tau=tau0)
    assert np.abs(st['force'].numpy() - np.asarray(g['sampling']['init_Langevin']['force'])).max() < 5e-5
    for step in range(4):
        kp, ka = J.split(J.prng_key(step), 2)
        st, acc = langevin_step(wfg, R, mol.charges, st, torch.as_tensor(J.normal(kp, (10, 4, 3))), torch.as_tensor(J.uniform(ka, (10,))), 0.57, None)
    gs = g['sampling']['sample_Langevin']
    assert np.abs(st['r'].numpy() - np.asarray(gs['smpl_state:r'])).max() < 5e-6  # the float32-tainted force moves the walkers
    assert np.abs(st['force'].numpy() - np.asarray(gs['smpl_state:force'])).max() < 5e-5
    assert st['age'].tolist() == gs['smpl_state:age'] and abs(st['tau'].item() - gs['smpl_state:tau']) < 1e-12
    assert acc.item() == gs['stats:sampling/acceptance']


def test_multi_nuclear_geometry_sampler_reproduces_reference_fixture(g, lih):
    """TestMultimoleculeSampling: two copies of LiH; init keys = split(PRNGKey(0), 2), every molecule's walkers from
    split(key_m, 10); sample keys = split(PRNGKey(step), (2, 2)) -- electron keys first, nuclear keys second
    (sampling/combined_samplers.py:128-141,173-199)."""
    mol, spec, pt, R, _ = lih
    wfb = _wf_batch(spec, pt, R)
    states = []
    for m, km in enumerate(J.split(J.prng_key(0), 2)):
        r0 = np.stack([PJ.atom_centered_initializer(k, mol.charges, mol.charges, mol.coords, 2, 2) for k in J.split(km, 10)])
        assert np.abs(r0 - np.asarray(g['sampling_multi']['init']['elec:r'])[m]).max() < 1e-14
        s, l = wfb(torch.as_tensor(r0))
        states.append(dict(r=torch.as_tensor(r0), sign=s, log=l, age=torch.zeros(10, dtype=torch.int32), tau=torch.tensor(0.1, dtype=torch.float64)))
    for step in range(4):
        ks = J.split(J.prng_key(step), 4)
        for m in range(2):
            kp, ka = J.split(ks[m], 2)
            states[m], _ = metropolis_step(wfb, states[m], torch.as_tensor(J.normal(kp, (10, 4, 3))), torch.as_tensor(J.uniform(ka, (10,))), 0.57, None)
    gs = g['sampling_multi']['sample']
    for m in range(2):
        assert np.abs(states[m]['r'].numpy() - np.asarray(gs['smpl_state:elec:r'])[m]).max() < 1e-12
        assert states[m]['age'].tolist() == gs['smpl_state:elec:age'][m] and abs(states[m]['tau'].item() - gs['smpl_state:elec:tau'][m]) < 1e-12


def test_lih_ccecp_fixtures(g):
    """LiH with the ccECP on lithium (tests/test_hamil.py 'Molecular+PP', tests/test_potential.py LiH / ccECP): valence
    charges and mask, the five recorded walkers, local and non-local potential and the local energy.  (The last digits of
    one lithium coefficient were fixed with the V_loc fixture, see oracle/hamil.py; everything else is independent of it.)"""
    mol = Molecule.from_name('LiH')
    oh = OracleHamiltonian(mol, ecp_type='ccECP')
    gi = g['hamil_init']['Molecular_PP']
    assert (oh.n_up, oh.n_down) == (gi['n_up'], gi['n_down']) and oh.ns_valence.tolist() == gi['ns_valence'] and oh.ecp_mask.tolist() == gi['pp_mask']
    rs = np.stack([PJ.atom_centered_initializer(k, mol.charges, oh.ns_valence, mol.coords, 1, 1) for k in J.split(J.prng_key(0), 5)])
    assert np.abs(rs - np.asarray(g['init_sample_Molecular_PP']['rs'])).max() < 1e-14
    r = torch.as_tensor(PJ.atom_centered_initializer(J.prng_key(0), mol.charges, oh.ns_valence, mol.coords, 1, 1))
    R = torch.as_tensor(mol.coords)
    assert abs(oh.local_potential(r, R).item() - g['potential_LiH_ccECP']['local_potential']) < 1e-9
    spec = paulinet_spec(oh)
    pt = wf.to_torch(J.haiku_init_conv_gnn_ansatz(spec, seed=0))
    f = lambda x: wf.log_psi(spec, pt, x, R)
    tw = torch.as_tensor(PJ.ecp_quadrature_twists(J.prng_key(0), 1, spec.n_elec))
    assert abs(oh.nonloc_potential(r, R, f, tw).item() / g['potential_LiH_ccECP']['nonlocal_potential'] - 1) < 1e-6
    e_loc, _ = oh.local_energy(f, r, R, phi_random=tw)
    assert abs(e_loc.item() - g['local_energy_Molecular_PP']['E_loc']) < 2e-6
