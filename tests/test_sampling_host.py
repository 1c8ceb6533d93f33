"""Host-side sampler orchestration that needs no GPU: molecule index batches, electron warps and the
multi-geometry sampler's bookkeeping (reference: sampling/combined_samplers.py:17-55,93-214,
sampling/nuclei_samplers.py:16-36,175-213) with a stand-in electron sampler."""
import numpy as np
import torch

from deepqmc_b200.sampling import (IdleNucleiSampler, MoleculeIdxSampler, MultiNuclearGeometrySampler, nn_elec_warp)
from deepqmc_b200.types import PhysicalConfiguration


def test_molecule_idx_sampler_cycles_and_wraps():
    s = MoleculeIdxSampler(0, n_mols=5, batch_size=2)
    got = [s.sample().tolist() for _ in range(5)]
    assert got == [[0, 1], [2, 3], [4, 0], [1, 2], [3, 4]]  # wrap-around fills from the next pass (:37-46)
    s = MoleculeIdxSampler(3, n_mols=6, batch_size=3, shuffle='once')
    a, b, c, d = (s.sample().tolist() for _ in range(4))
    assert sorted(a + b) == list(range(6)) and (a, b) == (c, d)  # 'once': every pass uses the same permutation
    s = MoleculeIdxSampler(3, n_mols=6, batch_size=6, shuffle='always')
    assert sorted(s.sample().tolist()) == list(range(6))


def test_nn_elec_warp_moves_electrons_with_their_nearest_nucleus():
    R_old = torch.tensor([[0.0, 0, 0], [4.0, 0, 0]], dtype=torch.float64)
    dR = torch.tensor([[0.0, 0.5, 0], [1.0, 0, 0]], dtype=torch.float64)
    r = torch.tensor([[[0.3, 0.1, 0.0], [3.8, -0.2, 0.1], [1.9, 0.0, 0.0]]], dtype=torch.float64)
    st = nn_elec_warp(0, R_old + dR, dR, {'r': r.clone()})
    assert torch.allclose(st['r'][0, 0], r[0, 0] + dR[0]) and torch.allclose(st['r'][0, 1], r[0, 1] + dR[1])
    assert torch.allclose(st['r'][0, 2], r[0, 2] + dR[0])  # 1.9 is closer to the nucleus at 0 than to the one at 4


class _FakeElecSampler:
    """Two 'electronic states', walkers random-walk; records the geometry every call saw."""

    def __init__(self):
        self.seen, self.updates = [], 0

    def init(self, rng, params, n, R):
        g = torch.Generator().manual_seed(int(rng))
        return [{'r': torch.randn(n, 2, 3, generator=g, dtype=torch.float64)} for _ in range(2)]

    def update(self, state, params, R):
        self.updates += 1
        return state

    def sample(self, rng, state, params, R):
        self.seen.append(R.clone())
        state = [{'r': s['r'] + 0.01} for s in state]
        r = torch.stack([s['r'] for s in state])
        pc = PhysicalConfiguration(R, r, torch.zeros(r.shape[:2], dtype=torch.int32))
        return state, pc, {'sampling/acceptance': torch.tensor([0.5, 0.6])}


class _ShiftNuclei(IdleNucleiSampler):
    def sample(self, rng, state):
        dR = torch.full_like(state['R'], 0.1)
        return {'R': state['R'] + dR}, dR, {}


def test_multi_nuclear_geometry_sampler_layout_and_update_period():
    R = torch.tensor([[[0.0, 0, 0], [1.5, 0, 0]], [[0.0, 0, 0], [2.0, 0, 0]], [[0.0, 0, 0], [2.5, 0, 0]]], dtype=torch.float64)
    es = _FakeElecSampler()
    smp = MultiNuclearGeometrySampler(es, _ShiftNuclei(), nn_elec_warp, update_nuc_period=2, elec_equilibration_steps=1)
    state = smp.init(0, None, 4, R)
    assert len(state['elec']) == 3 and state['update_nuc_counter'].tolist() == [0, 0, 0]
    state, pc, stats = smp.sample(1, state, None, np.array([2, 0]))
    assert pc.r.shape == (2, 2, 4, 2, 3) and pc.R.shape == (2, 2, 4, 2, 3) and pc.batch_shape == (2, 2, 4)
    assert pc.mol_idx[:, 0, 0].tolist() == [2, 0] and torch.equal(pc.R[0, 1, 3], R[2]) and torch.equal(pc.R[1, 0, 0], R[0])
    assert stats['sampling/acceptance'].shape == (2, 2) and state['update_nuc_counter'].tolist() == [1, 0, 1]
    assert es.updates == 0
    # second visit of molecule 2: counter == period - 1 -> nuclei move, electrons are warped, psi refreshed, 1 equilibration sweep
    r_before = state['elec'][2][0]['r'].clone()
    state, pc, _ = smp.sample(2, state, None, [2])
    assert torch.allclose(state['nuc'][2]['R'], R[2] + 0.1) and state['update_nuc_counter'].tolist() == [1, 0, 0]
    assert es.updates == 1 and torch.allclose(pc.R[0, 0, 0], R[2] + 0.1)
    assert torch.allclose(state['elec'][2][0]['r'], r_before + 0.1 + 0.02)  # warp + equilibration sweep + sample sweep
    state = smp.update(state, None)
    assert es.updates == 4


def test_chain_and_combine_samplers_follow_the_reference_composition():
    """chain(DecorrSampler(length), MetropolisSampler(...)) (sampling_utils.py:31-69): the Decorr link only carries the
    number of sub-steps; the combined sampler is the last link."""
    from functools import partial

    from deepqmc_b200.sampling import DecorrSampler, LangevinSampler, MetropolisSampler, chain, combine_samplers

    class _Ansatz:
        def apply(self, *a):
            raise AssertionError('not evaluated here')

    hamil, wf = object(), _Ansatz().apply
    s = chain(DecorrSampler(length=30), MetropolisSampler(hamil, wf, tau=1.0, target_acceptance=0.57))
    assert isinstance(s, MetropolisSampler) and s.length == 30 and s.initial_tau == 1.0
    s = combine_samplers([DecorrSampler(length=10), partial(LangevinSampler, tau=0.1)], hamil, wf)
    assert isinstance(s, LangevinSampler) and s.length == 10 and s.hamil is hamil
    assert DecorrSampler(hamil, wf, length=5).length == 5  # the merged form used elsewhere in this repository


def test_equilibrate_stops_when_the_criterion_is_stationary():
    from deepqmc_b200.sampling import equilibrate

    class _Sampler:
        def __init__(self):
            self.n = 0

        def sample(self, rng, state, params, mol_idxs):
            self.n += 1
            return state, self.n, {'step': self.n}

    # criterion decays to a noisy plateau: 1/n + alternating ripple
    crit = lambda n: 1.0 / n + (0.01 if n % 2 else -0.01)
    smp = _Sampler()
    mols = MoleculeIdxSampler(0, 3, 1)
    out = list(equilibrate(0, None, mols, smp, {}, crit, range(1000), block_size=4, n_blocks=3))
    assert 12 <= len(out) < 1000 and out[0][0] == 0 and out[-1][3]['step'] == len(out)
    assert [o[2].tolist() for o in out[:4]] == [[0], [1], [2], [0]]
    smp2 = _Sampler()
    assert len(list(equilibrate(0, None, MoleculeIdxSampler(0, 3, 1), smp2, {}, crit, range(20), block_size=4, n_blocks=3,
                                allow_early_stopping=False))) == 20


def test_jax_compatible_initializer_gives_the_reference_walkers():
    """sampler.init(seed) with JaxCompatibleElectronInitializer == the reference's sampler.init(PRNGKey(seed)): the recorded
    walkers of tests/test_sampling/test_sampler_init_Metropolis_.npz (and of every n = 1 fixture) come out bit for bit."""
    import json
    import os

    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.sampling import JaxCompatibleElectronInitializer

    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_goldens.json')))
    mol = Molecule.from_name('LiH')
    r = JaxCompatibleElectronInitializer().walkers(0, 10, mol.charges, mol.charges, mol.coords, 2, 2)
    assert np.abs(r - np.asarray(g['sampling']['init_Metropolis']['r'])).max() < 1e-14
    assert np.abs(r[:5] - np.asarray(g['init_sample_Molecular']['rs'])).max() < 1e-14


def test_parameter_exchange_with_a_reference_run(tmp_path):
    """':'-flattened Haiku names <-> npz: nested Haiku tree in, flat tree out, shape / name checks."""
    import pytest

    from deepqmc_b200 import params as PN
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.spec import paulinet_spec, psiformer_spec

    hamil = MolecularHamiltonian(mol=Molecule.from_name('LiH'))
    spec = paulinet_spec(hamil)
    flat = PN.init_params(spec, 0)
    nested = PN.unflatten_haiku_tree(flat)
    assert all(isinstance(v, dict) for v in nested.values()) and PN.flatten_haiku_tree(nested).keys() == flat.keys()
    PN.save_params(str(tmp_path / 'p.npz'), nested)  # what np.savez(..., **flatten_pytree(params)) writes on the reference side
    back = PN.load_params(str(tmp_path / 'p.npz'), spec)
    assert back.keys() == flat.keys() and all(np.array_equal(back[k], flat[k]) for k in flat)
    with pytest.raises(ValueError):
        PN.load_params(str(tmp_path / 'p.npz'), psiformer_spec(hamil))  # another ansatz: names differ


def test_chain_with_spin_exchange_link():
    """conf/task/sampler_factory/elec_sampler/decorr_spin_exchange_metropolis.yaml: Decorr(33), spin exchange with probability
    0.1, Metropolis -- the exchange link only sets the probability on the combined sampler."""
    from functools import partial

    import pytest

    from deepqmc_b200.sampling import DecorrSampler, LangevinSampler, MetropolisSampler, OppositeSpinExchangeSampler, combine_samplers

    class _Ansatz:
        def apply(self, *a):
            raise AssertionError('not evaluated here')

    hamil, wf = object(), _Ansatz().apply
    s = combine_samplers([DecorrSampler(length=33), OppositeSpinExchangeSampler(exchange_step_probability=0.1),
                          partial(MetropolisSampler, tau=1.0, max_age=None)], hamil, wf)
    assert isinstance(s, MetropolisSampler) and s.length == 33 and s.exchange_step_probability == 0.1
    with pytest.raises(AssertionError):
        combine_samplers([OppositeSpinExchangeSampler(exchange_step_probability=0.1), partial(LangevinSampler, tau=0.1)], hamil, wf)
    with pytest.raises(NotImplementedError):
        OppositeSpinExchangeSampler(exchange_step_probability=0.1, up_logits_fn=lambda r: r)
