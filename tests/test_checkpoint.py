"""Checkpoint layout of the reference (log.py:45-126, types.py:99-104) read / written without JAX.  There is no reference
checkpoint and no JAX in this image: the reader is exercised on a pickle produced by stand-ins that pickle the way JAX arrays
and the reference's NamedTuples do (GLOBAL jax._src.array._reconstruct_array around the numpy __reduce__; GLOBAL
deepqmc.types.TrainState / Psi), the writer on a plain pickle.load with those stand-ins installed -- the way
CheckpointStore.load would see it."""
import pickle
import sys
import types
from typing import NamedTuple

import numpy as np
import torch

from deepqmc_b200 import checkpoint as ck
from deepqmc_b200 import params as PN
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.spec import psiformer_spec


def _install_standins():
    jax, src, arr = types.ModuleType('jax'), types.ModuleType('jax._src'), types.ModuleType('jax._src.array')

    def _reconstruct_array(fun, args, arr_state, aval_state):  # what jax does, minus the device_put
        v = fun(*args)
        v.__setstate__(arr_state)
        return FakeArray(v)

    class FakeArray:
        def __init__(self, value):
            self._value = np.asarray(value)

        def __reduce__(self):  # jax._src.array.ArrayImpl.__reduce__
            fun, args, arr_state = self._value.__reduce__()
            return (_reconstruct_array, (fun, args, arr_state, {'weak_type': False}))

    _reconstruct_array.__module__ = 'jax._src.array'
    _reconstruct_array.__qualname__ = '_reconstruct_array'
    arr._reconstruct_array, arr.FakeArray = _reconstruct_array, FakeArray
    jax._src, src.array = src, arr
    dq, dqt = types.ModuleType('deepqmc'), types.ModuleType('deepqmc.types')
    TS = NamedTuple('TrainState', [('sampler', object), ('params', object), ('opt', object)])
    PS = NamedTuple('Psi', [('sign', object), ('log', object)])
    TS.__module__ = PS.__module__ = 'deepqmc.types'
    dqt.TrainState, dqt.Psi, dq.types = TS, PS, dqt
    mods = {'jax': jax, 'jax._src': src, 'jax._src.array': arr, 'deepqmc': dq, 'deepqmc.types': dqt}
    sys.modules.update(mods)
    return FakeArray, TS, PS, list(mods)


def _problem():
    hamil = MolecularHamiltonian(mol=Molecule.from_name('LiH'))
    spec = psiformer_spec(hamil, embedding_dim=16, n_layers=1, n_heads=2, n_determinants=2)
    per_state = [PN.init_params(spec, s) for s in range(2)]
    rng = np.random.default_rng(0)
    sampler = {'elec': {'r': rng.normal(size=(1, 2, 5, 4, 3)), 'age': rng.integers(0, 3, size=(1, 2, 5)).astype(np.int32),
                        'tau': np.array([[0.3, 0.4]]), 'psi': (np.ones((1, 2, 5)), rng.normal(size=(1, 2, 5)))},
               'nuc': {'R': hamil.mol.coords[None]}, 'update_nuc_counter': np.zeros(1)}
    return spec, per_state, sampler


def test_reader_on_a_jax_style_pickle(tmp_path):
    spec, per_state, sampler = _problem()
    FakeArray, TS, PS, names = _install_standins()
    try:
        wrap = lambda t: {k: (wrap(v) if isinstance(v, dict) else FakeArray(v)) for k, v in t.items()}
        smp = {'elec': {**{k: FakeArray(v) for k, v in sampler['elec'].items() if k != 'psi'},
                        'psi': PS(FakeArray(sampler['elec']['psi'][0]), FakeArray(sampler['elec']['psi'][1]))},
               'nuc': {'R': FakeArray(sampler['nuc']['R'])}, 'update_nuc_counter': FakeArray(sampler['update_nuc_counter'])}
        blob = pickle.dumps((1234, TS(smp, wrap(ck.stack_params(per_state)), {'opt': FakeArray(np.arange(3.0))})))
    finally:
        for n in names:
            del sys.modules[n]
    assert b'jax._src.array' in blob and b'deepqmc.types' in blob and 'jax' not in sys.modules
    (tmp_path / 'chkpt-1234.pt').write_bytes(blob)
    step, state = ck.load_reference_checkpoint(str(tmp_path / 'chkpt-1234.pt'))
    assert step == 1234 and isinstance(state, ck.TrainState) and isinstance(state.sampler['elec']['psi'], ck.Psi)
    back = ck.params_per_state(state.params)
    assert len(back) == 2 and all(np.array_equal(back[s][k], per_state[s][k]) for s in range(2) for k in per_state[s])
    assert set(back[0]) == set(PN.param_shapes(spec))
    assert np.array_equal(state.sampler['elec']['r'], sampler['elec']['r']) and np.array_equal(state.opt['opt'], np.arange(3.0))
    mirror = ck.sampler_state_to_mirror(state.sampler)
    assert len(mirror['elec']) == 1 and len(mirror['elec'][0]) == 2 and mirror['elec'][0][1]['r'].shape == (5, 4, 3)
    assert mirror['elec'][0][1]['tau'].item() == 0.4 and mirror['elec'][0][0]['age'].dtype == torch.int32
    again = ck.mirror_to_sampler_state(mirror)
    assert np.array_equal(again['elec']['r'], sampler['elec']['r']) and np.allclose(again['elec']['tau'], sampler['elec']['tau'])


def test_writer_is_loadable_the_way_the_reference_loads_it(tmp_path):
    spec, per_state, sampler = _problem()
    path = str(tmp_path / 'chkpt-7.pt')
    ck.save_reference_checkpoint(path, 7, sampler, ck.stack_params(per_state), opt=None)
    assert 'deepqmc' not in sys.modules  # the writer cleans up its stand-in modules
    FakeArray, TS, PS, names = _install_standins()
    try:
        with open(path, 'rb') as f:
            step, state = pickle.load(f)  # CheckpointStore.load (log.py:122-126)
    finally:
        for n in names:
            del sys.modules[n]
    assert step == 7 and type(state).__name__ == 'TrainState' and type(state).__module__ == 'deepqmc.types'
    assert type(state.sampler['elec']['psi']).__name__ == 'Psi'
    flat = PN.flatten_haiku_tree(state.params)
    assert all(v.shape[0] == 2 for v in flat.values()) and np.array_equal(flat[PN.BF_UP + ':w'][1], per_state[1][PN.BF_UP + ':w'])
    assert state.sampler['elec']['r'].shape == (1, 2, 5, 4, 3)  # [mol, state, walkers of all devices, N, 3] (log.py:45-58)
    step2, state2 = ck.load_reference_checkpoint(path)
    assert step2 == 7 and np.array_equal(state2.sampler['elec']['tau'], sampler['elec']['tau'])
