"""Training checkpoints of the reference, read and written without JAX (checkpoint part).

The reference stores ``pickle.dump((step, serialize_train_state(train_state)), f)`` in ``chkpt-<step>.pt``
(src/deepqmc/log.py:45-58,113-126): ``TrainState(sampler, params, opt)`` (types.py:99-104) with
  * ``params``: the Haiku tree ``{module path: {name: array[S, ...]}}`` with a leading electronic-state axis (wf/base.py:27-44,
    device axis already removed by ``select_one_device``),
  * ``sampler``: ``{'elec': {'r': [mol, S, B, N, 3], 'psi': Psi(sign, log)[mol, S, B], 'age': [mol, S, B], 'tau': [mol, S],
    (+ 'force')}, 'nuc': {'R': [mol, M, 3]}, 'update_nuc_counter': [mol]}`` with the walkers of all devices gathered,
  * ``opt``: the optimiser state (optax / kfac pytrees), opaque here.
Leaves are ``jax.Array`` objects, which pickle as ``jax._src.array._reconstruct_array(fun, args, arr_state, aval_state)`` around
the ``__reduce__`` of the underlying numpy array.  The reader resolves exactly those globals with stand-ins (no JAX, no deepqmc
import; unknown classes of the optimiser state become inert ``Opaque`` records), the writer emits the same layout with numpy
leaves and GLOBAL references to ``deepqmc.types.TrainState`` / ``Psi``, which the reference's ``CheckpointStore.load``
resolves to its own classes (``deserialize_train_state`` only applies jnp functions to the leaves, which accept numpy).

STATUS: written from the reference's source and from JAX's documented pickling protocol; there is neither a reference
checkpoint nor JAX in this image, so the round trip is tested against a faithful stand-in only (tests/test_checkpoint.py).
The result.h5 schema (log.py:204-275) needs h5py, which is not installed here: not built.
"""
from __future__ import annotations

import io
import pickle
import sys
import types
from typing import NamedTuple

import numpy as np

from . import params as PN


class TrainState(NamedTuple):
    sampler: object
    params: object
    opt: object


class Psi(NamedTuple):
    sign: object
    log: object


class Opaque:
    """Inert record for classes of the optimiser state that are not importable here."""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs, self.state = args, kwargs, None

    def __setstate__(self, state):
        self.state = state

    def __reduce__(self):
        return (Opaque, self.args)


def _reconstruct_array(fun, args, arr_state, aval_state=None):
    """jax._src.array._reconstruct_array without the device_put: rebuild the numpy payload."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == '_reconstruct_array' and module.startswith('jax'):
            return _reconstruct_array
        if module == 'deepqmc.types':
            return {'TrainState': TrainState, 'Psi': Psi}.get(name, Opaque)
        if module.split('.')[0] in ('numpy', 'builtins', 'collections', 'copyreg', '_codecs'):
            return super().find_class(module, name)
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (Opaque,), {'__module__': module})


def load_reference_checkpoint(path: str):
    """-> (step, TrainState) with numpy leaves; ``TrainState.params`` is the nested Haiku tree with its leading state axis."""
    with open(path, 'rb') as f:
        step, state = _Unpickler(f).load()
    if not isinstance(state, TrainState):
        state = TrainState(*state)
    return int(step), state


def params_per_state(params_tree: dict) -> list[dict]:
    """Haiku tree with a leading electronic-state axis -> one ':'-flattened tree per state (what B200Ansatz.apply takes)."""
    flat = PN.flatten_haiku_tree(params_tree)
    n_state = {v.shape[0] for v in flat.values()}
    assert len(n_state) == 1, 'every parameter carries the electronic-state axis'
    return [{k: np.asarray(v[s], dtype=np.float64) for k, v in flat.items()} for s in range(n_state.pop())]


def stack_params(per_state: list[dict]) -> dict:
    """inverse of params_per_state: list of flat trees -> nested Haiku tree with the leading state axis"""
    keys = per_state[0].keys()
    return PN.unflatten_haiku_tree({k: np.stack([np.asarray(p[k]) for p in per_state]) for k in keys})


def sampler_state_to_mirror(sampler: dict, device=None, dtype=None):
    """Reference sampler state -> the state layout of deepqmc_b200.sampling.MultiNuclearGeometrySampler over a
    MultiElectronicStateSampler (lists over molecules and states; psi is recomputed by ``sampler.update``)."""
    import torch

    elec = sampler['elec']
    r = np.asarray(elec['r'])
    n_mol, n_state = r.shape[:2]
    tau = np.asarray(elec['tau']).reshape(n_mol, n_state)
    age = np.asarray(elec['age'])
    mk = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
    fdt = dtype or torch.float64
    out = {'nuc': [{'R': torch.as_tensor(np.asarray(sampler['nuc']['R'])[m], dtype=torch.float64)} for m in range(n_mol)],
           'elec': [], 'update_nuc_counter': torch.as_tensor(np.asarray(sampler.get('update_nuc_counter', np.zeros(n_mol))), dtype=torch.int64)}
    for m in range(n_mol):
        out['elec'].append([{'r': mk(r[m, s], fdt), 'age': mk(age[m, s], torch.int32), 'tau': mk(tau[m, s].reshape(1), fdt)}
                            for s in range(n_state)])
    return out


def mirror_to_sampler_state(state: dict) -> dict:
    """inverse of sampler_state_to_mirror (psi taken from the mirror state if present)"""
    def arr(x):
        return np.asarray(x.detach().cpu()) if hasattr(x, 'detach') else np.asarray(x)

    elec = state['elec']
    per = [[e] if isinstance(e, dict) else e for e in elec]
    stack = lambda f: np.stack([np.stack([f(s) for s in mol]) for mol in per])
    out = {'r': stack(lambda s: arr(s['r'])), 'age': stack(lambda s: arr(s['age'])), 'tau': stack(lambda s: arr(s['tau']).reshape(()))}
    if all('psi' in s for mol in per for s in mol):
        out['psi'] = Psi(stack(lambda s: arr(s['psi'].sign)), stack(lambda s: arr(s['psi'].log)))
    if all('force' in s for mol in per for s in mol):
        out['force'] = stack(lambda s: arr(s['force']))
    return {'elec': out, 'nuc': {'R': np.stack([arr(n['R']) for n in state['nuc']])},
            'update_nuc_counter': arr(state['update_nuc_counter'])}


def _reference_globals():
    """Stand-in classes whose pickled GLOBAL reference is deepqmc.types.TrainState / Psi."""
    mod = sys.modules.get('deepqmc.types')
    if mod is not None and hasattr(mod, 'TrainState'):
        return mod.TrainState, mod.Psi, None
    pkg, mod = types.ModuleType('deepqmc'), types.ModuleType('deepqmc.types')
    ts = NamedTuple('TrainState', [('sampler', object), ('params', object), ('opt', object)])
    ps = NamedTuple('Psi', [('sign', object), ('log', object)])
    ts.__module__ = ps.__module__ = 'deepqmc.types'
    mod.TrainState, mod.Psi = ts, ps
    pkg.types = mod
    return ts, ps, (pkg, mod)


def save_reference_checkpoint(path: str, step: int, sampler: dict, params_tree: dict, opt=None) -> None:
    """Write ``(step, TrainState(sampler, params, opt))`` in the reference's checkpoint layout with numpy leaves."""
    ts, ps, fake = _reference_globals()
    installed = []
    if fake is not None:
        for name, m in (('deepqmc', fake[0]), ('deepqmc.types', fake[1])):
            if name not in sys.modules:
                sys.modules[name] = m
                installed.append(name)
    try:
        sampler = dict(sampler)
        elec = dict(sampler['elec'])
        if 'psi' in elec and not isinstance(elec['psi'], ps):
            elec['psi'] = ps(np.asarray(elec['psi'][0]), np.asarray(elec['psi'][1]))
        sampler['elec'] = elec
        buf = io.BytesIO()
        pickle.dump((int(step), ts(sampler, params_tree, opt)), buf)
    finally:
        for name in installed:
            del sys.modules[name]
    with open(path, 'wb') as f:
        f.write(buf.getvalue())
