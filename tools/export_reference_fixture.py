#!/usr/bin/env python
"""Run this where the reference (deepqmc + jax + haiku + hydra) IS installed -- not in the build container -- to record a
fixture for an ansatz the reference's own test-suite does not cover (Psiformer, FermiNet, TransPsiformer, default.yaml):

    JAX_ENABLE_X64=1 python tools/export_reference_fixture.py psiformer LiH tests/golden/external/psiformer_LiH.npz

It instantiates the reference's hydra config ``conf/ansatz/<name>.yaml`` (optionally with smaller sizes), initialises it with
PRNGKey(0), evaluates log|psi|, its Laplacian / gradient and the local energy on the reference's standard LiH-style walker and
stores parameters (':'-flattened Haiku names) and outputs.  tests/test_external_fixtures.py then checks the CPU oracle and the
CUDA engine against every file found under tests/golden/external/ (the directory is empty in this repository: the build
container has no JAX).
"""
import sys


def main():
    name, mol_name, out = sys.argv[1], sys.argv[2], sys.argv[3]
    overrides = sys.argv[4:]
    import os

    import haiku as hk
    import jax
    import jax.numpy as jnp
    import numpy as np
    from hydra import compose, initialize_config_dir
    from hydra.utils import instantiate

    import deepqmc
    from deepqmc.hamil import MolecularHamiltonian
    from deepqmc.molecule import Molecule
    from deepqmc.physics import reverse_forward_laplacian
    from deepqmc.sampling.electron_sample_initializers import AtomCenteredElectronInitializer, ShellBasedDistribution
    from deepqmc.types import PhysicalConfiguration

    jax.config.update('jax_enable_x64', True)
    mol = Molecule.from_name(mol_name)
    hamil = MolecularHamiltonian(mol=mol)
    conf_dir = os.path.join(os.path.dirname(deepqmc.__file__), 'conf', 'ansatz')
    with initialize_config_dir(version_base=None, config_dir=conf_dir):
        cfg = compose(config_name=name, overrides=overrides)
    ansatz_factory = instantiate(cfg, _recursive_=True)
    ansatz = hk.without_apply_rng(hk.transform(lambda pc: ansatz_factory(hamil)(pc)))
    r = AtomCenteredElectronInitializer(atom_centered_distribution=ShellBasedDistribution())(
        jax.random.PRNGKey(0), mol.charges, hamil.ns_valence, mol.coords, hamil.n_up, hamil.n_down)
    pc = PhysicalConfiguration(R=mol.coords, r=r, mol_idx=jnp.zeros(1))
    params = ansatz.init(jax.random.PRNGKey(0), pc)
    psi = ansatz.apply(params, pc)
    lap, grad = reverse_forward_laplacian(lambda x: ansatz.apply(params, pc._replace(r=x.reshape(-1, 3)) if hasattr(pc, '_replace')
                                                                 else PhysicalConfiguration(R=pc.R, r=x.reshape(-1, 3), mol_idx=pc.mol_idx)).log)(r.flatten())
    e_loc, stats = hamil.local_energy(ansatz.apply)(jax.random.PRNGKey(0), params, pc)
    flat = {}
    for mod, leaves in params.items():
        for leaf, v in leaves.items():
            flat[f'param/{mod}:{leaf}'] = np.asarray(v, dtype=np.float64)
    np.savez(out, ansatz=name, molecule=mol_name, overrides=np.array(overrides, dtype=str), r=np.asarray(r), R=np.asarray(mol.coords),
             sign=np.asarray(psi.sign), log=np.asarray(psi.log), lap=np.asarray(lap), grad=np.asarray(grad), e_loc=np.asarray(e_loc),
             **{f'stat/{k}': np.asarray(v) for k, v in stats.items()}, **flat)
    print('wrote', out, 'log|psi| =', float(psi.log), 'E_loc =', float(e_loc))


if __name__ == '__main__':
    main()
