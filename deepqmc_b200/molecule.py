"""Molecule container and the named geometries the hot-path configs use.

Mirrors the reference's ``Molecule`` dataclass (reference: src/deepqmc/molecule.py:33-78):
coordinates are stored in bohr, ``unit='angstrom'`` inputs are converted with the CODATA
constants from scipy (reference: src/deepqmc/units.py:17-22).  Only data lives here; no
wave-function arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy import constants

__all__ = ['Molecule', 'angstrom_to_bohr']


def angstrom_to_bohr(x):
    return np.asarray(x, dtype=np.float64) * constants.angstrom / constants.value(
        'atomic unit of length'
    )


def _benzene():
    # D6h benzene, C-C 1.39 A, C-H 1.09 A (not shipped by the reference,
    # there it would come through conf/hamil/mol/from_file.yaml).
    rc, rh = 1.39, 1.39 + 1.09
    ang = np.arange(6) * np.pi / 3
    c = np.stack([rc * np.cos(ang), rc * np.sin(ang), np.zeros(6)], -1)
    h = np.stack([rh * np.cos(ang), rh * np.sin(ang), np.zeros(6)], -1)
    return dict(
        coords=np.concatenate([c, h]).tolist(),
        charges=[6] * 6 + [1] * 6,
        charge=0,
        spin=0,
        unit='angstrom',
    )


# Geometries as given by the reference's conf/hamil/mol/<name>.yaml files (angstrom).
_NAMED = {
    'H2': dict(coords=[[0.0, 0.0, 0.0], [0.742, 0.0, 0.0]], charges=[1, 1], charge=0, spin=0, unit='angstrom'),
    'LiH': dict(coords=[[0.0, 0.0, 0.0], [1.595, 0.0, 0.0]], charges=[3, 1], charge=0, spin=0, unit='angstrom'),
    'C': dict(coords=[[0.0, 0.0, 0.0]], charges=[6], charge=0, spin=2, unit='angstrom'),
    'N2': dict(coords=[[-2.13534, 0.0, 0.0], [2.13534, 0.0, 0.0]], charges=[7, 7], charge=0, spin=0, unit='angstrom'),
    'H2O': dict(coords=[[0.0, 0.0, 0.0], [0.75695, 0.58588, 0.0], [-0.75695, 0.58588, 0.0]], charges=[8, 1, 1], charge=0, spin=0, unit='angstrom'),
    'cyclobutadiene_square': dict(
        coords=[[0.0, 0.0, 0.0], [2.74199, 0.0, 0.0], [2.74199, 2.74199, 0.0], [0.0, 2.74199, 0.0],
                [-1.44047, -1.44047, 0.0], [4.18246, -1.44047, 0.0], [4.18246, 4.18246, 0.0], [-1.44047, 4.18246, 0.0]],
        charges=[6, 6, 6, 6, 1, 1, 1, 1], charge=0, spin=0, unit='angstrom'),
    'benzene': _benzene(),
}

# Further named geometries of the reference's conf/hamil/mol/<name>.yaml files (geometry DATA only), so that
# Molecule.from_name accepts every name the reference accepts.
_NAMED.update({
    'B': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[5], charge=0, spin=1, unit='angstrom'),
    'B2': dict(coords=[[-0.7951, 0.0, 0.0], [0.7951, 0.0, 0.0]],
        charges=[5, 5], charge=0, spin=2, unit='angstrom'),
    'Be': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[4], charge=0, spin=0, unit='angstrom'),
    'Be2': dict(coords=[[-1.23, 0.0, 0.0], [1.23, 0.0, 0.0]],
        charges=[4, 4], charge=0, spin=0, unit='angstrom'),
    'C2': dict(coords=[[-0.621265, 0.0, 0.0], [0.621265, 0.0, 0.0]],
        charges=[6, 6], charge=0, spin=0, unit='angstrom'),
    'CH2O': dict(coords=[[-2.288269281387329, 0.0, 0.0], [0.0, 0.0, 0.0], [1.125103235244751, -1.7935682535171509, 0.0], [1.125103235244751, 1.7935682535171509, 0.0]],
        charges=[8, 6, 1, 1], charge=0, spin=0, unit='bohr'),
    'CH4': dict(coords=[[0.0, 0.0, 0.0], [0.62912, 0.62912, 0.62912], [-0.62912, -0.62912, 0.62912], [0.62912, -0.62912, -0.62912], [-0.62912, 0.62912, -0.62912]],
        charges=[6, 1, 1, 1, 1], charge=0, spin=0, unit='angstrom'),
    'CO': dict(coords=[[-0.575169, 0.0, 0.0], [0.575169, 0.0, 0.0]],
        charges=[6, 8], charge=0, spin=0, unit='angstrom'),
    'CO2': dict(coords=[[-1.161, 0.0, 0.0], [0.0, 0.0, 0.0], [1.161, 0.0, 0.0]],
        charges=[8, 6, 8], charge=0, spin=0, unit='angstrom'),
    'Cl': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[17], charge=0, spin=1, unit='angstrom'),
    'Cr': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[24], charge=0, spin=6, unit='angstrom'),
    'Fe': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[26], charge=0, spin=0, unit='angstrom'),
    'H': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[1], charge=0, spin=1, unit='angstrom'),
    'H10': dict(coords=[[0.0, 0.0, 0.0], [0.95305, 0.0, 0.0], [1.9061, 0.0, 0.0], [2.85914, 0.0, 0.0], [3.81219, 0.0, 0.0], [4.76524, 0.0, 0.0], [5.71829, 0.0, 0.0], [6.67134, 0.0, 0.0], [7.62439, 0.0, 0.0], [8.57743, 0.0, 0.0]],
        charges=[1, 1, 1, 1, 1, 1, 1, 1, 1, 1], charge=0, spin=0, unit='angstrom'),
    'H2+': dict(coords=[[-0.52918, 0.0, 0.0], [0.52918, 0.0, 0.0]],
        charges=[1, 1], charge=1, spin=1, unit='angstrom'),
    'H2O3': dict(coords=[[2.256039344065506, 0.07979422300310074, 2.2237440085628726], [-2.9879436323137165e-16, -9.67142599187823e-16, 1.267380970753913], [0.0, 0.0, 0.0], [2.418987015584533, 1.0503857463931796, 2.774087487340858], [1.0922959761531792, -6.938893903907228e-16, 1.6065553188661499]],
        charges=[8, 8, 8, 1, 1], charge=0, spin=2, unit='angstrom'),
    'He': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[2], charge=0, spin=0, unit='angstrom'),
    'Li2': dict(coords=[[-1.3364, 0.0, 0.0], [1.3364, 0.0, 0.0]],
        charges=[3, 3], charge=0, spin=0, unit='angstrom'),
    'LiNH3_6': dict(coords=[[0.0, 0.0, 0.0], [0.0, 0.0, 2.078], [-0.9795785942037634, 1.696679895167815, -0.6926666666666663], [-0.9795785942037646, -1.6966798951678144, -0.6926666666666662], [1.9591571884075278, -4.798551159619614e-16, -0.6926666666666663], [0.47015020731013546, 0.8143240462501947, 2.473265898181145], [-0.9403004146202705, 1.1515358930016434e-16, 2.473265898181145], [0.47015020731013457, -0.8143240462501952, 2.473265898181145], [-1.009191989750838, 1.7479718008399818, -1.7109456987677352], [-1.9494924043711082, 1.7479718008399818, -0.3811600997067045], [-0.5390417824407028, 2.5622958470901764, -0.38116009970670406], [-1.009191989750839, -1.7479718008399812, -1.7109456987677352], [-0.5390417824407039, -2.562295847090176, -0.38116009970670445], [-1.9494924043711093, -1.747971800839981, -0.38116009970670395], [2.0183839795016767, -4.943614959894052e-16, -1.7109456987677352], [2.4885341868118123, 0.8143240462501944, -0.3811600997067045], [2.488534186811812, -0.814324046250195, -0.38116009970670406]],
        charges=[3, 7, 7, 7, 7, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1], charge=0, spin=1, unit='bohr'),
    'NH3': dict(coords=[[0.0, 0.0, 0.116488], [0.0, 0.93973, -0.27181], [0.81383, -0.46986, -0.27181], [-0.81383, -0.46986, -0.27181]],
        charges=[7, 1, 1, 1], charge=0, spin=0, unit='angstrom'),
    'Sc': dict(coords=[[0.0, 0.0, 0.0]],
        charges=[21], charge=0, spin=1, unit='angstrom'),
    'ScO': dict(coords=[[0.0, 0.0, 0.0], [1.668, 0.0, 0.0]],
        charges=[21, 8], charge=0, spin=1, unit='angstrom'),
    'bicyclobutane': dict(coords=[[0.7507, 0.0, -0.3193], [-0.7507, 0.0, -0.3193], [0.0, 1.135, 0.3153], [0.0, -1.135, 0.3153], [1.4194, 0.0, -1.1631], [-1.4194, 0.0, -1.1631], [0.0, 2.082, -0.2148], [0.0, -2.082, -0.2148], [0.0, 1.2163, 1.402], [0.0, -1.2163, 1.402]],
        charges=[6, 6, 6, 6, 1, 1, 1, 1, 1, 1], charge=0, spin=0, unit='bohr'),
})


@dataclass(frozen=True)
class Molecule:
    coords: np.ndarray
    charges: np.ndarray
    charge: int
    spin: int
    data: dict | None = None
    unit: str = 'bohr'
    n_atom_types: int = field(init=False)

    all_names = frozenset(_NAMED)

    def __post_init__(self):
        conv = {'bohr': lambda x: np.asarray(x, dtype=np.float64), 'angstrom': angstrom_to_bohr}[self.unit]
        object.__setattr__(self, 'coords', conv(self.coords).reshape(-1, 3))
        object.__setattr__(self, 'charges', np.asarray(self.charges, dtype=np.float64))
        object.__setattr__(self, 'data', self.data or {})
        object.__setattr__(self, 'unit', 'bohr')
        object.__setattr__(self, 'n_atom_types', len(np.unique(self.charges)))

    def __len__(self):
        return len(self.charges)

    def __iter__(self):
        yield from zip(self.coords, self.charges)

    @classmethod
    def from_name(cls, name: str) -> 'Molecule':
        if name not in _NAMED:
            raise ValueError(f'Unknown molecule name: {name}')
        return cls(**_NAMED[name])

    @classmethod
    def from_file(cls, file: str) -> 'Molecule':
        import yaml

        with open(file) as stream:
            return cls(**yaml.safe_load(stream))
