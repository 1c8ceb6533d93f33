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
