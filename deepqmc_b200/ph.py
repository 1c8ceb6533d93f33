"""Host side of the pseudo-Hamiltonian (reference: src/deepqmc/ecp/pseudo_hamiltonian.py).

The pseudo-Hamiltonian replaces the semi-local effective core potential of P, S, Cl and the 3d metals Cr-Zn by a
fully local operator with a position-dependent mass tensor, so the 12 N N_ecp quadrature forwards of the
Gaussian-type ECP disappear.  This module only *loads the tables* (QMCPACK pseudopotential XML, the OPH23 set the
reference ships as ``deepqmc/ecp/ph_data/<El>.<suffix>.xml``) and hands them to libdqmc_b200.so
(``dqmc_set_pseudo_hamiltonian``); interpolation, the coefficients A(r), b(r), the Cholesky seeds of the
forward-Laplacian pass and the energy assembly are CUDA (csrc/common.cuh ``ph_coeff_kernel`` / ``PhMetric``).

The tables are not part of this repository: pass ``ph_data_dir`` (or set ``DQMC_PH_DATA``) to the directory holding
the XML files, e.g. ``<site-packages>/deepqmc/ecp/ph_data``.
"""
from __future__ import annotations

import os
from xml.etree import ElementTree

import numpy as np

# atomic number -> (symbol, default variant); reference pseudo_hamiltonian.py:18-29
PH_ELEMENTS = {15: ('P', 'cc'), 16: ('S', 'cc'), 17: ('Cl', 'cc'), 24: ('Cr', 'cc'), 25: ('Mn', 'hf'),
               26: ('Fe', 'cc'), 27: ('Co', 'cc'), 28: ('Ni', 'hf'), 29: ('Cu', 'hf'), 30: ('Zn', 'cc')}


def read_ph_xml(path: str):
    """-> (r*V_loc [G], r*V_L2 [G], n_valence, r_max).

    QMCPACK layout: <pseudo><header zval=…/><grid/><semilocal format="r*V"><vps l="s|p|d|…"><radfunc><grid ri rf
    npts/><data>…</data>.  With the s and d channels r*V_s, r*V_d the reference forms (pseudo_hamiltonian.py:59-68)
    r*V_loc = r*V_s + Z_val (the effective-charge offset cancels the -Z_val/r Coulomb tail the Hamiltonian adds) and
    r*V_L2 = -(r*V_s - r*V_d)/6.
    """
    root = ElementTree.parse(path).getroot()
    zval = float(root.find('header').attrib['zval'])
    semi = root.find('semilocal')
    assert semi is not None and semi.attrib.get('format', 'r*V') == 'r*V', 'expected a semilocal block in r*V format'
    chan = {}
    r_max = None
    for vps in semi.findall('vps'):
        rad = vps.find('radfunc')
        grid = rad.find('grid')
        assert grid.attrib.get('type', 'linear') == 'linear' and float(grid.attrib.get('ri', 0.0)) == 0.0
        data = np.array(rad.find('data').text.split(), dtype=np.float64)
        assert len(data) == int(grid.attrib['npts'])
        r_max = float(grid.attrib['rf'])
        chan[vps.attrib['l']] = data
    rv_s, rv_d = chan['s'], chan['d']
    return rv_s + zval, (rv_d - rv_s) / 6.0, zval, r_max


class PseudoHamiltonianPotential:
    """Potential record of the PH (reference :165-178, load_PH_functions :71-112): ``ns_valence`` per nucleus, one
    table pair per distinct element, ``tab_of_nuc[M]`` (-1: plain Coulomb nucleus)."""

    def __init__(self, charges, ecp_type, ecp_mask, ph_data_dir=None):
        ph_data_dir = ph_data_dir or os.environ.get('DQMC_PH_DATA')
        if not ph_data_dir:
            raise ValueError('pseudo-Hamiltonian tables needed: pass ph_data_dir or set DQMC_PH_DATA '
                             '(the reference ships them as deepqmc/ecp/ph_data/*.xml)')
        suffix = str(ecp_type).removeprefix('PH') or None
        names, tabs, nsv, tab_of_nuc = [], [], [], []
        self.r_max = None
        for z, m in zip(charges, ecp_mask):
            z = int(z)
            if not m:
                nsv.append(float(z)); tab_of_nuc.append(-1)
                continue
            if z not in PH_ELEMENTS:
                raise ValueError(f'Pseudo-Hamiltonian for atomic number {z} not found (probably does not exist!)')
            sym, default = PH_ELEMENTS[z]
            if sym not in names:
                loc, l2, zval, r_max = read_ph_xml(os.path.join(ph_data_dir, f'{sym}.{suffix or default}.xml'))
                assert self.r_max in (None, r_max) and (not tabs or len(loc) == tabs[0].shape[1]), 'tables must share one grid'
                self.r_max = r_max
                names.append(sym); tabs.append(np.stack([loc, l2])); self._zval = {**getattr(self, '_zval', {}), sym: zval}
            nsv.append(self._zval[sym]); tab_of_nuc.append(names.index(sym))
        self.ns_valence = np.asarray(nsv, dtype=np.float64)
        self.tables = np.ascontiguousarray(np.stack(tabs), dtype=np.float64)  # [n_tab][2][G]
        self.tab_of_nuc = np.asarray(tab_of_nuc, dtype=np.int32)
        self.nuc_with_nl_pot = np.zeros(0, dtype=np.int64)  # fully local: no quadrature (physics.py:52-76)
