"""Synthetic pseudo-Hamiltonian tables in the QMCPACK XML layout of the reference's ecp/ph_data/*.xml files.

The real OPH23 tables ship with the reference package and are not part of this repository; the parity tests only need
*some* smooth radial functions in the same file format (header zval, semilocal r*V block with the s, p, d, ... channels
on the linear grid [0, 10] with 10001 points).  The functions below keep the mass tensor A(r) positive definite.
"""
import os

import numpy as np

SPECS = {  # symbol -> (file suffix, zval, amplitude / exponent of the s channel bump, of the d - s difference)
    'P': ('cc', 5.0, (1.7, 0.9), (1.3, 1.0)),
    'S': ('cc', 6.0, (2.1, 0.8), (-1.1, 1.2)),
    'Cl': ('cc', 7.0, (2.4, 0.7), (1.5, 1.1)),
}


def channels(sym, l2_scale=1.0):
    _, zval, (a_s, e_s), (a_l, e_l) = SPECS[sym]
    a_l = a_l * l2_scale
    r = np.linspace(0.0, 10.0, 10001)
    rv_s = -zval * (1 - np.exp(-1.5 * r * r)) + a_s * r * r * np.exp(-e_s * r * r)
    rv_d = rv_s + a_l * r * r * np.exp(-e_l * r * r)
    rv_p = rv_s + (rv_d - rv_s) / 3  # the tables satisfy 2 (V_s - V_d) = 3 (V_p - V_d) (pseudo_hamiltonian.py:63-65)
    return r, {'s': rv_s, 'p': rv_p, 'd': rv_d}


def write_synthetic_ph(directory, symbols=('P', 'S', 'Cl'), l2_scale=1.0):
    os.makedirs(directory, exist_ok=True)
    for sym in symbols:
        suffix, zval = SPECS[sym][0], SPECS[sym][1]
        _, ch = channels(sym, l2_scale)
        grid = '<grid type="linear" units="bohr" ri="0.0" rf="10.0" npts="10001"/>'
        out = ['<?xml version="1.0" encoding="UTF-8"?>', '<pseudo version="0.5">',
               f'  <header symbol="{sym}" atomic-number="0" zval="{zval:g}" relativistic="unknown"/>', f'  {grid}',
               '  <semilocal units="hartree" format="r*V" npots-down="3" npots-up="0" l-local="2">']
        for l, data in ch.items():
            out.append(f'    <vps principal-n="0" l="{l}" spin="-1" cutoff="2.0" occupation="unknown">')
            out.append(f'      <radfunc>\n        {grid}\n        <data>')
            for i in range(0, len(data), 3):
                out.append('          ' + '  '.join(f'{v:.14e}' for v in data[i:i + 3]))
            out.append('        </data>\n      </radfunc>\n    </vps>')
        out += ['  </semilocal>', '</pseudo>']
        with open(os.path.join(directory, f'{sym}.{suffix}.xml'), 'w') as f:
            f.write('\n'.join(out) + '\n')
    return directory
