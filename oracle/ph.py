"""Restatement of the reference's pseudo-Hamiltonian (TEST INFRASTRUCTURE, see oracle/__init__).

Follows src/deepqmc/ecp/pseudo_hamiltonian.py of the reference: parse_xml (:32-68), load_PH_functions (:71-112,
tables on linspace(0, 10, 10001) read through jax.scipy.interpolate.RegularGridInterpolator, linear, fill_value 0),
PseudoHamiltonian.local_potential (:180-196), compute_coefficients_of_differential_operators (:198-233),
kinetic_term (:235-278) with compute_differential_operator_using_laplacian (:115-148).
torch.float64 on CPU, single walker.  The second-order term is evaluated exactly as the reference states it:
the Laplacian of v -> log|psi(Q v)| in the transformed coordinates (here: trace of the autograd Hessian).

The XML tables themselves (QMCPACK format, OPH23 set) ship with the reference package (ecp/ph_data/*.xml) and are
not part of this repository: callers pass the directory that holds them.
"""
from __future__ import annotations

import os
from xml.etree import ElementTree

import numpy as np
import torch

F64 = torch.float64

# element -> (symbol, default file suffix), reference pseudo_hamiltonian.py:18-29
ELEMENTS_WITH_EXISTING_PH = {15: ('P', 'cc'), 16: ('S', 'cc'), 17: ('Cl', 'cc'), 24: ('Cr', 'cc'), 25: ('Mn', 'hf'),
                             26: ('Fe', 'cc'), 27: ('Co', 'cc'), 28: ('Ni', 'hf'), 29: ('Cu', 'hf'), 30: ('Zn', 'cc')}


def parse_xml(xml_file):
    """reference :32-68 (positional access: 3rd child of the root = semilocal block, its children 0 / 2 = s / d
    channels, each radfunc -> data).  -> (r V_loc [G], r V_L2 [G], n_valence)"""
    root = ElementTree.parse(xml_file).getroot()
    n_valence = float(root.find('header').attrib['zval'])

    def channel(index):
        return np.array(root[2][index][0][1].text.split(), dtype=np.float64)

    s_arr, d_arr = channel(0), channel(2)
    v0_nl = s_arr - d_arr
    return d_arr + v0_nl + n_valence, -v0_nl / 6, n_valence


def interp_linear(grid: np.ndarray, values: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """RegularGridInterpolator([grid], values, method='linear', fill_value=0.0) at the points x."""
    g = torch.as_tensor(grid, dtype=F64)
    idx = torch.clamp(torch.searchsorted(g, x.detach().contiguous(), right=True) - 1, 0, len(grid) - 2)
    w = (x - g[idx]) / (g[idx + 1] - g[idx])
    val = (1 - w) * values[idx] + w * values[idx + 1]
    inside = (x >= g[0]) & (x <= g[-1])
    return torch.where(inside, val, torch.zeros_like(val))


class OraclePseudoHamiltonian:
    """reference :165-278"""

    def __init__(self, charges, ecp_type, ecp_mask, ph_dir):
        self.ecp_mask = np.asarray(ecp_mask, dtype=bool)
        suffix = str(ecp_type).removeprefix('PH') or None
        self.grid = np.linspace(0, 10.0, 10001)
        ns_valence, self.rV_loc, self.rV_L2 = [], [], []
        cache = {}
        for z, m in zip(charges, self.ecp_mask):
            z = int(z)
            if m:
                assert z in ELEMENTS_WITH_EXISTING_PH, f'Pseudo-Hamiltonian for atomic number {z} not found'
                name, default_suffix = ELEMENTS_WITH_EXISTING_PH[z]
                if name not in cache:
                    cache[name] = parse_xml(os.path.join(ph_dir, f'{name}.{suffix or default_suffix}.xml'))
                loc, l2, nv = cache[name]
                self.rV_loc.append(torch.as_tensor(loc, dtype=F64))
                self.rV_L2.append(torch.as_tensor(l2, dtype=F64))
                ns_valence.append(nv)
            else:
                ns_valence.append(z)
        self.ns_valence = np.asarray(ns_valence, dtype=np.float64)

    def _columns(self, tables, dists_ph):
        return torch.stack([interp_linear(self.grid, t, dists_ph[:, j]) for j, t in enumerate(tables)], dim=1)

    def local_potential(self, r, R):
        dists = torch.linalg.norm(r[:, None] - R[None], dim=-1)
        v = -(torch.as_tensor(self.ns_valence) / dists).sum()
        d_ph = dists[:, torch.as_tensor(self.ecp_mask)]
        return v + (self._columns(self.rV_loc, d_ph) / d_ph).sum()

    def coefficients(self, r, R):
        mask = torch.as_tensor(self.ecp_mask)
        diffs = (r[:, None] - R[None])[:, mask]
        d_ph = torch.linalg.norm(r[:, None] - R[None], dim=-1)[:, mask]
        rv = self._columns(self.rV_L2, d_ph)
        v = rv / d_ph
        b = (2 * v[..., None] * diffs).sum(-2)
        eye = torch.eye(3, dtype=F64)
        diag = (rv * d_ph)[..., None, None] * eye
        nondiag = v[..., None, None] * diffs[..., :, None] * diffs[..., None, :]
        return (diag - nondiag).sum(-3) + 0.5 * eye, b

    def kinetic_term(self, log_psi, r, R):
        """log_psi(r[N,3]) -> log|psi|.  -> (E_kin-like term, lap, quantum_force) as the reference's tuple."""
        A, b = self.coefficients(r.detach(), R)
        Q = torch.linalg.cholesky(A)  # lower: A = Q Q^T
        v = torch.linalg.solve_triangular(Q, r.detach()[..., None], upper=False)[..., 0]
        f = lambda vf: log_psi(torch.einsum('nxy,ny->nx', Q, vf.reshape(-1, 3)))
        vf = v.reshape(-1)
        H = torch.autograd.functional.hessian(f, vf)
        x = vf.clone().requires_grad_(True)
        jac_v, = torch.autograd.grad(f(x), x)
        jac_v = jac_v.reshape(-1, 3)
        lap = torch.diagonal(H).sum()
        jac_r = torch.linalg.solve_triangular(Q.transpose(-1, -2), jac_v[..., None], upper=True)[..., 0]
        first = (b * jac_r).sum()
        qf = (jac_v * jac_v).sum()
        return first - (lap + qf), lap, qf
