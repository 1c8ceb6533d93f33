"""Input/output structs of the hot path (reference: src/deepqmc/types.py:20-56).

``PhysicalConfiguration`` and ``Psi`` keep the reference's field names.  Arrays are torch
tensors (device-memory plumbing); a leading walker axis is allowed everywhere the reference
would ``vmap``.
"""
from __future__ import annotations

import dataclasses
from typing import NamedTuple

import torch

__all__ = ['Psi', 'PhysicalConfiguration']


class Psi(NamedTuple):
    sign: torch.Tensor
    log: torch.Tensor


@dataclasses.dataclass(frozen=True)
class PhysicalConfiguration:
    R: torch.Tensor  # [..., M, 3]
    r: torch.Tensor  # [..., N, 3]
    mol_idx: torch.Tensor  # [...]

    def __getitem__(self, idx):
        return PhysicalConfiguration(self.R[idx], self.r[idx], self.mol_idx[idx])

    def __len__(self):
        return len(self.r)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)

    @property
    def batch_shape(self):
        assert self.r.shape[:-2] == self.R.shape[:-2] == self.mol_idx.shape
        return self.r.shape[:-2]
