"""Ansatz hyper-parameter record shared by the CUDA engine, the host shims and the tests.

The fields restate the knobs of the reference's hydra ansatz YAMLs that fix kernel shapes
(reference: src/deepqmc/conf/ansatz/psiformer.yaml, ferminet.yaml).
Pure data: no arithmetic.
"""
from __future__ import annotations

import dataclasses

__all__ = ['AnsatzSpec', 'psiformer_spec', 'ferminet_spec', 'transpsiformer_spec']


@dataclasses.dataclass(frozen=True)
class AnsatzSpec:
    kind: str  # 'psiformer' | 'ferminet' | 'transpsiformer'
    n_up: int
    n_down: int
    n_nuc: int
    embedding_dim: int = 256
    n_layers: int = 4
    n_heads: int = 4
    n_determinants: int = 16
    edge_dim: int = 32  # two_particle_stream_dim (FermiNet)
    # e-e cusp (reference: conf/ansatz/psiformer.yaml:18-26); 'none' for FermiNet
    cusp: str = 'psiformer'
    cusp_same_scale: float = 0.25
    cusp_anti_scale: float = 0.5
    # TransPsiformer (reference: conf/ansatz/transpsiformer.yaml): nuclei are extra attention tokens
    # (elec_to_nuc = false) and the envelope exponents are read out of the nuclear embeddings
    n_env_per_nuc: int = 1  # SimplifiedNucleusDependentEnvelopes.n_envelope_per_nucleus (3)
    nuc_edge_dim: int = 32  # NucleiEmbedding edge_mlp width (gnn/electron_gnn.py:476-484)
    charges: tuple = ()     # nuclear charges (atom-type one-hot of the nuclear embedding)

    @property
    def n_elec(self):
        return self.n_up + self.n_down

    @property
    def head_dim(self):
        return self.embedding_dim // self.n_heads

    @property
    def n_feat_in(self):
        # Psiformer: [log1p|r_iI|, d_iI log1p/r] for all I (+) spin; FermiNet: [|r_iI|, d_iI]
        return 4 * self.n_nuc + (1 if self.kind in ('psiformer', 'transpsiformer') else 0)


def psiformer_spec(hamil, **kw):
    """reference: src/deepqmc/conf/ansatz/psiformer.yaml"""
    return AnsatzSpec('psiformer', hamil.n_up, hamil.n_down, hamil.n_nuc, **kw)


def ferminet_spec(hamil, **kw):
    """reference: src/deepqmc/conf/ansatz/ferminet.yaml"""
    kw.setdefault('cusp', 'none')
    return AnsatzSpec('ferminet', hamil.n_up, hamil.n_down, hamil.n_nuc, **kw)


def transpsiformer_spec(hamil, **kw):
    """reference: src/deepqmc/conf/ansatz/transpsiformer.yaml"""
    kw.setdefault('n_env_per_nuc', 3)
    kw.setdefault('charges', tuple(float(z) for z in hamil.mol.charges))
    return AnsatzSpec('transpsiformer', hamil.n_up, hamil.n_down, hamil.n_nuc, **kw)
