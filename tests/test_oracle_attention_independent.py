"""CPU: the oracle's Psiformer trunk (oracle/wf.py psiformer_embeddings: its own restatement of hk.MultiHeadAttention + the
tanh MLP + residuals, reference gnn/update_features.py:241-286, hkext.py:22-137, 215-253) against a THIRD implementation of the
attention algebra that shares no code with it: torch.nn.functional.multi_head_attention_forward fed with the same weights
(head h = columns [h dh, (h + 1) dh) of the projections, logits scaled by 1 / sqrt(dh), softmax over keys, heads concatenated
before the output projection -- the conventions haiku's module and torch's agree on).  The reference ships no fixture for the
attention trunks (its tests evaluate the conv-GNN ansatz only), so this pins the head split / scaling / softmax-axis choices of
the oracle independently of both the product and the oracle's own einsums."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deepqmc_b200 import params as PN
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.spec import psiformer_spec
from oracle import names as P
from oracle import wf


@pytest.mark.parametrize('mol_name,hyper', [('LiH', dict(embedding_dim=32, n_layers=2, n_heads=4)),
                                            ('H2O', dict(embedding_dim=48, n_layers=3, n_heads=3)),
                                            ('N2', dict(embedding_dim=256, n_layers=4, n_heads=4))])
def test_oracle_psiformer_trunk_equals_torch_multi_head_attention(mol_name, hyper):
    mol = Molecule.from_name(mol_name)
    hamil = MolecularHamiltonian(mol=mol)
    spec = psiformer_spec(hamil, n_determinants=2, **hyper)
    pt = wf.to_torch(PN.perturb_params(PN.init_params(spec, 0)))
    rng = np.random.default_rng(3)
    N, d, H = spec.n_elec, spec.embedding_dim, spec.n_heads
    r = torch.as_tensor(mol.coords[rng.integers(0, len(mol.coords), size=N)] + rng.normal(size=(N, 3)))
    R = torch.as_tensor(mol.coords)
    x_oracle = wf.psiformer_embeddings(spec, pt, r, R)

    def W(name):
        return torch.as_tensor(pt[name], dtype=torch.float64)

    # the input projection is not under test: take the oracle's own features for it
    feats, _ = wf.ne_features(r, R, True)
    spins = torch.cat([torch.ones(spec.n_up), -torch.ones(spec.n_down)]).double()[:, None]
    x = torch.cat([feats, spins], 1) @ W(P.GNN + 'electron_embedding/linear:w')
    for l in range(spec.n_layers):
        a = P.attn_prefix(l)
        in_proj = torch.cat([W(a + f'multi_head_attention/{n}:w').T for n in ('query', 'key', 'value')], 0)  # [3d, d]
        out, _ = F.multi_head_attention_forward(
            x[:, None, :], x[:, None, :], x[:, None, :], d, H, in_proj, None, None, None, False, 0.0,
            W(a + 'multi_head_attention/linear:w').T, None, training=False, need_weights=False)
        att = x + out[:, 0, :]
        m = torch.tanh(att @ W(a + 'mlp/linear_0:w') + W(a + 'mlp/linear_0:b'))
        m = torch.tanh(m @ W(a + 'mlp/linear_1:w') + W(a + 'mlp/linear_1:b'))
        x = att + m
    assert x.shape == x_oracle.shape == (N, d)
    assert (x - x_oracle).abs().max().item() < 1e-11 * max(1.0, x_oracle.abs().max().item())


@pytest.mark.parametrize('mol_name,hyper', [('LiH', dict(embedding_dim=32, n_layers=2, n_heads=4)),
                                            ('cyclobutadiene_square', dict(embedding_dim=64, n_layers=3, n_heads=4))])
def test_oracle_transpsiformer_trunk_equals_torch_masked_multi_head_attention(mol_name, hyper):
    """Joint attention over [nuclei; electrons] tokens with the nuclei masked from attending electrons (reference
    gnn/update_features.py:385-451) == torch's multi-head attention with the corresponding boolean attn_mask."""
    from deepqmc_b200.spec import transpsiformer_spec

    mol = Molecule.from_name(mol_name)
    hamil = MolecularHamiltonian(mol=mol)
    spec = transpsiformer_spec(hamil, n_determinants=2, **hyper)
    pt = wf.to_torch(PN.perturb_params(PN.init_params(spec, 0)))
    rng = np.random.default_rng(4)
    N, M, d, H = spec.n_elec, spec.n_nuc, spec.embedding_dim, spec.n_heads
    r = torch.as_tensor(mol.coords[rng.integers(0, len(mol.coords), size=N)] + rng.normal(size=(N, 3)))
    R = torch.as_tensor(mol.coords)
    xe_oracle, xn_oracle = wf.transpsiformer_embeddings(spec, pt, r, R)

    def W(name):
        return torch.as_tensor(pt[name], dtype=torch.float64)

    feats, _ = wf.ne_features(r, R, True)
    spins = torch.cat([torch.ones(spec.n_up), -torch.ones(spec.n_down)]).double()[:, None]
    xe = torch.cat([feats, spins], 1) @ W(P.GNN + 'electron_embedding/linear:w')
    h = torch.cat([wf.nuclei_embedding(spec, pt, R), xe], 0)  # token order of the reference: nuclei first
    not_allowed = torch.zeros(M + N, M + N, dtype=torch.bool)
    not_allowed[:M, M:] = True  # a nucleus never attends an electron
    for l in range(spec.n_layers):
        a = P.comb_prefix(l)
        in_proj = torch.cat([W(a + f'multi_head_attention/{n}:w').T for n in ('query', 'key', 'value')], 0)
        out, _ = F.multi_head_attention_forward(
            h[:, None, :], h[:, None, :], h[:, None, :], d, H, in_proj, None, None, None, False, 0.0,
            W(a + 'multi_head_attention/linear:w').T, None, training=False, need_weights=False, attn_mask=not_allowed)
        att = h + out[:, 0, :]
        m = torch.tanh(att @ W(a + 'mlp/linear_0:w') + W(a + 'mlp/linear_0:b'))
        m = torch.tanh(m @ W(a + 'mlp/linear_1:w') + W(a + 'mlp/linear_1:b'))
        h = att + m
    scale = max(1.0, xe_oracle.abs().max().item())
    assert (h[M:] - xe_oracle).abs().max().item() < 1e-11 * scale and (h[:M] - xn_oracle).abs().max().item() < 1e-11 * scale
