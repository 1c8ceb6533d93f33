"""Walker-independent nuclear stream of the TransPsiformer ansatz (host side, once per
(parameters, geometry) upload).

In the reference's transferable ansatz (src/deepqmc/conf/ansatz/transpsiformer.yaml) the nuclei are
extra attention tokens with ``elec_to_nuc: false`` (gnn/update_features.py:428-434): nuclei never
attend electrons, so their embeddings depend on the parameters and the geometry R only -- not on
the walkers.  The per-walker CUDA engine therefore receives them as *constants*: per layer the key
and value rows of the M nuclear tokens, and the envelope exponents the ``NuclearGNNHead`` reads
out of the final nuclear embeddings (wf/omni.py:181-211, wf/env.py:111-226).  This module evaluates
that O(M^2 d) stream in float64 numpy when parameters are uploaded; everything that depends on
electron positions runs in libdqmc_b200.so.

reference lines: gnn/electron_gnn.py:435-537 (NucleiEmbedding with 'nn' edge features),
gnn/edge_features.py:21-78, hkext.py:22-113 (MLP), :165-202 (GLU, LayerNorm(-1, False, False)).
"""
from __future__ import annotations

import numpy as np

from . import params as PN
from .spec import AnsatzSpec


def _silu(x):
    return x / (1.0 + np.exp(-x))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _layer_norm(x, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps)


def nuclei_embedding(spec: AnsatzSpec, params: dict, R: np.ndarray) -> np.ndarray:
    M = spec.n_nuc
    g = lambda k: np.asarray(params[k], dtype=np.float64)
    d = R[None, :, :] - R[:, None, :]  # [sender, receiver, 3] = receiver - sender
    rr = np.sqrt(np.finfo(np.float64).eps + (d * d).sum(-1))
    lg = np.log1p(rr)
    feats = np.concatenate([lg[..., None], d * (lg / rr)[..., None]], -1)
    inv = np.unique(np.asarray(spec.charges), return_inverse=True)[1]
    onehot = np.eye(M)[inv]  # atom type of the sender
    x = np.concatenate([feats, np.broadcast_to(onehot[:, None, :], (M, M, M))], -1)
    e = _silu(x @ g(PN.NUC_EMB + 'edge_mlp/linear_0:w') + g(PN.NUC_EMB + 'edge_mlp/linear_0:b'))
    e = e @ g(PN.NUC_EMB + 'edge_mlp/linear_1:w') + g(PN.NUC_EMB + 'edge_mlp/linear_1:b')
    h = _silu(e.sum(0) @ g(PN.NUC_EMB + 'embed_mlp/linear_0:w') + g(PN.NUC_EMB + 'embed_mlp/linear_0:b'))
    return h @ g(PN.NUC_EMB + 'embed_mlp/linear_1:w') + g(PN.NUC_EMB + 'embed_mlp/linear_1:b')


def nuclear_stream(spec: AnsatzSpec, params: dict, R) -> dict:
    """-> {'kn': [L][M, d], 'vn': [L][M, d], 'zetas_up': [M, K, E], 'zetas_down': [M, K, E]}"""
    R = np.asarray(R, dtype=np.float64)
    assert R.shape == (spec.n_nuc, 3), 'the TransPsiformer engine takes one geometry per handle'
    g = lambda k: np.asarray(params[k], dtype=np.float64)
    M, d, H = spec.n_nuc, spec.embedding_dim, spec.n_heads
    dh = d // H
    h = nuclei_embedding(spec, params, R)
    kn, vn = [], []
    for l in range(spec.n_layers):
        a = PN.comb_prefix(l)
        q = (h @ g(a + 'multi_head_attention/query:w')).reshape(M, H, dh)
        k = h @ g(a + 'multi_head_attention/key:w')
        v = h @ g(a + 'multi_head_attention/value:w')
        kn.append(k)
        vn.append(v)
        logits = np.einsum('thd,Thd->htT', q, k.reshape(M, H, dh)) / np.sqrt(dh)
        w = np.exp(logits - logits.max(-1, keepdims=True))
        w /= w.sum(-1, keepdims=True)
        o = np.einsum('htT,Thd->thd', w, v.reshape(M, H, dh)).reshape(M, d)
        att = h + o @ g(a + 'multi_head_attention/linear:w')
        m = np.tanh(att @ g(a + 'mlp/linear_0:w') + g(a + 'mlp/linear_0:b'))
        m = np.tanh(m @ g(a + 'mlp/linear_1:w') + g(a + 'mlp/linear_1:b'))
        h = att + m
    x = _layer_norm(h)
    out = {'kn': kn, 'vn': vn}
    K, E = spec.n_determinants, spec.n_env_per_nuc
    for glu, spin in (('zetas_readout_glu', 'up'), ('zetas_readout_glu_1', 'down')):
        y = _sigmoid(x @ g(PN.HEAD + glu + '/W:w') + g(PN.HEAD + glu + '/W:b')) * (
            x @ g(PN.HEAD + glu + '/V:w') + g(PN.HEAD + glu + '/V:b'))
        out[f'zetas_{spin}'] = y.reshape(M, K, E) + g(PN.HEAD + f':zetas_bias_{spin}')
    return out


def nuclear_stream_vjp(spec: AnsatzSpec, params: dict, R, cot: dict) -> dict:
    """Parameter gradient contribution of the nuclear stream: `cot` holds the cotangents the CUDA reverse pass
    accumulated for the stream's outputs ({'kn': [L][M, d], 'vn': [L][M, d], 'zetas_up', 'zetas_down': [M, K, E]});
    the stream itself (O(M^2 d), walker-independent) is differentiated on the host with torch autograd in float64.
    Returns {haiku name: gradient} for every parameter the stream touches."""
    import torch

    R_t = torch.as_tensor(np.asarray(R, dtype=np.float64))
    pt = {k: torch.as_tensor(np.asarray(v, dtype=np.float64)).requires_grad_(True) for k, v in params.items()}
    g = lambda k: pt[k]
    M, d, H = spec.n_nuc, spec.embedding_dim, spec.n_heads
    dh = d // H
    dd = R_t[None, :, :] - R_t[:, None, :]
    rr = torch.sqrt(torch.finfo(torch.float64).eps + (dd * dd).sum(-1))
    lg = torch.log1p(rr)
    feats = torch.cat([lg[..., None], dd * (lg / rr)[..., None]], -1)
    inv = np.unique(np.asarray(spec.charges), return_inverse=True)[1]
    onehot = torch.eye(M, dtype=torch.float64)[torch.as_tensor(inv)]
    x = torch.cat([feats, onehot[:, None, :].expand(M, M, M)], -1)
    silu = torch.nn.functional.silu
    e = silu(x @ g(PN.NUC_EMB + 'edge_mlp/linear_0:w') + g(PN.NUC_EMB + 'edge_mlp/linear_0:b'))
    e = e @ g(PN.NUC_EMB + 'edge_mlp/linear_1:w') + g(PN.NUC_EMB + 'edge_mlp/linear_1:b')
    h = silu(e.sum(0) @ g(PN.NUC_EMB + 'embed_mlp/linear_0:w') + g(PN.NUC_EMB + 'embed_mlp/linear_0:b'))
    h = h @ g(PN.NUC_EMB + 'embed_mlp/linear_1:w') + g(PN.NUC_EMB + 'embed_mlp/linear_1:b')
    total = torch.zeros((), dtype=torch.float64)
    as_t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    for l in range(spec.n_layers):
        a = PN.comb_prefix(l)
        q = (h @ g(a + 'multi_head_attention/query:w')).reshape(M, H, dh)
        k = h @ g(a + 'multi_head_attention/key:w')
        v = h @ g(a + 'multi_head_attention/value:w')
        total = total + (k * as_t(cot['kn'][l])).sum() + (v * as_t(cot['vn'][l])).sum()
        logits = torch.einsum('thd,Thd->htT', q, k.reshape(M, H, dh)) / np.sqrt(dh)
        w = torch.softmax(logits, -1)
        o = torch.einsum('htT,Thd->thd', w, v.reshape(M, H, dh)).reshape(M, d)
        att = h + o @ g(a + 'multi_head_attention/linear:w')
        m = torch.tanh(att @ g(a + 'mlp/linear_0:w') + g(a + 'mlp/linear_0:b'))
        m = torch.tanh(m @ g(a + 'mlp/linear_1:w') + g(a + 'mlp/linear_1:b'))
        h = att + m
    mu = h.mean(-1, keepdim=True)
    xn = (h - mu) / torch.sqrt(((h - mu) ** 2).mean(-1, keepdim=True) + 1e-5)
    K, E = spec.n_determinants, spec.n_env_per_nuc
    for glu, spin in (('zetas_readout_glu', 'up'), ('zetas_readout_glu_1', 'down')):
        y = torch.sigmoid(xn @ g(PN.HEAD + glu + '/W:w') + g(PN.HEAD + glu + '/W:b')) * (
            xn @ g(PN.HEAD + glu + '/V:w') + g(PN.HEAD + glu + '/V:b'))
        z = y.reshape(M, K, E) + g(PN.HEAD + f':zetas_bias_{spin}')
        total = total + (z * as_t(cot[f'zetas_{spin}'])).sum()
    total.backward()
    return {k: v.grad for k, v in pt.items() if v.grad is not None}
