"""Numpy restatement of the pieces of jax.random / dm-haiku initialisation that the reference's own tests use to
create parameters (TEST INFRASTRUCTURE, see oracle/__init__): with it the goldens of /root/reference/tests that depend
on ``hk.transform(...).init(jax.random.PRNGKey(0), ...)`` parameter VALUES can be regenerated without JAX, which pins
the network trunk of the oracle against the reference's recorded outputs.

Third-party algorithms restated (neither library is in this image):
  * Threefry-2x32, 20 rounds (Salmon et al., SC'11) -- jax's default PRNG implementation;
  * jax.random.split / bits / uniform / normal / truncated_normal with ``jax_threefry_partitionable = True`` (the default
    since JAX 0.5; the reference's CHANGELOG 1.3.0 "Compatibility with the PRNG key changes in JAX v0.5.0") and, for
    cross-checking, the older counter layout;
  * haiku's PRNGSequence (one ``split(key, 2)`` per requested key) and the TruncatedNormal / VarianceScaling initialisers.
Known answers used to check this file are in tests/test_jaxrand.py.
"""
from __future__ import annotations

import numpy as np

from deepqmc_b200.jaxrand import *  # noqa: F401,F403  (Threefry, jax.random samplers, electron initialiser: shared with the product)
from deepqmc_b200.jaxrand import U32, prng_key, split, normal, truncated_normal  # noqa: F401


class PRNGSequence:
    """haiku.PRNGSequence: every requested key costs one split(key, 2) of the running key."""

    def __init__(self, key, partitionable=True):
        self.key, self.partitionable, self.count = np.asarray(key, dtype=U32), partitionable, 0

    def next(self):
        new = split(self.key, 2, self.partitionable)
        self.key = new[0]
        self.count += 1
        return new[1]


def hk_truncated_normal(seq: PRNGSequence, shape, stddev, dtype=np.float64):
    """hk.initializers.TruncatedNormal(stddev)(shape, dtype)"""
    return np.dtype(dtype).type(stddev) * truncated_normal(seq.next(), -2.0, 2.0, shape, dtype, seq.partitionable)


def hk_variance_scaling_normal(seq: PRNGSequence, shape, scale=1.0, fan='fan_in', dtype=np.float64):
    """hk.initializers.VarianceScaling(scale, fan, 'normal')(shape, dtype) for 2-d (fan_in, fan_out) weights and 1-d biases
    (haiku computes the fans of a 1-d shape as fan_in = fan_out = shape[0])."""
    fan_in, fan_out = (shape[0], shape[0]) if len(shape) == 1 else (shape[-2], shape[-1])
    n = {'fan_in': fan_in, 'fan_out': fan_out, 'fan_avg': (fan_in + fan_out) / 2}[fan]
    return np.dtype(dtype).type(np.sqrt(scale / max(1.0, n))) * normal(seq.next(), shape, dtype, seq.partitionable)


def haiku_init_conv_gnn_ansatz(spec, seed: int = 0, partitionable: bool = True, gnn_only: bool = False, g_layers: int = 1):
    """Parameters of the reference's conv-GNN test ansatz exactly as ``hk.transform(...).init(jax.random.PRNGKey(seed), ...)``
    creates them under ``jax_enable_x64`` (tests/conftest.py:20,121-136 of the reference), in haiku's creation order:

      1. envelope ``pi`` in the constructor of ExponentialEnvelopes (wf/env.py:37-42,75-81): ones + VarianceScaling(1.0) drawn
         in float32 (hk.get_parameter's default dtype); ``zetas`` = z / (k + 1), no random numbers;
      2. on the first call (wf/nn_wave_function.py:127-133 -> wf/omni.py:157-178 -> gnn/electron_gnn.py:374-432):
         nuclear hk.Embed table, electron hk.Embed table (float32, TruncatedNormal(1));
      3. per layer, per edge type in the yaml's order same, anti, ne: the filter MLP w_t then the node MLP h_t
         (gnn/update_features.py:196-209), then g_conv_same / anti / ne (electron_gnn.py:243-259);
      4. Jastrow MLP, Backflow (up) MLP, Backflow_1 (down) MLP (wf/omni.py:168-177); conf_coeff = ones.
    hkext.MLP(init='default'): w ~ VarianceScaling(1, fan_in, truncated_normal), b = 0 (hkext.py:63-78).  A hk.Linear
    takes the dtype of its input, so the h MLPs of the FIRST layer (input: float32 embeddings) draw float32 numbers.
    ``spec``: deepqmc_b200.spec.paulinet_spec(...) ('featurewise' update, hk.Embed embeddings, no deep edge features).
    ``gnn_only`` / ``g_layers``: the bare ElectronGNN of tests/conf/gnn.yaml (tests/test_gnn.py TestGNN.test_embedding):
    steps 2 and 3 only, g_t MLPs with ['log', g_layers] layers.
    Returns {haiku path: float64 array}.
    """
    from deepqmc_b200 import params as PN

    assert spec.kind == 'paulinet' and spec.gnn_update == 'featurewise' and spec.gnn_embedding == 'embed' and not spec.gnn_deep_edges
    seq = PRNGSequence(prng_key(seed), partitionable)
    K, N, M, d, e = spec.n_determinants, spec.n_elec, spec.n_nuc, spec.embedding_dim, spec.edge_dim

    def vs(shape, dtype):  # VarianceScaling(1.0, 'fan_in', 'truncated_normal')
        return hk_truncated_normal(seq, shape, np.sqrt(1.0 / max(1.0, shape[0])) / .87962566103423978, dtype).astype(np.float64)

    out = {}
    if not gnn_only:
        n_env = len(spec.env_centers)
        out[f'{PN.ENV}:pi'] = 1.0 + vs((K * N, n_env), np.float32)
        out[f'{PN.ENV}:zetas'] = np.asarray(spec.env_zeta_init, dtype=np.float64)
    types = PN.EDGE_TYPES if spec.gnn_conv_ne else PN.EDGE_TYPES[:2]
    if spec.gnn_conv_ne:
        out[PN.GNN + 'nuclei_embedding/~/embed:embeddings'] = hk_truncated_normal(seq, (M, d), 1.0, np.float32).astype(np.float64)
    n_types = 1 if spec.n_up == spec.n_down else 2
    out[PN.GNN + 'electron_embedding/ElectronicEmbedding:embeddings'] = hk_truncated_normal(seq, (n_types, d), 1.0, np.float32).astype(np.float64)
    nl = spec.gnn_subnet_layers
    for l in range(spec.n_layers):
        c, lp = PN.conv_prefix(l), PN.layer_prefix(l)
        x_dtype = np.float32 if l == 0 else np.float64  # electron embeddings are float64 after the first residual update
        for t in types:
            dw = [4] + PN.log_dims(4, e, nl)
            for i in range(nl):
                out[c + f'w_{t}/linear_{i}:w'] = vs((dw[i], dw[i + 1]), np.float64)
            dh = [d] + PN.log_dims(d, e, nl)
            h_dtype = np.float32 if t == 'ne' else x_dtype  # nuclear embeddings are never updated: float32 in every layer
            for i in range(nl):
                out[c + f'h_{t}/linear_{i}:w'] = vs((dh[i], dh[i + 1]), h_dtype)
                out[c + f'h_{t}/linear_{i}:b'] = np.zeros(dh[i + 1])
        for t in types:
            dg = [e] + PN.log_dims(e, d, g_layers)
            for i in range(g_layers):
                out[lp + f'g_conv_{t}/linear_{i}:w'] = vs((dg[i], dg[i + 1]), np.float64)
                out[lp + f'g_conv_{t}/linear_{i}:b'] = np.zeros(dg[i + 1])
    if gnn_only:
        return out
    dj = [d] + PN.log_dims(d, 1, spec.jastrow_layers) if spec.jastrow_layers else []
    for i in range(len(dj) - 1):
        out[PN.JASTROW + f'linear_{i}:w'] = vs((dj[i], dj[i + 1]), np.float64)
        if i < len(dj) - 2:
            out[PN.JASTROW + f'linear_{i}:b'] = np.zeros(dj[i + 1])
    for pre, n_spin in ((PN.BF_UP, spec.n_up), (PN.BF_DN, spec.n_down)):
        db = [d] + PN.backflow_dims(spec, n_spin)
        base = pre.rsplit('linear_0', 1)[0]
        for i in range(len(db) - 1):
            out[base + f'linear_{i}:w'] = vs((db[i], db[i + 1]), np.float64)
            if spec.backflow_bias:
                out[base + f'linear_{i}:b'] = np.zeros(db[i + 1])
    if spec.conf_coeff == 'linear':
        out[PN.CONF + ':w'] = np.ones((K, 1))
    return out


