"""Parameter trees of the supported ansatzes: names, shapes, initialisation, packing.

Names follow the Haiku module paths the reference produces (visible for the test ansatz in
the reference's tests/test_wf/test_grad_psi.npz) so that parameters dumped from a reference
run with ``tests/conftest.py:39-52 flatten_pytree`` (':'-joined) load unchanged:
``<module path>:<param>``.  Initialisers restate the reference's choices
(src/deepqmc/hkext.py:66-80 'ferminet' init = VarianceScaling(1, fan_in, normal) weights and
VarianceScaling(1, fan_out, normal) biases; conf/ansatz/psiformer.yaml:10 envelopes to ones;
gnn/update_features.py:276 attention weights VarianceScaling(1, fan_in, normal)).
JAX's threefry stream cannot be reproduced, so values differ from a reference run with the
same seed; distributions match.
"""
from __future__ import annotations

import numpy as np

from .spec import AnsatzSpec

P = 'neural_network_wave_function/~/'
ENV = P + 'exponential_envelopes'
CUSP = P + 'electronic_cusp_asymptotic'
GNN = P + 'omni_net/~/electron_gnn/~/'
BF_UP = P + 'omni_net/~/Backflow/~/mlp/linear_0'
BF_DN = P + 'omni_net/~/Backflow_1/~/mlp/linear_0'


NUC_EMB = GNN + 'nuclei_embedding/'
HEAD = P + 'omni_net/~/nuclear_gnn_head/'


def comb_prefix(l):
    return layer_prefix(l) + 'combined_node_attention_update_feature/'


def layer_prefix(l):
    return GNN + ('electron_gnn_layer' if l == 0 else f'electron_gnn_layer_{l}') + '/~/'


def attn_prefix(l):
    return layer_prefix(l) + 'node_attention_electron_update_feature/'


def param_shapes(spec: AnsatzSpec) -> dict[str, tuple[int, ...]]:
    N, M, d, K = spec.n_elec, spec.n_nuc, spec.embedding_dim, spec.n_determinants
    s: dict[str, tuple[int, ...]] = {}
    if spec.kind != 'transpsiformer':
        for nm in ('pi_up', 'pi_down', 'zetas_up', 'zetas_down'):
            s[f'{ENV}:{nm}'] = (K * N, M)
    if spec.cusp == 'psiformer':
        s[f'{CUSP}:same_alpha'] = ()
        s[f'{CUSP}:anti_alpha'] = ()
    if spec.kind == 'psiformer':
        s[GNN + 'electron_embedding/linear:w'] = (spec.n_feat_in, d)
        for l in range(spec.n_layers):
            a = attn_prefix(l)
            for nm in ('query', 'key', 'value', 'linear'):
                s[a + f'multi_head_attention/{nm}:w'] = (d, d)
            for i in range(2):
                s[a + f'mlp/linear_{i}:w'] = (d, d)
                s[a + f'mlp/linear_{i}:b'] = (d,)
    elif spec.kind == 'ferminet':
        de = spec.edge_dim
        d_in, e_in = spec.n_feat_in, 4
        for l in range(spec.n_layers):
            lp = layer_prefix(l)
            s[lp + 'g/linear_0:w'] = (3 * d_in + 2 * e_in, d)
            s[lp + 'g/linear_0:b'] = (d,)
            if l < spec.n_layers - 1:
                s[lp + 'u/linear_0:w'] = (e_in, de)
                s[lp + 'u/linear_0:b'] = (de,)
            d_in, e_in = d, de
    elif spec.kind == 'transpsiformer':
        # reference: conf/ansatz/transpsiformer.yaml; gnn/electron_gnn.py:435-537 NucleiEmbedding,
        # gnn/update_features.py:385-451 CombinedNodeAttentionUpdateFeature, wf/omni.py:181-211 head
        de, E = spec.nuc_edge_dim, spec.n_env_per_nuc
        s[GNN + 'electron_embedding/linear:w'] = (spec.n_feat_in, d)
        for nm, (i, o) in (('edge_mlp/linear_0', (4 + M, de)), ('edge_mlp/linear_1', (de, de)),
                           ('embed_mlp/linear_0', (de, d)), ('embed_mlp/linear_1', (d, d))):
            s[NUC_EMB + nm + ':w'] = (i, o)
            s[NUC_EMB + nm + ':b'] = (o,)
        for l in range(spec.n_layers):
            a = comb_prefix(l)
            for nm in ('query', 'key', 'value', 'linear'):
                s[a + f'multi_head_attention/{nm}:w'] = (d, d)
            for i in range(2):
                s[a + f'mlp/linear_{i}:w'] = (d, d)
                s[a + f'mlp/linear_{i}:b'] = (d,)
        for glu, spin in (('zetas_readout_glu', 'up'), ('zetas_readout_glu_1', 'down')):
            for lin in ('W', 'V'):
                s[HEAD + f'{glu}/{lin}:w'] = (d, K * E)
                s[HEAD + f'{glu}/{lin}:b'] = (K * E,)
            s[HEAD + f':zetas_bias_{spin}'] = (M, K, E)
    else:
        raise ValueError(spec.kind)
    s[BF_UP + ':w'] = (d, K * N)
    s[BF_DN + ':w'] = (d, K * N)
    return s


def init_params(spec: AnsatzSpec, seed: int = 0) -> dict[str, np.ndarray]:
    """Ansatz.init equivalent (reference: src/deepqmc/types.py:119-131)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in param_shapes(spec).items():
        leaf = name.rsplit(':', 1)[1]
        if name.startswith(ENV):
            v = np.ones(shape)
        elif name.startswith(CUSP):
            v = np.ones(shape)
        elif leaf.startswith('zetas_bias'):
            v = 2 * np.ones(shape)  # wf/omni.py:199-203
        elif leaf == 'w':
            v = rng.standard_normal(shape) / np.sqrt(shape[0])
        elif leaf == 'b':
            v = rng.standard_normal(shape) / np.sqrt(shape[0])
        else:
            raise AssertionError(name)
        out[name] = np.asarray(v, dtype=np.float64)
    return out


def perturb_params(params, seed=1, scale=0.2):
    """Randomise the parameters that initialise to constants (envelopes, cusp alphas) so
    parity tests exercise them; used by tests and the benchmark's synthetic weights."""
    rng = np.random.default_rng(seed)
    out = dict(params)
    for k, v in params.items():
        if k.startswith(ENV) or k.startswith(CUSP) or 'zetas_bias' in k:
            out[k] = v * (1 + scale * rng.uniform(-1, 1, size=v.shape))
    return out


def n_params(spec):
    return int(sum(int(np.prod(s)) for s in param_shapes(spec).values()))
