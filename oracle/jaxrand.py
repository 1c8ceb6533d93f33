"""jax.random streams and dm-haiku parameter initialisation restated in numpy -- the ORACLE's own copy (test infrastructure).

Why it exists: the parameter-dependent fixtures of the reference's tests (tests/test_wf/*.npz, tests/test_hamil/test_local_energy_*.npz,
tests/test_gnn/test_embedding.npz, tests/test_sampling/*.npz) were recorded with parameters from
``hk.transform(...).init(jax.random.PRNGKey(0), phys_conf)`` and noise from ``jax.random``; JAX and haiku are not installable here, so
the streams are restated: Threefry-2x32 (Salmon et al., SC'11; 20 rounds, the Random123 key schedule) with JAX's key layout,
``split`` / ``uniform`` / ``normal`` / ``truncated_normal`` as jax/_src/random.py derives them from the raw bits, haiku's
``PRNGSequence`` and initialisers, and the parameter-creation order of the reference's conv-GNN test ansatz.  Known answers:
tests/test_jaxrand.py (Random123 vectors, values printed in the JAX documentation).

Deliberately independent of deepqmc_b200/jaxrand.py (the product's restatement, which drives its JAX-compatible walker
initialiser): the block function below works on 64-bit integers with explicit masks where the product's wraps uint32 arrays, and
tests/test_jaxrand.py checks the two against each other as well as against the known answers.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf, erfinv

from . import names as HN

U32 = np.uint32
_M32 = np.uint64(0xFFFFFFFF)
_ROTATIONS = (13, 15, 26, 6, 17, 29, 16, 24)   # Threefry-2x32 rotation constants, rounds 0-3 / 4-7, repeated
_PARITY = 0x1BD11BDA                            # key-schedule parity constant (Skein)


def threefry2x32(key, c0, c1):
    """Threefry-2x32-20: key [..., 2] uint32, counters c0 / c1 (uint32 arrays, broadcastable against the key) -> two uint32 arrays."""
    key = np.asarray(key, dtype=np.uint64)
    ks = [key[..., 0] & _M32, key[..., 1] & _M32]
    ks.append((ks[0] ^ ks[1] ^ np.uint64(_PARITY)) & _M32)
    x = (np.asarray(c0, dtype=np.uint64) + ks[0]) & _M32
    y = (np.asarray(c1, dtype=np.uint64) + ks[1]) & _M32
    for rnd in range(20):
        r = np.uint64(_ROTATIONS[rnd % 8])
        x = (x + y) & _M32
        y = ((y << r) | (y >> (np.uint64(32) - r))) & _M32
        y = y ^ x
        if rnd % 4 == 3:  # key injection after every fourth round
            s = rnd // 4 + 1
            x = (x + ks[s % 3]) & _M32
            y = (y + ks[(s + 1) % 3] + np.uint64(s)) & _M32
    return x.astype(U32), y.astype(U32)


def prng_key(seed: int):
    """jax.random.PRNGKey(seed): (high word, low word)"""
    return np.array([(int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF], dtype=U32)


def _counters(n):
    i = np.arange(n, dtype=np.uint64)
    return (i >> np.uint64(32)).astype(U32), (i & _M32).astype(U32)


def _bits32_pairs(key, n32):
    """the pre-0.5 layout: 32-bit counters 0 .. n32-1 (zero-padded to an even count) split into two halves"""
    pad = n32 % 2
    cnt = np.concatenate([np.arange(n32, dtype=U32), np.zeros(pad, dtype=U32)])
    h = (n32 + pad) // 2
    a, b = threefry2x32(key, cnt[:h], cnt[h:])
    return np.concatenate([a, b])[:n32]


def split(key, num=2, partitionable=True):
    """jax.random.split.  partitionable (jax_threefry_partitionable, default since JAX 0.5): key i = block(key, 64-bit counter i)."""
    if partitionable:
        a, b = threefry2x32(key, *_counters(num))
        return np.stack([a, b], -1)
    return _bits32_pairs(key, 2 * num).reshape(num, 2)


def random_bits(key, bit_width, shape, partitionable=True):
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if partitionable:
        a, b = threefry2x32(key, *_counters(n))
        out = ((a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)) if bit_width == 64 else (a ^ b).astype(U32)
        return out.reshape(shape)
    bits = _bits32_pairs(key, n * (bit_width // 32))
    out = ((bits[:n].astype(np.uint64) << np.uint64(32)) | bits[n:].astype(np.uint64)) if bit_width == 64 else bits
    return out.reshape(shape)


def uniform(key, shape, dtype=np.float64, minval=0.0, maxval=1.0, partitionable=True):
    """jax.random.uniform: mantissa bits under the exponent of 1.0, minus 1, scaled; clamped from below."""
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        m = random_bits(key, 64, shape, partitionable) >> np.uint64(12)
        f = (m | np.float64(1.0).view(np.uint64)).view(np.float64) - 1.0
    else:
        m = random_bits(key, 32, shape, partitionable) >> U32(9)
        f = (m | np.float32(1.0).view(U32)).view(np.float32) - np.float32(1.0)
    lo, hi = dtype.type(minval), dtype.type(maxval)
    return np.maximum(lo, (f * (hi - lo) + lo).astype(dtype))


def _erfinv_f32(x):
    """XLA's float32 erf_inv (Giles, 'Approximating the erfinv function', 2010: two degree-8 polynomials in w = -log(1 - x^2))."""
    f = np.float32
    x = x.astype(f)
    w = -np.log((f(1) - x) * (f(1) + x)).astype(f)
    central = w < f(5)
    t = np.where(central, w - f(2.5), np.sqrt(np.maximum(w, 0)).astype(f) - f(3)).astype(f)
    c_central = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164,
                 0.246640727, 1.50140941)
    c_tail = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
              1.00167406, 2.83297682)
    acc = np.where(central, f(c_central[0]), f(c_tail[0])).astype(f)
    for a, b in zip(c_central[1:], c_tail[1:]):
        acc = (np.where(central, f(a), f(b)) + acc * t).astype(f)
    return (acc * x).astype(f)


def _inv_erf(u):
    return _erfinv_f32(u) if u.dtype == np.float32 else erfinv(u)


def normal(key, shape, dtype=np.float64, partitionable=True):
    dtype = np.dtype(dtype)
    u = uniform(key, shape, dtype, np.nextafter(dtype.type(-1), dtype.type(0)), 1.0, partitionable)
    return (dtype.type(np.sqrt(2)) * _inv_erf(u)).astype(dtype)


def truncated_normal(key, lower, upper, shape, dtype=np.float64, partitionable=True):
    dtype = np.dtype(dtype)
    r2 = dtype.type(np.sqrt(2))
    a, b = dtype.type(erf(dtype.type(lower) / r2)), dtype.type(erf(dtype.type(upper) / r2))
    z = (r2 * _inv_erf(uniform(key, shape, dtype, a, b, partitionable))).astype(dtype)
    return np.clip(z, np.nextafter(dtype.type(lower), dtype.type(np.inf)), np.nextafter(dtype.type(upper), dtype.type(-np.inf)))


# ---- dm-haiku parameter initialisation on top of these streams --------------------------------------------------------------
class PRNGSequence:
    """haiku.PRNGSequence: every requested key costs one split(key, 2) of the running key."""

    def __init__(self, key, partitionable=True):
        self.key, self.partitionable, self.count = np.asarray(key, dtype=U32), partitionable, 0

    def next(self):
        pair = split(self.key, 2, self.partitionable)
        self.key = pair[0]
        self.count += 1
        return pair[1]


def hk_truncated_normal(seq: PRNGSequence, shape, stddev, dtype=np.float64):
    """hk.initializers.TruncatedNormal(stddev)(shape, dtype)"""
    return np.dtype(dtype).type(stddev) * truncated_normal(seq.next(), -2.0, 2.0, shape, dtype, seq.partitionable)


def hk_variance_scaling_normal(seq: PRNGSequence, shape, scale=1.0, fan='fan_in', dtype=np.float64):
    """hk.initializers.VarianceScaling(scale, fan, 'normal')(shape, dtype) for 2-d (fan_in, fan_out) weights and 1-d biases
    (haiku computes the fans of a 1-d shape as fan_in = fan_out = shape[0])."""
    fan_in, fan_out = (shape[0], shape[0]) if len(shape) == 1 else (shape[-2], shape[-1])
    n = {'fan_in': fan_in, 'fan_out': fan_out, 'fan_avg': (fan_in + fan_out) / 2}[fan]
    return np.dtype(dtype).type(np.sqrt(scale / max(1.0, n))) * normal(seq.next(), shape, dtype, seq.partitionable)


def log_widths(d_in, d_out, n):
    """hidden_layers ['log', n] of the reference's MLP (hkext.py:95-99): n widths interpolated geometrically, ending in d_out."""
    return [round(d_in ** (1 - k / n) * d_out ** (k / n)) for k in range(1, n + 1)]


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
    ``spec``: the hyper-parameter record of tests/conf/ansatz.yaml ('featurewise' update, hk.Embed embeddings, no deep edge
    features; any object with the attributes read below).
    ``gnn_only`` / ``g_layers``: the bare ElectronGNN of tests/conf/gnn.yaml (tests/test_gnn.py TestGNN.test_embedding):
    steps 2 and 3 only, g_t MLPs with ['log', g_layers] layers.
    Returns {haiku path: float64 array}.
    """
    assert spec.kind == 'paulinet' and spec.gnn_update == 'featurewise' and spec.gnn_embedding == 'embed' and not spec.gnn_deep_edges
    seq = PRNGSequence(prng_key(seed), partitionable)
    K, N, M, d, e = spec.n_determinants, spec.n_elec, spec.n_nuc, spec.embedding_dim, spec.edge_dim

    def vs(shape, dtype):  # VarianceScaling(1.0, 'fan_in', 'truncated_normal'): the std of a unit normal cut at +-2 is 0.8796...
        return hk_truncated_normal(seq, shape, np.sqrt(1.0 / max(1.0, shape[0])) / .87962566103423978, dtype).astype(np.float64)

    out = {}
    if not gnn_only:
        out[f'{HN.ENV}:pi'] = 1.0 + vs((K * N, len(spec.env_centers)), np.float32)
        out[f'{HN.ENV}:zetas'] = np.asarray(spec.env_zeta_init, dtype=np.float64)
    types = ('same', 'anti', 'ne') if spec.gnn_conv_ne else ('same', 'anti')
    if spec.gnn_conv_ne:
        out[HN.GNN + 'nuclei_embedding/~/embed:embeddings'] = hk_truncated_normal(seq, (M, d), 1.0, np.float32).astype(np.float64)
    n_types = 1 if spec.n_up == spec.n_down else 2
    out[HN.GNN + 'electron_embedding/ElectronicEmbedding:embeddings'] = hk_truncated_normal(seq, (n_types, d), 1.0, np.float32).astype(np.float64)
    nl = spec.gnn_subnet_layers
    for l in range(spec.n_layers):
        c, lp = HN.conv_prefix(l), HN.layer_prefix(l)
        x_dtype = np.float32 if l == 0 else np.float64  # electron embeddings are float64 after the first residual update
        for t in types:
            dw = [4] + log_widths(4, e, nl)
            for i in range(nl):
                out[c + f'w_{t}/linear_{i}:w'] = vs((dw[i], dw[i + 1]), np.float64)
            dh = [d] + log_widths(d, e, nl)
            h_dtype = np.float32 if t == 'ne' else x_dtype  # nuclear embeddings are never updated: float32 in every layer
            for i in range(nl):
                out[c + f'h_{t}/linear_{i}:w'] = vs((dh[i], dh[i + 1]), h_dtype)
                out[c + f'h_{t}/linear_{i}:b'] = np.zeros(dh[i + 1])
        for t in types:
            dg = [e] + log_widths(e, d, g_layers)
            for i in range(g_layers):
                out[lp + f'g_conv_{t}/linear_{i}:w'] = vs((dg[i], dg[i + 1]), np.float64)
                out[lp + f'g_conv_{t}/linear_{i}:b'] = np.zeros(dg[i + 1])
    if gnn_only:
        return out
    dj = [d] + log_widths(d, 1, spec.jastrow_layers) if spec.jastrow_layers else []
    for i in range(len(dj) - 1):
        out[HN.JASTROW + f'linear_{i}:w'] = vs((dj[i], dj[i + 1]), np.float64)
        if i < len(dj) - 2:
            out[HN.JASTROW + f'linear_{i}:b'] = np.zeros(dj[i + 1])
    for pre, n_spin in ((HN.BF_UP, spec.n_up), (HN.BF_DN, spec.n_down)):
        n_orb = spec.n_elec if spec.full_determinant else n_spin
        db = [d] + log_widths(d, K * n_orb, spec.backflow_layers)
        base = pre.rsplit('linear_0', 1)[0]
        for i in range(len(db) - 1):
            out[base + f'linear_{i}:w'] = vs((db[i], db[i + 1]), np.float64)
            if spec.backflow_bias:
                out[base + f'linear_{i}:b'] = np.zeros(db[i + 1])
    if spec.conf_coeff == 'linear':
        out[HN.CONF + ':w'] = np.ones((K, 1))
    return out
