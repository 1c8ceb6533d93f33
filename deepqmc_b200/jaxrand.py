"""Numpy restatement of the jax.random streams the reference draws its walkers from (Threefry-2x32, JAX >= 0.5 key layout).

Why this is in the product: the reference initialises its walkers with ``AtomCenteredElectronInitializer`` driven by
``jax.random`` (sampling/electron_sample_initializers.py:43-288).  With this module the samplers of this package can be
seeded so that ``sampler.init(seed, ...)`` returns the SAME walkers as the reference's ``sampler.init(PRNGKey(seed), ...)``
(``deepqmc_b200.sampling.JaxCompatibleElectronInitializer``) -- checked bit-for-bit against the reference's recorded walkers
(tests/test_reference_fixtures.py).  Host-side, one-off work; no wave-function arithmetic.
(The restatement of haiku's parameter initialisation, which only the tests need to regenerate the reference's fixtures, lives
with the oracle: oracle/jaxrand.py -- an independent implementation of the same streams, cross-checked in tests/test_jaxrand.py.)
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf, erfinv

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, d):
    return ((x << U32(d)) | (x >> U32(32 - d))).astype(U32)


def threefry2x32(key, c0, c1):
    """key: 2 uint32; c0, c1: uint32 arrays of equal shape -> two uint32 arrays."""
    with np.errstate(over='ignore'):
        key = np.asarray(key, dtype=U32)
        k0, k1 = key[..., 0], key[..., 1]  # a single key or an array of keys [..., 2] broadcasting against the counters
        ks = (k0, k1, (k0 ^ k1 ^ U32(0x1BD11BDA)).astype(U32))
        x0 = (np.asarray(c0, dtype=U32) + ks[0]).astype(U32)
        x1 = (np.asarray(c1, dtype=U32) + ks[1]).astype(U32)
        for r in range(5):
            for d in _ROT[r % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, d)
                x1 = (x1 ^ x0).astype(U32)
            x0 = (x0 + ks[(r + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(r + 2) % 3] + U32(r + 1)).astype(U32)
    return x0, x1


def prng_key(seed: int):
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _iota_2x32(n):
    i = np.arange(n, dtype=np.uint64)
    return (i >> np.uint64(32)).astype(U32), (i & np.uint64(0xFFFFFFFF)).astype(U32)


def split(key, num=2, partitionable=True):
    if partitionable:  # key_i = threefry(key, 64-bit counter i as (hi, lo))
        b1, b2 = threefry2x32(key, *_iota_2x32(num))
        return np.stack([b1, b2], -1)
    cnt = np.arange(2 * num, dtype=U32)  # original layout: counters split in two halves
    o0, o1 = threefry2x32(key, cnt[:num], cnt[num:])
    return np.concatenate([o0, o1]).reshape(num, 2)


def random_bits(key, bit_width, shape, partitionable=True):
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if partitionable:
        b1, b2 = threefry2x32(key, *_iota_2x32(size))
        if bit_width == 64:
            out = (b1.astype(np.uint64) << np.uint64(32)) | b2.astype(np.uint64)
        else:
            out = (b1 ^ b2).astype(U32)
        return out.reshape(shape)
    n32 = size * (bit_width // 32)
    pad = n32 % 2
    cnt = np.concatenate([np.arange(n32, dtype=U32), np.zeros(pad, dtype=U32)])  # odd counts are padded with a 0 counter
    half = (n32 + pad) // 2
    o0, o1 = threefry2x32(key, cnt[:half], cnt[half:])
    bits = np.concatenate([o0, o1])[:n32]
    if bit_width == 64:
        out = (bits[:size].astype(np.uint64) << np.uint64(32)) | bits[size:].astype(np.uint64)
    else:
        out = bits
    return out.reshape(shape)


def uniform(key, shape, dtype=np.float64, minval=0.0, maxval=1.0, partitionable=True):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        bits = random_bits(key, 64, shape, partitionable)
        fb = (bits >> np.uint64(64 - 52)) | np.float64(1.0).view(np.uint64)
        f = fb.view(np.float64) - 1.0
    else:
        bits = random_bits(key, 32, shape, partitionable)
        fb = (bits >> U32(32 - 23)) | np.float32(1.0).view(U32)
        f = fb.view(np.float32) - np.float32(1.0)
    minval, maxval = dtype.type(minval), dtype.type(maxval)
    return np.maximum(minval, (f * (maxval - minval) + minval).astype(dtype))


def _erfinv32(x):
    """XLA's single-precision erf_inv (Giles 2010 polynomial), evaluated in float32."""
    x = x.astype(np.float32)
    w = -np.log((np.float32(1) - x) * (np.float32(1) + x)).astype(np.float32)
    lt = w < np.float32(5)
    wa = np.where(lt, w - np.float32(2.5), np.sqrt(np.maximum(w, 0)).astype(np.float32) - np.float32(3)).astype(np.float32)
    ca = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164,
          0.246640727, 1.50140941]
    cb = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
          1.00167406, 2.83297682]
    p = np.where(lt, np.float32(ca[0]), np.float32(cb[0])).astype(np.float32)
    for a, b in zip(ca[1:], cb[1:]):
        p = (np.where(lt, np.float32(a), np.float32(b)) + p * wa).astype(np.float32)
    return (p * x).astype(np.float32)


def _erfinv(u):
    return _erfinv32(u) if u.dtype == np.float32 else erfinv(u)


def normal(key, shape, dtype=np.float64, partitionable=True):
    dtype = np.dtype(dtype)
    lo = np.nextafter(dtype.type(-1), dtype.type(0))
    u = uniform(key, shape, dtype, lo, 1.0, partitionable)
    return (dtype.type(np.sqrt(2)) * _erfinv(u)).astype(dtype)


def truncated_normal(key, lower, upper, shape, dtype=np.float64, partitionable=True):
    dtype = np.dtype(dtype)
    s2 = dtype.type(np.sqrt(2))
    a, b = dtype.type(erf(dtype.type(lower) / s2)), dtype.type(erf(dtype.type(upper) / s2))
    u = uniform(key, shape, dtype, a, b, partitionable)
    out = (s2 * _erfinv(u)).astype(dtype)
    return np.clip(out, np.nextafter(dtype.type(lower), dtype.type(np.inf)), np.nextafter(dtype.type(upper), dtype.type(-np.inf)))


# ---- electron initialiser of the reference (sampling/electron_sample_initializers.py) with jax.random streams ----------
def exponential(key, shape, partitionable=True):
    return -np.log1p(-uniform(key, shape, np.float64, 0.0, 1.0, partitionable))


def gumbel(key, shape, partitionable=True):
    u = uniform(key, shape, np.float64, np.finfo(np.float64).tiny, 1.0, partitionable)
    return -np.log(-np.log(u))


def categorical(key, logits, partitionable=True):
    logits = np.asarray(logits, dtype=np.float64)
    return int(np.argmax(gumbel(key, logits.shape, partitionable) + logits))


def orthogonal3(key, partitionable=True):
    """jax.random.orthogonal(key, 3): QR of a Gaussian matrix, columns signed so that diag(R) > 0 (Haar measure)."""
    z = normal(key, (3, 3), np.float64, partitionable)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))[None, :]


def _shell_positions(key, charges, counts, partitionable=True):
    """ShellBasedDistribution.__call__ (:198-251): |r| ~ Exp / (2 zeta), zeta = Z x {1, 1/2, 1/3, 1/4} by shell."""
    total = len(charges)
    marks = np.zeros(total + 1)
    cs = np.cumsum(counts)
    for c_, n_ in zip(cs, counts):
        if c_ < total:
            marks[c_] = n_
    spin_idx = np.arange(total) - np.cumsum(marks[:total])
    factor = np.where(spin_idx < 1, 1.0, np.where(spin_idx < 5, 0.5, np.where(spin_idx < 9, 1 / 3, 0.25)))
    zetas = np.asarray(charges, dtype=np.float64) * factor
    pos = np.zeros((total, 3))
    for i, (k, z) in enumerate(zip(split(key, total, partitionable), zetas)):
        k_r, k_dir = split(k, 2, partitionable)
        pos[i] = exponential(k_r, (), partitionable) / (2 * z) * orthogonal3(k_dir, partitionable)[:, 0]
    return pos


def atom_centered_initializer(key, charges, ns_valence, R, n_up, n_down, partitionable=True):
    """AtomCenteredElectronInitializer(ShellBasedDistribution())(rng, ...) (:254-288) -> r[n_up + n_down, 3]."""
    charges, ns_valence, R = (np.asarray(a, dtype=np.float64) for a in (charges, ns_valence, R))
    M = len(charges)
    k_assign, k_spin, k_up, k_dn = split(key, 4, partitionable)
    # assign_electrons_to_nuclei (:43-81)
    charge = ns_valence.sum() - n_up - n_down
    valence = ns_valence - charge / M
    el = np.floor(valence).astype(int)
    rng = k_assign
    while ns_valence.sum() - charge - el.sum() > 0:
        rng, k_cat = split(rng, 2, partitionable)
        el[categorical(k_cat, valence - el, partitionable)] += 1
    # assign_spins_to_nuclei (:84-155)
    up, down = np.zeros(M, dtype=int), np.zeros(M, dtype=int)
    for i in range(int(el.max())):
        mask = el >= 2 * (i + 1)
        inc = np.where(mask & (mask.sum() + down.sum() <= n_down), 1, 0)
        up, down = up + inc, down + inc
    dists = np.linalg.norm(R[:, None] - R[None], axis=-1)
    np.fill_diagonal(dists, np.inf)
    nn = np.argsort(dists, axis=-1, kind='stable')
    rem = el - up - down
    center = categorical(k_spin, np.where(rem == rem.max(), 0.0, -np.inf), partitionable)
    i = 0
    while (up + down < el).any():
        is_down = (i % 2) & int(down.sum() < n_down)
        up[center] += 1 - is_down
        down[center] += is_down
        ordering = nn[center]
        has_rem = (el - up - down)[ordering] > 0
        center = ordering[int(np.argmax(has_rem))]
        i += 1
    idx = lambda counts, total: (np.cumsum(counts)[:, None] <= np.arange(total)).sum(0)
    up_idx, dn_idx = idx(up, n_up), idx(down, n_down)
    r_up = R[up_idx] + _shell_positions(k_up, charges[up_idx], up, partitionable)
    r_dn = R[dn_idx] + _shell_positions(k_dn, charges[dn_idx], down, partitionable)
    return np.concatenate([r_up, r_dn])


def fold_in(key, data: int):
    """jax.random.fold_in(key, data) for 32-bit data: threefry(key, (0, data))."""
    o0, o1 = threefry2x32(key, np.array([0], dtype=U32), np.array([data & 0xFFFFFFFF], dtype=U32))
    return np.array([o0[0], o1[0]], dtype=U32)


def ecp_quadrature_twists(key, n_nl_nuclei: int, n_elec: int, partitionable=True):
    """phi_random[j, i] of the reference's non-local ECP quadrature: uniform(fold_in(fold_in(rng, j), i), (), 0, pi / 5)
    (ecp/gaussian_type_ecp.py:224, ecp/ecp_utils.py:52)."""
    out = np.zeros((n_nl_nuclei, n_elec))
    for j in range(n_nl_nuclei):
        kj = fold_in(key, j)
        for i in range(n_elec):
            out[j, i] = uniform(fold_in(kj, i), (), np.float64, 0.0, np.pi / 5, partitionable)
    return out


def fold_in_many(keys, data):
    """fold_in for an array of keys [..., 2] and broadcastable integer data -> keys [..., 2]."""
    keys = np.asarray(keys, dtype=U32)
    data = np.asarray(data)
    shape = np.broadcast_shapes(keys.shape[:-1], data.shape)
    kb = np.broadcast_to(keys, shape + (2,))
    o0, o1 = threefry2x32(kb, np.zeros(shape, dtype=U32), np.broadcast_to(data, shape).astype(U32))
    return np.stack([o0, o1], -1)


def ecp_quadrature_twists_batch(keys, n_nl_nuclei: int, n_elec: int):
    """phi_random[b, j, i] for per-walker keys [B, 2] (the reference splits its key over the batch shape, loss/energy.py:43,
    then folds in the nucleus slot j and the electron i, gaussian_type_ecp.py:224; JAX >= 0.5 bit layout)."""
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2)
    kj = fold_in_many(keys[:, None, :], np.arange(n_nl_nuclei)[None, :])             # [B, J, 2]
    kji = fold_in_many(kj[:, :, None, :], np.arange(n_elec)[None, None, :])           # [B, J, N, 2]
    zero = np.zeros(kji.shape[:-1], dtype=U32)
    b1, b2 = threefry2x32(kji, zero, zero)                                            # random_bits(key, 64, ()) per key
    bits = (b1.astype(np.uint64) << np.uint64(32)) | b2.astype(np.uint64)
    f = ((bits >> np.uint64(12)) | np.float64(1.0).view(np.uint64)).view(np.float64) - 1.0
    return np.maximum(0.0, f * (np.pi / 5))
