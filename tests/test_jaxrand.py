"""Known answers for the two independent numpy restatements of jax.random -- oracle/jaxrand.py (test infrastructure: also haiku's
initialisers, used to regenerate the reference's fixtures) and deepqmc_b200/jaxrand.py (product: the JAX-compatible walker
initialiser) -- and their agreement with each other; neither JAX nor haiku is installed here.  Values: the Threefry-2x32 known-answer vectors of
the Random123 distribution (also used by JAX's own test-suite) and the outputs of jax.random for PRNGKey(0) as printed
in the JAX documentation before and after the key-layout change of JAX 0.5."""
import numpy as np
import pytest

from deepqmc_b200 import jaxrand as PJ
from oracle import jaxrand as OJ

J = OJ


def _tf(key, ctr, J=OJ):
    a, b = J.threefry2x32(np.array(key, dtype=np.uint32), np.array([ctr[0]], dtype=np.uint32), np.array([ctr[1]], dtype=np.uint32))
    return int(a[0]), int(b[0])


@pytest.mark.parametrize('J', [OJ, PJ], ids=['oracle', 'product'])
def test_threefry2x32_known_answers(J):
    assert _tf((0, 0), (0, 0), J) == (0x6B200159, 0x99BA4EFE)
    assert _tf((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), J) == (0x1CB996FC, 0xBB002BE7)
    assert _tf((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), J) == (0xC4923A9C, 0x483DF7A0)


def test_oracle_and_product_streams_agree():
    """two independent implementations (64-bit masked arithmetic vs wrapping uint32 arrays), both key layouts, both widths"""
    for seed in (0, 42, 2**40 + 17):
        k = OJ.prng_key(seed)
        assert np.array_equal(k, PJ.prng_key(seed))
        for part in (True, False):
            assert np.array_equal(OJ.split(k, 5, part), PJ.split(k, 5, part))
            for dt in (np.float64, np.float32):
                assert np.array_equal(OJ.normal(k, (7, 3), dt, part), PJ.normal(k, (7, 3), dt, part))
                assert np.array_equal(OJ.uniform(k, (5,), dt, 0.0, 1.0, part), PJ.uniform(k, (5,), dt, 0.0, 1.0, part))
                assert np.array_equal(OJ.truncated_normal(k, -2.0, 2.0, (9,), dt, part), PJ.truncated_normal(k, -2.0, 2.0, (9,), dt, part))


@pytest.mark.parametrize('J', [OJ, PJ], ids=['oracle', 'product'])
def test_split_and_samplers_match_documented_jax_outputs(J):
    k0 = J.prng_key(0)
    assert k0.tolist() == [0, 0]
    assert J.split(k0, 2, partitionable=True).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]  # JAX >= 0.5
    assert J.split(k0, 2, partitionable=False).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]  # JAX < 0.5
    assert abs(float(J.normal(k0, (), np.float32, partitionable=True)) - 1.6226422) < 1e-6
    assert abs(float(J.normal(k0, (1,), np.float32, partitionable=False)[0]) - (-0.20584226)) < 1e-7
    assert abs(float(J.uniform(k0, (1,), np.float32, partitionable=False)[0]) - 0.41845703) < 1e-7


def test_sampler_properties():
    k = J.split(J.prng_key(7), 3)[2]
    x = J.truncated_normal(k, -2.0, 2.0, (4000,), np.float64)
    assert x.min() > -2 and x.max() < 2 and abs(x.mean()) < 0.05 and abs(x.std() - 0.87962566103423978) < 0.03
    n = J.normal(k, (4000,), np.float64)
    assert abs(n.mean()) < 0.06 and abs(n.std() - 1) < 0.04
    u = J.uniform(k, (1000,), np.float32)
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1
    seq = J.PRNGSequence(J.prng_key(0))
    a, b = seq.next(), seq.next()
    assert a.tolist() == [928981903, 3453687069] and seq.count == 2 and a.tolist() != b.tolist()
