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
from deepqmc_b200.jaxrand import PRNGSequence, haiku_init_conv_gnn_ansatz, hk_truncated_normal, hk_variance_scaling_normal  # noqa: F401
