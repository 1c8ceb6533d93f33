"""jax.random / dm-haiku initialisation restated in numpy -- re-exported from deepqmc_b200/jaxrand.py, where the code lives
because the product offers seed-compatible walkers, noise streams and (for the conv-GNN test ansatz) parameters with it.
The tests use it through this name to regenerate the reference's ``hk.transform(...).init(jax.random.PRNGKey(0), ...)``
parameters and random streams without JAX, which pins the oracle against the reference's recorded fixtures
(tests/test_oracle_goldens.py, tests/test_reference_fixtures.py).  Known answers for the generator itself: tests/test_jaxrand.py.
"""
from __future__ import annotations

import numpy as np

from deepqmc_b200.jaxrand import *  # noqa: F401,F403  (Threefry, jax.random samplers, electron initialiser: shared with the product)
from deepqmc_b200.jaxrand import U32, prng_key, split, normal, truncated_normal  # noqa: F401
from deepqmc_b200.jaxrand import PRNGSequence, haiku_init_conv_gnn_ansatz, hk_truncated_normal, hk_variance_scaling_normal  # noqa: F401
