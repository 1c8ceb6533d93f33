"""CPU restatement (torch.float64) of the reference's local-energy hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepqmc_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / CPU baseline.

Parity pinning: the reference (JAX/Haiku) cannot be imported in the build container (jax,
haiku, folx, pyscf absent).  The oracle is pinned against every *param-free*
golden the reference's tests hold (tests/golden/reference_goldens.json, extracted by
tools/extract_reference_goldens.py): geometry, electron counts, LiH walker, Coulomb terms,
E_loc assembly identity.  Goldens that need Haiku-initialised parameters (psi, Laplacian,
E_loc of the test ansatz) cannot be reproduced here: for those the oracle is "parity
unpinned" and says so in DESIGN.md.  Independent internal checks: Laplacian via Hessian
trace vs jvp-of-grad loop (reference: src/deepqmc/physics.py:144-156).
"""
