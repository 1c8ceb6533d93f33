"""CPU restatement (torch.float64) of the reference's local-energy hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepqmc_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / CPU baseline.

Parity pinning: the reference (JAX/Haiku) cannot be imported in the build container (jax,
haiku, folx, pyscf absent).  The oracle is pinned against every *param-free*
golden the reference's tests hold (tests/golden/reference_goldens.json, extracted by
tools/extract_reference_goldens.py): geometry, electron counts, LiH walker, Coulomb terms,
E_loc assembly identity.  The goldens that need Haiku-initialised parameters (psi, its
parameter gradient, Laplacian / quantum force and E_loc of the reference's test ansatz,
tests/test_wf.py, tests/test_hamil.py) are PINNED as well: oracle/jaxrand.py regenerates
``hk.transform(...).init(PRNGKey(0), ...)`` in numpy (Threefry-2x32, jax.random samplers,
haiku initialisers in creation order) and the oracle reproduces all of them to ~3e-7, the
limit set by the reference's own float32 sub-computations
(tests/test_oracle_goldens.py::test_oracle_reproduces_parameter_dependent_reference_goldens).
Not pinned by a reference output: the ansatz kinds the reference's tests never evaluate
(Psiformer / FermiNet / TransPsiformer trunks: hk.MultiHeadAttention and the FermiNet
aggregation are restated from the cited source lines; everything downstream of the trunk --
envelopes, backflow, determinants, cusp, Laplacian, E_loc assembly -- is shared with the
pinned ansatz), the ccECP table and the pseudo-Hamiltonian.  Independent
internal checks: Laplacian via Hessian trace vs jvp-of-grad loop (reference:
src/deepqmc/physics.py:144-156).
"""
