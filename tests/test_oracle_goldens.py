"""CPU: pin the oracle (and the host-side Hamiltonian mirror) against every param-free golden
of the reference's own test-suite (tests/golden/reference_goldens.json)."""
import numpy as np
import pytest
import torch

from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from oracle import wf as owf
from oracle.hamil import OracleHamiltonian, pairwise_self_distance
from oracle.laplacian import laplacian_hessian, laplacian_jvp_loop


@pytest.mark.parametrize('name', ['LiH', 'C', 'H2O', 'NH3', 'H10', 'ScO', 'bicyclobutane'])
def test_molecule_geometry(goldens, name):
    # reference: tests/test_molecule.py (from_name) -- angstrom -> bohr conversion
    g = goldens['molecule'][name]
    mol = Molecule.from_name(name)
    np.testing.assert_allclose(mol.coords, np.asarray(g['coords']).reshape(-1, 3), rtol=0, atol=2e-8)
    np.testing.assert_allclose(mol.charges, g['charges'])
    assert mol.charge == g['charge'] and mol.spin == g['spin']


@pytest.mark.parametrize('cls', [OracleHamiltonian, lambda mol, ecp_type=None: MolecularHamiltonian(mol=mol, ecp_type=ecp_type)])
def test_hamil_init(goldens, cls):
    # reference: tests/test_hamil.py:16-26 (n_up, n_down, ns_valence, pp_mask)
    g = goldens['hamil_init']['Molecular']
    h = cls(Molecule.from_name('LiH'))
    assert (h.n_up, h.n_down) == (g['n_up'], g['n_down'])
    np.testing.assert_allclose(h.ns_valence, g['ns_valence'])
    np.testing.assert_array_equal(h.ecp_mask, g['pp_mask'])


@pytest.mark.parametrize('cls', [OracleHamiltonian, lambda mol, ecp_type=None: MolecularHamiltonian(mol=mol, ecp_type=ecp_type)])
def test_hamil_init_with_pseudopotential(goldens, cls):
    # reference: tests/test_hamil.py 'Molecular+PP' (LiH, ccECP on Li: two core electrons removed)
    g = goldens['hamil_init']['Molecular_PP']
    h = cls(Molecule.from_name('LiH'), ecp_type='ccECP')
    assert (h.n_up, h.n_down) == (g['n_up'], g['n_down'])
    np.testing.assert_allclose(h.ns_valence, g['ns_valence'])
    np.testing.assert_array_equal(h.ecp_mask, g['pp_mask'])


def _lih_walker(goldens):
    # r = ne[0] because R_Li = 0 and edges are receiver - sender (reference gnn/graph.py:24)
    ne = np.asarray(goldens['edge_builder_LiH']['ne'])
    return ne[0], ne


def test_edge_convention_and_walker(goldens):
    r, ne = _lih_walker(goldens)
    R = Molecule.from_name('LiH').coords
    np.testing.assert_allclose(ne[1] + R[1], r, atol=1e-14)
    same = np.asarray(goldens['edge_builder_LiH']['same'])
    np.testing.assert_allclose(same[0], r[0] - r[1], atol=1e-14)  # receiver - sender, uu block
    anti = np.asarray(goldens['edge_builder_LiH']['anti'])
    # du block first: senders = down electrons, receivers = up electrons (graph.py:146-149)
    np.testing.assert_allclose(anti[0], r[0] - r[2], atol=1e-14)


def test_graph_edge_builder_goldens(goldens):
    # reference: tests/test_gnn.py:7-9,20-32; compute_edges = receiver - sender, optional diagonal filter
    nodes = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 2.0], [0.0, 0.0, 6.0]])
    full = nodes[None, :, :] - nodes[:, None, :]  # [sender, receiver]
    np.testing.assert_allclose(full, np.asarray(goldens['graph_edge_builder_mask_self_False']['graph_edges']))
    n = 3
    sender_idx = (np.arange(n)[None, :] <= np.arange(n - 1)[:, None]) + np.arange(n - 1)[:, None]
    recv_idx = np.broadcast_to(np.arange(n)[None], (n - 1, n))
    np.testing.assert_allclose(full[sender_idx, recv_idx], np.asarray(goldens['graph_edge_builder_mask_self_True']['graph_edges']))


def test_coulomb_known_answers(goldens):
    # reference: tests/test_physics.py:7-17, tests/test_geom.py:8-18
    k = goldens['coulomb_kat']
    mol = Molecule(coords=k['R'], charges=[1, 1], charge=0, spin=0)
    h = OracleHamiltonian(mol)
    R, r = torch.as_tensor(k['R'], dtype=torch.float64), torch.as_tensor(k['r'], dtype=torch.float64)
    assert abs(h.nuclear_energy(R).item() - k['nuclear_energy']) < 1e-12
    assert abs(h.electronic_potential(r).item() - k['electronic_potential']) < 1e-12
    np.testing.assert_allclose(pairwise_self_distance(R).numpy(), [1.4], atol=1e-12)
    d = torch.linalg.norm(r[:, None] - R[None], dim=-1).numpy()
    np.testing.assert_allclose(d, k['pairwise_distance'], atol=1e-12)


def test_lih_potentials_and_eloc_assembly(goldens):
    """V_loc is bit-comparable to the reference golden; the four goldens together pin the
    E_loc assembly (hamil.py:165-172, physics.py:108)."""
    r, _ = _lih_walker(goldens)
    mol = Molecule.from_name('LiH')
    h = OracleHamiltonian(mol)
    rt, Rt = torch.as_tensor(r), torch.as_tensor(mol.coords)
    v_loc = h.local_potential(rt, Rt).item()
    assert abs(v_loc - goldens['potential_LiH_None']['local_potential']) < 1e-11
    v_el, e_nuc = h.electronic_potential(rt).item(), h.nuclear_energy(Rt).item()
    lap = goldens['wf_laplace']['lap_log_psis']
    qf = np.asarray(goldens['wf_laplace']['quantum_force'])
    e_kin = -0.5 * (lap + (qf**2).sum())
    assert abs(e_kin + v_loc + v_el + e_nuc - goldens['local_energy_Molecular']['E_loc']) < 1e-11


def test_oracle_laplacian_two_ways():
    """Hessian trace == jvp-of-grad loop (reference physics.py:144-156) for both ansatz families."""
    from deepqmc_b200 import params as PN
    from deepqmc_b200.spec import ferminet_spec, psiformer_spec

    mol = Molecule.from_name('LiH')
    h = OracleHamiltonian(mol)
    R = torch.as_tensor(mol.coords)
    torch.manual_seed(1)
    r = torch.randn(4, 3, dtype=torch.float64)
    for mk in (psiformer_spec, ferminet_spec):
        spec = mk(h, embedding_dim=16, n_layers=2, n_heads=2, n_determinants=2)
        p = owf.to_torch(PN.perturb_params(PN.init_params(spec, 0)))
        f = lambda x: owf.log_psi(spec, p, x.reshape(-1, 3), R)[1]
        a, ga = laplacian_hessian(f, r.reshape(-1))
        b, gb = laplacian_jvp_loop(f, r.reshape(-1))
        assert abs(a.item() - b.item()) < 1e-9 * max(1, abs(a.item()))
        assert torch.allclose(ga, gb, atol=1e-12)


def test_psiformer_param_count():
    # reference logs 1 610 498 parameters for LiH Psiformer (doc/examples/ground_state_lih.ipynb:67)
    from deepqmc_b200 import params as PN
    from deepqmc_b200.spec import psiformer_spec

    h = OracleHamiltonian(Molecule.from_name('LiH'))
    assert PN.n_params(psiformer_spec(h)) == 1610498


def test_antisymmetry_of_oracle():
    """Exchange of two same-spin electrons flips the sign and keeps log|psi| (fermionic WF)."""
    from deepqmc_b200 import params as PN
    from deepqmc_b200.spec import psiformer_spec

    mol = Molecule.from_name('LiH')
    h = OracleHamiltonian(mol)
    spec = psiformer_spec(h, embedding_dim=16, n_layers=1, n_heads=2, n_determinants=2)
    p = owf.to_torch(PN.perturb_params(PN.init_params(spec, 0)))
    R = torch.as_tensor(mol.coords)
    torch.manual_seed(0)
    r = torch.randn(4, 3, dtype=torch.float64)
    s0, l0 = owf.log_psi(spec, p, r, R)
    s1, l1 = owf.log_psi(spec, p, r[[1, 0, 2, 3]], R)
    assert s0.item() == -s1.item() and abs(l0.item() - l1.item()) < 1e-10


def test_oracle_reproduces_parameter_dependent_reference_goldens(goldens):
    """The reference's own regression fixtures for its test ansatz (tests/conf/ansatz.yaml on LiH) with the parameters
    ``hk.transform(...).init(jax.random.PRNGKey(0), phys_conf)`` creates: test_wf/test_psi.npz, test_grad_psi.npz,
    test_laplace_psi.npz and test_hamil/test_local_energy_Molecular_.npz.  The parameters are regenerated without JAX by
    oracle/jaxrand.py (Threefry + haiku initialisers in creation order), the walker comes from the edge-builder golden.
    Agreement is limited to ~3e-7 by the reference itself: it evaluates the first layer's node MLPs in float32 (hk.Embed
    tables and the hk.Linear layers fed by them are float32 under jax_enable_x64), the oracle uses float64 throughout.
    The reference's own tolerances for these fixtures are rtol 1e-4 / atol 1e-6 (grad) and rtol 2e-4 (E_loc)."""
    from deepqmc_b200 import params as PN
    from deepqmc_b200.spec import paulinet_spec
    from oracle import jaxrand
    from oracle.laplacian import laplacian_hessian

    mol = Molecule.from_name('LiH')
    oh = OracleHamiltonian(mol)
    spec = paulinet_spec(oh)
    p = jaxrand.haiku_init_conv_gnn_ansatz(spec, seed=0)
    assert {k: list(v.shape) for k, v in p.items()} == goldens['test_ansatz_param_shapes']
    r = torch.as_tensor(np.asarray(goldens['edge_builder_LiH']['ne'])[0])  # edges are r - R and R_Li = 0
    R = torch.as_tensor(mol.coords)
    pt = {k: torch.as_tensor(v).requires_grad_(True) for k, v in p.items()}
    sign, log = owf.log_psi(spec, pt, r, R)
    assert sign.item() == goldens['wf_psi']['sign'] and abs(log.item() - goldens['wf_psi']['log']) < 1e-6
    log.backward()
    for k, ref in goldens['wf_grad_psi'].items():
        ref = np.asarray(ref, dtype=np.float64)
        assert np.allclose(pt[k].grad.numpy(), ref, rtol=1e-5, atol=1e-6 * max(1.0, np.abs(ref).max())), k
    p64 = owf.to_torch(p)
    f = lambda x: owf.log_psi(spec, p64, x, R)
    lap, grad = laplacian_hessian(lambda x: f(x.reshape(-1, 3))[1], r.reshape(-1))
    assert abs(lap.item() - goldens['wf_laplace']['lap_log_psis']) < 1e-6 * abs(lap.item())
    assert np.allclose(grad.numpy(), np.asarray(goldens['wf_laplace']['quantum_force']), rtol=1e-5, atol=1e-6)
    e_loc, _ = oh.local_energy(f, r, R)
    assert abs(e_loc.item() - goldens['local_energy_Molecular']['E_loc']) < 2e-6


def test_oracle_gnn_reproduces_reference_embedding_fixture(goldens):
    """tests/test_gnn.py TestGNN.test_embedding: the bare ElectronGNN of tests/conf/gnn.yaml (4 interactions, three-layer
    'log' MLPs for w / h / g) with its Haiku-initialised parameters -- pins the multi-layer message passing of the oracle
    (layers 2-4 run on float64 embeddings in the reference as well)."""
    from deepqmc_b200.spec import paulinet_spec
    from oracle import jaxrand

    mol = Molecule.from_name('LiH')
    spec = paulinet_spec(OracleHamiltonian(mol), n_layers=4, gnn_subnet_layers=3)
    p = jaxrand.haiku_init_conv_gnn_ansatz(spec, seed=0, gnn_only=True, g_layers=3)
    r = torch.as_tensor(np.asarray(goldens['edge_builder_LiH']['ne'])[0])
    x = owf.paulinet_embeddings(spec, owf.to_torch(p), r, torch.as_tensor(mol.coords))
    ref = np.asarray(goldens['gnn_embedding']['embedding'])
    assert x.shape == ref.shape == (4, 8) and np.abs(x.numpy() - ref).max() < 2e-6  # reference tolerance: rtol 1e-4


def test_additive_backflow_oracle_antisymmetry_and_mult_limit():
    """BackflowOp with an additive branch (nn_wave_function.py:14-33): still antisymmetric; with a zero additive head
    'both' reduces to 'mult'; the cutoff polynomial is C^2 at R = 1 (value 1, zero slope / curvature)."""
    from deepqmc_b200 import params as PN
    from deepqmc_b200.spec import psiformer_spec

    mol = Molecule.from_name('LiH')
    h = OracleHamiltonian(mol)
    hy = dict(embedding_dim=16, n_layers=1, n_heads=2, n_determinants=2)
    both, mult = psiformer_spec(h, backflow_transform='both', **hy), psiformer_spec(h, **hy)
    pb = PN.perturb_params(PN.init_params(both, 0))
    assert set(pb) - set(PN.init_params(mult, 0)) == {PN.BF_UP_ADD + ':w', PN.BF_DN_ADD + ':w'}
    p = owf.to_torch(pb)
    R = torch.as_tensor(mol.coords)
    torch.manual_seed(0)
    r = 0.5 * torch.randn(4, 3, dtype=torch.float64)
    s0, l0 = owf.log_psi(both, p, r, R)
    s1, l1 = owf.log_psi(both, p, r[[1, 0, 2, 3]], R)
    assert s0.item() == -s1.item() and abs(l0.item() - l1.item()) < 1e-10
    sm, lm = owf.log_psi(mult, p, r, R)
    assert abs(lm.item() - l0.item()) > 1e-6  # the additive head matters ...
    p0 = {**p, PN.BF_UP_ADD + ':w': torch.zeros_like(p[PN.BF_UP_ADD + ':w']), PN.BF_DN_ADD + ':w': torch.zeros_like(p[PN.BF_DN_ADD + ':w'])}
    sz, lz = owf.log_psi(both, p0, r, R)
    assert sz.item() == sm.item() and abs(lz.item() - lm.item()) < 1e-12  # ... and vanishes with it
    c = lambda x: x**2 * (6 - 8 * x + 3 * x**2)
    x = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    g, = torch.autograd.grad(c(x), x, create_graph=True)
    g2, = torch.autograd.grad(g, x)
    assert c(x).item() == 1.0 and abs(g.item()) < 1e-12 and abs(g2.item()) < 1e-12


def test_transpsiformer_nuclear_stream_matches_oracle():
    """Host-side nuclear stream of the product (numpy, deepqmc_b200/nuclear.py) against the oracle's
    independent torch restatement: final nuclear embeddings -> envelope exponents, and antisymmetry
    of the resulting wave function (reference: conf/ansatz/transpsiformer.yaml)."""
    import numpy as np
    import torch

    from deepqmc_b200 import params as PN
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.nuclear import nuclear_stream
    from deepqmc_b200.spec import transpsiformer_spec
    from oracle import wf

    mol = Molecule.from_name('H2O')
    h = MolecularHamiltonian(mol=mol)
    spec = transpsiformer_spec(h, embedding_dim=16, n_layers=2, n_heads=2, n_determinants=3)
    params = PN.perturb_params(PN.init_params(spec, 0))
    pt = wf.to_torch(params)
    rng = np.random.default_rng(0)
    R = torch.as_tensor(mol.coords)
    r = torch.as_tensor(mol.coords[rng.integers(0, 3, size=10)] + rng.normal(size=(10, 3)))
    _, nuc = wf.transpsiformer_embeddings(spec, pt, r, R)
    z = wf.nuclear_head_zetas(spec, pt, nuc)
    ns = nuclear_stream(spec, params, mol.coords)
    for s in ('up', 'down'):
        assert np.allclose(ns[f'zetas_{s}'], z[s].numpy(), rtol=1e-12, atol=1e-13)
    assert len(ns['kn']) == spec.n_layers and ns['kn'][0].shape == (3, 16)
    s0, l0 = wf.log_psi(spec, pt, r, R)
    perm = list(range(10))
    perm[0], perm[1] = 1, 0
    s1, l1 = wf.log_psi(spec, pt, r[perm], R)
    assert s0.item() == -s1.item() and abs(l0.item() - l1.item()) < 1e-10
    # parameter count of the full-size cyclobutadiene ansatz is finite and the table is consistent
    assert PN.n_params(spec) == sum(int(np.prod(v.shape)) for v in params.values())


def test_test_ansatz_parameter_table_matches_reference_golden():
    """The parameter tree of the 'paulinet' mirror (names AND shapes) equals the Haiku tree of the reference's
    test ansatz on LiH, read off the keys of tests/test_wf/test_grad_psi.npz (tests/golden/reference_goldens.json):
    pins the embedding tables, the three w/h/g convolution subnets, the 'log'-width Jastrow (8-4-2-1) and backflow
    (8-6-5-4) MLPs, the per-shell envelope count (3 for LiH) and the hk.Linear conf_coeff."""
    import json
    import os

    from deepqmc_b200 import params as PN
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.spec import paulinet_spec

    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_goldens.json')))
    ref = {k: tuple(v) for k, v in g['test_ansatz_param_shapes'].items()}
    h = MolecularHamiltonian(mol=Molecule.from_name('LiH'))
    mine = {k: tuple(v) for k, v in PN.param_shapes(paulinet_spec(h)).items()}
    assert mine == ref


def test_paulinet_oracle_antisymmetry_and_laplacian_self_check():
    import numpy as np
    import torch

    from deepqmc_b200 import params as PN
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.spec import paulinet_spec
    from oracle import wf
    from oracle.laplacian import laplacian_hessian, laplacian_jvp_loop

    mol = Molecule.from_name('C')
    h = MolecularHamiltonian(mol=mol)
    spec = paulinet_spec(h)
    pt = wf.to_torch(PN.perturb_params(PN.init_params(spec, 0)))
    rng = np.random.default_rng(1)
    R = torch.as_tensor(mol.coords)
    r = torch.as_tensor(rng.normal(size=(6, 3)))
    s0, l0 = wf.log_psi(spec, pt, r, R)
    perm = [1, 0, 2, 3, 4, 5]
    s1, l1 = wf.log_psi(spec, pt, r[perm], R)
    assert s0.item() == -s1.item() and abs(l0.item() - l1.item()) < 1e-10
    f = lambda x: wf.log_psi(spec, pt, x.reshape(-1, 3), R)[1]
    la, ga = laplacian_hessian(f, r.reshape(-1))
    lb, gb = laplacian_jvp_loop(f, r.reshape(-1))
    assert abs(la.item() - lb.item()) < 1e-8 * max(1, abs(la.item())) and torch.allclose(ga, gb)


def test_paulinet_default_yaml_parameter_count_and_antisymmetry():
    """conf/ansatz/default.yaml on LiH: layer widths from the 'log' rule (w: 4-11-32 then 32-32-32; h: 8-16-32 then
    128-64-32; g: 88->128 then 448->128) and antisymmetry of the oracle restatement."""
    import numpy as np
    import torch

    from deepqmc_b200 import params as PN
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.spec import paulinet_default_spec
    from oracle import wf

    mol = Molecule.from_name('LiH')
    h = MolecularHamiltonian(mol=mol)
    spec = paulinet_default_spec(h)
    sh = PN.param_shapes(spec)
    c0 = PN.conv_prefix(0)
    assert sh[c0 + 'w_same/linear_0:w'] == (4, 11) and sh[c0 + 'w_same/linear_1:w'] == (11, 32)
    assert sh[c0 + 'h_anti/linear_0:w'] == (8, 16) and sh[PN.conv_prefix(1) + 'h_anti/linear_0:w'] == (128, 64)
    assert sh[PN.layer_prefix(0) + 'g/linear_0:w'] == (88, 128) and sh[PN.layer_prefix(2) + 'g/linear_0:w'] == (448, 128)
    assert PN.layer_prefix(2) + 'u/linear_0:w' not in sh  # no edge update in the last layer
    small = paulinet_default_spec(h, embedding_dim=16, n_determinants=3, edge_dim=8)
    pt = wf.to_torch(PN.perturb_params(PN.init_params(small, 0)))
    rng = np.random.default_rng(2)
    R = torch.as_tensor(mol.coords)
    r = torch.as_tensor(rng.normal(size=(4, 3)))
    s0, l0 = wf.log_psi(small, pt, r, R)
    s1, l1 = wf.log_psi(small, pt, r[[1, 0, 2, 3]], R)
    assert s0.item() == -s1.item() and abs(l0.item() - l1.item()) < 1e-10


def test_overlap_tangent_cotangents_match_autograd_of_the_penalty():
    """compute_mean_overlap_tangent (reference loss/overlap.py:182-229): with a toy 'ansatz' whose log|psi| is linear
    in the parameters the returned gradient must equal the reference's tangent formula evaluated directly:
    d/d theta_j of  sum_{i<j} 2 <ratio_ji> < (ratio_ij - <ratio_ij>) d log psi_j > ."""
    import torch

    from deepqmc_b200.overlap import compute_mean_overlap_tangent
    from deepqmc_b200.types import PhysicalConfiguration, Psi

    g = torch.Generator().manual_seed(0)
    n, B, P = 3, 7, 4
    ratio = torch.randn(n, n, B, generator=g, dtype=torch.float64)
    feats = torch.randn(n, B, P, generator=g, dtype=torch.float64)  # d log psi_j(r_b) / d theta_j

    class Toy:
        def log_psi_vjp(self, params, pc, cot):
            return Psi(torch.ones_like(cot), torch.zeros_like(cot)), {'theta': (cot[:, None] * pc.r).sum(0)}

    pc = PhysicalConfiguration(torch.zeros(1, 3), feats, torch.zeros(n, B))
    grads = compute_mean_overlap_tangent(ratio, None, None, Toy(), [None] * n, pc)
    assert grads[0] is None
    mean = ratio.mean(-1)
    for j in (1, 2):
        ref = torch.zeros(P, dtype=torch.float64)
        for i in range(j):
            ref += 2 * mean[j, i] * ((ratio[i, j] - mean[i, j])[:, None] * feats[j]).mean(0)
        assert torch.allclose(grads[j]['theta'], ref)
