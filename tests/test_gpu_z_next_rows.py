"""GPU parity tests for the rows SURVEY.md 8(f) marks "next": Langevin sampler (N3) and the TransPsiformer parameter
VJP (N1).  Kept in a file that sorts after the hot-path parity tests so that `pytest -x` reaches those first."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from deepqmc_b200.types import PhysicalConfiguration
from test_gpu_parity import DEV, make


def test_langevin_injected_noise_matches_oracle():
    """LangevinSampler (reference electron_samplers.py:176-232, clean_force sampling_utils.py:71-101; SURVEY.md 8(f)
    N3) with identical injected random numbers: update (value + cleaned drift), three sub-steps with rejections,
    age / tau bookkeeping."""
    from deepqmc_b200.sampling import LangevinSampler
    from oracle import wf
    from oracle.sampling import clean_force, langevin_step

    mol, hamil, oh, ansatz, params, r, R = make('H2O', B=5, embedding_dim=32, n_layers=1, n_heads=2, n_determinants=2)
    pt = wf.to_torch(params)
    Rc = R.cpu()
    B, N = 5, 10

    def wfg(rr):
        ss, ll, gg = [], [], []
        for b in range(len(rr)):
            x = rr[b].clone().requires_grad_(True)
            s, l = wf.log_psi(ansatz.spec, pt, x, Rc)
            g, = torch.autograd.grad(l, x)
            ss.append(s.detach()); ll.append(l.detach()); gg.append(g)
        return torch.stack(ss), torch.stack(ll), torch.stack(gg)

    tau0 = 1.5
    smp = LangevinSampler(hamil, ansatz.apply, tau=tau0, max_age=2, length=3)
    state = {'r': r.clone(), 'age': torch.zeros(B, dtype=torch.int32, device=DEV),
             'tau': torch.tensor([tau0], dtype=torch.float64, device=DEV)}
    state = smp.update(state, params, R)
    s0, l0, g0 = wfg(r.cpu())
    f0 = clean_force(g0, r.cpu(), Rc, mol.charges, torch.tensor(tau0, dtype=torch.float64))
    assert torch.allclose(state['psi'].log.cpu(), l0, rtol=0, atol=1e-9)
    assert torch.allclose(state['force'].cpu(), f0, rtol=1e-8, atol=1e-9)
    rng = np.random.default_rng(4)
    nn = torch.as_tensor(rng.normal(size=(3, B, N, 3)), device=DEV)
    nu = torch.as_tensor(rng.uniform(size=(3, B)), device=DEV)
    new, pc, stats = smp.sample(0, state, params, R, noise_normal=nn, noise_uniform=nu)
    ost = dict(r=r.cpu().clone(), sign=s0, log=l0, force=f0, age=torch.zeros(B, dtype=torch.int32),
               tau=torch.tensor(tau0, dtype=torch.float64))
    for s in range(3):
        ost, acc = langevin_step(wfg, Rc, mol.charges, ost, nn[s].cpu(), nu[s].cpu(), 0.57, 2)
    assert torch.allclose(new['r'].cpu(), ost['r'], rtol=0, atol=1e-9)
    assert torch.allclose(new['psi'].log.cpu(), ost['log'], rtol=0, atol=1e-9)
    assert torch.allclose(new['force'].cpu(), ost['force'], rtol=1e-8, atol=1e-9)
    assert new['age'].cpu().tolist() == ost['age'].tolist()
    assert abs(new['tau'].item() - ost['tau'].item()) < 1e-12 and abs(stats['sampling/acceptance'].item() - acc.item()) < 1e-12


def test_parameter_vjp_transpsiformer_matches_autograd_fp64():
    """Reverse pass of the TransPsiformer: electron stream in the CUDA engine (attention with nuclear tokens, cotangents
    of their keys / values and of the envelope exponents accumulated over walkers), nuclear stream differentiated on the
    host -- every parameter against torch autograd through the oracle."""
    from oracle import wf

    mol, hamil, oh, ansatz, params, r, R = make('H2O', B=3, kind='transpsiformer', embedding_dim=32, n_layers=2, n_heads=2,
                                                n_determinants=3)
    w = torch.as_tensor(np.random.default_rng(5).normal(size=3), device=DEV)
    psi, grads = ansatz.log_psi_vjp(params, PhysicalConfiguration(R, r, torch.zeros(3, device=DEV)), w)
    pt = {k: torch.as_tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in params.items()}
    tot = 0
    for b in range(3):
        s, l = wf.log_psi(ansatz.spec, pt, r[b].cpu(), R.cpu())
        assert abs(psi.log[b].item() - l.item()) <= 1e-10 * max(1, abs(l.item()))
        tot = tot + w[b].cpu() * l
    tot.backward()
    assert set(grads) == set(pt)
    for k, v in pt.items():
        ref = v.grad
        assert torch.allclose(grads[k].cpu().reshape(ref.shape), ref, rtol=1e-8, atol=1e-9 * max(1.0, ref.abs().max().item())), k


@pytest.mark.parametrize('kind,charges,spin,dtype', [
    ('psiformer', [17, 1], 0, 'float64'),        # HCl-like: one pseudo-Hamiltonian centre, 8 valence electrons
    ('psiformer', [15, 17], 0, 'float64'),       # two centres with different tables, 12 electrons
    ('transpsiformer', [16, 1, 1], 0, 'float64'),
    ('psiformer', [16, 1, 1], 0, 'float32'),
    ('ferminet', [16, 1, 1], 0, 'float64'),      # edge features carry the metrics of both electrons of a pair
    ('paulinet', [16, 1, 1], 0, 'float64'),      # compact (value, d/dv_i, d/dv_j, Laplacian) edge state
    ('paulinet_default', [17, 1], 0, 'float64'),
])
def test_pseudo_hamiltonian_local_energy(tmp_path, kind, charges, spin, dtype):
    """PseudoHamiltonian (reference ecp/pseudo_hamiltonian.py:165-278): mass-tensor kinetic term
    through Cholesky-seeded forward-Laplacian tangents, first-order term, tabulated local potential -- E_loc, all six
    statistics and grad_r log|psi| against the oracle (autograd Hessian in the transformed coordinates).  Tables:
    synthetic files in the reference's XML layout (tests/ph_fixture.py)."""
    from deepqmc_b200.hamil import STAT_KEYS, MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200 import params as PN
    from deepqmc_b200.ansatz import B200Ansatz
    from oracle import wf
    from oracle.hamil import OracleHamiltonian
    from ph_fixture import write_synthetic_ph

    d = write_synthetic_ph(str(tmp_path))
    coords = [[0.0, 0.0, 0.0], [2.4, 0.0, 0.0], [0.0, 2.6, 0.3]][:len(charges)]
    mol = Molecule(coords=coords, charges=charges, charge=0, spin=spin)
    hamil = MolecularHamiltonian(mol=mol, ecp_type='PH', ph_data_dir=d)
    oh = OracleHamiltonian(mol, ecp_type='PH', ph_dir=d)
    hyper = {'ferminet': dict(embedding_dim=32, n_layers=2, n_determinants=3, edge_dim=8), 'paulinet': dict(n_layers=2),
             'paulinet_default': dict(embedding_dim=16, n_layers=2, n_determinants=3, edge_dim=8)}.get(
                 kind, dict(embedding_dim=32, n_layers=2, n_determinants=3, n_heads=2))
    ansatz = B200Ansatz(hamil, kind, dtype=dtype, **hyper)
    params = PN.perturb_params(ansatz.init(0))
    pt = wf.to_torch(params)
    B, N = 3, hamil.n_up + hamil.n_down
    rng = np.random.default_rng(0)
    r64 = torch.as_tensor(mol.coords[rng.integers(0, len(mol.coords), size=(B, N))] + 0.8 * rng.normal(size=(B, N, 3)))
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    R = torch.as_tensor(mol.coords, dtype=tdt, device=DEV)
    pc = PhysicalConfiguration(R, r64.to(tdt).to(DEV), torch.zeros(B, device=DEV))
    E, st, grad = hamil.local_energy(ansatz.apply)(None, params, pc, return_grad=True)
    for b in range(B):
        f = lambda x: wf.log_psi(ansatz.spec, pt, x, R.cpu().double())
        eo, so = oh.local_energy(f, r64[b], R.cpu().double())
        x = r64[b].clone().requires_grad_(True)
        g, = torch.autograd.grad(f(x)[1], x)
        if dtype == 'float64':
            tol = 1e-8 * max(1.0, abs(eo.item()), 0.5 * abs(so['hamil/lap'].item()))
            gtol = 1e-8 * max(1.0, g.abs().max().item())
        else:  # the reference's own fp32 regression tolerance (tests/test_hamil.py:37-40) on the largest term
            tol = 2e-4 * max(1.0, abs(eo.item()), abs(so['hamil/lap'].item()), abs(so['hamil/quantum_force'].item()))
            gtol = 2e-4 * max(1.0, g.abs().max().item())
        assert abs(E[b].item() - eo.item()) <= tol
        for k in STAT_KEYS:
            assert abs(st[k][b].item() - so[k].item()) <= tol * (10 if k in ('hamil/lap', 'hamil/quantum_force') else 1), k
        assert (grad[b].reshape(N, 3).cpu().double() - g).abs().max().item() <= gtol


@pytest.mark.parametrize('mol_name,hyper', [
    ('H2O', dict(embedding_dim=32, n_layers=3, n_determinants=3, edge_dim=8)),
    ('LiH', dict(embedding_dim=16, n_layers=1, n_determinants=2, edge_dim=8)),
])
def test_parameter_vjp_ferminet_matches_autograd_fp64(mol_name, hyper):
    """Reverse pass of the FermiNet trunk (node update on concat[h, spin means, spin means of the incoming edges], shared
    edge MLP, residuals / sqrt(2); reference gnn/electron_gnn.py:160-259, update_features.py:47-159) -- every parameter
    against torch autograd through the oracle."""
    from oracle import wf

    mol, hamil, oh, ansatz, params, r, R = make(mol_name, B=3, kind='ferminet', **hyper)
    w = torch.as_tensor(np.random.default_rng(6).normal(size=3), device=DEV)
    psi, grads = ansatz.log_psi_vjp(params, PhysicalConfiguration(R, r, torch.zeros(3, device=DEV)), w)
    pt = {k: torch.as_tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in params.items()}
    tot = 0
    for b in range(3):
        s, l = wf.log_psi(ansatz.spec, pt, r[b].cpu(), R.cpu())
        assert psi.sign[b].item() == s.item() and abs(psi.log[b].item() - l.item()) <= 1e-10 * max(1, abs(l.item()))
        tot = tot + w[b].cpu() * l
    tot.backward()
    assert set(grads) == set(pt)
    for k, v in pt.items():
        ref = v.grad
        assert torch.allclose(grads[k].cpu().reshape(ref.shape), ref, rtol=1e-8, atol=1e-9 * max(1.0, ref.abs().max().item())), k


def test_empty_and_single_walker_batches():
    """Edge cases of the batched entry points: an empty batch is a no-op (empty outputs, zero parameter gradient),
    a single walker equals its row of a larger batch (walkers are independent), and the samplers
    refuse an empty walker set."""
    mol, hamil, oh, ansatz, params, r, R = make('LiH', B=3, embedding_dim=32, n_layers=2, n_heads=4, n_determinants=4)
    eng = ansatz.engine_for(hamil, params)
    E3, st3, s3, l3, g3 = eng.local_energy(r, R, want_grad=True)
    E1, st1, s1, l1, g1 = eng.local_energy(r[1:2], R, want_grad=True)
    same = lambda a, b: torch.allclose(a, b, rtol=1e-11, atol=1e-11)
    assert same(E1, E3[1:2]) and same(st1, st3[:, 1:2]) and same(l1, l3[1:2]) and same(g1, g3[1:2])
    E0, st0, s0, l0, g0 = eng.local_energy(r[:0], R, want_grad=True)
    assert E0.shape == (0,) and st0.shape == (6, 0) and l0.shape == (0,) and g0.shape == (0, 12)
    sf, lf = eng.wf_forward(r[:0], R)
    assert sf.shape == (0,) and lf.shape == (0,)
    _, _, grads = eng.vjp_params(r[:0], R, torch.zeros(0, dtype=r.dtype, device=DEV))
    assert all(float(v.abs().sum()) == 0.0 for v in grads.values())
    state = dict(r=r[:0].clone(), sign=s3[:0].clone(), log=l3[:0].clone(), age=torch.zeros(0, dtype=torch.int32, device=DEV),
                 tau=torch.tensor([0.1], dtype=r.dtype, device=DEV))
    with pytest.raises(RuntimeError):
        eng.mcmc_sweep(state, R, 2)


def test_multi_nuclear_geometry_sampler_two_lih_geometries():
    """MultiNuclearGeometrySampler (reference sampling/combined_samplers.py:93-214) over two bond lengths of LiH with one
    parameter tree: every geometry's walkers are advanced with that geometry's nuclei, the returned batch has the
    reference layout [mol, state, walker] with R tiled per walker, and compute_local_energy on it equals the direct
    evaluation per geometry."""
    from deepqmc_b200.energy import compute_local_energy
    from deepqmc_b200.sampling import DecorrSampler, MultiNuclearGeometrySampler

    mol, hamil, oh, ansatz, params, r, R = make('LiH', B=4, embedding_dim=32, n_layers=2, n_heads=4, n_determinants=4)
    Rs = torch.stack([R.cpu(), R.cpu() * 1.3])
    smp = MultiNuclearGeometrySampler(DecorrSampler(hamil, ansatz.apply, length=3, tau=0.3))
    state = smp.init(3, params, 4, Rs)
    r0 = [state['elec'][m]['r'].clone() for m in range(2)]
    state, pc, stats = smp.sample(5, state, params, [1, 0])
    assert pc.batch_shape == (2, 1, 4) and pc.mol_idx[:, 0, 0].tolist() == [1, 0]
    assert torch.equal(pc.R[0, 0, 2].cpu(), Rs[1]) and torch.equal(pc.R[1, 0, 0].cpu(), Rs[0])
    assert stats['sampling/acceptance'].shape[0] == 2
    eng = ansatz.engine_for(hamil, params)
    for k, m in enumerate([1, 0]):
        assert not torch.equal(pc.r[k, 0], r0[m])  # walkers moved
        s, l = eng.wf_forward(pc.r[k, 0], Rs[m].to(DEV))
        assert torch.allclose(l, state['elec'][m]['psi'].log, rtol=0, atol=1e-10)  # psi of the state belongs to ITS geometry
    E, st = compute_local_energy(None, hamil, ansatz.apply, params, pc)
    assert E.shape == (2, 1, 4)
    for k, m in enumerate([1, 0]):
        Ed = eng.local_energy(pc.r[k, 0], Rs[m].to(DEV))[0]
        assert torch.allclose(E[k, 0], Ed, rtol=0, atol=1e-10)
    from oracle import wf

    pt = wf.to_torch(params)
    f = lambda x: wf.log_psi(ansatz.spec, pt, x, Rs[1])
    eo, _ = oh.local_energy(f, pc.r[0, 0, 0].cpu(), Rs[1])
    assert abs(E[0, 0, 0].item() - eo.item()) <= 1e-8 * max(1.0, abs(eo.item()))


@pytest.mark.parametrize('kind,mol_name,hyper', [
    ('psiformer', 'H2O', dict(embedding_dim=32, n_layers=1, n_heads=2, n_determinants=3)),
    ('ferminet', 'LiH', dict(embedding_dim=16, n_layers=2, n_determinants=2, edge_dim=8)),
    ('transpsiformer', 'LiH', dict(embedding_dim=32, n_layers=1, n_heads=2, n_determinants=2)),
    ('paulinet', 'LiH', dict()),                 # spin-factorised determinants: n_up x n_up and n_down x n_down blocks
    ('paulinet_default', 'H2O', dict(embedding_dim=16, n_layers=1, n_determinants=2, edge_dim=8)),
])
def test_return_mos_matches_oracle(kind, mol_name, hyper):
    """Ansatz.apply(params, phys_conf, return_mos=True) (reference types.py:133-150, nn_wave_function.py:131-142): the
    per-spin orbital matrices envelope * mult_act(backflow); their determinants reproduce psi."""
    from oracle import wf

    mol, hamil, oh, ansatz, params, r, R = make(mol_name, B=3, kind=kind, **hyper)
    pc = PhysicalConfiguration(R, r, torch.zeros(3, device=DEV))
    up, dn = ansatz.apply(params, pc, return_mos=True)
    psi = ansatz.apply(params, pc)
    pt = wf.to_torch(params)
    for b in range(3):
        ou, od = wf.molecular_orbitals(ansatz.spec, pt, r[b].cpu(), R.cpu())
        assert up[b].shape == ou.shape and dn[b].shape == od.shape
        assert torch.allclose(up[b].cpu(), ou, rtol=1e-9, atol=1e-11) and torch.allclose(dn[b].cpu(), od, rtol=1e-9, atol=1e-11)
    su, du = ansatz.apply(params, pc[0], return_mos=True)  # single-sample call, as in the reference
    assert su.shape == up[0].shape and torch.equal(su, up[0]) and torch.equal(du, dn[0])
    if ansatz.spec.full_determinant and ansatz.spec.cusp == 'none' and kind == 'ferminet':
        s, l = torch.linalg.slogdet(torch.cat([up, dn], 2))  # [B, K]
        tot = (s * torch.exp(l - l.max(-1, keepdim=True).values)).sum(-1)
        assert torch.allclose(torch.log(tot.abs()) + l.max(-1).values, psi.log, rtol=0, atol=1e-9)


def test_laplacian_factory_seam_matches_oracle():
    """The LaplacianFactory seam (reference physics.py:24-33,144-156): Laplacian and gradient of log|psi| with respect to
    the 3N electron coordinates, against the oracle's Hessian trace / autograd gradient."""
    from oracle import wf
    from oracle.laplacian import laplacian_hessian

    mol, hamil, oh, ansatz, params, r, R = make('H2O', B=3, embedding_dim=32, n_layers=2, n_heads=2, n_determinants=3)
    lap_fn = hamil.laplacian(ansatz.apply)
    lap, grad = lap_fn(params, PhysicalConfiguration(R, r, torch.zeros(3, device=DEV)))
    pt = wf.to_torch(params)
    for b in range(3):
        lo, go = laplacian_hessian(lambda x: wf.log_psi(ansatz.spec, pt, x.reshape(-1, 3), R.cpu())[1], r[b].cpu().reshape(-1))
        assert abs(lap[b].item() - lo.item()) <= 1e-8 * max(1.0, abs(lo.item()))
        assert torch.allclose(grad[b].cpu(), go, rtol=1e-8, atol=1e-9)
    l1, g1 = lap_fn(params, PhysicalConfiguration(R, r[1], torch.zeros((), device=DEV)))  # single sample, as the reference calls it
    assert torch.allclose(l1, lap[1]) and torch.allclose(g1, grad[1])


@pytest.mark.parametrize('kind,mode,mol_name,dtype', [
    ('psiformer', 'both', 'H2O', 'float64'),
    ('psiformer', 'add', 'LiH', 'float64'),
    ('ferminet', 'both', 'LiH', 'float64'),
    ('psiformer', 'both', 'H2O', 'float32'),
])
def test_additive_backflow_branch(kind, mode, mol_name, dtype):
    """backflow_transform = 'add' / 'both' of the BackflowOp (reference wf/nn_wave_function.py:14-33,111-125): orbitals
    = envelope * f_mult + cutoff(nearest-nucleus distance) * |envelope_i| * 0.1 tanh(f_add / 4).  Walkers are drawn close
    to the nuclei so the cutoff polynomial is active.  Value, orbital matrices and the local energy against the oracle."""
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200 import params as PN
    from deepqmc_b200.ansatz import B200Ansatz
    from oracle import wf
    from oracle.hamil import OracleHamiltonian

    mol = Molecule.from_name(mol_name)
    hamil, oh = MolecularHamiltonian(mol=mol), OracleHamiltonian(mol)
    hyper = dict(embedding_dim=16, n_layers=1, n_determinants=3, backflow_transform=mode)
    hyper.update(dict(n_heads=2) if kind == 'psiformer' else dict(edge_dim=8))
    ansatz = B200Ansatz(hamil, kind, dtype=dtype, **hyper)
    params = PN.perturb_params(ansatz.init(0))
    pt = wf.to_torch(params)
    B, N = 3, hamil.n_up + hamil.n_down
    rng = np.random.default_rng(0)
    r64 = torch.as_tensor(mol.coords[rng.integers(0, len(mol.coords), size=(B, N))] + 0.4 * rng.normal(size=(B, N, 3)))
    R64 = torch.as_tensor(mol.coords)
    assert ((r64[:, :, None] - R64[None, None]).norm(dim=-1).min(-1).values < 0.5).any()  # cutoff region is sampled
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    pc = PhysicalConfiguration(R64.to(tdt).to(DEV), r64.to(tdt).to(DEV), torch.zeros(B, device=DEV))
    E, st = hamil.local_energy(ansatz.apply)(None, params, pc)
    psi = ansatz.apply(params, pc)
    up, dn = ansatz.apply(params, pc, return_mos=True)
    tol = 1e-8 if dtype == 'float64' else 2e-4
    for b in range(B):
        f = lambda x: wf.log_psi(ansatz.spec, pt, x, R64)
        so, lo = f(r64[b])
        eo, sto = oh.local_energy(f, r64[b], R64)
        ou, od = wf.molecular_orbitals(ansatz.spec, pt, r64[b], R64)
        scale = max(1.0, abs(eo.item()), 0.5 * abs(sto['hamil/lap'].item()), 0.5 * sto['hamil/quantum_force'].item())
        assert psi.sign[b].item() == so.item() and abs(psi.log[b].item() - lo.item()) <= (1e-10 if dtype == 'float64' else 2e-4) * max(1, abs(lo.item()))
        assert abs(E[b].item() - eo.item()) <= tol * scale
        mt = 1e-10 if dtype == 'float64' else 1e-4
        assert torch.allclose(up[b].cpu().double(), ou, rtol=mt, atol=mt) and torch.allclose(dn[b].cpu().double(), od, rtol=mt, atol=mt)


@pytest.mark.parametrize('kind,form', [('psiformer', 'psiformer'), ('ferminet', 'deepqmc')])
def test_parameter_vjp_with_trainable_nuclear_cusp(kind, form):
    """Reverse pass with the NuclearCuspAsymptotic factor switched on (reference wf/cusp.py:81-101): the gradient of its
    trainable exponent and of every other parameter against torch autograd through the oracle."""
    from oracle import wf

    hyper = dict(embedding_dim=16, n_layers=1, n_determinants=2, cusp_nuclei=form, cusp_nuclei_alpha=0.7)
    hyper.update(dict(n_heads=2) if kind == 'psiformer' else dict(edge_dim=8))
    mol, hamil, oh, ansatz, params, r, R = make('LiH', B=3, kind=kind, **hyper)
    w = torch.as_tensor(np.random.default_rng(8).normal(size=3), device=DEV)
    psi, grads = ansatz.log_psi_vjp(params, PhysicalConfiguration(R, r, torch.zeros(3, device=DEV)), w)
    pt = {k: torch.as_tensor(v, dtype=torch.float64).requires_grad_(True) for k, v in params.items()}
    tot = 0
    for b in range(3):
        s, l = wf.log_psi(ansatz.spec, pt, r[b].cpu(), R.cpu())
        assert abs(psi.log[b].item() - l.item()) <= 1e-10 * max(1, abs(l.item()))
        tot = tot + w[b].cpu() * l
    tot.backward()
    assert set(grads) == set(pt) and any('nuc_alpha' in k for k in grads)
    for k, v in pt.items():
        ref = v.grad
        assert torch.allclose(grads[k].cpu().reshape(ref.shape), ref, rtol=1e-8, atol=1e-9 * max(1.0, ref.abs().max().item())), k
