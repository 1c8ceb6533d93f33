"""CPU, gloo, world_size = 2: walker sharding and the fused per-step statistics exchange
(deepqmc_b200/parallel.py; reference: src/deepqmc/parallel.py, loss/energy.py:63-74)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from deepqmc_b200 import parallel

    r, w = parallel.init_from_env('gloo')
    assert (r, w) == (rank, world)
    B = 12
    g = torch.Generator().manual_seed(0)
    E_all = torch.randn(B, generator=g, dtype=torch.float64)
    stats_all = {'hamil/E_kin': torch.randn(B, generator=g, dtype=torch.float64),
                 'hamil/V_el': torch.randn(B, generator=g, dtype=torch.float64)}
    lo, hi = parallel.shard_bounds(B)
    out = parallel.energy_statistics(E_all[lo:hi], {k: v[lo:hi] for k, v in stats_all.items()})
    gathered = parallel.all_gather_walkers(E_all[lo:hi])
    ok = (
        abs(out['energy/mean'].item() - E_all.mean().item()) < 1e-12
        and abs(out['energy/var'].item() - E_all.var(unbiased=False).item()) < 1e-12
        and abs(out['energy/max'].item() - E_all.max().item()) < 1e-15
        and abs(out['energy/min'].item() - E_all.min().item()) < 1e-15
        and abs(out['hamil/E_kin'].item() - stats_all['hamil/E_kin'].mean().item()) < 1e-12
        and torch.equal(gathered, E_all)
        and parallel.rank_seed(5) == 5 + rank
    )
    # global-median clipping + energy-gradient cotangents / gradient all-reduce (loss/clip.py:73-98, loss/energy.py:77-102)
    from deepqmc_b200.energy import compute_mean_energy_tangent, median_clip_and_mask
    from deepqmc_b200.types import PhysicalConfiguration, Psi

    Ec, mask = median_clip_and_mask(E_all[lo:hi], 1.0, exclude_width=2.0)
    center = torch.as_tensor(np.median(E_all.numpy()))  # jnp.median convention: mean of the two middle values
    mad = (E_all - center).abs().mean()
    ok = ok and torch.allclose(Ec, torch.clamp(E_all[lo:hi], center - mad, center + mad))
    ok = ok and torch.equal(mask, (E_all[lo:hi] - center).abs() < 2.0)

    # soft squeeze and wave-function-ratio clipping use all-walker medians / quantiles as well (loss/clip.py:101-174)
    from deepqmc_b200.energy import median_log_squeeze_and_mask
    from deepqmc_b200.overlap import psi_ratio_clip_and_mask

    xs, ms = median_log_squeeze_and_mask(E_all[lo:hi], clip_width=1.5, quantile=0.9, exclude_width=2.0)
    qv = torch.quantile((E_all - center).abs(), 0.9)
    z = (E_all[lo:hi] - center) / (3.0 * qv)
    ref_sq = center + 3.0 * qv * torch.sign(z) * torch.log1p((z.abs() + 0.5 * z**2 + z.abs() ** 3) / (1 + z**2))
    ok = ok and torch.allclose(xs, ref_sq) and torch.equal(ms, (E_all[lo:hi] - center).abs() / qv < 2.0)
    rc, mr = psi_ratio_clip_and_mask(E_all[lo:hi], clip_width=2.0, exclude_width=1.0)
    sig = torch.as_tensor(np.median((E_all - center).abs().numpy()))
    ok = ok and torch.allclose(rc, torch.clamp(E_all[lo:hi], center - 2 * sig, center + 2 * sig))
    ok = ok and torch.equal(mr, (E_all[lo:hi] - center).abs() < 1.0)

    class FakeAnsatz:  # d log psi_b / d theta = feature vector f_b: the VJP is sum_b cot_b f_b
        def log_psi_vjp(self, params, pc, cot):
            return Psi(torch.ones_like(cot), torch.zeros_like(cot)), {'theta': (cot[:, None] * pc.r).sum(0)}

    feats = torch.randn(B, 3, generator=g, dtype=torch.float64)
    pc = PhysicalConfiguration(torch.zeros(1, 3), feats[lo:hi], torch.zeros(hi - lo))
    grads = compute_mean_energy_tangent(E_all[lo:hi], None, mask, FakeAnsatz(), None, pc)
    mask_all = (E_all - center).abs() < 2.0
    ref = (((E_all - E_all.mean()) * mask_all)[:, None] * feats).sum(0) / mask_all.sum()
    ok = ok and torch.allclose(grads['theta'], ref)
    q.put((rank, bool(ok)))
    torch.distributed.destroy_process_group()


def test_sharded_statistics_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_contract():
    from deepqmc_b200 import parallel

    assert parallel.shard_bounds(4096, 3, 8) == (1536, 2048)
    with pytest.raises(ValueError):  # electron_batch_size % device_count == 0 (validate_kwargs.py:45-48)
        parallel.shard_bounds(10, 0, 4)
    E = torch.arange(6, dtype=torch.float64)
    out = parallel.energy_statistics(E)  # single process: no collective
    assert out['energy/mean'].item() == 2.5 and out['energy/count'].item() == 6


def test_initializer_and_host_hamiltonian():
    """Walker initialiser (host, one-off): right shapes, spins on the right atoms, charge-neutral
    assignment (reference: electron_sample_initializers.py:43-160; tests/test_hamil.py:28-31)."""
    import numpy as np

    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.sampling import AtomCenteredElectronInitializer

    for name, ecp in (('LiH', None), ('N2', None), ('benzene', 'ccECP'), ('C', 'ccECP')):
        h = MolecularHamiltonian(mol=Molecule.from_name(name), ecp_type=ecp)
        init = AtomCenteredElectronInitializer()
        g = np.random.default_rng(0)
        r = init(g, h.mol.charges, h.ns_valence, h.mol.coords, h.n_up, h.n_down)
        assert r.shape == (h.n_up + h.n_down, 3) and np.isfinite(r).all()
        el = init.assign_electrons(g, h.ns_valence, h.n_up, h.n_down)
        assert el.sum() == h.n_up + h.n_down
        up, dn = init.assign_spins(g, el, h.mol.coords, h.n_up, h.n_down)
        assert up.sum() == h.n_up and dn.sum() == h.n_down and ((up + dn) == el).all()
    hb = MolecularHamiltonian(mol=Molecule.from_name('benzene'), ecp_type='ccECP')
    assert (hb.n_up, hb.n_down) == (15, 15) and list(hb.ns_valence) == [4.0] * 6 + [1.0] * 6
    assert list(hb.pot.nuc_with_nl_pot) == [0, 1, 2, 3, 4, 5]


def test_clipping_functions_follow_numpy_median_and_quantile_conventions():
    """loss/clip.py:73-174 on one process: medians are jnp.median (mean of the two middle values for an even count),
    quantiles linear-interpolated; restated here with numpy."""
    from deepqmc_b200.energy import clip_local_energy, median_clip_and_mask, median_log_squeeze_and_mask
    from deepqmc_b200.overlap import clip_psi_ratio, psi_ratio_clip_and_mask

    rng = np.random.default_rng(0)
    x = rng.standard_cauchy(size=10)  # even count, heavy tails
    xt = torch.as_tensor(x)
    med = np.median(x)
    assert med != np.sort(x)[4]  # the lower-median convention would differ here
    mad = np.abs(x - med).mean()
    xc, m = median_clip_and_mask(xt, 5.0, exclude_width=3.0)
    assert np.allclose(xc.numpy(), np.clip(x, med - 5 * mad, med + 5 * mad)) and (m.numpy() == (np.abs(x - med) < 3.0)).all()
    xc, m = median_clip_and_mask(xt, 2.0, median_center=False)
    mad_mean = np.abs(x - x.mean()).mean()
    assert np.allclose(xc.numpy(), np.clip(x, x.mean() - 2 * mad_mean, x.mean() + 2 * mad_mean)) and m.all()
    # log-squeeze
    q = np.quantile(np.abs(x - med), 0.95)
    z = (x - med) / (2 * 1.5 * q)
    ls = np.sign(z) * np.log1p((np.abs(z) + 0.5 * z**2 + np.abs(z) ** 3) / (1 + z**2))
    xs, m = median_log_squeeze_and_mask(xt, clip_width=1.5, quantile=0.95, exclude_width=0.8)
    assert np.allclose(xs.numpy(), med + 2 * 1.5 * q * ls) and (m.numpy() == (np.abs(x - med) / q < 0.8)).all()
    # wave-function ratios: median absolute deviation as sigma
    sig = np.median(np.abs(x - med))
    rc, m = psi_ratio_clip_and_mask(xt, clip_width=2.0, exclude_width=4.0)
    assert np.allclose(rc.numpy(), np.clip(x, med - 2 * sig, med + 2 * sig)) and (m.numpy() == (np.abs(x - med) < 4.0)).all()
    # batched application over the leading axes
    E = torch.as_tensor(rng.normal(size=(2, 3, 8)))
    Ec, M = clip_local_energy(lambda e: median_clip_and_mask(e, 1.0), E)
    assert Ec.shape == E.shape and M.shape == E.shape
    assert torch.allclose(Ec[1, 2], median_clip_and_mask(E[1, 2], 1.0)[0])
    Rr = torch.as_tensor(rng.normal(size=(1, 2, 2, 8)))
    Rc, Mr = clip_psi_ratio(psi_ratio_clip_and_mask, Rr)
    assert Rc.shape == Rr.shape and torch.allclose(Rc[0, 1, 0], psi_ratio_clip_and_mask(Rr[0, 1, 0])[0])
