"""Host-side mirror of the reference's electron samplers (reference:
src/deepqmc/sampling/base.py:14-79 ElectronSampler protocol, electron_samplers.py:32-173
MetropolisSampler, :333-357 DecorrSampler, electron_sample_initializers.py:43-288 initialiser).

``sampler.init(rng, params, n, R)``, ``sampler.sample(rng, state, params, R) -> (state,
phys_conf, stats)`` and ``sampler.update(state, params, R)`` keep the reference's meaning; the
sweep itself (proposal, wave-function forward, accept/reject, age, tau adaptation, statistics)
runs in ``dqmc_mcmc_sweep``.  State dict: r[B,N,3], psi=Psi(sign[B], log[B]), age[B] int32, tau.
``rng`` is an integer seed (Philox key) instead of a JAX key.
"""
from __future__ import annotations

import numpy as np
import torch

from .types import PhysicalConfiguration, Psi

STAT_NAMES = ('sampling/acceptance', 'sampling/tau', 'sampling/age/mean', 'sampling/age/max',
              'sampling/log_psi/mean', 'sampling/log_psi/std', 'sampling/dists/mean')


class ShellBasedDistribution:
    """|r| ~ Exp / (2 zeta), zeta = Z * shell factor, uniform direction
    (reference: electron_sample_initializers.py:196-255)."""

    @staticmethod
    def shell_factor(spin_idx):
        return np.where(spin_idx < 1, 1.0, np.where(spin_idx < 5, 0.5, np.where(spin_idx < 9, 1 / 3, 0.25)))

    def __call__(self, rng: np.random.Generator, charges, counts_per_nucleus):
        total = len(charges)
        # index of each electron among the same-spin electrons of its nucleus
        spin_idx = np.concatenate([np.arange(c) for c in counts_per_nucleus if c > 0]) if total else np.zeros(0)
        zetas = np.asarray(charges) * self.shell_factor(spin_idx)
        dist = rng.exponential(size=total) / (2 * zetas)
        v = rng.normal(size=(total, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        return dist[:, None] * v


class AtomCenteredElectronInitializer:
    """Places electrons around nuclei (reference: electron_sample_initializers.py:258-288 with the
    assignment heuristics of :43-160: electrons per nucleus by valence charge, electron pairs
    first, remaining spins alternating between nearest neighbours)."""

    def __init__(self, atom_centered_distribution=None):
        self.dist = atom_centered_distribution or ShellBasedDistribution()

    @staticmethod
    def assign_electrons(rng, ns_valence, n_up, n_down):
        ns_valence = np.asarray(ns_valence, dtype=float)
        charge = ns_valence.sum() - n_up - n_down
        valence = ns_valence - charge / len(ns_valence)
        el = np.floor(valence).astype(int)
        while ns_valence.sum() - charge - el.sum() > 0:
            logits = valence - el
            p = np.exp(logits - logits.max())
            el[rng.choice(len(el), p=p / p.sum())] += 1
        return el

    @staticmethod
    def assign_spins(rng, el, coords, n_up, n_down):
        up, down = np.zeros_like(el), np.zeros_like(el)
        for i in range(int(el.max()) if len(el) else 0):
            mask = el >= 2 * (i + 1)
            inc = np.where(mask & (mask.sum() + down.sum() <= n_down), 1, 0)
            up, down = up + inc, down + inc
        d = np.linalg.norm(coords[:, None] - coords[None], axis=-1)
        np.fill_diagonal(d, np.inf)
        nn = np.argsort(d, axis=-1)
        rem = el - up - down
        if (rem > 0).any():
            cands = np.flatnonzero(rem == rem.max())
            center = int(rng.choice(cands))
            i = 0
            while (up + down < el).any():
                is_down = (i % 2) and (down.sum() < n_down)
                if is_down:
                    down[center] += 1
                else:
                    up[center] += 1
                order = nn[center]
                has = (el - up - down)[order] > 0
                center = int(order[int(np.argmax(has))])
                i += 1
        return up, down

    def __call__(self, rng, charges, ns_valence, nuclear_coordinates, n_up, n_down):
        coords = np.asarray(nuclear_coordinates)
        el = self.assign_electrons(rng, ns_valence, n_up, n_down)
        up, down = self.assign_spins(rng, el, coords, n_up, n_down)
        up_idx, dn_idx = np.repeat(np.arange(len(el)), up)[:n_up], np.repeat(np.arange(len(el)), down)[:n_down]
        ch = np.asarray(charges)
        r_up = coords[up_idx] + self.dist(rng, ch[up_idx], up)
        r_dn = coords[dn_idx] + self.dist(rng, ch[dn_idx], down)
        return np.concatenate([r_up, r_dn])


def _as_jax_key(rng):
    """int seed -> PRNGKey(seed); a uint32[2] key (numpy) passes through (keys derived by split in combined samplers)."""
    from . import jaxrand

    if isinstance(rng, np.ndarray) and rng.dtype == np.uint32 and rng.shape == (2,):
        return rng
    return jaxrand.prng_key(int(rng))


def _rank_key(key):
    """This rank's key of a uint32[2] JAX key: the reference pmaps its samplers over per-device keys split(rng, device_count)
    (sampling_utils.py:236-262, parallel.py:126-133); with one process per GPU that is split(key, world)[rank]."""
    from . import jaxrand, parallel

    rank, world = parallel.world()
    return key if world == 1 else jaxrand.split(key, world)[rank]


def _default_offset(walker_offset, n_local):
    """Global index of this rank's first walker: keys the in-kernel Philox streams, so ranks that are handed the SAME integer
    seed still draw independent noise (contiguous walker blocks, deepqmc_b200.parallel.shard_bounds)."""
    if walker_offset is not None:
        return walker_offset
    from . import parallel

    return parallel.world()[0] * n_local


class JaxCompatibleElectronInitializer:
    """The reference's AtomCenteredElectronInitializer(ShellBasedDistribution()) driven by a numpy restatement of its
    jax.random streams (deepqmc_b200/jaxrand.py): ``sampler.init(seed, ...)`` then returns the SAME walkers as the
    reference's ``sampler.init(jax.random.PRNGKey(seed), ...)`` (walker b from key ``split(PRNGKey(seed), n)[b]``,
    electron_samplers.py:84-100).  Slower than the default numpy initialiser (a few ms per walker, one-off)."""

    jax_compatible = True

    def walkers(self, seed, n, charges, ns_valence, R, n_up, n_down):
        from . import jaxrand

        keys = jaxrand.split(_as_jax_key(seed), n)
        return np.stack([jaxrand.atom_centered_initializer(k, charges, ns_valence, R, n_up, n_down) for k in keys])


class MetropolisSampler:
    """reference: sampling/electron_samplers.py:32-173."""

    WALKER_STATE = ['r', 'psi', 'age']

    def __init__(self, hamil, wf, *, sample_initializer=None, tau=1.0, target_acceptance=0.57, max_age=None,
                 jax_compatible_noise=False):
        self.hamil = hamil
        self.wf = wf  # bound B200Ansatz.apply
        self.ansatz = wf.__self__
        self.sample_initializer = sample_initializer or AtomCenteredElectronInitializer()
        self.initial_tau, self.target_acceptance, self.max_age = tau, target_acceptance, max_age
        self.length = 1
        self._step = 0
        # True: sample(seed, ...) draws its proposal / acceptance numbers from the reference's jax.random streams for
        # PRNGKey(seed) (host-side restatement, deepqmc_b200/jaxrand.py) instead of the in-kernel Philox generator, i.e. the
        # Markov chain of the reference's sampler.sample(PRNGKey(seed), ...) -- a parity mode, not the fast path
        self.jax_compatible_noise = jax_compatible_noise
        self.exchange_step_probability = 0.0  # set by chain(..., OppositeSpinExchangeSampler(...), MetropolisSampler(...))

    def _jax_noise_exchange(self, rng, shape_r, device, dtype):
        """The reference's streams with an OppositeSpinExchangeSampler in the chain (electron_samplers.py:286-330): every step
        splits its key three ways (exchange decision, proposal, acceptance); an exchange step draws the up / down indices from
        categorical(split(rng_prop)) with uniform logits.  -> (normal, uniform, flags[length], idx[length, B, 2])"""
        from . import jaxrand

        key = _rank_key(_as_jax_key(rng))
        subkeys = jaxrand.split(key, self.length) if self.length > 1 else [key]
        B, n_up, n_dn = shape_r[0], self.hamil.n_up, self.hamil.n_down
        nn, nu, flags, idx = [], [], [], []
        for k in subkeys:
            ke, kp, ka = jaxrand.split(k, 3)
            ex = bool(jaxrand.uniform(ke, ()) < self.exchange_step_probability)
            flags.append(int(ex))
            if ex:
                ku, kd = jaxrand.split(kp, 2)
                up = np.argmax(jaxrand.gumbel(ku, (B, n_up)), -1)
                dn = np.argmax(jaxrand.gumbel(kd, (B, n_dn)), -1)
                idx.append(np.stack([up, dn], -1))
                nn.append(np.zeros(tuple(shape_r)))
            else:
                idx.append(np.zeros((B, 2), dtype=np.int64))
                nn.append(jaxrand.normal(kp, tuple(shape_r)))
            nu.append(jaxrand.uniform(ka, (B,)))
        return (torch.as_tensor(np.stack(nn), device=device, dtype=dtype), torch.as_tensor(np.stack(nu), device=device, dtype=dtype),
                flags, torch.as_tensor(np.stack(idx), dtype=torch.int32, device=device))

    def _jax_noise(self, rng, shape_r, device, dtype):
        """(normal[length, B, N, 3], uniform[length, B]) exactly as the reference draws them: DecorrSampler scans over
        split(rng, length); every Metropolis / Langevin step splits its key into (proposal, acceptance)
        (electron_samplers.py:140-147,347-353)."""
        from . import jaxrand

        key = _rank_key(_as_jax_key(rng))
        subkeys = jaxrand.split(key, self.length) if self.length > 1 else [key]
        B = shape_r[0]
        nn = np.stack([jaxrand.normal(jaxrand.split(k, 2)[0], tuple(shape_r)) for k in subkeys])
        nu = np.stack([jaxrand.uniform(jaxrand.split(k, 2)[1], (B,)) for k in subkeys])
        return torch.as_tensor(nn, device=device, dtype=dtype), torch.as_tensor(nu, device=device, dtype=dtype)

    def phys_conf(self, R, r):
        if r.dim() == 2:
            return PhysicalConfiguration(R, r, torch.zeros((), device=r.device))
        return PhysicalConfiguration(R, r, torch.zeros(len(r), dtype=torch.int32, device=r.device))

    def _engine(self, params):
        return self.ansatz.engine_for(self.hamil, params)

    def update(self, state, params, R):
        eng = self._engine(params)
        sign, log = eng.wf_forward(state['r'], R)
        return {**state, 'psi': Psi(sign, log)}

    def init(self, rng, params, n, R):
        eng = self._engine(params)
        jaxc = getattr(self.sample_initializer, 'jax_compatible', False)
        g = None if jaxc else np.random.default_rng(int(rng))
        h = self.hamil
        Rn = np.asarray(R.detach().cpu() if torch.is_tensor(R) else R, dtype=np.float64)
        if jaxc:  # the reference's walkers for the same seed
            r = self.sample_initializer.walkers(rng, n, h.mol.charges, h.ns_valence, Rn, h.n_up, h.n_down)
        else:
            r = np.stack([self.sample_initializer(g, h.mol.charges, h.ns_valence, Rn, h.n_up, h.n_down) for _ in range(n)])
        state = {
            'r': torch.as_tensor(r, dtype=eng.dtype, device=eng.device),
            'age': torch.zeros(n, dtype=torch.int32, device=eng.device),
            'tau': torch.tensor([self.initial_tau], dtype=eng.dtype, device=eng.device),
        }
        return self.update(state, params, torch.as_tensor(Rn, dtype=eng.dtype, device=eng.device))

    def sample(self, rng, state, params, R, *, walker_offset=None, noise_normal=None, noise_uniform=None, exchange_flags=None,
               exchange_idx=None):
        """``rng`` may be the same on every rank: the Philox streams are keyed by the global walker index (``walker_offset``
        defaults to rank * local batch), a uint32[2] key is split per rank as the reference's pmap does."""
        eng = self._engine(params)
        walker_offset = _default_offset(walker_offset, state['r'].shape[0])
        xkw = {}
        if self.exchange_step_probability > 0.0:
            xkw = {'exchange_probability': self.exchange_step_probability, 'exchange_flags': exchange_flags, 'exchange_idx': exchange_idx}
        if self.jax_compatible_noise and noise_normal is None:
            if self.exchange_step_probability > 0.0:
                noise_normal, noise_uniform, fl, ix = self._jax_noise_exchange(rng, state['r'].shape, state['r'].device, state['r'].dtype)
                xkw.update(exchange_flags=fl, exchange_idx=ix)
            else:
                noise_normal, noise_uniform = self._jax_noise(rng, state['r'].shape, state['r'].device, state['r'].dtype)
        st = {'r': state['r'], 'sign': state['psi'].sign, 'log': state['psi'].log, 'age': state['age'], 'tau': state['tau']}
        stats = eng.mcmc_sweep(st, R, self.length, target_acceptance=self.target_acceptance, max_age=self.max_age,
                               seed=0 if isinstance(rng, np.ndarray) else int(rng), step0=self._step, walker_offset=walker_offset,
                               noise_normal=noise_normal, noise_uniform=noise_uniform, **xkw)
        self._step += self.length
        new = {'r': st['r'], 'psi': Psi(st['sign'], st['log']), 'age': st['age'], 'tau': st['tau']}
        return new, self.phys_conf(R, new['r']), dict(zip(STAT_NAMES, stats))


class DecorrSampler(MetropolisSampler):
    """DecorrSampler(length) chained in front of MetropolisSampler (reference:
    electron_samplers.py:333-357; `chain(DecorrSampler(length=30), MetropolisSampler(...))` in
    conf/task/sampler_factory/elec_sampler/decorr_metropolis_psiformer.yaml): `length` sub-steps per
    sample(), statistics of the last one.  The whole scan is one dqmc_mcmc_sweep call."""

    def __init__(self, hamil=None, wf=None, *, length, **kw):
        if hamil is None:  # reference form DecorrSampler(length=...): a link for chain(), not a sampler by itself
            self.hamil, self.length = None, int(length)
            return
        super().__init__(hamil, wf, **kw)
        self.length = int(length)


class OppositeSpinExchangeSampler:
    """Link for chain(): with probability ``exchange_step_probability`` a sampling step proposes, for every walker, to swap the
    positions of a random spin-up / spin-down electron pair instead of the Gaussian move (reference:
    sampling/electron_samplers.py:235-330; conf/task/sampler_factory/elec_sampler/decorr_spin_exchange_metropolis.yaml).
    Uniform pair selection (the reference's default logits); custom logits functions are not supported."""

    def __init__(self, *, exchange_step_probability: float, up_logits_fn=None, down_logits_fn=None):
        if up_logits_fn is not None or down_logits_fn is not None:
            raise NotImplementedError('custom spin-exchange logits')
        self.exchange_step_probability = float(exchange_step_probability)


def chain(*samplers):
    """chain(DecorrSampler(length=30), MetropolisSampler(hamil, wf, tau=1.0)) -> a sampler that returns every 30th
    Metropolis step (reference: sampling/sampling_utils.py:31-54, conf/task/sampler_factory/elec_sampler/*.yaml).  The last
    link must be a Metropolis or Langevin sampler; the engine runs the whole scan as one sweep call."""
    last = samplers[-1]
    assert isinstance(last, MetropolisSampler) and last.hamil is not None, 'the last link must be a Metropolis / Langevin sampler'
    for link in samplers[:-1]:
        if isinstance(link, OppositeSpinExchangeSampler):
            assert not isinstance(last, LangevinSampler), 'spin-exchange steps are implemented for the Metropolis sampler'
            last.exchange_step_probability = link.exchange_step_probability
            continue
        assert isinstance(link, DecorrSampler) and link.hamil is None, 'only DecorrSampler(length=...) links can be chained in front'
        last.length = last.length * link.length
    return last


def combine_samplers(samplers, hamil, wf):
    """reference: sampling/sampling_utils.py:57-69 -- the last entry is a constructor taking (hamil, wf)"""
    return chain(*samplers[:-1], samplers[-1](hamil, wf))


def equilibrate(rng, params, molecule_idx_sampler, sampler, state, criterion, steps, *, block_size, n_blocks=5,
                allow_early_stopping=True):
    """Generator: sample until ``criterion(phys_conf)`` has stabilised -- once block_size * n_blocks values are buffered,
    stop when the means of the first and last block differ by less than the smaller of their standard deviations
    (reference: sampling/sampling_utils.py:104-162).  Yields (step, state, mol_idxs, stats)."""
    from statistics import mean, stdev

    buffer_size = block_size * n_blocks
    buffer: list = []
    for k, step in enumerate(steps):
        mol_idxs = molecule_idx_sampler.sample()
        state, phys_conf, stats = sampler.sample(int(rng) * 1000003 + k, state, params, mol_idxs)
        yield step, state, mol_idxs, stats
        if allow_early_stopping:
            buffer = [*buffer[-buffer_size + 1:], float(criterion(phys_conf))]
            if len(buffer) < buffer_size:
                continue
            b1, b2 = buffer[:block_size], buffer[-block_size:]
            if abs(mean(b1) - mean(b2)) < min(stdev(b1), stdev(b2)):
                break


class MultiElectronicStateSampler:
    """Sample from several electronic states side by side (reference:
    sampling/combined_samplers.py:58-90: vmap of the underlying sampler over the state axis).
    ``params`` is a sequence with one parameter tree per state; the state dicts are kept in a list;
    the returned PhysicalConfiguration has a leading state axis [n_state, B, ...]."""

    def __init__(self, sampler: MetropolisSampler, n_state: int):
        self.sampler, self.n_state = sampler, n_state

    def _seeds(self, rng):
        """per-state seeds: the reference's split(rng, n_state) keys in the seed-compatible mode, integers otherwise"""
        if isinstance(rng, np.ndarray) or getattr(self.sampler, 'jax_compatible_noise', False):
            from . import jaxrand

            return list(jaxrand.split(_as_jax_key(rng), self.n_state))
        return [int(rng) * self.n_state + s for s in range(self.n_state)]

    def init(self, rng, params, electron_batch_size, R):
        assert len(params) == self.n_state
        seeds = self._seeds(rng)
        return [self.sampler.init(seeds[s], params[s], electron_batch_size, R) for s in range(self.n_state)]

    def update(self, state, params, R):
        return [self.sampler.update(state[s], params[s], R) for s in range(self.n_state)]

    def sample(self, rng, state, params, R, **kw):
        new, rs, stats = [], [], []
        seeds = self._seeds(rng)
        for s in range(self.n_state):
            st, pc, stt = self.sampler.sample(seeds[s], state[s], params[s], R, **kw)
            new.append(st); rs.append(pc.r); stats.append(stt)
        r = torch.stack(rs)
        pc = PhysicalConfiguration(R, r, torch.zeros(r.shape[:2], dtype=torch.int32, device=r.device))
        return new, pc, {k: torch.stack([s[k] for s in stats]) for k in stats[0]}


class LangevinSampler(MetropolisSampler):
    """Metropolis-adjusted Langevin sampler (reference: sampling/electron_samplers.py:176-232): the walker state
    carries the drift ``force`` = clean_force(grad log|psi|) (sampling_utils.py:71-101); proposal, Green's-function
    acceptance and the force clean-up run in ``dqmc_langevin_sweep``; every sub-step costs one forward-Laplacian
    pass.  ``length`` sub-steps per sample() (use ``length > 1`` for a Decorr-chained sampler)."""

    WALKER_STATE = MetropolisSampler.WALKER_STATE + ['force']

    def __init__(self, hamil, wf, *, length=1, **kw):
        super().__init__(hamil, wf, **kw)
        self.length = int(length)

    def update(self, state, params, R):
        eng = self._engine(params)
        r = state['r']
        st = {'r': r, 'sign': torch.empty(len(r), dtype=eng.dtype, device=eng.device),
              'log': torch.empty(len(r), dtype=eng.dtype, device=eng.device), 'force': torch.empty_like(r),
              'age': state['age'], 'tau': state['tau']}
        eng.langevin_sweep(st, R, 0)
        return {**state, 'psi': Psi(st['sign'], st['log']), 'force': st['force']}

    def sample(self, rng, state, params, R, *, walker_offset=None, noise_normal=None, noise_uniform=None):
        eng = self._engine(params)
        walker_offset = _default_offset(walker_offset, state['r'].shape[0])
        if self.jax_compatible_noise and noise_normal is None:
            noise_normal, noise_uniform = self._jax_noise(rng, state['r'].shape, state['r'].device, state['r'].dtype)
        st = {'r': state['r'], 'sign': state['psi'].sign, 'log': state['psi'].log, 'force': state['force'],
              'age': state['age'], 'tau': state['tau']}
        stats = eng.langevin_sweep(st, R, self.length, target_acceptance=self.target_acceptance, max_age=self.max_age,
                                   seed=0 if isinstance(rng, np.ndarray) else int(rng), step0=self._step, walker_offset=walker_offset, noise_normal=noise_normal,
                                   noise_uniform=noise_uniform)
        self._step += self.length
        new = {'r': st['r'], 'psi': Psi(st['sign'], st['log']), 'force': st['force'], 'age': st['age'], 'tau': st['tau']}
        return new, self.phys_conf(R, new['r']), dict(zip(STAT_NAMES, stats))


class MoleculeIdxSampler:
    """Cycles through the molecule indices of a transferable run in batches of ``batch_size``, optionally through a
    random permutation drawn once or after every pass (reference: sampling/combined_samplers.py:17-55)."""

    def __init__(self, rng, n_mols: int, batch_size: int, shuffle=False):
        assert shuffle in [False, 'once', 'always']
        self.rng = np.random.default_rng(int(rng))
        self.n_mols, self.batch_size, self.shuffle, self.state = n_mols, batch_size, shuffle, 0
        self._once = None
        self.permutation = self.new_permutation()

    def new_permutation(self):
        if not self.shuffle:
            return np.arange(self.n_mols)
        if self.shuffle == 'once':  # the same permutation after every pass (the reference re-uses its key)
            if self._once is None:
                self._once = self.rng.permutation(self.n_mols)
            return self._once
        return self.rng.permutation(self.n_mols)

    def sample(self):
        idx = np.arange(self.state, min(self.state + self.batch_size, self.n_mols))
        value = [self.permutation[idx]]
        if len(idx) < self.batch_size:  # wrap around into the next pass
            self.permutation = self.new_permutation()
            value.append(self.permutation[np.arange(self.batch_size - len(idx))])
        self.state = (self.state + self.batch_size) % self.n_mols
        return np.concatenate(value)


class IdleNucleiSampler:
    """Keeps the nuclei where they are (reference: sampling/nuclei_samplers.py:16-36)."""

    def __init__(self, charges=None):
        pass

    def init(self, nuc_coords, *args, **kwargs):
        return {'R': nuc_coords}

    def sample(self, rng, state):
        return state, torch.zeros_like(state['R']), {}


def no_elec_warp(rng, R, dR, smpl_state):
    """reference: sampling/nuclei_samplers.py:175-190"""
    return smpl_state


def nn_elec_warp(rng, R, dR, smpl_state):
    """Move every electron with its nearest nucleus (reference: sampling/nuclei_samplers.py:193-213).  ``smpl_state`` is
    the state list of a MultiElectronicStateSampler (or a single sampler state)."""
    states = smpl_state if isinstance(smpl_state, list) else [smpl_state]
    R_old = R - dR
    for st in states:
        nearest = torch.cdist(st['r'], R_old.to(st['r'])[None].expand(len(st['r']), -1, -1)).argmin(-1)  # [B, N]
        st['r'] = st['r'] + dR.to(st['r'])[nearest]
    return smpl_state


class MultiNuclearGeometrySampler:
    """Electron samplers of several nuclear geometries side by side (reference: sampling/combined_samplers.py:93-214).

    The reference vmaps one sampler over the molecule axis; here every geometry keeps its own electron-sampler state
    (a list, like the state axis of ``MultiElectronicStateSampler``) and each (geometry, state) block is one sweep of
    the CUDA engine with that geometry's ``R``.  ``sample`` returns the reference's batch layout
    ``[mol_batch, n_state, B, ...]`` with ``R`` tiled per walker and ``mol_idx`` filled in.
    State: ``{'nuc': [nuc state per molecule], 'elec': [elec state per molecule], 'update_nuc_counter': int64[n_mol]}``.
    """

    def __init__(self, elec_sampler, nuc_sampler=None, warp_elec_fn=None, update_nuc_period=None,
                 elec_equilibration_steps=None):
        self.elec_sampler = elec_sampler
        self.nuc_sampler = nuc_sampler or IdleNucleiSampler()
        self.warp_elec_fn = warp_elec_fn or no_elec_warp
        self.update_nuc_period, self.elec_equilibration_steps = update_nuc_period, elec_equilibration_steps

    def init(self, rng, params, electron_batch_size, R):
        R = torch.as_tensor(np.asarray(R, dtype=np.float64)) if not torch.is_tensor(R) else R
        n_mol = len(R)
        if self._jax():  # reference key layout: one key per molecule from split(rng, n_mol) (combined_samplers.py:128-131)
            from . import jaxrand

            seeds = list(jaxrand.split(_as_jax_key(rng), n_mol))
        else:
            seeds = [int(rng) * n_mol + m for m in range(n_mol)]
        return {
            'nuc': [self.nuc_sampler.init(R[m]) for m in range(n_mol)],
            'elec': [self.elec_sampler.init(seeds[m], params, electron_batch_size, R[m]) for m in range(n_mol)],
            'update_nuc_counter': torch.zeros(n_mol, dtype=torch.int64),
        }

    def _jax(self):
        inner = getattr(self.elec_sampler, 'sampler', self.elec_sampler)  # through a MultiElectronicStateSampler
        return getattr(inner, 'jax_compatible_noise', False)

    def _R_dev(self, elec_state, R):
        st0 = elec_state[0] if isinstance(elec_state, list) else elec_state
        return R.to(device=st0['r'].device, dtype=st0['r'].dtype)

    def update_nuc(self, rng, nuc_state, elec_state, params):
        """One molecule: propose new nuclei, warp the electrons along, refresh psi, re-equilibrate (:130-160).  ``rng`` must be
        the same on every rank (the reference aligns it with align_rng_key_across_devices): the nuclear move is then identical
        everywhere, while the electron re-equilibration still draws per-rank noise through the walker offset."""
        nuc_state, dR, stats = self.nuc_sampler.sample(int(rng), nuc_state)
        R = nuc_state['R']
        elec_state = self.warp_elec_fn(int(rng) + 1, R, dR, elec_state)
        Rd = self._R_dev(elec_state, R)
        elec_state = self.elec_sampler.update(elec_state, params, Rd)
        for i in range(self.elec_equilibration_steps or 0):
            elec_state = self.elec_sampler.sample(int(rng) * 7919 + i, elec_state, params, Rd)[0]
        return nuc_state, elec_state, stats

    def sample(self, rng, smpl_state, params, mol_idxs, *, noise_normal=None, noise_uniform=None):
        """``noise_normal`` / ``noise_uniform``: optional per-sampled-molecule lists of injected random numbers handed to the
        electron sampler (parity tests replay the reference's streams)."""
        mol_idxs = [int(m) for m in np.asarray(mol_idxs).reshape(-1)]
        counter = smpl_state['update_nuc_counter']
        rs, Rs, stats = [], [], []
        if self._jax():  # rngs_elec, rngs_nuc = split(rng, (2, n)): electron keys first (combined_samplers.py:173)
            from . import jaxrand

            elec_seeds = list(jaxrand.split(_as_jax_key(rng), 2 * len(mol_idxs))[:len(mol_idxs)])
        else:
            elec_seeds = [int(rng) * len(mol_idxs) + k for k in range(len(mol_idxs))]
        for k, m in enumerate(mol_idxs):
            if self.update_nuc_period is not None:
                if int(counter[m]) == self.update_nuc_period - 1:
                    smpl_state['nuc'][m], smpl_state['elec'][m], _ = self.update_nuc(
                        (0 if isinstance(rng, np.ndarray) else int(rng)) * 104729 + k, smpl_state['nuc'][m], smpl_state['elec'][m], params)
                    counter[m] = 0
                else:
                    counter[m] += 1
            Rd = self._R_dev(smpl_state['elec'][m], smpl_state['nuc'][m]['R'])
            kw = {} if noise_normal is None else {'noise_normal': noise_normal[k], 'noise_uniform': noise_uniform[k]}
            smpl_state['elec'][m], pc, st = self.elec_sampler.sample(elec_seeds[k], smpl_state['elec'][m], params, Rd, **kw)
            r = pc.r if pc.r.dim() == 4 else pc.r[None]  # [n_state, B, N, 3]
            rs.append(r)
            Rs.append(Rd[None, None].expand(*r.shape[:2], *Rd.shape))
            stats.append(st)
        r, R = torch.stack(rs), torch.stack(Rs)
        mol_idx = torch.as_tensor(mol_idxs, dtype=torch.int32, device=r.device)[:, None, None].expand(r.shape[:3])
        out_stats = {k: torch.stack([torch.as_tensor(s[k]) for s in stats]) for k in stats[0]} if stats else {}
        return smpl_state, PhysicalConfiguration(R, r, mol_idx), out_stats

    def update(self, smpl_state, params):
        for m in range(len(smpl_state['elec'])):
            smpl_state['elec'][m] = self.elec_sampler.update(
                smpl_state['elec'][m], params, self._R_dev(smpl_state['elec'][m], smpl_state['nuc'][m]['R']))
        return smpl_state


def initialize_sampling(rng, hamil, ansatz, mols, electronic_states, molecule_batch_size, *, elec_sampler, nuc_sampler=None,
                        elec_warp_fn=None, update_nuc_period=None, elec_equilibration_steps=None):
    """The reference's SamplerFactory (sampling/sampling_utils.py:165-233, types.py:83-93): -> (MoleculeIdxSampler,
    MultiNuclearGeometrySampler over a MultiElectronicStateSampler).  ``elec_sampler``: a callable taking (hamil=, wf=), e.g.
    ``partial(combine_samplers, [DecorrSampler(length=30), partial(MetropolisSampler, tau=1.0)])``."""
    molecule_idx_sampler = MoleculeIdxSampler(rng if not isinstance(rng, np.ndarray) else int(rng[1]), len(mols), molecule_batch_size, 'once')
    elec = elec_sampler(hamil=hamil, wf=ansatz.apply)
    multi_state = MultiElectronicStateSampler(elec, electronic_states)
    nuc = (IdleNucleiSampler if nuc_sampler is None else nuc_sampler)(hamil.mol.charges)
    sampler = MultiNuclearGeometrySampler(multi_state, nuc, no_elec_warp if elec_warp_fn is None else elec_warp_fn,
                                          update_nuc_period, elec_equilibration_steps)
    return molecule_idx_sampler, sampler


def initialize_sampler_state(rng, sampler, params, electron_batch_size, nuc_coords):
    """sampler.init with this process's share of the walkers: the reference pmaps over its local devices
    (sampling_utils.py:236-262), here there is one process per GPU (deepqmc_b200.parallel)."""
    from . import parallel

    rank, world = parallel.world()
    assert electron_batch_size % world == 0, 'electron_batch_size must be divisible by the number of devices'  # validate_kwargs.py:45-48
    # per-rank walkers: integer seeds are offset by the rank (train.py:134), uint32[2] keys are split over the ranks
    seed = _rank_key(rng) if isinstance(rng, np.ndarray) else parallel.rank_seed(rng, rank)
    return sampler.init(seed, params, electron_batch_size // world, nuc_coords)

