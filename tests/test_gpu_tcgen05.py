"""GPU: the tcgen05 3xTF32 dense-layer GEMM (gemm_tcgen05.cuh) against fp64 matmul and the
CUDA-core kernel, then the whole fp32 engine with the tensor-core backend against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from deepqmc_b200 import params as PN
from deepqmc_b200.ansatz import B200Ansatz
from deepqmc_b200.hamil import MolecularHamiltonian
from deepqmc_b200.molecule import Molecule
from deepqmc_b200.types import PhysicalConfiguration

DEV = 'cuda:0'


def _engine(backend, mol='LiH', **hyper):
    hamil = MolecularHamiltonian(mol=Molecule.from_name(mol))
    a = B200Ansatz(hamil, 'psiformer', dtype='float32', gemm_backend=backend, **hyper)
    params = PN.perturb_params(a.init(0))
    return hamil, a, params, a.engine_for(hamil, params)


@pytest.mark.parametrize('rows,S', [(128, 1), (1000, 14), (56 * 300 + 5, 14), (77, 1)])
@pytest.mark.parametrize('weight,bias', [('L0.wqkv', None), ('L1.w1', 'L1.b1'), ('L2.wo', None)])
def test_gemm_3xtf32_matches_fp64(rows, S, weight, bias):
    hamil, a, params, eng = _engine(1)
    g = torch.Generator(device='cpu').manual_seed(rows + S)
    A = (torch.randn(rows, 256, generator=g) * torch.exp(2 * torch.randn(rows, 1, generator=g))).to(DEV)
    off, K, Nc = eng.entries[weight]
    flat = torch.as_tensor(eng._flat, device=DEV)
    W = flat[off:off + K * Nc].reshape(K, Nc).float().double()
    Res = torch.randn(rows, Nc, generator=g).to(DEV) if weight.endswith('wo') else None
    ref = A.double() @ W
    scale = (A.double().abs() @ W.abs()) + 1e-30  # per-element magnitude of the accumulated terms
    if bias:
        boff, _, bn = eng.entries[bias]
        bvec = flat[boff:boff + bn].float().double()
        ref[torch.arange(rows, device=DEV) % S == 0] += bvec
        scale += bvec.abs()
    if Res is not None:
        ref += Res.double()
        scale += Res.double().abs()
    C_tc = eng.debug_gemm(weight, A, bias=bias, Res=Res, S=S, backend=1)
    C_sm = eng.debug_gemm(weight, A, bias=bias, Res=Res, S=S, backend=0)
    torch.cuda.synchronize()
    err_tc = ((C_tc.double() - ref).abs() / scale).max().item()
    err_sm = ((C_sm.double() - ref).abs() / scale).max().item()
    assert err_sm < 5e-7, err_sm
    assert err_tc < 2e-6, (err_tc, err_sm)  # 3xTF32: ~2^-21 per product, fp32 accumulation


def test_gemm_sliced_backflow_heads():
    hamil, a, params, eng = _engine(1)
    B, N, S = 37, 4, 14
    A = torch.randn(B * N * S, 256, device=DEV)
    C_tc = eng.debug_gemm('bf.up', A, S=S, sliced=True, backend=1)
    C_sm = eng.debug_gemm('bf.up', A, S=S, sliced=True, backend=0)
    flat = torch.as_tensor(eng._flat, device=DEV)
    ws = []
    for nm in ('bf.up', 'bf.dn'):
        off, K, Nc = eng.entries[nm]
        ws.append(flat[off:off + K * Nc].reshape(K, Nc).float().double())
    A4 = A.double().reshape(B, N, S, 256)
    ref = torch.stack([A4[:, i] @ ws[0 if i < hamil.n_up else 1] for i in range(N)], 1).reshape(B * N * S, -1)
    assert (C_sm.double() - ref).abs().max().item() < 1e-4
    assert (C_tc.double() - ref).abs().max().item() < 1e-4


def test_engine_tcgen05_backend_parity():
    """fp32 engine, tensor-core backend: agrees with the CUDA-core fp32 engine and with the fp64
    oracle within the reference's fp32 tolerance (2e-4, tests/test_hamil.py:37-40)."""
    from oracle import wf
    from oracle.hamil import OracleHamiltonian

    hamil, a1, params, e1 = _engine(1)
    _, a0, _, e0 = _engine(0)
    e0.set_params(params)
    mol = hamil.mol
    rng = np.random.default_rng(0)
    B = 64
    r = torch.as_tensor(mol.coords[rng.integers(0, 2, size=(B, 4))] + rng.normal(size=(B, 4, 3)), device=DEV).float()
    R = torch.as_tensor(mol.coords, device=DEV).float()
    E1, st1, s1, l1, _ = e1.local_energy(r, R)
    E0, st0, s0, l0, _ = e0.local_energy(r, R)
    assert torch.equal(s0, s1)
    scale = torch.maximum(E0.abs(), st0[1].abs()).clamp(min=1)
    rel = (E1 - E0).abs() / scale  # two fp32 evaluations: round-off level, tail = ill-conditioned walkers
    assert rel.median().item() < 2e-5 and rel.max().item() < 2e-3, (rel.median().item(), rel.max().item())
    # Metropolis forward (S = 1) goes through the fused-activation epilogue as well
    sf1, lf1 = e1.wf_forward(r, R)
    sf0, lf0 = e0.wf_forward(r, R)
    assert torch.equal(sf0, sf1) and (lf1 - lf0).abs().max().item() < 1e-3
    oh = OracleHamiltonian(mol)
    pt = wf.to_torch(params)
    Rc = R.double().cpu()
    for b in range(3):
        f = lambda x: wf.log_psi(a1.spec, pt, x, Rc)
        eo, st = oh.local_energy(f, r[b].double().cpu(), Rc)
        # fp32 tolerance relative to the magnitudes that cancel in E_kin = -(lap + |grad|^2)/2
        scale_b = max(1, abs(eo.item()), 0.5 * abs(st['hamil/lap'].item()), 0.5 * st['hamil/quantum_force'].item())
        assert abs(E1[b].item() - eo.item()) <= 2e-4 * scale_b


def test_cta_pair_variant_matches_single_cta(monkeypatch):
    """The cta_group::2 variant of the kernel (clusters of 2 CTAs, 256-row pair tiles, half weight tile per CTA,
    multicast commits; engine switch DQMC_GEMM_2CTA read at handle creation) gives the same result as the
    single-CTA kernel: dense GEMM incl. ragged row count and residual, sliced backflow heads, and one fp32
    local-energy evaluation."""
    hamil, a1, params, eng1 = _engine(1)
    monkeypatch.setenv('DQMC_GEMM_2CTA', '1')
    hamil2, a2, _, eng2 = _engine(1)
    g = torch.Generator(device='cpu').manual_seed(5)
    for rows, S, weight, bias in [(56 * 300 + 5, 14, 'L0.wqkv', None), (1000, 14, 'L1.w1', 'L1.b1'), (77, 1, 'L2.wo', None)]:
        A = torch.randn(rows, 256, generator=g).to(DEV)
        Res = torch.randn(rows, 256, generator=g).to(DEV) if weight.endswith('wo') else None
        C1 = eng1.debug_gemm(weight, A, bias=bias, Res=Res, S=S, backend=1)
        C2 = eng2.debug_gemm(weight, A, bias=bias, Res=Res, S=S, backend=1)
        torch.cuda.synchronize()
        assert torch.allclose(C1, C2, rtol=1e-6, atol=1e-5), (weight, (C1 - C2).abs().max().item())
    A = torch.randn(37 * 4 * 14, 256, device=DEV)
    assert torch.allclose(eng1.debug_gemm('bf.up', A, S=14, sliced=True, backend=1),
                          eng2.debug_gemm('bf.up', A, S=14, sliced=True, backend=1), rtol=1e-6, atol=1e-5)
    rng = np.random.default_rng(0)
    mol = hamil.mol
    r = torch.as_tensor(mol.coords[rng.integers(0, 2, size=(64, 4))] + rng.normal(size=(64, 4, 3)), device=DEV, dtype=torch.float32)
    R = torch.as_tensor(mol.coords, device=DEV, dtype=torch.float32)
    pc = PhysicalConfiguration(R, r, torch.zeros(64, device=DEV))
    E1, _ = hamil.local_energy(a1.apply)(None, params, pc)
    E2, _ = hamil2.local_energy(a2.apply)(None, params, pc)
    assert torch.allclose(E1, E2, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('rows', [5, 128, 1000, 148 * 128 * 2 + 77])
def test_fused_mlp_block_matches_fp64(rows):
    """One launch of the fused plain-forward MLP block (fused_tc.cuh: A = X + O Wo, M1 = tanh(A W1 + b1), X' = A + tanh(M1 W2 + b2),
    half hi / lo operands on kind::f16, intermediates in TMEM / shared memory) against fp64; several tiles per CTA for the
    largest row count (barrier phases wrap), a ragged last tile, and in-place use (Out aliases O) as the engine calls it."""
    hamil, a, params, eng = _engine(1)
    g = torch.Generator(device='cpu').manual_seed(rows)
    O = torch.randn(rows, 256, generator=g).to(DEV)
    X = (3 * torch.randn(rows, 256, generator=g)).to(DEV)
    flat = torch.as_tensor(eng._flat, device=DEV)

    def W(name):
        off, K, Nc = eng.entries[name]
        return flat[off:off + K * Nc].reshape(K, Nc).float().double()

    out = eng.debug_mlp_block(2, O, X)
    torch.cuda.synchronize()
    A = X.double() + O.double() @ W('L2.wo')
    M1 = torch.tanh(A @ W('L2.w1') + W('L2.b1')[0])
    ref = A + torch.tanh(M1 @ W('L2.w2') + W('L2.b2')[0])
    err = (out.double() - ref).abs().max().item()
    assert err < 3e-5, err  # fp32 class: |X'| ~ 10, tanh approximation 3e-7, products 2^-22 relative
    # the unfused tensor-core layers (3xTF32 / half GEMMs + epilogues) agree as well
    A32 = eng.debug_gemm('L2.wo', O, Res=X, S=1, backend=1)
    assert (A32.double() - A).abs().max().item() < 1e-5


def test_plain_forward_half_operands_vs_3xtf32(monkeypatch):
    """The S = 1 path with half hi / lo operands + the fused MLP block against the same engine forced back to the 3xTF32
    row GEMMs (DQMC_TC_F16=0, read at handle creation): log|psi| of 256 walkers agrees at fp32 round-off level, and both
    agree with the fp64 CUDA-core engine; large activations (|x| up to ~1e3) stay in range of the scaled halves."""
    hamil, a1, params, e1 = _engine(1)
    monkeypatch.setenv('DQMC_TC_F16', '0')
    _, a0, _, e0 = _engine(1)
    monkeypatch.delenv('DQMC_TC_F16')
    e0.set_params(params)
    a64 = B200Ansatz(hamil, 'psiformer', dtype='float64')
    e64 = a64.engine_for(hamil, params)
    mol = hamil.mol
    rng = np.random.default_rng(3)
    B = 256
    r = torch.as_tensor(mol.coords[rng.integers(0, 2, size=(B, 4))] + rng.normal(size=(B, 4, 3)), device=DEV)
    R = torch.as_tensor(mol.coords, device=DEV)
    s1, l1 = e1.wf_forward(r.float(), R.float())
    s0, l0 = e0.wf_forward(r.float(), R.float())
    s64, l64 = e64.wf_forward(r, R)
    assert torch.equal(s1, s0) and torch.equal(s1.double(), s64)
    d10, d1, d0 = (l1 - l0).abs().max().item(), (l1.double() - l64).abs().max().item(), (l0.double() - l64).abs().max().item()
    assert d10 < 1e-3 and d1 < 1e-3 and d0 < 1e-3, (d10, d1, d0)
    assert d1 < 3 * d0 + 2e-5, (d1, d0)  # the half-operand path is not less accurate than 3xTF32
    A = (torch.randn(300, 256, device=DEV) * 1e3)
    off, K, Nc = e1.entries['L0.wqkv']
    Wq = torch.as_tensor(e1._flat, device=DEV)[off:off + K * Nc].reshape(K, Nc).float().double()
    C = e1.debug_gemm('L0.wqkv', A, S=1, backend=1)
    ref = A.double() @ Wq
    assert ((C.double() - ref).abs() / (A.double().abs() @ Wq.abs())).max().item() < 2e-6
