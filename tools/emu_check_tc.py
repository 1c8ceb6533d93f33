"""Development tool: run the tensor-core kernels (gemm_tcgen05.cuh, fused_tc.cuh) on the CPU model of mbarrier / TMA / tcgen05 /
TMEM in tools/cuda_emu/tcgen05_emu.h and compare with fp64.  NOT part of the product or of any parity claim: it checks warp
roles, barrier protocol, descriptor / swizzle arithmetic and epilogue indexing of the same kernel source before GPU minutes
are spent; the hardware tests are tests/test_gpu_tcgen05.py.

    python tools/emu_check_tc.py [--d 128] [--rows 300] [--nobuild]
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def build(out):
    cmd = ['g++', '-std=c++17', '-O2', '-g', '-DDQMC_EMU', '-x', 'c++', f'-I{ROOT}/tools/cuda_emu', f'-I{ROOT}/include',
           f'-I{ROOT}/deepqmc_b200/csrc', '-fPIC', '-shared', f'{ROOT}/deepqmc_b200/csrc/engine.cu', '-o', out]
    subprocess.check_call(cmd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--d', type=int, default=128)
    ap.add_argument('--rows', type=int, default=300)
    ap.add_argument('--lib', default='/tmp/libdqmc_emu_tc.so')
    ap.add_argument('--nobuild', action='store_true')
    ap.add_argument('--nsms', type=int, default=3)
    a = ap.parse_args()
    if not a.nobuild:
        build(a.lib)
    os.environ['DQMC_NSMS'] = str(a.nsms)  # few persistent CTAs: every CTA loops over several tiles
    from deepqmc_b200 import params as PN
    from deepqmc_b200.engine import Engine
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.spec import psiformer_spec

    mol = Molecule.from_name('LiH')
    h = MolecularHamiltonian(mol=mol)
    d = a.d
    spec = psiformer_spec(h, embedding_dim=d, n_layers=2, n_heads=2, n_determinants=4)
    params = PN.perturb_params(PN.init_params(spec, 0))
    eng = Engine(spec, h, dtype='float32', gemm_backend=1, _lib_path=a.lib)
    eng.set_params(params)
    flat = torch.as_tensor(eng._flat)
    g = torch.Generator().manual_seed(1)

    def W(name):
        off, K, Nc = eng.entries[name]
        return flat[off:off + K * Nc].reshape(K, Nc).float().double()

    worst = 0.0
    for rows, S, weight, bias in [(a.rows, 1, 'L0.wqkv', None), (a.rows, 1, 'L1.w1', 'L1.b1'), (a.rows, 1, 'L1.wo', None),
                                  (3 * 14 * 5, 14, 'L0.wqkv', None), (3 * 14 * 5, 14, 'L1.w1', 'L1.b1')]:
        A = torch.randn(rows, d, generator=g) * torch.exp(1.5 * torch.randn(rows, 1, generator=g))
        Res = torch.randn(rows, W(weight).shape[1], generator=g) if weight.endswith('wo') else None
        ref = A.double() @ W(weight)
        scale = A.double().abs() @ W(weight).abs() + 1e-30
        if bias:
            b = W(bias)[0]
            ref[torch.arange(rows) % S == 0] += b
            scale += b.abs()
        if Res is not None:
            ref += Res.double()
            scale += Res.double().abs()
        for be in (1, 0):
            C = eng.debug_gemm(weight, A, bias=bias, Res=Res, S=S, backend=be)
            err = ((C.double() - ref).abs() / scale).max().item()
            print(f'gemm {weight:8s} rows={rows:5d} S={S:2d} backend={be} ({"F16" if S == 1 and be else "TF32" if be else "SIMT"}): rel err {err:.2e}')
            assert err < 2e-6, err
            worst = max(worst, err)
    # sliced per-spin heads, S = 1 (F16) and S = 14 (TF32)
    for S in (1, 14):
        B, N = 9, spec.n_elec
        A = torch.randn(B * N * S, d, generator=g)
        C1 = eng.debug_gemm('bf.up', A, S=S, sliced=True, backend=1)
        ws = [W('bf.up'), W('bf.dn')]
        A4 = A.double().reshape(B, N, S, d)
        ref = torch.stack([A4[:, i] @ ws[0 if i < spec.n_up else 1] for i in range(N)], 1).reshape(B * N * S, -1)
        err = (C1.double() - ref).abs().max().item()
        print(f'sliced heads S={S}: max abs err {err:.2e}')
        assert err < 1e-4
    # fused MLP block
    if d in (128, 256):
        for rows in (a.rows, 128, 5):
            O = torch.randn(rows, d, generator=g)
            X = torch.randn(rows, d, generator=g) * 2
            out = eng.debug_mlp_block(1, O, X)
            A_ = X.double() + O.double() @ W('L1.wo')
            M1 = torch.tanh(A_ @ W('L1.w1') + W('L1.b1')[0])
            ref = A_ + torch.tanh(M1 @ W('L1.w2') + W('L1.b2')[0])
            err = (out.double() - ref).abs().max().item()
            print(f'fused MLP block rows={rows}: max abs err {err:.2e} (|ref| max {ref.abs().max().item():.2f})')
            assert err < 2e-5, err
    # whole engine: fp32 tensor-core path (emulated) vs the fp64 SIMT engine
    eng64 = Engine(spec, h, dtype='float64', _lib_path=a.lib)
    eng64.set_params(params)
    rng = np.random.default_rng(0)
    Bw = 40
    r = torch.as_tensor(mol.coords[rng.integers(0, 2, size=(Bw, spec.n_elec))] + rng.normal(size=(Bw, spec.n_elec, 3)))
    R = torch.as_tensor(mol.coords)
    s1, l1 = eng.wf_forward(r.float(), R.float())
    s0, l0 = eng64.wf_forward(r, R)
    err = (l1.double() - l0).abs().max().item()
    print(f'wf_forward fp32 tensor-core path vs fp64: max |dlog psi| {err:.2e}, signs equal: {bool((s1.double() == s0).all())}')
    assert err < 5e-4 and bool((s1.double() == s0).all())
    E1 = eng.local_energy(r[:4].float(), R.float())[0]
    E0 = eng64.local_energy(r[:4], R)[0]
    print('E_loc fp32 (tc):', E1.numpy(), ' fp64:', E0.numpy())
    print('all ok, worst gemm rel err', worst)


if __name__ == '__main__':
    main()
