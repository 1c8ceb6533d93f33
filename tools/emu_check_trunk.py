"""Development tool: run the whole-trunk tensor-core kernel (deepqmc_b200/csrc/trunk_tc.cuh) on the CPU model of mbarrier / TMA /
tcgen05 / TMEM / mma.sync (tools/cuda_emu) and compare with an fp64 torch restatement of the layers.  Checks warp roles, barrier
protocol, scratch / fragment layouts and swizzled addressing before GPU minutes are spent; the hardware test is
tests/test_gpu_tcgen05.py::test_fused_trunk_matches_fp64.

    python tools/emu_check_trunk.py [--mol LiH] [--walkers 70] [--layers 2] [--nobuild] [--nsms 2]
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def build(out):
    cmd = ['g++', '-std=c++17', '-O2', '-g', '-DDQMC_EMU', '-x', 'c++', f'-I{ROOT}/tools/cuda_emu', f'-I{ROOT}/include',
           f'-I{ROOT}/deepqmc_b200/csrc', '-fPIC', '-shared', f'{ROOT}/deepqmc_b200/csrc/engine.cu', '-o', out]
    subprocess.check_call(cmd)


def trunk_ref(eng, X0, N, L, H=4, dtype=torch.float64):
    """fp64 restatement: gnn/update_features.py:241-286 + hk.MultiHeadAttention + hkext.py MLP / residuals."""
    flat = torch.as_tensor(eng._flat)

    def W(name):
        off, K, Nc = eng.entries[name]
        return flat[off:off + K * Nc].reshape(K, Nc).float().to(dtype)

    X = X0.to(dtype)
    rows, d = X.shape
    B = rows // N
    dh = d // H
    for l in range(L):
        p = f'L{l}.'
        qkv = X @ W(p + 'wqkv')
        q, k, v = (t.reshape(B, N, H, dh).permute(0, 2, 1, 3) for t in qkv.split(d, dim=1))
        att = torch.softmax(q @ k.transpose(-1, -2) / dh ** 0.5, dim=-1)
        O = (att @ v).permute(0, 2, 1, 3).reshape(rows, d)
        A = X + O @ W(p + 'wo')
        M1 = torch.tanh(A @ W(p + 'w1') + W(p + 'b1')[0])
        X = A + torch.tanh(M1 @ W(p + 'w2') + W(p + 'b2')[0])
    return X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mol', default='LiH')
    ap.add_argument('--walkers', type=int, default=70)
    ap.add_argument('--layers', type=int, default=2)
    ap.add_argument('--lib', default='/tmp/libdqmc_emu_tc.so')
    ap.add_argument('--nobuild', action='store_true')
    ap.add_argument('--nsms', type=int, default=2)
    a = ap.parse_args()
    if not a.nobuild:
        build(a.lib)
    os.environ['DQMC_NSMS'] = str(a.nsms)
    from deepqmc_b200 import params as PN
    from deepqmc_b200.engine import Engine
    from deepqmc_b200.hamil import MolecularHamiltonian
    from deepqmc_b200.molecule import Molecule
    from deepqmc_b200.spec import psiformer_spec

    mol = Molecule.from_name(a.mol)
    h = MolecularHamiltonian(mol=mol, ecp_type='ccECP' if a.mol == 'benzene' else None)
    spec = psiformer_spec(h, embedding_dim=256, n_layers=a.layers, n_heads=4, n_determinants=2)
    params = PN.perturb_params(PN.init_params(spec, 0))
    eng = Engine(spec, h, dtype='float32', gemm_backend=1, _lib_path=a.lib)
    eng.set_params(params)
    N = spec.n_elec
    g = torch.Generator().manual_seed(2)
    X0 = 2 * torch.randn(a.walkers * N, 256, generator=g)
    out = eng.debug_trunk(X0)
    ref = trunk_ref(eng, X0, N, a.layers)
    ref32 = trunk_ref(eng, X0, N, a.layers, dtype=torch.float32)
    err = (out.double() - ref).abs().max().item()
    err32 = (ref32.double() - ref).abs().max().item()
    rms, rms32 = (out.double() - ref).pow(2).mean().sqrt().item(), (ref32.double() - ref).pow(2).mean().sqrt().item()
    print(f'fused trunk {a.mol} N={N} walkers={a.walkers} layers={a.layers}: max abs err {err:.2e} rms {rms:.2e} '
          f'(plain fp32 torch: max {err32:.2e} rms {rms32:.2e}; |ref| max {ref.abs().max().item():.2f})')
    assert err < 10 * err32 + 1e-5 and rms < 5 * rms32 + 1e-6, (err, err32, rms, rms32)
    print('ok')


if __name__ == '__main__':
    main()
