"""Thin Python owner of one ``dqmc_handle``: config marshalling, parameter packing, workspace.

Everything numerical happens inside libdqmc_b200.so; this module only moves pointers.  torch is
used for device memory and streams (plumbing).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import params as PN
from .spec import AnsatzSpec

MODE_FORWARD, MODE_LOCAL_ENERGY, MODE_VJP, MODE_MCMC, MODE_LANGEVIN = 0, 1, 2, 3, 4
_TORCH_DTYPE = {0: torch.float64, 1: torch.float32}


def paulinet_backflow_hidden(spec: AnsatzSpec) -> list[int]:
    """Hidden widths of the per-spin backflow MLPs, padded to the wider spin."""
    du, dd = PN.backflow_dims(spec, spec.n_up)[:-1], PN.backflow_dims(spec, spec.n_down)[:-1]
    return [max(a, b) for a, b in zip(du, dd)]


def paulinet_env_rep(spec: AnsatzSpec) -> int:
    """Envelope terms per nucleus in the engine layout = the largest number of shells on one nucleus."""
    return max(max((spec.env_centers.count(c) for c in set(spec.env_centers)), default=1), 1)


def _pack_haiku_params(spec: AnsatzSpec, params: dict, R=None) -> dict[str, np.ndarray]:
    """Haiku-named tree (deepqmc_b200.params) -> the engine's packed entries."""
    g = lambda k: np.asarray(params[k], dtype=np.float64)
    out = {}
    if spec.kind in ('psiformer', 'transpsiformer'):
        out['emb.w'] = g(PN.GNN + 'electron_embedding/linear:w')
        for l in range(spec.n_layers):
            a = PN.attn_prefix(l) if spec.kind == 'psiformer' else PN.comb_prefix(l)
            out[f'L{l}.wqkv'] = np.concatenate(
                [g(a + f'multi_head_attention/{n}:w') for n in ('query', 'key', 'value')], axis=1
            )
            out[f'L{l}.wo'] = g(a + 'multi_head_attention/linear:w')
            out[f'L{l}.w1'], out[f'L{l}.b1'] = g(a + 'mlp/linear_0:w'), g(a + 'mlp/linear_0:b')[None]
            out[f'L{l}.w2'], out[f'L{l}.b2'] = g(a + 'mlp/linear_1:w'), g(a + 'mlp/linear_1:b')[None]
    elif spec.kind == 'paulinet':
        N, K, M, d, n_up = spec.n_elec, spec.n_determinants, spec.n_nuc, spec.embedding_dim, spec.n_up
        nl = spec.gnn_subnet_layers
        if spec.gnn_embedding == 'embed':
            out['emb.table'] = g(PN.GNN + 'electron_embedding/ElectronicEmbedding:embeddings')
        types = PN.EDGE_TYPES if spec.gnn_conv_ne else PN.EDGE_TYPES[:2]
        xn = g(PN.GNN + 'nuclei_embedding/~/embed:embeddings') if spec.gnn_conv_ne else None
        w_bias = spec.gnn_update == 'concatenate'
        for l in range(spec.n_layers):
            c, lp = PN.conv_prefix(l), PN.layer_prefix(l)
            for t in types:
                for i in range(nl):
                    out[f'G{l}.w_{t}.{i}.w'] = g(c + f'w_{t}/linear_{i}:w')
                    if w_bias:
                        out[f'G{l}.w_{t}.{i}.b'] = g(c + f'w_{t}/linear_{i}:b')[None]
                if t != 'ne':
                    for i in range(nl):
                        out[f'G{l}.h_{t}.{i}.w'] = g(c + f'h_{t}/linear_{i}:w')
                        out[f'G{l}.h_{t}.{i}.b'] = g(c + f'h_{t}/linear_{i}:b')[None]
                else:
                    # nuclear embeddings are an hk.Embed lookup (gnn/electron_gnn.py:514): h_ne of them is walker-independent
                    hn = xn
                    for i in range(nl):
                        hn = np.tanh(hn @ g(c + f'h_ne/linear_{i}:w') + g(c + f'h_ne/linear_{i}:b'))
                    out[f'G{l}.hne'] = hn
                if spec.gnn_update == 'featurewise':
                    out[f'G{l}.g_{t}.w'] = g(lp + f'g_conv_{t}/linear_0:w')
                    out[f'G{l}.g_{t}.b'] = g(lp + f'g_conv_{t}/linear_0:b')[None]
            if spec.gnn_update == 'concatenate':
                out[f'G{l}.g.w'] = g(lp + 'g/linear_0:w')
                if spec.gnn_g_bias:
                    out[f'G{l}.g.b'] = g(lp + 'g/linear_0:b')[None]
            if spec.gnn_deep_edges and l < spec.n_layers - 1:
                for i in range(nl):
                    out[f'G{l}.u.{i}.w'] = g(lp + f'u/linear_{i}:w')
                    out[f'G{l}.u.{i}.b'] = g(lp + f'u/linear_{i}:b')[None]
        for i in range(spec.jastrow_layers):
            out[f'J{i}.w'] = g(PN.JASTROW + f'linear_{i}:w')
            if i < spec.jastrow_layers - 1:
                out[f'J{i}.b'] = g(PN.JASTROW + f'linear_{i}:b')[None]
        hid = paulinet_backflow_hidden(spec)
        dims_pad = [d] + hid
        for tag, pre, n_spin, off in (('up', PN.BF_UP, n_up, 0), ('dn', PN.BF_DN, spec.n_down, 0 if spec.full_determinant else n_up)):
            base = pre.rsplit('linear_0', 1)[0]
            nl = spec.backflow_layers
            zb = lambda k, n: g(k) if spec.backflow_bias else np.zeros(n)
            for i in range(nl - 1):  # hidden layers, zero-padded to the wider spin (ssp(0) = 0 keeps the padding inert)
                w = g(base + f'linear_{i}:w')
                b = zb(base + f'linear_{i}:b', w.shape[1])
                wp, bp = np.zeros((dims_pad[i], dims_pad[i + 1])), np.zeros((1, dims_pad[i + 1]))
                wp[:w.shape[0], :w.shape[1]], bp[0, :b.shape[0]] = w, b
                out[f'bfh{i}.{tag}'], out[f'bfb{i}.{tag}'] = wp, bp
            w = g(base + f'linear_{nl - 1}:w')
            b = zb(base + f'linear_{nl - 1}:b', w.shape[1])
            n_orb = N if spec.full_determinant else n_spin
            cols = (np.arange(K)[:, None] * N + off + np.arange(n_orb)[None, :]).ravel()  # (k, mu') -> k N + mu
            wp, bp = np.zeros((dims_pad[-1], K * N)), np.zeros((1, K * N))
            wp[:w.shape[0], cols], bp[0, cols] = w, b
            out[f'bf.{tag}'], out[f'bfb.{tag}'] = wp, bp
        if spec.env_per_shell:
            # per-shell spin-restricted envelopes (wf/env.py:26-75) -> engine layout [K N][M rep], unused terms pi = 0
            rep = paulinet_env_rep(spec)
            pi, zeta = g(f'{PN.ENV}:pi'), g(f'{PN.ENV}:zetas')
            pe, ze = np.zeros((K * N, M * rep)), np.ones((K * N, M * rep))
            seen = {}
            for j, c in enumerate(spec.env_centers):
                sh = seen.get(c, 0)
                seen[c] = sh + 1
                pe[:, c * rep + sh], ze[:, c * rep + sh] = pi[:, j], zeta[j]
            for t in ('up', 'dn'):
                out[f'env.pi_{t}'], out[f'env.zeta_{t}'] = pe, ze
        else:
            for s_, t in (('up', 'up'), ('down', 'dn')):
                out[f'env.pi_{t}'] = g(f'{PN.ENV}:pi_{s_}')
                out[f'env.zeta_{t}'] = g(f'{PN.ENV}:zetas_{s_}')
        out['cusp.alpha'] = np.array([[spec.cusp_alpha, spec.cusp_alpha]], dtype=np.float64)
        if spec.conf_coeff == 'linear':
            out['conf.w'] = g(PN.CONF + ':w').reshape(1, K)
        return out
    elif spec.kind == 'ferminet':
        for l in range(spec.n_layers):
            lp = PN.layer_prefix(l)
            out[f'F{l}.wg'], out[f'F{l}.bg'] = g(lp + 'g/linear_0:w'), g(lp + 'g/linear_0:b')[None]
            if l < spec.n_layers - 1:
                out[f'F{l}.wu'], out[f'F{l}.bu'] = g(lp + 'u/linear_0:w'), g(lp + 'u/linear_0:b')[None]
    else:
        raise NotImplementedError(spec.kind)
    out['bf.up'], out['bf.dn'] = g(PN.BF_UP + ':w'), g(PN.BF_DN + ':w')
    if spec.backflow_transform == 'both':  # [multiplicative head | additive head]
        out['bf.up'] = np.concatenate([out['bf.up'], g(PN.BF_UP_ADD + ':w')], axis=1)
        out['bf.dn'] = np.concatenate([out['bf.dn'], g(PN.BF_DN_ADD + ':w')], axis=1)
    if spec.kind == 'transpsiformer':
        # walker-independent nuclear stream (deepqmc_b200/nuclear.py): per-layer key/value rows of the
        # nuclear tokens and the envelope exponents zetas[M, K, E] -> engine layout [K N][M E], pi = 1
        from .nuclear import nuclear_stream

        ns = nuclear_stream(spec, params, R)
        N, K, M, E = spec.n_elec, spec.n_determinants, spec.n_nuc, spec.n_env_per_nuc
        for l in range(spec.n_layers):
            out[f'L{l}.kn'], out[f'L{l}.vn'] = ns['kn'][l], ns['vn'][l]
        for s, t in (('up', 'up'), ('down', 'dn')):
            z = np.transpose(ns[f'zetas_{s}'], (1, 0, 2)).reshape(K, 1, M * E)  # [K][m E + e]
            out[f'env.zeta_{t}'] = np.broadcast_to(z, (K, N, M * E)).reshape(K * N, M * E).copy()
            out[f'env.pi_{t}'] = np.ones((K * N, M * E))
    else:
        for s, t in (('up', 'up'), ('down', 'dn')):
            out[f'env.pi_{t}'] = g(f'{PN.ENV}:pi_{s}')
            out[f'env.zeta_{t}'] = g(f'{PN.ENV}:zetas_{s}')
    if spec.cusp == 'psiformer':
        out['cusp.alpha'] = np.array([[g(f'{PN.CUSP}:same_alpha'), g(f'{PN.CUSP}:anti_alpha')]], dtype=np.float64)
    elif spec.cusp == 'deepqmc':
        out['cusp.alpha'] = np.array([[spec.cusp_alpha, spec.cusp_alpha]], dtype=np.float64)
    else:
        out['cusp.alpha'] = np.ones((1, 2))
    return out


def _unpack_psiformer_grads(spec: AnsatzSpec, entries: dict, flat) -> dict:
    """Engine-layout gradient vector -> Haiku-named tree (inverse of _pack_haiku_params for the Psiformer)."""
    def e(name):
        off, rows, cols = entries[name]
        return flat[off:off + rows * cols].reshape(rows, cols)

    d = spec.embedding_dim
    out = {PN.GNN + 'electron_embedding/linear:w': e('emb.w')}
    for l in range(spec.n_layers):
        a = PN.attn_prefix(l) if spec.kind == 'psiformer' else PN.comb_prefix(l)
        qkv = e(f'L{l}.wqkv')
        for j, n in enumerate(('query', 'key', 'value')):
            out[a + f'multi_head_attention/{n}:w'] = qkv[:, j * d:(j + 1) * d]
        out[a + 'multi_head_attention/linear:w'] = e(f'L{l}.wo')
        out[a + 'mlp/linear_0:w'], out[a + 'mlp/linear_0:b'] = e(f'L{l}.w1'), e(f'L{l}.b1')[0]
        out[a + 'mlp/linear_1:w'], out[a + 'mlp/linear_1:b'] = e(f'L{l}.w2'), e(f'L{l}.b2')[0]
    out[PN.BF_UP + ':w'], out[PN.BF_DN + ':w'] = e('bf.up'), e('bf.dn')
    if spec.kind == 'psiformer':
        for s_, t in (('up', 'up'), ('down', 'dn')):
            out[f'{PN.ENV}:pi_{s_}'] = e(f'env.pi_{t}')
            out[f'{PN.ENV}:zetas_{s_}'] = e(f'env.zeta_{t}')
    if spec.cusp == 'psiformer':
        ca = e('cusp.alpha')
        out[f'{PN.CUSP}:same_alpha'], out[f'{PN.CUSP}:anti_alpha'] = ca[0, 0], ca[0, 1]
    return out


def _unpack_paulinet_grads(spec: AnsatzSpec, entries: dict, flat, params: dict) -> dict:
    """Engine-layout gradient vector -> Haiku-named tree for the conv-GNN test ansatz (inverse of _pack_haiku_params for
    spec.kind == 'paulinet', featurewise / hk.Embed variant).  The walker-independent h_ne(nuclear embedding) rows were
    evaluated on the host at upload; their cotangents G<l>.hne are pulled back through that small tanh MLP here."""
    def e(name):
        off, rows, cols = entries[name]
        return flat[off:off + rows * cols].reshape(rows, cols)

    N, K, M, d, n_up = spec.n_elec, spec.n_determinants, spec.n_nuc, spec.embedding_dim, spec.n_up
    nl = spec.gnn_subnet_layers
    out = {}
    if spec.gnn_embedding == 'embed':
        out[PN.GNN + 'electron_embedding/ElectronicEmbedding:embeddings'] = e('emb.table')
    types = PN.EDGE_TYPES if spec.gnn_conv_ne else PN.EDGE_TYPES[:2]
    dxn = None
    for l in range(spec.n_layers):
        c, lp = PN.conv_prefix(l), PN.layer_prefix(l)
        for t in types:
            for i in range(nl):
                out[c + f'w_{t}/linear_{i}:w'] = e(f'G{l}.w_{t}.{i}.w')
                if spec.gnn_update == 'concatenate':
                    out[c + f'w_{t}/linear_{i}:b'] = e(f'G{l}.w_{t}.{i}.b')[0]
                if t != 'ne':
                    out[c + f'h_{t}/linear_{i}:w'] = e(f'G{l}.h_{t}.{i}.w')
                    out[c + f'h_{t}/linear_{i}:b'] = e(f'G{l}.h_{t}.{i}.b')[0]
            if spec.gnn_update == 'featurewise':
                out[lp + f'g_conv_{t}/linear_0:w'] = e(f'G{l}.g_{t}.w')
                out[lp + f'g_conv_{t}/linear_0:b'] = e(f'G{l}.g_{t}.b')[0]
        if spec.gnn_update == 'concatenate':
            out[lp + 'g/linear_0:w'] = e(f'G{l}.g.w')
            if spec.gnn_g_bias:
                out[lp + 'g/linear_0:b'] = e(f'G{l}.g.b')[0]
        if spec.gnn_deep_edges and l < spec.n_layers - 1:
            for i in range(nl):
                out[lp + f'u/linear_{i}:w'] = e(f'G{l}.u.{i}.w')
                out[lp + f'u/linear_{i}:b'] = e(f'G{l}.u.{i}.b')[0]
        if spec.gnn_conv_ne:  # host: hne = tanh MLP(xn) -> gradients of the nuclear table and of h_ne
            leaves = {k: torch.as_tensor(np.asarray(params[k], dtype=np.float64)).requires_grad_(True)
                      for k in [PN.GNN + 'nuclei_embedding/~/embed:embeddings']
                      + [c + f'h_ne/linear_{i}:{wb}' for i in range(nl) for wb in 'wb']}
            hn = leaves[PN.GNN + 'nuclei_embedding/~/embed:embeddings']
            for i in range(nl):
                hn = torch.tanh(hn @ leaves[c + f'h_ne/linear_{i}:w'] + leaves[c + f'h_ne/linear_{i}:b'])
            (hn * e(f'G{l}.hne').detach().cpu().double()).sum().backward()
            for k, v in leaves.items():
                g = v.grad.to(device=flat.device, dtype=flat.dtype)
                if k.endswith('embed:embeddings'):
                    dxn = g if dxn is None else dxn + g
                else:
                    out[k] = g
    if dxn is not None:
        out[PN.GNN + 'nuclei_embedding/~/embed:embeddings'] = dxn
    for i in range(spec.jastrow_layers):
        out[PN.JASTROW + f'linear_{i}:w'] = e(f'J{i}.w')
        if i < spec.jastrow_layers - 1:
            out[PN.JASTROW + f'linear_{i}:b'] = e(f'J{i}.b')[0]
    for tag, pre, n_spin, off in (('up', PN.BF_UP, n_up, 0), ('dn', PN.BF_DN, spec.n_down, 0 if spec.full_determinant else n_up)):
        base = pre.rsplit('linear_0', 1)[0]
        nb = spec.backflow_layers
        for i in range(nb - 1):
            shp = np.asarray(params[base + f'linear_{i}:w']).shape
            out[base + f'linear_{i}:w'] = e(f'bfh{i}.{tag}')[:shp[0], :shp[1]]
            if spec.backflow_bias:
                out[base + f'linear_{i}:b'] = e(f'bfb{i}.{tag}')[0, :shp[1]]
        shp = np.asarray(params[base + f'linear_{nb - 1}:w']).shape
        n_orb = N if spec.full_determinant else n_spin
        cols = torch.as_tensor((np.arange(K)[:, None] * N + off + np.arange(n_orb)[None, :]).ravel(), device=flat.device)
        out[base + f'linear_{nb - 1}:w'] = e(f'bf.{tag}')[:shp[0]][:, cols]
        if spec.backflow_bias:
            out[base + f'linear_{nb - 1}:b'] = e(f'bfb.{tag}')[0, cols]
    if spec.env_per_shell:  # packed [K N][M rep] (both spins share the parameters) -> pi[K N, n_env], zetas[n_env]
        rep = paulinet_env_rep(spec)
        idx, seen = [], {}
        for c_ in spec.env_centers:
            sh = seen.get(c_, 0)
            seen[c_] = sh + 1
            idx.append(c_ * rep + sh)
        idx = torch.as_tensor(idx, device=flat.device)
        dpi = e('env.pi_up') + e('env.pi_dn')
        dze = e('env.zeta_up') + e('env.zeta_dn')
        out[f'{PN.ENV}:pi'] = dpi[:, idx]
        out[f'{PN.ENV}:zetas'] = dze[:, idx].sum(0)
    else:
        for s_, t in (('up', 'up'), ('down', 'dn')):
            out[f'{PN.ENV}:pi_{s_}'] = e(f'env.pi_{t}')
            out[f'{PN.ENV}:zetas_{s_}'] = e(f'env.zeta_{t}')
    if spec.conf_coeff == 'linear':
        out[PN.CONF + ':w'] = e('conf.w').reshape(spec.n_determinants, 1)
    return out


def _unpack_ferminet_grads(spec: AnsatzSpec, entries: dict, flat) -> dict:
    """Engine-layout gradient vector -> Haiku-named tree for the FermiNet (inverse of _pack_haiku_params)."""
    def e(name):
        off, rows, cols = entries[name]
        return flat[off:off + rows * cols].reshape(rows, cols)

    out = {}
    for l in range(spec.n_layers):
        lp = PN.layer_prefix(l)
        out[lp + 'g/linear_0:w'], out[lp + 'g/linear_0:b'] = e(f'F{l}.wg'), e(f'F{l}.bg')[0]
        if l < spec.n_layers - 1:
            out[lp + 'u/linear_0:w'], out[lp + 'u/linear_0:b'] = e(f'F{l}.wu'), e(f'F{l}.bu')[0]
    out[PN.BF_UP + ':w'], out[PN.BF_DN + ':w'] = e('bf.up'), e('bf.dn')
    for s_, t in (('up', 'up'), ('down', 'dn')):
        out[f'{PN.ENV}:pi_{s_}'] = e(f'env.pi_{t}')
        out[f'{PN.ENV}:zetas_{s_}'] = e(f'env.zeta_{t}')
    if spec.cusp == 'psiformer':
        ca = e('cusp.alpha')
        out[f'{PN.CUSP}:same_alpha'], out[f'{PN.CUSP}:anti_alpha'] = ca[0, 0], ca[0, 1]
    return out


class Engine:
    """One engine per (ansatz spec, Hamiltonian constants, dtype, device)."""

    def __init__(self, spec: AnsatzSpec, hamil, dtype: str = 'float64', device: int | None = None,
                 gemm_backend: int = 0, _lib_path: str | None = None, plan_only: bool = False):
        """``plan_only``: a handle created with device = -1 -- no CUDA context; only the parameter table, the workspace
        sizes and ``debug_plan`` work (tests/test_plan.py checks the workspace planner on the CPU with it)."""
        self._host = _lib_path is not None  # emulator build: "device" pointers are host pointers
        self.lib = _lib.load(_lib_path)
        self.plan_only = plan_only
        if not self._host and not plan_only and not torch.cuda.is_available():
            raise RuntimeError('deepqmc_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        self.spec, self.hamil = spec, hamil
        self.dtype_code = {'float64': 0, 'float32': 1}[dtype]
        self.dtype = _TORCH_DTYPE[self.dtype_code]
        if plan_only:
            self.device_index, self.device = -1, torch.device('cpu')
        else:
            self.device_index = 0 if self._host else (torch.cuda.current_device() if device is None else device)
            self.device = torch.device('cpu') if self._host else torch.device('cuda', self.device_index)
        cfg = _lib.DqmcConfig()
        cfg.kind = {'psiformer': 0, 'ferminet': 1, 'transpsiformer': 2, 'paulinet': 3}[spec.kind]
        cfg.dtype, cfg.gemm_backend = self.dtype_code, gemm_backend
        cfg.n_up, cfg.n_down, cfg.n_nuc = spec.n_up, spec.n_down, spec.n_nuc
        cfg.embedding_dim, cfg.n_layers, cfg.n_heads = spec.embedding_dim, spec.n_layers, spec.n_heads
        cfg.n_determinants, cfg.edge_dim = spec.n_determinants, spec.edge_dim
        cfg.n_env_per_nuc = spec.n_env_per_nuc
        cfg.n_nuc_tokens = spec.n_nuc if spec.kind == 'transpsiformer' else 0
        if spec.kind == 'paulinet':
            cfg.n_env_per_nuc = paulinet_env_rep(spec) if spec.env_per_shell else 1
            cfg.gnn_features = 1 if spec.gnn_embedding == 'features' else 0
            cfg.gnn_concat = 1 if spec.gnn_update == 'concatenate' else 0
            cfg.gnn_conv_ne = 1 if spec.gnn_conv_ne else 0
            cfg.gnn_sub_n = spec.gnn_subnet_layers
            cfg.gnn_deep_edges = 1 if spec.gnn_deep_edges else 0
            cfg.gnn_res_norm = 1 if spec.gnn_residual_normalize else 0
            cfg.gnn_g_bias = 1 if spec.gnn_g_bias else 0
            cfg.gnn_w_bias = 1 if spec.gnn_update == 'concatenate' else 0
            assert spec.n_layers <= 8 and spec.gnn_subnet_layers <= 4
            d_in, e_in = (spec.embedding_dim if spec.gnn_embedding == 'embed' else 4 * spec.n_nuc), 4
            for l in range(spec.n_layers):
                for i, v in enumerate(PN.log_dims(e_in, spec.edge_dim, spec.gnn_subnet_layers)):
                    cfg.gnn_w_dims[l * 4 + i] = v
                    cfg.gnn_u_dims[l * 4 + i] = v
                for i, v in enumerate(PN.log_dims(d_in, spec.edge_dim, spec.gnn_subnet_layers)):
                    cfg.gnn_h_dims[l * 4 + i] = v
                if spec.gnn_deep_edges and l < spec.n_layers - 1:
                    e_in = spec.edge_dim
                d_in = spec.embedding_dim
            cfg.factorized_det = 0 if spec.full_determinant else 1
            cfg.conf_linear = 1 if spec.conf_coeff == 'linear' else 0
            cfg.mult_act = 1 if spec.mult_act == 'default' else 0
            cfg.n_elec_types = 1 if spec.n_up == spec.n_down else 2
            jd = PN.log_dims(spec.embedding_dim, 1, spec.jastrow_layers) if spec.jastrow_layers else []
            cfg.jastrow_n = len(jd)
            for i, v in enumerate(jd):
                cfg.jastrow_dims[i] = v
            hid = paulinet_backflow_hidden(spec)
            cfg.backflow_n = len(hid)
            for i, v in enumerate(hid):
                cfg.backflow_dims[i] = v
        cfg.cusp_kind = {'psiformer': 1, 'deepqmc': 2}.get(spec.cusp, 0)
        cfg.backflow_add = {'mult': 0, 'add': 1, 'both': 2}[spec.backflow_transform]
        if cfg.backflow_add and spec.kind not in ('psiformer', 'ferminet'):
            raise NotImplementedError('additive backflow branch: Psiformer / FermiNet kinds only')
        cfg.nuc_cusp_kind = {'psiformer': 1, 'deepqmc': 2}.get(spec.cusp_nuclei, 0)
        for m in range(spec.n_nuc):
            cfg.z_nuclear[m] = float(hamil.mol.charges[m])
        cfg.cusp_same_scale, cfg.cusp_anti_scale = spec.cusp_same_scale, spec.cusp_anti_scale
        M = spec.n_nuc
        assert M <= _lib.MAX_NUC
        for m in range(M):
            cfg.z_valence[m] = float(hamil.ns_valence[m])
            cfg.ecp_mask[m] = int(hamil.ecp_mask[m])
        lp, nl = getattr(hamil, 'loc_params', None), getattr(hamil, 'nl_params', None)
        if lp is not None and lp.shape[-1] > 0:
            Tm = lp.shape[-1]
            assert Tm <= _lib.MAX_T
            cfg.ecp_loc_terms = Tm
            arr = np.zeros((_lib.MAX_NUC, 3, 2, _lib.MAX_T))
            arr[:M, :, :, :Tm] = lp
            cfg.ecp_loc[:] = arr.ravel().tolist()
        if nl is not None and nl.size > 0:
            L, Tn = nl.shape[1], nl.shape[3]
            assert L <= _lib.MAX_L and Tn <= _lib.MAX_T
            cfg.ecp_nl_lmax_p1, cfg.ecp_nl_terms = L, Tn
            arr = np.zeros((_lib.MAX_NUC, _lib.MAX_L, 2, _lib.MAX_T))
            arr[:M, :L, :, :Tn] = nl
            cfg.ecp_nl[:] = arr.ravel().tolist()
        self._cfg = cfg
        h = C.c_void_p()
        rc = self.lib.dqmc_create(C.byref(cfg), self.device_index, C.byref(h))
        if rc != 0:
            raise RuntimeError(f'dqmc_create failed with status {rc}')
        self.h = h
        ph = getattr(hamil, 'ph', None)
        if ph is not None:  # pseudo-Hamiltonian tables (deepqmc_b200/ph.py) -> device
            n_tab, _, G = ph.tables.shape
            rc = self.lib.dqmc_set_pseudo_hamiltonian(h, n_tab, G, float(ph.r_max),
                                                      ph.tables.ctypes.data_as(C.POINTER(C.c_double)),
                                                      ph.tab_of_nuc.ctypes.data_as(C.POINTER(C.c_int32)))
            self._check(rc, 'dqmc_set_pseudo_hamiltonian')
        self.entries = {}
        name = C.create_string_buffer(64)
        off, rows, cols = C.c_int64(), C.c_int32(), C.c_int32()
        for i in range(self.lib.dqmc_param_count(h)):
            self.lib.dqmc_param_entry(h, i, name, 64, C.byref(off), C.byref(rows), C.byref(cols))
            self.entries[name.value.decode()] = (off.value, rows.value, cols.value)
        self.n_packed = self.lib.dqmc_param_total(h)
        self._ws = None
        self._ws_ok = set()  # (n_walkers, mode, cap) requests the current workspace is known to satisfy
        self._params_version = None
        self._nuc_R_dev = None

    # ------------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f'{what} failed ({rc}): {self.lib.dqmc_last_error(self.h).decode()}')

    def _stream(self):
        return C.c_void_p(0 if self._host else torch.cuda.current_stream(self.device).cuda_stream)

    def set_params(self, params: dict, R=None):
        """Upload a parameter tree.  TransPsiformer: ``R`` (default: the Hamiltonian's geometry) fixes
        the walker-independent nuclear stream that is uploaded with the parameters."""
        if self.spec.kind == 'transpsiformer':
            R = self.hamil.mol.coords if R is None else R
            R = np.asarray(R.detach().cpu() if torch.is_tensor(R) else R, dtype=np.float64)
            self._nuc_R, self._nuc_R_dev = R, None
        self._params = params
        packed = _pack_haiku_params(self.spec, params, R)
        if self.spec.cusp_nuclei != 'none':  # NuclearCuspAsymptotic: alpha (trainable or fixed) + the nuclear charges
            al = (float(np.asarray(params[f'{PN.NUC_CUSP}:nuc_alpha'])) if self.spec.cusp_nuclei_trainable
                  else self.spec.cusp_nuclei_alpha)
            packed['cusp.nuc'] = np.array([[al, *[float(z) for z in self.hamil.mol.charges]]], dtype=np.float64)
        flat = np.zeros(self.n_packed, dtype=np.float64)
        for k, (off, rows, cols) in self.entries.items():
            v = packed[k]
            assert v.shape == (rows, cols), (k, v.shape, rows, cols)
            flat[off:off + rows * cols] = v.ravel()
        self._flat = flat  # keep alive until the async copy is done
        rc = self.lib.dqmc_set_params(self.h, flat.ctypes.data_as(C.POINTER(C.c_double)), self.n_packed, self._stream())
        self._check(rc, 'dqmc_set_params')
        if not self._host:
            torch.cuda.current_stream(self.device).synchronize()

    def workspace(self, n_walkers: int, mode: int, max_bytes: int | None = None):
        """The engine's scratch buffer for one call of the entry point ``mode`` names: dqmc_workspace_bytes (a dry pass of the
        code that carves it), capped at ``max_bytes`` / 60 % of the free HBM -- the engine then chunks the walkers -- but
        never below dqmc_workspace_bytes_min."""
        key = (n_walkers, mode, max_bytes)
        if self._ws is not None and key in self._ws_ok:  # hot path: no driver queries per call
            return self._ws
        need = self.lib.dqmc_workspace_bytes(self.h, n_walkers, mode)
        if max_bytes is None and not self._host:
            free = torch.cuda.mem_get_info(self.device)[0] + (self._ws.numel() if self._ws is not None else 0)
            max_bytes = int(0.6 * free)
        if max_bytes is not None:
            need = min(need, max(max_bytes, self.lib.dqmc_workspace_bytes_min(self.h, n_walkers, mode)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_ok = set()
        self._ws_ok.add(key)
        return self._ws

    def _prep(self, x):
        x = torch.as_tensor(x, dtype=self.dtype, device=self.device)
        return x.contiguous()

    def _R(self, R, B):
        R = self._prep(R)
        assert R.dim() in (2, 3) and tuple(R.shape[-2:]) == (self.spec.n_nuc, 3), f'R must be [M, 3] or [B, M, 3], got {tuple(R.shape)}'
        batched = 1 if R.dim() == 3 else 0
        if batched:
            assert R.shape[0] == B
        if self.spec.kind == 'transpsiformer':
            if batched:
                raise NotImplementedError('TransPsiformer engine: one geometry per call (unbatched R)')
            # the nuclear stream depends on the geometry: compare the CONTENTS with the geometry it was evaluated for (never a
            # pointer / version key: temporaries of different geometries reuse addresses in the caching allocator)
            if self._nuc_R_dev is None or self._nuc_R_dev.device != R.device or not torch.equal(R.double(), self._nuc_R_dev):
                Rh = R.detach().cpu().double().numpy()
                if not np.allclose(Rh, self._nuc_R, rtol=0, atol=1e-12):
                    self.set_params(self._params, Rh)
                self._nuc_R_dev = R.detach().double().clone()
        return R, batched

    # ------------------------------------------------------------------------------------
    def wf_forward(self, r, R, max_ws_bytes=None):
        r = self._prep(r)
        B = r.shape[0]
        R, Rb = self._R(R, B)
        sign = torch.empty(B, dtype=self.dtype, device=self.device)
        log = torch.empty(B, dtype=self.dtype, device=self.device)
        ws = self.workspace(B, MODE_FORWARD, max_ws_bytes)
        rc = self.lib.dqmc_wf_forward(self.h, r.data_ptr(), R.data_ptr(), Rb, B, sign.data_ptr(), log.data_ptr(),
                                      ws.data_ptr(), ws.numel(), self._stream())
        self._check(rc, 'dqmc_wf_forward')
        return sign, log

    def wf_orbitals(self, r, R, max_ws_bytes=None):
        """-> (orb_up[B, K, n_up, n_orb], orb_down[B, K, n_down, n_orb]): envelope * mult_act(backflow), the matrices whose
        determinants make up psi; n_orb = N for full determinants, the spin's electron count otherwise
        (reference: Ansatz.apply(..., return_mos=True), wf/nn_wave_function.py:131-142)."""
        r = self._prep(r)
        B, N = r.shape[0], r.shape[1]
        R, Rb = self._R(R, B)
        K, n_up = self.spec.n_determinants, self.spec.n_up
        out = torch.empty(B, K, N, N, dtype=self.dtype, device=self.device)
        ws = self.workspace(B, MODE_FORWARD, max_ws_bytes)
        rc = self.lib.dqmc_wf_orbitals(self.h, r.data_ptr(), R.data_ptr(), Rb, B, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       self._stream())
        self._check(rc, 'dqmc_wf_orbitals')
        if self.spec.full_determinant:
            return out[:, :, :n_up, :], out[:, :, n_up:, :]
        return out[:, :, :n_up, :n_up], out[:, :, n_up:, n_up:]

    def local_energy(self, r, R, seed=0, ecp_twist=None, want_grad=False, max_ws_bytes=None):
        r = self._prep(r)
        B, N = r.shape[0], r.shape[1]
        R, Rb = self._R(R, B)
        mk = lambda *s: torch.empty(*s, dtype=self.dtype, device=self.device)
        E, stats, sign, log = mk(B), mk(6, B), mk(B), mk(B)
        grad = mk(B, 3 * N) if want_grad else None
        tw = self._prep(ecp_twist) if ecp_twist is not None else None
        ws = self.workspace(B, MODE_LOCAL_ENERGY, max_ws_bytes)
        rc = self.lib.dqmc_local_energy(
            self.h, r.data_ptr(), R.data_ptr(), Rb, B, seed, tw.data_ptr() if tw is not None else None,
            E.data_ptr(), stats.data_ptr(), sign.data_ptr(), log.data_ptr(),
            grad.data_ptr() if grad is not None else None, ws.data_ptr(), ws.numel(), self._stream())
        self._check(rc, 'dqmc_local_energy')
        return E, stats, sign, log, grad

    def vjp_params(self, r, R, weights, max_ws_bytes=None):
        """-> (sign[B], log[B], grads): grads = d/dparams sum_b weights[b] log|psi(r_b)| as a Haiku-named dict
        (reference: loss/loss_function.py:53-82).  Every ansatz kind (the additive backflow branch excepted)."""
        r = self._prep(r)
        B = r.shape[0]
        R, Rb = self._R(R, B)
        w = self._prep(weights)
        assert w.shape == (B,)
        sign = torch.empty(B, dtype=self.dtype, device=self.device)
        log = torch.empty(B, dtype=self.dtype, device=self.device)
        flat = torch.empty(self.n_packed, dtype=self.dtype, device=self.device)
        ws = self.workspace(B, MODE_VJP, max_ws_bytes)
        rc = self.lib.dqmc_wf_vjp_params(self.h, r.data_ptr(), R.data_ptr(), Rb, B, w.data_ptr(), sign.data_ptr(), log.data_ptr(),
                                         flat.data_ptr(), ws.data_ptr(), ws.numel(), self._stream())
        self._check(rc, 'dqmc_wf_vjp_params')
        if self.spec.kind == 'paulinet':
            grads = _unpack_paulinet_grads(self.spec, self.entries, flat, self._params)
        else:
            unpack = _unpack_ferminet_grads if self.spec.kind == 'ferminet' else _unpack_psiformer_grads
            grads = unpack(self.spec, self.entries, flat)
        if self.spec.cusp_nuclei != 'none' and self.spec.cusp_nuclei_trainable:
            grads[f'{PN.NUC_CUSP}:nuc_alpha'] = flat[self.entries['cusp.nuc'][0]]
        if self.spec.kind == 'transpsiformer':
            # the walker-independent nuclear stream is differentiated on the host: the engine accumulated the
            # cotangents of its outputs (keys / values of the nuclear tokens, envelope exponents)
            from .nuclear import nuclear_stream_vjp

            def e(name):
                off, rows, cols = self.entries[name]
                return flat[off:off + rows * cols].reshape(rows, cols).detach().cpu().double().numpy()

            N, K, M, E = self.spec.n_elec, self.spec.n_determinants, self.spec.n_nuc, self.spec.n_env_per_nuc
            cot = {'kn': [e(f'L{l}.kn') for l in range(self.spec.n_layers)],
                   'vn': [e(f'L{l}.vn') for l in range(self.spec.n_layers)]}
            for s_, t in (('up', 'up'), ('down', 'dn')):  # engine [K N][M E] -> zetas[M, K, E] (shared by the orbitals)
                cot[f'zetas_{s_}'] = e(f'env.zeta_{t}').reshape(K, N, M, E).sum(1).transpose(1, 0, 2)
            host = nuclear_stream_vjp(self.spec, self._params, self._nuc_R, cot)
            for k, v in host.items():
                v = v.to(device=self.device, dtype=self.dtype)
                grads[k] = grads[k] + v.reshape(grads[k].shape) if k in grads else v
        return sign, log, grads

    def mcmc_sweep(self, state, R, n_sub, target_acceptance=0.57, max_age=None, seed=0, step0=0, walker_offset=0,
                   noise_normal=None, noise_uniform=None, max_ws_bytes=None, exchange_probability=0.0, exchange_flags=None,
                   exchange_idx=None):
        """state: dict(r[B,N,3], sign[B], log[B], age[B] int32, tau[1]) -- updated IN PLACE.
        exchange_probability > 0 (or injected ``exchange_flags[n_sub]`` / ``exchange_idx[n_sub, B, 2]``): spin-exchange
        sub-steps mixed in (dqmc_mcmc_sweep_exchange; reference OppositeSpinExchangeSampler)."""
        r = state['r']
        B = r.shape[0]
        for k in ('r', 'sign', 'log', 'tau'):
            assert state[k].dtype == self.dtype and state[k].is_contiguous() and state[k].device == self.device, k
        assert state['age'].dtype == torch.int32
        R, Rb = self._R(R, B)
        nn = self._prep(noise_normal) if noise_normal is not None else None
        nu = self._prep(noise_uniform) if noise_uniform is not None else None
        stats = torch.zeros(7, dtype=self.dtype, device=self.device)
        ws = self.workspace(B, MODE_MCMC, max_ws_bytes)  # proposal buffers + the plain-forward chunk
        if exchange_probability > 0.0 or exchange_flags is not None:
            flags = None
            if exchange_flags is not None:
                flags = (C.c_int32 * n_sub)(*[int(f) for f in exchange_flags])
            xi = None
            if exchange_idx is not None:
                xi = exchange_idx.to(device=self.device, dtype=torch.int32).contiguous()
                assert xi.shape == (n_sub, B, 2)
            rc = self.lib.dqmc_mcmc_sweep_exchange(
                self.h, r.data_ptr(), state['sign'].data_ptr(), state['log'].data_ptr(), state['age'].data_ptr(),
                state['tau'].data_ptr(), R.data_ptr(), Rb, B, n_sub, float(target_acceptance if target_acceptance else 0.0),
                -1 if max_age is None else int(max_age), seed, step0, walker_offset,
                nn.data_ptr() if nn is not None else None, nu.data_ptr() if nu is not None else None,
                float(exchange_probability), flags, xi.data_ptr() if xi is not None else None,
                stats.data_ptr(), ws.data_ptr(), ws.numel(), self._stream())
            self._check(rc, 'dqmc_mcmc_sweep_exchange')
            return stats
        rc = self.lib.dqmc_mcmc_sweep(
            self.h, r.data_ptr(), state['sign'].data_ptr(), state['log'].data_ptr(), state['age'].data_ptr(),
            state['tau'].data_ptr(), R.data_ptr(), Rb, B, n_sub, float(target_acceptance if target_acceptance else 0.0),
            -1 if max_age is None else int(max_age), seed, step0, walker_offset,
            nn.data_ptr() if nn is not None else None, nu.data_ptr() if nu is not None else None,
            stats.data_ptr(), ws.data_ptr(), ws.numel(), self._stream())
        self._check(rc, 'dqmc_mcmc_sweep')
        return stats

    def langevin_sweep(self, state, R, n_sub, target_acceptance=0.57, max_age=None, seed=0, step0=0, walker_offset=0,
                       noise_normal=None, noise_uniform=None):
        """state: dict(r[B,N,3], sign[B], log[B], force[B,N,3], age[B] int32, tau[1]) -- updated IN PLACE.
        n_sub = 0: recompute sign / log / force of the current walkers (sampler update)."""
        r = state['r']
        B, N = r.shape[0], r.shape[1]
        for k in ('r', 'sign', 'log', 'force', 'tau'):
            assert state[k].dtype == self.dtype and state[k].is_contiguous() and state[k].device == self.device, k
        assert state['age'].dtype == torch.int32
        R, Rb = self._R(R, B)
        nn = self._prep(noise_normal) if noise_normal is not None else None
        nu = self._prep(noise_uniform) if noise_uniform is not None else None
        stats = torch.zeros(7, dtype=self.dtype, device=self.device)
        ws = self.workspace(B, MODE_LANGEVIN)  # proposal / force buffers + the forward-Laplacian chunk
        rc = self.lib.dqmc_langevin_sweep(
            self.h, r.data_ptr(), state['sign'].data_ptr(), state['log'].data_ptr(), state['force'].data_ptr(),
            state['age'].data_ptr(), state['tau'].data_ptr(), R.data_ptr(), Rb, B, n_sub,
            float(target_acceptance if target_acceptance else 0.0), -1 if max_age is None else int(max_age), seed, step0,
            walker_offset, nn.data_ptr() if nn is not None else None, nu.data_ptr() if nu is not None else None,
            stats.data_ptr(), ws.data_ptr(), ws.numel(), self._stream())
        self._check(rc, 'dqmc_langevin_sweep')
        return stats

    def debug_gemm(self, weight, A, bias=None, Res=None, S=1, sliced=False, backend=0):
        off, K, Nc = self.entries[weight]
        A = self._prep(A)
        rows = A.shape[0]
        out_rows = rows
        Cout = torch.zeros(out_rows, Nc, dtype=self.dtype, device=self.device)
        Res = self._prep(Res) if Res is not None else None
        rc = self.lib.dqmc_debug_gemm(self.h, weight.encode(), bias.encode() if bias else None, A.data_ptr(),
                                      Res.data_ptr() if Res is not None else None, Cout.data_ptr(),
                                      rows // self.spec.n_elec if sliced else rows, S, int(sliced),
                                      backend, self._stream())
        self._check(rc, 'dqmc_debug_gemm')
        return Cout

    def stats_pack(self, E, stats=None):
        """-> float64[11] on the device: sum E, sum E^2, B, sums of the six stats rows, max E, -min E (one launch)."""
        out = torch.empty(11, dtype=torch.float64, device=self.device)
        assert E.dtype == self.dtype and E.is_contiguous() and (stats is None or (stats.is_contiguous() and stats.shape == (6, E.shape[0])))
        rc = self.lib.dqmc_stats_pack(self.h, E.data_ptr(), stats.data_ptr() if stats is not None else None, E.shape[0],
                                      out.data_ptr(), self._stream())
        self._check(rc, 'dqmc_stats_pack')
        return out

    def debug_plan(self, n_walkers: int, mode: int, workspace_bytes: int = 0):
        """-> (planned, carved): dqmc_workspace_bytes and the highest offset the entry point of ``mode`` carves when given
        ``workspace_bytes`` (0: the planned size); host-only."""
        pl, cv = C.c_int64(), C.c_int64()
        rc = self.lib.dqmc_debug_plan(self.h, n_walkers, mode, workspace_bytes, C.byref(pl), C.byref(cv))
        self._check(rc, 'dqmc_debug_plan')
        return pl.value, cv.value

    def workspace_bytes(self, n_walkers: int, mode: int) -> int:
        return self.lib.dqmc_workspace_bytes(self.h, n_walkers, mode)

    def workspace_bytes_min(self, n_walkers: int, mode: int) -> int:
        return self.lib.dqmc_workspace_bytes_min(self.h, n_walkers, mode)

    def debug_mlp_block(self, layer, O, X):
        """One launch of the fused plain-forward MLP block of `layer` -> X' [rows, d] (self-test hook)."""
        O, X = self._prep(O), self._prep(X)
        out = torch.empty_like(O)
        rc = self.lib.dqmc_debug_mlp_block(self.h, layer, O.data_ptr(), X.data_ptr(), out.data_ptr(), O.shape[0], self._stream())
        self._check(rc, 'dqmc_debug_mlp_block')
        return out

    def debug_trunk(self, X0):
        """One launch of the whole-trunk kernel: all attention layers applied to the embedding rows -> [rows, d] (self-test hook)."""
        X0 = self._prep(X0)
        out = torch.empty_like(X0)
        rc = self.lib.dqmc_debug_trunk(self.h, X0.data_ptr(), out.data_ptr(), X0.shape[0], self._stream())
        self._check(rc, 'dqmc_debug_trunk')
        return out

    def profile_begin(self):
        self.lib.dqmc_profile_begin(self.h)

    def profile_end(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self.lib.dqmc_profile_end(self.h, C.byref(ms), C.byref(fl), C.byref(n))
        return ms.value, fl.value, n.value

    PROFILE_CLASSES = ('row_gemm', 'mlp_block', 'trunk')

    def profile_end_classes(self):
        """{class: (ms, algorithmic flops, launches)} for the three tensor-core kernel classes."""
        ms, fl, n = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int64 * 3)()
        self.lib.dqmc_profile_end_classes(self.h, ms, fl, n)
        return {k: (ms[i], fl[i], n[i]) for i, k in enumerate(self.PROFILE_CLASSES)}

    @property
    def launch_count(self):
        return self.lib.dqmc_launch_count(self.h)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.dqmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
