"""Restatement of the reference's neural wave functions (TEST INFRASTRUCTURE).

Single-walker, functional, torch.float64; differentiable so the Laplacian can be taken by
autograd (oracle/laplacian.py).  Follows
  src/deepqmc/wf/nn_wave_function.py:127-173  (assembly, slogdet, exp-normalised det sum, cusp)
  src/deepqmc/wf/env.py:57-108                (ExponentialEnvelopes, isotropic per-orbital)
  src/deepqmc/wf/omni.py:43-88                (Backflow reshape/swap order)
  src/deepqmc/wf/cusp.py:17-26,49-78          (PsiformerCusp)
  src/deepqmc/gnn/electron_gnn.py:596-619     (positional embedding + spin + projection)
  src/deepqmc/gnn/update_features.py:241-286  (attention + MLP, residuals) with
  hk.MultiHeadAttention's algebra (restated in src/deepqmc/hkext.py:215-253)
  src/deepqmc/gnn/update_features.py:47-159, electron_gnn.py:160-259 (FermiNet layer)
  src/deepqmc/gnn/edge_features.py:21-78, gnn/graph.py:23-31 (features, receiver - sender)
"""
from __future__ import annotations

import math

import torch

from . import names as P  # the oracle's own table of the reference's Haiku parameter names

from .hamil import safe_norm


def _t(params, name):
    return params[name]


def ne_features(r, R, log_rescale):
    diffs = r[:, None] - R[None]  # receiver(electron) - sender(nucleus)
    rr = safe_norm(diffs)
    if log_rescale:
        lg = torch.log1p(rr)
        feats = torch.cat([lg[..., None], diffs * (lg / rr)[..., None]], -1)
    else:
        feats = torch.cat([rr[..., None], diffs], -1)
    return feats.reshape(r.shape[0], -1), rr


def psiformer_embeddings(spec, params, r, R):
    N, d, H = spec.n_elec, spec.embedding_dim, spec.n_heads
    dh = d // H
    feats, _ = ne_features(r, R, True)
    spins = torch.cat([torch.ones(spec.n_up), -torch.ones(spec.n_down)]).to(r.dtype)[:, None]
    x = torch.cat([feats, spins], 1) @ _t(params, P.GNN + 'electron_embedding/linear:w')
    for l in range(spec.n_layers):
        a = P.attn_prefix(l)
        q = (x @ _t(params, a + 'multi_head_attention/query:w')).reshape(N, H, dh)
        k = (x @ _t(params, a + 'multi_head_attention/key:w')).reshape(N, H, dh)
        v = (x @ _t(params, a + 'multi_head_attention/value:w')).reshape(N, H, dh)
        logits = torch.einsum('thd,Thd->htT', q, k) / math.sqrt(dh)
        w = torch.softmax(logits, -1)
        o = torch.einsum('htT,Thd->thd', w, v).reshape(N, d)
        att = x + o @ _t(params, a + 'multi_head_attention/linear:w')
        m = torch.tanh(att @ _t(params, a + 'mlp/linear_0:w') + _t(params, a + 'mlp/linear_0:b'))
        m = torch.tanh(m @ _t(params, a + 'mlp/linear_1:w') + _t(params, a + 'mlp/linear_1:b'))
        x = att + m
    return x


def ferminet_embeddings(spec, params, r, R):
    n_up = spec.n_up
    x, _ = ne_features(r, R, False)  # [N, 4M], no projection

    def edge_feats(sender):
        d = r[None, :, :] - sender[:, None, :]  # [S, N, 3] receiver - sender
        return torch.cat([safe_norm(d)[..., None], d], -1)

    e_up, e_dn = edge_feats(r[:n_up]), edge_feats(r[n_up:])
    sq2 = math.sqrt(2.0)
    for l in range(spec.n_layers):
        lp = P.layer_prefix(l)
        f = torch.cat(
            [
                x,
                x[:n_up].mean(0, keepdim=True).expand(x.shape[0], -1),
                x[n_up:].mean(0, keepdim=True).expand(x.shape[0], -1),
                e_up.mean(0),
                e_dn.mean(0),
            ],
            -1,
        )
        upd = torch.tanh(f @ _t(params, lp + 'g/linear_0:w') + _t(params, lp + 'g/linear_0:b'))
        x_new = (x + upd) / sq2 if upd.shape == x.shape else upd
        if l < spec.n_layers - 1:
            wu, bu = _t(params, lp + 'u/linear_0:w'), _t(params, lp + 'u/linear_0:b')
            nu, nd = torch.tanh(e_up @ wu + bu), torch.tanh(e_dn @ wu + bu)
            if nu.shape == e_up.shape:
                nu, nd = (e_up + nu) / sq2, (e_dn + nd) / sq2
            e_up, e_dn = nu, nd
        x = x_new
    return x


def layer_norm(x, eps=1e-5):
    # hk.LayerNorm(-1, create_scale=False, create_offset=False) (reference: hkext.py:200-201)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def nuclei_embedding(spec, params, R):
    """reference: gnn/electron_gnn.py:435-537 (edge_features given, 'nn' edges with
    self-interaction, receiver - sender), edge features gnn/edge_features.py:21-78 log-rescaled."""
    M = spec.n_nuc
    d = R[None, :, :] - R[:, None, :]  # [sender, receiver, 3]
    rr = safe_norm(d)
    lg = torch.log1p(rr)
    feats = torch.cat([lg[..., None], d * (lg / rr)[..., None]], -1)  # [M, M, 4]
    ch = torch.as_tensor(spec.charges, dtype=R.dtype)
    inv = torch.unique(ch, return_inverse=True)[1]
    onehot = torch.nn.functional.one_hot(inv, M).to(R.dtype)  # type of the SENDER (axis 0)
    x = torch.cat([feats, onehot[:, None, :].expand(M, M, M)], -1)
    silu = torch.nn.functional.silu
    g = lambda nm: (_t(params, P.NUC_EMB + nm + ':w'), _t(params, P.NUC_EMB + nm + ':b'))
    w0, b0 = g('edge_mlp/linear_0'); w1, b1 = g('edge_mlp/linear_1')
    e = silu(x @ w0 + b0) @ w1 + b1
    v0, c0 = g('embed_mlp/linear_0'); v1, c1 = g('embed_mlp/linear_1')
    return silu(e.sum(0) @ v0 + c0) @ v1 + c1  # [M, d]


def transpsiformer_embeddings(spec, params, r, R):
    """Electron AND nucleus embeddings (reference: gnn/update_features.py:385-451: attention over
    [nuclei; electrons] tokens, nuclei masked from attending electrons; residual, tanh MLP, residual)."""
    N, M, d, H = spec.n_elec, spec.n_nuc, spec.embedding_dim, spec.n_heads
    dh = d // H
    feats, _ = ne_features(r, R, True)
    spins = torch.cat([torch.ones(spec.n_up), -torch.ones(spec.n_down)]).to(r.dtype)[:, None]
    xe = torch.cat([feats, spins], 1) @ _t(params, P.GNN + 'electron_embedding/linear:w')
    h = torch.cat([nuclei_embedding(spec, params, R), xe], 0)  # [M + N, d]
    mask = torch.ones(M + N, M + N, dtype=torch.bool)
    mask[:M, M:] = False
    for l in range(spec.n_layers):
        a = P.comb_prefix(l)
        q = (h @ _t(params, a + 'multi_head_attention/query:w')).reshape(M + N, H, dh)
        k = (h @ _t(params, a + 'multi_head_attention/key:w')).reshape(M + N, H, dh)
        v = (h @ _t(params, a + 'multi_head_attention/value:w')).reshape(M + N, H, dh)
        logits = torch.einsum('thd,Thd->htT', q, k) / math.sqrt(dh)
        logits = torch.where(mask[None], logits, torch.full_like(logits, -1e30))
        w = torch.softmax(logits, -1)
        o = torch.einsum('htT,Thd->thd', w, v).reshape(M + N, d)
        att = h + o @ _t(params, a + 'multi_head_attention/linear:w')
        m = torch.tanh(att @ _t(params, a + 'mlp/linear_0:w') + _t(params, a + 'mlp/linear_0:b'))
        m = torch.tanh(m @ _t(params, a + 'mlp/linear_1:w') + _t(params, a + 'mlp/linear_1:b'))
        h = att + m
    return h[M:], h[:M]


def nuclear_head_zetas(spec, params, nuc_emb):
    """reference: wf/omni.py:181-211 NuclearGNNHead + hkext.py:165-202 GLU (LayerNorm before,
    sigmoid gate) + bias initialised to 2 -> zetas_{up,down}[M, K, E]."""
    K, E = spec.n_determinants, spec.n_env_per_nuc
    x = layer_norm(nuc_emb)
    out = {}
    for glu, spin in (('zetas_readout_glu', 'up'), ('zetas_readout_glu_1', 'down')):
        W, bW = _t(params, P.HEAD + glu + '/W:w'), _t(params, P.HEAD + glu + '/W:b')
        V, bV = _t(params, P.HEAD + glu + '/V:w'), _t(params, P.HEAD + glu + '/V:b')
        y = torch.sigmoid(x @ W + bW) * (x @ V + bV)
        out[spin] = y.reshape(-1, K, E) + _t(params, P.HEAD + f':zetas_bias_{spin}')
    return out


def orbitals_nucdep(spec, params, emb, zetas, r, R):
    """SimplifiedNucleusDependentEnvelopes (reference: wf/env.py:111-226; fixed pi = 1,
    per_orbital_exponent = false: the envelope does not depend on the orbital index) (*) backflow."""
    N, K, n_up = spec.n_elec, spec.n_determinants, spec.n_up
    dist = safe_norm(r[:, None] - R[None])  # [N, M]

    def env(spin, sl):
        ex = torch.abs(dist[sl][:, :, None, None] * zetas[spin][None])  # [n, M, K, E]
        return torch.exp(-ex).sum((1, 3)).permute(1, 0)[:, :, None]  # [K, n, 1]

    def bf(w, sl):
        return (emb[sl] @ w).reshape(-1, K, N).permute(1, 0, 2)

    up, dn = slice(None, n_up), slice(n_up, None)
    a_up = env('up', up) * bf(_t(params, P.BF_UP + ':w'), up)
    a_dn = env('down', dn) * bf(_t(params, P.BF_DN + ':w'), dn)
    return torch.cat([a_up, a_dn], 1)


def ssp(x):
    # reference: hkext.py:11-19 shifted softplus
    return torch.nn.functional.softplus(x) + math.log(0.5)


def paulinet_embeddings(spec, params, r, R):
    """conv-GNN of the reference's test ansatz (tests/conf/ansatz.yaml) and of conf/ansatz/default.yaml:
    initial embeddings = hk.Embed lookup (gnn/electron_gnn.py:620-624; one electron type if n_up == n_down,
    :337-343) or the raw nucleus-electron features [|d|, d] (:596-611); per layer
    conv_t(i) = sum_senders w_t(e) * h_t(x_sender) for t in same / anti (/ ne) (gnn/update_features.py:162-238,
    graph.py:226-335; edges = receiver - sender, no self-interaction; features [|d| eps-safe, d],
    edge_features.py:21-78); update 'featurewise' sum_t g_t(conv_t) or 'concatenate'
    g([h, mean_up h, mean_down h, conv_same, conv_anti]) (update_features.py:47-121, electron_gnn.py:243-259) with
    (normalised) residual (hkext.py:116-137); optional shared edge MLP u with normalised residual between layers
    (electron_gnn.py:160-192)."""
    N, n_up = spec.n_elec, spec.n_up
    if spec.gnn_embedding == 'embed':
        emb = _t(params, P.GNN + 'electron_embedding/ElectronicEmbedding:embeddings')
        types = [0] * n_up + [int(spec.n_up != spec.n_down)] * spec.n_down
        x = emb[types]  # [N, d]
    else:
        x, _ = ne_features(r, R, False)  # [N, 4M]
    xn = _t(params, P.GNN + 'nuclei_embedding/~/embed:embeddings') if spec.gnn_conv_ne else None

    def feats(d):
        return torch.cat([safe_norm(d)[..., None], d], -1)

    def mlp(base, h, n, bias=True):
        for i in range(n):
            h = h @ _t(params, base + f'linear_{i}:w')
            if bias:
                h = h + _t(params, base + f'linear_{i}:b')
            h = torch.tanh(h)  # last_linear = false
        return h

    up = torch.arange(N) < n_up
    same = (up[:, None] == up[None, :]) & ~torch.eye(N, dtype=torch.bool)  # [sender j, receiver i]
    anti = up[:, None] != up[None, :]
    e_ee = feats(r[None, :, :] - r[:, None, :])  # [j, i, 4] receiver - sender
    e_ne = feats(r[None, :, :] - R[:, None, :]) if spec.gnn_conv_ne else None  # [I, i, 4]
    nl = spec.gnn_subnet_layers
    sq2 = math.sqrt(2.0)
    for l in range(spec.n_layers):
        c, lp = P.conv_prefix(l), P.layer_prefix(l)
        convs = []
        kinds = [('same', e_ee, same, x), ('anti', e_ee, anti, x)] + ([('ne', e_ne, None, xn)] if spec.gnn_conv_ne else [])
        for t, edges, mask, send in kinds:
            we = mlp(c + f'w_{t}/', edges, nl, bias=spec.gnn_update == 'concatenate')  # [senders, N, e]
            hx = mlp(c + f'h_{t}/', send, nl)
            prod = we * hx[:, None, :]
            if mask is not None:
                prod = prod * mask[:, :, None].to(prod.dtype)
            convs.append((t, prod.sum(0)))  # [N, e]
        if spec.gnn_update == 'featurewise':
            upd = 0
            for t, cv in convs:  # g_t: hkext.MLP, last_linear = false (one layer in tests/conf/ansatz.yaml, three in gnn.yaml)
                n_g = sum(1 for i in range(8) if (lp + f'g_conv_{t}/linear_{i}:w') in params)
                upd = upd + mlp(lp + f'g_conv_{t}/', cv, n_g)
        else:
            f = torch.cat([x, x[:n_up].mean(0, keepdim=True).expand(N, -1), x[n_up:].mean(0, keepdim=True).expand(N, -1)]
                          + [cv for _, cv in convs], -1)
            upd = f @ _t(params, lp + 'g/linear_0:w')
            if spec.gnn_g_bias:
                upd = upd + _t(params, lp + 'g/linear_0:b')
            upd = torch.tanh(upd)
        if upd.shape == x.shape:
            x = (x + upd) / sq2 if spec.gnn_residual_normalize else x + upd
        else:
            x = upd
        if spec.gnn_deep_edges and l < spec.n_layers - 1:  # shared edge MLP, same + anti edges alike
            ne_ = mlp(lp + 'u/', e_ee, nl)
            e_ee = (e_ee + ne_) / sq2 if ne_.shape == e_ee.shape else ne_
    return x


def paulinet_log_psi(spec, params, r, R, return_mos=False):
    """reference: wf/nn_wave_function.py:127-173 with full_determinant = False, mult backflow with the
    default mult_act, hk.Linear conf_coeff, DeepQMCCusp, Jastrow (wf/omni.py:13-40, sum_first)."""
    N, K, n_up, n_dn = spec.n_elec, spec.n_determinants, spec.n_up, spec.n_down
    x = paulinet_embeddings(spec, params, r, R)

    def mlp(base, h, n_lin, act, bias_last, bias=True):
        for i in range(n_lin):
            h = h @ _t(params, base + f'linear_{i}:w')
            if bias and (i < n_lin - 1 or bias_last):
                h = h + _t(params, base + f'linear_{i}:b')
            if i < n_lin - 1:
                h = act(h)
        return h

    jastrow = mlp(P.JASTROW, x.sum(0), spec.jastrow_layers, ssp, False).squeeze(-1) if spec.jastrow_layers else 0.0
    # envelopes (wf/env.py:57-75: per_shell, shared exponents, spin-restricted) -> [K, N_el, N_orb]
    if spec.env_per_shell:
        centers = list(spec.env_centers)
        dist = safe_norm(r[:, None] - R[None])[:, centers]  # [N, n_env]
        zeta, pi = _t(params, f'{P.ENV}:zetas'), _t(params, f'{P.ENV}:pi')
        orb = (pi[None] * torch.exp(-torch.abs(zeta * dist))[:, None, :]).sum(-1)  # [N, K*N]
        orb = orb.reshape(N, K, N).permute(1, 0, 2)
    else:  # per-orbital exponents, one shell per nucleus, spin-unrestricted (conf/ansatz/default.yaml:3-11)
        dist = safe_norm(r[:, None] - R[None])  # [N, M]
        rows = []
        for spin, sl in (('up', slice(0, n_up)), ('down', slice(n_up, N))):
            zeta, pi = _t(params, f'{P.ENV}:zetas_{spin}'), _t(params, f'{P.ENV}:pi_{spin}')
            rows.append((pi[None] * torch.exp(-torch.abs(zeta[None] * dist[sl][:, None, :]))).sum(-1))  # [n, K*N]
        orb = torch.cat(rows, 0).reshape(N, K, N).permute(1, 0, 2)
    mult = (lambda v: 1 + 2 * torch.tanh(v / 4)) if spec.mult_act == 'default' else (lambda v: v)
    signs, logs, blocks = 1.0, 0.0, []
    for sl, osl, pre, n in ((slice(0, n_up), slice(0, n_up), P.BF_UP, n_up), (slice(n_up, N), slice(n_up, N), P.BF_DN, n_dn)):
        base = pre.rsplit('linear_0', 1)[0]
        n_orb = N if spec.full_determinant else n
        f = mlp(base, x[sl], spec.backflow_layers, ssp, True, bias=spec.backflow_bias)  # [n, K * n_orb]
        f = f.reshape(n, K, n_orb).permute(1, 0, 2)  # wf/omni.py:78-88
        if spec.full_determinant:
            osl = slice(0, N)
        a = orb[:, sl, osl] * mult(f)
        if spec.full_determinant or return_mos:
            blocks.append(a)
            continue
        s, l = torch.linalg.slogdet(a) if n > 0 else (torch.ones(K, dtype=r.dtype), torch.zeros(K, dtype=r.dtype))
        signs, logs = signs * s, logs + l
    if return_mos:
        return tuple(blocks)
    if spec.full_determinant:
        signs, logs = torch.linalg.slogdet(torch.cat(blocks, 1))
    shift = logs.max().detach()
    if torch.isinf(shift):
        shift = torch.zeros_like(shift)
    xs = signs * torch.exp(logs - shift)
    psi = (xs @ _t(params, P.CONF + ':w')).squeeze() if spec.conf_coeff == 'linear' else xs.sum()
    log = torch.log(torch.abs(psi)) + shift
    sgn = torch.sign(psi).detach()
    if spec.cusp == 'deepqmc':  # wf/cusp.py:5-14
        al = spec.cusp_alpha
        for i in range(N):
            for j in range(i + 1, N):
                sc = spec.cusp_same_scale if (i < n_up) == (j < n_up) else spec.cusp_anti_scale
                log = log - sc / (al * (1 + al * safe_norm(r[i] - r[j])))
    return sgn, log + jastrow


def orbitals(spec, params, emb, r, R):
    """envelopes (*) backflow -> A[K, N, N] (electron i, orbital mu)."""
    N, K, n_up = spec.n_elec, spec.n_determinants, spec.n_up
    dist = safe_norm(r[:, None] - R[None])  # [N, M]

    def env(spin, sl):
        zeta, pi = _t(params, f'{P.ENV}:zetas_{spin}'), _t(params, f'{P.ENV}:pi_{spin}')
        ex = torch.abs(zeta[None] * dist[sl][:, None, :])  # [n, K*N, M]
        orb = (pi[None] * torch.exp(-ex)).sum(-1)  # [n, K*N]
        return orb.reshape(-1, K, N).permute(1, 0, 2)  # [K, n, N]

    def bf(w, sl):
        return (emb[sl] @ w).reshape(-1, K, N).permute(1, 0, 2)

    up, dn = slice(None, n_up), slice(n_up, None)
    mode = spec.backflow_transform
    if mode == 'mult':
        a_up = env('up', up) * bf(_t(params, P.BF_UP + ':w'), up)
        a_dn = env('down', dn) * bf(_t(params, P.BF_DN + ':w'), dn)
        return torch.cat([a_up, a_dn], 1)
    # additive branch of the BackflowOp (wf/nn_wave_function.py:14-33,111-125); Backflow(multi_head=True) builds one
    # net per transform: 'mlp' (first) and 'mlp_1' (wf/omni.py:69-73)
    dists_nuc = torch.sqrt(((r[:, None] - R[None]) ** 2).sum(-1))  # plain norm, nn_wave_function.py:128-129
    out = []
    for spin, sl, first, second in (('up', up, P.BF_UP, P.BF_UP_ADD), ('down', dn, P.BF_DN, P.BF_DN_ADD)):
        f_mult = bf(_t(params, first + ':w'), sl) if mode == 'both' else None
        f_add = bf(_t(params, (second if mode == 'both' else first) + ':w'), sl)
        out.append(backflow_op(env(spin, sl), f_mult, f_add, dists_nuc[sl]))
    return torch.cat(out, 1)


def backflow_op(xs, f_mult, f_add, dists_nuc):
    """reference: wf/nn_wave_function.py:14-33 with the identity mult_act of the Psiformer / FermiNet configs, the default
    add_act 0.1 tanh(x / 4) and with_envelope = True.  xs[K, n, n_orb], dists_nuc[n, M]."""
    envel = torch.sqrt((xs ** 2).sum((-1, -3), keepdim=True))
    if f_mult is not None:
        xs = xs * f_mult
    if f_add is not None:
        Rr = dists_nuc.min(-1).values / 0.5
        cutoff = torch.where(Rr < 1, Rr ** 2 * (6 - 8 * Rr + 3 * Rr ** 2), torch.ones_like(Rr))
        xs = xs + cutoff[None, :, None] * envel * (0.1 * torch.tanh(f_add / 4))
    return xs


def psiformer_cusp(spec, params, r):
    n_up, N = spec.n_up, spec.n_elec
    a_s, a_a = _t(params, f'{P.CUSP}:same_alpha'), _t(params, f'{P.CUSP}:anti_alpha')
    out = torch.zeros((), dtype=r.dtype)
    for i in range(N):
        for j in range(i + 1, N):
            dij = safe_norm(r[i] - r[j])
            same = (i < n_up) == (j < n_up)
            if same:
                out = out - spec.cusp_same_scale * a_s**2 / (a_s + dij)
            else:
                out = out - spec.cusp_anti_scale * a_a**2 / (a_a + dij)
    return out


def nuclear_cusp(spec, params, r, R):
    """reference: wf/cusp.py:81-101 on dists_nuc = plain norm (wf/nn_wave_function.py:129,169-170)"""
    if spec.cusp_nuclei == 'none':
        return 0.0
    al = _t(params, f'{P.NUC_CUSP}:nuc_alpha') if spec.cusp_nuclei_trainable else spec.cusp_nuclei_alpha
    z = torch.as_tensor(spec.charges, dtype=r.dtype)
    dist = torch.sqrt(((r[:, None] - R[None]) ** 2).sum(-1))  # [N, M]
    if spec.cusp_nuclei == 'psiformer':
        return -((z[None] * al**2) / (al + dist)).sum()
    return -(z[None] / (al * (1 + al * dist))).sum()


def log_psi(spec, params, r, R):
    """ansatz.apply for one walker -> (sign, log|psi|); reference nn_wave_function.py:127-173"""
    s, l = _log_psi(spec, params, r, R)
    return s, l + nuclear_cusp(spec, params, r, R)


def molecular_orbitals(spec, params, r, R):
    """ansatz.apply(..., return_mos=True) for one walker -> (orb_up[K, n_up, n_orb], orb_down[K, n_down, n_orb]);
    reference nn_wave_function.py:131-142"""
    return _log_psi(spec, params, r, R, return_mos=True)


def _log_psi(spec, params, r, R, return_mos=False):
    if spec.kind == 'paulinet':
        return paulinet_log_psi(spec, params, r, R, return_mos)
    if spec.kind == 'transpsiformer':
        emb, nuc = transpsiformer_embeddings(spec, params, r, R)
        A = orbitals_nucdep(spec, params, emb, nuclear_head_zetas(spec, params, nuc), r, R)
    else:
        emb = (psiformer_embeddings if spec.kind == 'psiformer' else ferminet_embeddings)(spec, params, r, R)
        A = orbitals(spec, params, emb, r, R)
    if return_mos:
        return A[:, :spec.n_up], A[:, spec.n_up:]
    sign, ld = torch.linalg.slogdet(A)
    shift = ld.max().detach()
    if torch.isinf(shift):
        shift = torch.zeros_like(shift)
    psi = (sign * torch.exp(ld - shift)).sum()
    log = torch.log(torch.abs(psi)) + shift
    sgn = torch.sign(psi).detach()
    if spec.cusp == 'psiformer':
        log = log + psiformer_cusp(spec, params, r)
    return sgn, log


def to_torch(params, dtype=torch.float64):
    return {k: torch.as_tensor(v, dtype=dtype) for k, v in params.items()}
