"""No-gradient paths built from the same kernels: deterministic posterior mean
(``predict_Y``, code/nmgp_dsvi.py:666-722) and the Monte-Carlo ELBO (``compute_ELBO``,
code/nmgp_dsvi.py:303-404)."""
from __future__ import annotations

import numpy as np
import torch

from . import _ops as ops
from . import dsvi_step as _step
from .dsvi_step import (EPS, HYPER_ORDER, H_LEN_ELL, H_LEN_L0, H_LEN_L1, H_S2_ELL, H_S2_L0, H_S2_L1, MODE_U, MODE_W,
                        packed_pair_index)

F64 = torch.float64


def _stationary_systems(x, Z, hyp, want_c=False):
    B, Q = x.shape[0], Z.shape[0]
    out = {}
    for name, is2, ilen in (("ell", H_S2_ELL, H_LEN_ELL), ("L0", H_S2_L0, H_LEN_L0), ("L1", H_S2_L1, H_LEN_L1)):
        A = ops.rbf_build_fwd(Z, Z, hyp, is2, ilen, EPS).reshape(1, Q, Q)
        R, hldR = ops.potrf(A, 0.0)
        K12 = ops.rbf_build_fwd(x, Z, hyp, is2, ilen, 0.0).reshape(1, B, Q)
        P, c = ops.solve_rows_fwd(K12, R)
        out[name] = dict(R=R, hldR=hldR, K12=K12, P=P, c=c)
    return out


def posterior_mean(p, Z, x, I):
    """Y*[n] = sum_{j<=I[n]} Lhat[I[n],j](x_n) Ghat[j](x_n) with Lhat/Ghat the conditional means through the
    inducing points (MGP_mu, code/utils.py:149-157); only the (row, own-output) entries the reference gathers
    at :721-722 are evaluated.  x sorted by output id, I int32."""
    D, Q = p["mu_W"].shape
    B = x.shape[0]
    dev = x.device
    zeros = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
    hyp = ops.hyper_exp(torch.stack([p[k].detach().reshape(()) for k in HYPER_ORDER]))
    sysm = _stationary_systems(x, Z, hyp)
    mu_v = p["mu_v"].detach().contiguous()
    v, ellZ = ops.sample_v_fwd(mu_v, zeros(Q, Q), zeros(1, Q))               # v = mu_v, ellZ = exp(mu_v)
    ellx = ops.ell_rows_fwd(sysm["ell"]["P"][0], v, zeros(1, B), zeros(B))     # exp(P mu_v)
    flat = packed_pair_index(D, dev)
    muU = p["mu_U"].detach().reshape(D * D, Q).index_select(0, flat)
    mU = ops.pair_means(sysm["L0"]["P"], sysm["L1"]["P"], I, muU, D, MODE_U)
    lhat = ops.coef_sample_fwd(mU[0], zeros(B, D), zeros(1, B, D), I)          # exp on the diagonal entry
    A_G = ops.gibbs_build_fwd(Z, Z, ellZ, ellZ, EPS)
    R_G, _ = ops.potrf(A_G, 0.0)
    KG = ops.gibbs_build_fwd(x, Z, ellx, ellZ, 0.0)
    PG, _ = ops.solve_rows_fwd(KG, R_G)
    ghat = ops.pair_means(PG, PG, I, p["mu_W"].detach().contiguous(), D, MODE_W)
    return ops.rowdot_live(lhat, ghat, I)[0]


_q5_cache = {}


def _transposed_packing(D, dev):
    """Packing that makes the step kernels evaluate the reference's compute_ELBO gather (quirk q5,
    code/nmgp_dsvi.py:361: ``sampled_L.permute(2,1,0)[n, I[n]]`` = column I[n] of L(x_n), entries i >= I[n]).
    With outputs and latents relabelled i' = D-1-i the column gather becomes the row gather the kernels implement:
    slot(I', j' <= I') holds the original pair (i = D-1-j', c = D-1-I')."""
    key = (D, str(dev))
    if key not in _q5_cache:
        idx = [(D - 1 - ip) * D + (D - 1 - ip) for ip in range(D)]
        idx += [(D - 1 - jp) * D + (D - 1 - ip) for ip in range(D) for jp in range(ip)]
        _q5_cache[key] = (torch.tensor(idx, dtype=torch.int64, device=dev),
                          torch.arange(D - 1, -1, -1, dtype=torch.int64, device=dev))
    return _q5_cache[key]


def mc_elbo(model, inputs_list, outputs_list, index=None, n_sample=1000, verbose=False, noise=None, chunk=32):
    """code/nmgp_dsvi.py:303-404 compute_ELBO with its quirks (q5): transposed coefficient gather, no N/B factor,
    KL_W evaluated with the LAST draw's Gibbs K22, mean over draws of the expected log-likelihood."""
    from .nmgp_dsvi import _rows_from_lists, TensorType
    D, Q = model.D, model.M
    dev = model.device
    x, y, I, perm = _rows_from_lists(inputs_list, outputs_list, D, index)
    B = x.shape[0]
    if model.N != B and model.N != 1:
        raise RuntimeError("compute_ELBO: the reference sizes K_tildeell_11_diag with self.N=%d, which does not "
                           "broadcast against %d rows (code/nmgp_dsvi.py:334)" % (model.N, B))
    order = np.arange(B) if perm is None else perm                      # caller row -> sorted position
    Is = I[order]
    rev = order[::-1].copy()                                            # relabelled problem: rows by I' = D-1-I ascending
    Ip = (D - 1 - I[rev]).astype(np.int32)
    pair_index, latent_order = _transposed_packing(D, dev)
    up = lambda t: torch.as_tensor(t).to(dev, dtype=F64).contiguous()
    xr, yr = up(x[torch.from_numpy(rev)]), up(y[torch.from_numpy(rev)])
    Ipd = torch.from_numpy(Ip).to(dev)
    p = {k: getattr(model, k).detach() for k in _step.PARAM_NAMES}
    noise = noise or model.noise
    rows_of = [torch.from_numpy(np.nonzero(I[rev] == c)[0]) for c in range(D)]   # relabelled rows whose output is c
    rev_t = torch.from_numpy(rev)
    Rs, klW_last, kl_v, kl_U = [], None, None, None
    for s0 in range(0, n_sample, chunk):
        ns = min(chunk, n_sample - s0)
        if noise == "reference":
            zv = torch.empty(ns, Q, dtype=F64); zell = torch.empty(ns, B, dtype=F64); zL = torch.zeros(ns, B, D, dtype=F64)
            for s in range(ns):
                zv[s] = torch.randn(Q).type(TensorType)
                zell[s] = torch.randn(B).type(TensorType)[rev_t]
                for i in range(D):
                    for j in range(i + 1):
                        z = torch.randn(B).type(TensorType)[rev_t]      # draw for pair (i, j); used by rows of output j
                        if rows_of[j].numel():
                            zL[s, rows_of[j], D - 1 - i] = z[rows_of[j]]
        else:
            zv = torch.randn(ns, Q, device=dev, dtype=torch.float32).to(F64)
            zell = torch.randn(ns, B, device=dev, dtype=torch.float32).to(F64)
            zL = torch.randn(ns, B, D, device=dev, dtype=torch.float32).to(F64)
        aux = {}
        _step.dsvi_step(p, model.Z.reshape(-1), xr, yr, Ipd, B, up(zv), up(zell), up(zL), want_grads=False,
                        pair_index=pair_index, latent_order=latent_order, aux=aux)
        Rs.append(aux["Rsum"])
        klW_last, kl_v, kl_U = aux["kl_W"][-1].sum(), aux["kl_v"], aux["kl_U"]
        if verbose:
            print("Monte Carlo index:", s0 + ns)
    return torch.cat(Rs).mean() - klW_last - kl_v - kl_U


# ------------------------------------------------------------------------------------------------------------
# posterior sampling (code/nmgp_dsvi.py:406-580): same kernels as the training step, no gradients
def _sampling_setup(model, x, I):
    """Sample-independent part: stationary systems, coefficient statistics, variational covariances."""
    D, Q = model.D, model.M
    dev = x.device
    p = {k: getattr(model, k).detach() for k in _step.PARAM_NAMES}
    hyp = ops.hyper_exp(torch.stack([p[k].reshape(()) for k in HYPER_ORDER]))
    sysm = _stationary_systems(x, model.Z.reshape(-1), hyp)
    flat = packed_pair_index(D, dev)
    SU = p["sqrt_U"].reshape(D * D, Q, Q).index_select(0, flat)
    muU = p["mu_U"].reshape(D * D, Q).index_select(0, flat)
    Sig_U = ops.tril_syrk_fwd(SU)
    seg = ops.segment_offsets(I, D)
    qU, mU = ops.quadform_fwd(sysm["L0"]["P"], sysm["L1"]["P"], I, Sig_U, muU, D, MODE_U, seg=seg)
    sdU = ops.coef_sd_fwd(qU[0], sysm["L0"]["c"][0], sysm["L1"]["c"][0], I, hyp)
    C_v, _ = ops.potrf(ops.tril_syrk_fwd(p["sqrt_v"].reshape(1, Q, Q)), EPS)
    return dict(p=p, hyp=hyp, sysm=sysm, mU=mU[0], sdU=sdU, C_v=C_v[0], Sig_W=ops.tril_syrk_fwd(p["sqrt_W"]),
                mu_W=p["mu_W"].contiguous(), mu_v=p["mu_v"].contiguous())


def _draw_latents(st, Z, xg, ellx, ellZ, zG):
    """G ~ q(g(x)) marginal per (point, latent): all D latents for every point (MGP_d with batched mu_W, Sigma_W)."""
    D = st["mu_W"].shape[0]
    Bg = xg.shape[0]
    A_G = ops.gibbs_build_fwd(Z, Z, ellZ, ellZ, EPS)
    R_G, _ = ops.potrf(A_G, 0.0)
    KG = ops.gibbs_build_fwd(xg, Z, ellx, ellZ, 0.0)
    PG, cG = ops.solve_rows_fwd(KG, R_G)
    I_all = torch.full((Bg,), D - 1, dtype=torch.int32, device=xg.device)
    qg, mg = ops.quadform_fwd(PG, PG, I_all, st["Sig_W"], st["mu_W"], D, MODE_W)
    s2g = (1.0 - cG.unsqueeze(-1) + qg).contiguous()
    return ops.reparam_diag(mg.contiguous(), s2g, zG.contiguous())           # [ns, Bg, D]


def sample_Y(model, inputs_list, index=None, n_sample=1000, noise=None, chunk=16):
    """code/nmgp_dsvi.py:406-491: posterior draws of Y, the used row of L, G and tilde-ell at the given rows."""
    from .nmgp_dsvi import _rows_from_lists, TensorType
    D, Q = model.D, model.M
    dev = model.device
    x, _, I, perm = _rows_from_lists(inputs_list, None, D, index)
    B = x.shape[0]
    order = np.arange(B) if perm is None else perm
    ot = torch.from_numpy(order)
    inv = torch.empty(B, dtype=torch.int64); inv[ot] = torch.arange(B)
    xs = x[ot].to(dev, dtype=F64).contiguous()
    Is_np = I[order]
    Id = torch.from_numpy(Is_np.astype(np.int32)).to(dev)
    st = _sampling_setup(model, xs, Id)
    sd_ell = ops.ell_sd_fwd(st["sysm"]["ell"]["c"][0], st["hyp"])
    Z = model.Z.reshape(-1)
    s2e = st["hyp"][_step.H_S2_ERR]
    noise = noise or model.noise
    sel = [torch.from_numpy(np.nonzero(Is_np == i)[0]) for i in range(D)]
    Ys, Ls, Gs, Es = [], [], [], []
    for s0 in range(0, n_sample, chunk):
        ns = min(chunk, n_sample - s0)
        if noise == "reference":
            zv = torch.empty(ns, Q, dtype=F64); zell = torch.empty(ns, B, dtype=F64); zL = torch.zeros(ns, B, D, dtype=F64)
            zG = torch.empty(ns, B, D, dtype=F64); zF = torch.empty(ns, B, dtype=F64)
            for s in range(ns):
                zv[s] = torch.randn(Q).type(TensorType)
                zell[s] = torch.randn(B).type(TensorType)[ot]
                for i in range(D):
                    for j in range(i + 1):
                        z = torch.randn(B).type(TensorType)[ot]
                        if sel[i].numel():
                            zL[s, sel[i], j] = z[sel[i]]
                zG[s] = torch.randn(D, B).type(TensorType)[:, ot].t()
                zF[s] = torch.randn(B).type(TensorType)[ot]
            zv, zell, zL, zG, zF = (t.to(dev).contiguous() for t in (zv, zell, zL, zG, zF))
        else:
            rn = lambda *sh: torch.randn(*sh, device=dev, dtype=torch.float32).to(F64)
            zv, zell, zL, zG, zF = rn(ns, Q), rn(ns, B), rn(ns, B, D), rn(ns, B, D), rn(ns, B)
        v, ellZ = ops.sample_v_fwd(st["mu_v"], st["C_v"], zv)
        ellx = ops.ell_rows_fwd(st["sysm"]["ell"]["P"][0], v, zell, sd_ell)
        l = ops.coef_sample_fwd(st["mU"], st["sdU"], zL, Id)
        G = _draw_latents(st, Z, xs, ellx, ellZ, zG)
        F = ops.rowdot_live(l, G, Id)
        Y = ops.reparam_diag(F.contiguous(), s2e.expand_as(F).contiguous(), zF)
        Ys.append(Y[:, inv.to(dev)]); Ls.append(l[:, inv.to(dev)]); Gs.append(G[:, inv.to(dev)].permute(0, 2, 1))
        Es.append(torch.log(ellx)[:, inv.to(dev)])
    return torch.cat(Ys), torch.cat(Ls), torch.cat(Gs), torch.cat(Es)


def sample_FY(model, inputs, n_sample=1000, noise=None, chunk=16):
    """code/nmgp_dsvi.py:493-580: draws of tilde-ell, Y = L G for ALL outputs, and the output correlation matrices
    corr(L L^T) on a common grid.  Returns (tilde_ells [S,B], Ys [S,B,D], corrs [S,B,D,D])."""
    from .nmgp_dsvi import TensorType
    D, Q = model.D, model.M
    dev = model.device
    xg = inputs.reshape(-1).to(dev, dtype=F64).contiguous()
    B = xg.shape[0]
    xrep = xg.repeat(D).contiguous()                                           # row i*B + n  <->  (output i, point n)
    Irep = torch.arange(D, dtype=torch.int32, device=dev).repeat_interleave(B).contiguous()
    st = _sampling_setup(model, xrep, Irep)
    Z = model.Z.reshape(-1)
    # tilde-ell is per point: inducing solve on the grid itself
    sys_g = _stationary_systems(xg, Z, st["hyp"])
    sd_ell = ops.ell_sd_fwd(sys_g["ell"]["c"][0], st["hyp"])
    s2e = st["hyp"][_step.H_S2_ERR]
    noise = noise or model.noise
    Es, Ys, Cs = [], [], []
    for s0 in range(0, n_sample, chunk):
        ns = min(chunk, n_sample - s0)
        if noise == "reference":
            zv = torch.empty(ns, Q, dtype=F64); zell = torch.empty(ns, B, dtype=F64); zL = torch.zeros(ns, D * B, D, dtype=F64)
            zG = torch.empty(ns, B, D, dtype=F64); zF = torch.empty(ns, B, D, dtype=F64)
            for s in range(ns):
                zv[s] = torch.randn(Q).type(TensorType)
                zell[s] = torch.randn(B).type(TensorType)
                for i in range(D):
                    for j in range(i + 1):
                        zL[s, i * B:(i + 1) * B, j] = torch.randn(B).type(TensorType)
                zG[s] = torch.randn(D, B).type(TensorType).t()
                zF[s] = torch.randn(B, D).type(TensorType)
            zv, zell, zL, zG, zF = (t.to(dev).contiguous() for t in (zv, zell, zL, zG, zF))
        else:
            rn = lambda *sh: torch.randn(*sh, device=dev, dtype=torch.float32).to(F64)
            zv, zell, zL, zG, zF = rn(ns, Q), rn(ns, B), rn(ns, D * B, D), rn(ns, B, D), rn(ns, B, D)
        v, ellZ = ops.sample_v_fwd(st["mu_v"], st["C_v"], zv)
        ellx = ops.ell_rows_fwd(sys_g["ell"]["P"][0], v, zell, sd_ell)
        l = ops.coef_sample_fwd(st["mU"], st["sdU"], zL, Irep)                 # [ns, D*B, D]: row I of L at every point
        G = _draw_latents(st, Z, xg, ellx, ellZ, zG)                           # [ns, B, D]
        F = ops.rowdot_live(l, G.repeat(1, D, 1).contiguous(), Irep)           # [ns, D*B]
        F = F.reshape(ns, D, B).permute(0, 2, 1).contiguous()                  # [ns, B, D]
        Y = ops.reparam_diag(F, s2e.expand_as(F).contiguous(), zF)
        Lfull = l.reshape(ns, D, B, D).permute(0, 2, 1, 3).contiguous()        # [ns, B, D(i), D(j)]
        corr = ops.lcorr(Lfull.reshape(ns * B, D, D)).reshape(ns, B, D, D)
        Es.append(torch.log(ellx)); Ys.append(Y); Cs.append(corr)
    return torch.cat(Es), torch.cat(Ys), torch.cat(Cs)
