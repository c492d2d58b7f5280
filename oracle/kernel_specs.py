"""Per-kernel CPU specifications -- TEST INFRASTRUCTURE ONLY.

One plain torch-CPU float64 function per entry point of the C ABI
(include/nmgp_b200.h), with the same tensor-level signature as the thin
wrappers in ``collaborative_nonstationary_multivariate_gaussian_process_b200/_ops.py``.
They serve two purposes:

* GPU unit tests compare every CUDA kernel against its spec on seeded inputs;
* CPU tests monkeypatch ``_ops`` with these functions to check the hand-written
  adjoint orchestration (``dsvi_step.py``) against the autograd oracle without a GPU.

The product never imports this module.  Formulas follow SURVEY.md Appendix A
(derived from code/utils.py and code/nmgp_dsvi.py of the reference).
"""
import torch

F64 = torch.float64
EPS = 1e-4                      # code/utils.py:7
MODE_W, MODE_U = 0, 1
# hyper-parameter slots (values = exp of the reference's *_log parameters)
H_S2_ELL, H_LEN_ELL, H_S2_L0, H_LEN_L0, H_S2_L1, H_LEN_L1, H_S2_ERR = range(7)


def hyper_exp(logs):
    return torch.exp(logs)


def segment_offsets(I, D):
    """seg[d] = first row with I >= d (rows sorted by output id); seg[D] = B."""
    return torch.searchsorted(I.long().contiguous(), torch.arange(D + 1)).to(torch.int32)


# ---- small batched Q x Q ---------------------------------------------------------------
def tril_syrk_fwd(S):
    L = torch.tril(S)
    return L @ L.transpose(-1, -2)


def tril_syrk_bwd(S, SigBar):
    L = torch.tril(S)
    return torch.tril((SigBar + SigBar.transpose(-1, -2)) @ L)


def raise_if_not_pd(info):
    if int(info.item()) != 0:
        raise RuntimeError("cholesky: not positive-definite")


def potrf(A, jitter=0.0, info=None):
    """Same protocol as _ops.potrf: raises like torch.cholesky, or -- with ``info`` -- records 1 + index of the first
    failing matrix there and returns (deferred check, dsvi_step)."""
    n = A.shape[-1]
    C, bad = torch.linalg.cholesky_ex(A + jitter * torch.eye(n, dtype=F64))
    if bool((bad != 0).any()):
        first = int(torch.nonzero(bad.reshape(-1) != 0)[0])
        if info is None:
            raise RuntimeError("cholesky: matrix %d of a batch is not positive-definite" % first)
        if int(info) == 0:
            info.fill_(1 + first)
    return C, C.diagonal(dim1=-2, dim2=-1).log().sum(-1)


def potrf_bwd(C, Cbar, hldbar):
    """Abar (symmetric) for C = chol(A); hld = sum log diag C folded in."""
    Cb = torch.tril(Cbar) + torch.diag_embed(hldbar.unsqueeze(-1) / C.diagonal(dim1=-2, dim2=-1))
    Pm = torch.tril(C.transpose(-1, -2) @ Cb)
    Pm = Pm - 0.5 * torch.diag_embed(Pm.diagonal(dim1=-2, dim2=-1))
    X = torch.linalg.solve_triangular(C.transpose(-1, -2), Pm, upper=True)             # C^-T Phi
    Y = torch.linalg.solve_triangular(C.transpose(-1, -2), X.transpose(-1, -2), upper=True)  # C^-T (C^-T Phi)^T
    Ab = Y.transpose(-1, -2)                                                            # C^-T Phi C^-1
    return 0.5 * (Ab + Ab.transpose(-1, -2))


def kl_fwd(CS, hldS, mu, R, hldR, exact=False):
    """kl[p,b] = hldR[p] - hldS[b] + 0.5 (term2 + ||R_p^-1 mu_b||^2 - Q); term2 is the
    reference's diagonal-only form (quirk q10) unless exact."""
    np_, nb, Q = R.shape[0], CS.shape[0], CS.shape[-1]
    # t[p,b,:] = R_p^-1 mu_b
    t = torch.linalg.solve_triangular(R, mu.t().unsqueeze(0).expand(np_, Q, nb), upper=False).transpose(1, 2)
    term3 = (t * t).sum(-1)
    if exact:
        sol = torch.linalg.solve_triangular(R.unsqueeze(1), CS.unsqueeze(0), upper=False)
        term2 = (sol * sol).sum((-1, -2))
    else:
        rs = (CS * CS).sum(-1)                                  # [nb,Q]
        w = 1.0 / R.diagonal(dim1=-2, dim2=-1) ** 2             # [np,Q]
        term2 = w @ rs.t()
    kl = hldR.unsqueeze(1) - hldS.unsqueeze(0) + 0.5 * (term2 + term3 - Q)
    return kl, t.contiguous()


def atb(A, Bm, C, sign=1.0):
    C += sign * A.transpose(1, 2) @ Bm
    return C


def kl_rbar(R, G, Rbar):
    Rbar -= torch.tril(G @ R)
    return Rbar


def kl_bwd(klbar, CS, mu, R, t, exact=False):
    if exact:                                # autograd of the exact forward (test infrastructure may be slow)
        with torch.enable_grad():
            CSv = CS.clone().requires_grad_(True); muv = mu.clone().requires_grad_(True); Rv = R.clone().requires_grad_(True)
            hS = torch.zeros(CS.shape[0], dtype=F64, requires_grad=True); hR = torch.zeros(R.shape[0], dtype=F64, requires_grad=True)
            kl, _ = kl_fwd(CSv, hS, muv, Rv, hR, exact=True)
            g = torch.autograd.grad((kl * klbar).sum(), [CSv, hS, muv, Rv, hR])
        return torch.tril(g[0]), g[1], g[2], torch.tril(g[3]), g[4]
    np_, nb, Q = R.shape[0], CS.shape[0], CS.shape[-1]
    d = R.diagonal(dim1=-2, dim2=-1)                            # [np,Q]
    w = 1.0 / d ** 2
    rs = (CS * CS).sum(-1)
    hldRbar = klbar.sum(1)
    hldSbar = -klbar.sum(0)
    # term2: 0.5 * sum_a w[p,a] rs[b,a]
    wbar = 0.5 * klbar @ rs                                     # [np,Q]
    rsbar = 0.5 * klbar.t() @ w                                 # [nb,Q]
    CSbar = torch.tril(2.0 * rsbar.unsqueeze(-1) * CS)
    Rbar = torch.diag_embed(wbar * (-2.0) / d ** 3)
    # term3: 0.5 * ||t||^2, t = R^-1 mu
    tbar = klbar.unsqueeze(-1) * t                              # [np,nb,Q]
    mb = torch.linalg.solve_triangular(R.transpose(-1, -2), tbar.transpose(1, 2), upper=True)   # [np,Q,nb] = R^-T tbar
    mubar = mb.sum(0).t().contiguous()
    Rbar = Rbar - torch.tril(mb @ t)                            # -sum_b (R^-T tbar_b) t_b^T
    return CSbar, hldSbar, mubar, Rbar, hldRbar


# ---- kernel builds ---------------------------------------------------------------------
def rbf_build_fwd(x, z, hyp, is2, ilen, jitter=0.0):
    r = x.view(-1, 1) / hyp[ilen] - z.view(1, -1) / hyp[ilen]
    K = hyp[is2] * torch.exp(-0.5 * r * r)
    if jitter:
        K = K + jitter * torch.eye(K.shape[0], K.shape[1], dtype=F64)
    return K


def rbf_build_bwd(x, z, hyp, is2, ilen, Kbar, ghyp):
    r = x.view(-1, 1) / hyp[ilen] - z.view(1, -1) / hyp[ilen]
    K = hyp[is2] * torch.exp(-0.5 * r * r)
    ghyp[is2] += (Kbar * K).sum()
    ghyp[ilen] += (Kbar * K * r * r).sum()


def gibbs_build_fwd(x, z, ellx, ellz, jitter=0.0):
    r2 = (x.view(1, -1, 1) - z.view(1, 1, -1)) ** 2
    a = ellx.unsqueeze(2); b = ellz.unsqueeze(1)
    den = a * a + b * b
    K = torch.sqrt(2 * (a * b) / den) * torch.exp(-r2 / den)
    if jitter:
        K = K + jitter * torch.eye(K.shape[1], K.shape[2], dtype=F64)
    return K


def gibbs_build_bwd(x, z, ellx, ellz, Kbar, ellxbar, ellzbar, Kfwd=None):
    """ellxbar[ns,B] = ; ellzbar[ns,Q] += (cotangents w.r.t. ell itself, not its log).  Kfwd: the forward values
    (jitter 0) when the caller kept them -- the kernel then skips their recomputation; same result."""
    r2 = (x.view(1, -1, 1) - z.view(1, 1, -1)) ** 2
    a = ellx.unsqueeze(2); b = ellz.unsqueeze(1)
    den = a * a + b * b
    K = torch.sqrt(2 * (a * b) / den) * torch.exp(-r2 / den)
    G = Kbar * K
    da = G * (0.5 / a - a / den + 2 * a * r2 / den ** 2)
    db = G * (0.5 / b - b / den + 2 * b * r2 / den ** 2)
    ellxbar.copy_(da.sum(2))
    ellzbar += db.sum(1)


# ---- row solves through the Cholesky factor of K22 + eps I ---------------------------------
def solve_rows_fwd(K, R):
    """P = K A^-1 (A = R R^T), c = rowsum(P o K)."""
    Y = torch.linalg.solve_triangular(R, K.transpose(1, 2), upper=False)
    P = torch.linalg.solve_triangular(R.transpose(1, 2), Y, upper=True).transpose(1, 2).contiguous()
    return P, (P * K).sum(-1)


def solve_rows_bwd(Pbar, cbar, K, P, R, Abar):
    g = Pbar + cbar.unsqueeze(-1) * K
    Y = torch.linalg.solve_triangular(R, g.transpose(1, 2), upper=False)
    t = torch.linalg.solve_triangular(R.transpose(1, 2), Y, upper=True).transpose(1, 2)
    Abar -= t.transpose(1, 2) @ P
    return (t + cbar.unsqueeze(-1) * P).contiguous()


# ---- quadratic forms through the variational covariances ----------------------------------
def _pair_index(I, j, mode, D):
    """Matrix slot used by row n for latent/coefficient j: MODE_W -> j; MODE_U -> the
    packed pair (I[n], j): diagonal pairs occupy slots 0..D-1, strictly-lower pairs
    (i, j<i) slot D + i(i-1)/2 + j (see dsvi_step.packed_pair_index)."""
    Il = I.long()
    if mode == MODE_W:
        return torch.full_like(Il, j)
    return torch.where(Il == j, Il, D + (Il * (Il - 1)) // 2 + j)


def quadform_fwd(Pa, Pb, I, Sig, Mu, D, mode, seg=None, rec=None):
    ns, B, Q = Pa.shape
    q = torch.zeros(ns, B, D, dtype=F64); m = torch.zeros(ns, B, D, dtype=F64)
    Il = I.long()
    for j in range(D):
        rows = torch.nonzero(Il >= j).view(-1)
        if rows.numel() == 0:
            continue
        idx = _pair_index(I[rows], j, mode, D)
        p = Pa[:, rows, :]
        if mode == MODE_U:
            diag = (Il[rows] == j).view(1, -1, 1)
            p = torch.where(diag, Pb[:, rows, :], p)
        S = Sig[idx]                                           # [r,Q,Q]
        V = torch.einsum("sra,rab->srb", p, S)
        q[:, rows, j] = (V * p).sum(-1)
        m[:, rows, j] = (p * Mu[idx].unsqueeze(0)).sum(-1)
    return q, m


def quadform_bwd(Pa, Pb, I, Sig, Mu, qbar, mbar, mode, seg=None, rec=None):
    ns, B, Q = Pa.shape
    D = qbar.shape[-1]
    Pabar = torch.zeros_like(Pa)
    Pbbar = torch.zeros_like(Pa) if mode == MODE_U else None
    Il = I.long()
    for j in range(D):
        rows = torch.nonzero(Il >= j).view(-1)
        if rows.numel() == 0:
            continue
        idx = _pair_index(I[rows], j, mode, D)
        p = Pa[:, rows, :]
        diag = (Il[rows] == j).view(1, -1, 1)
        if mode == MODE_U:
            p = torch.where(diag, Pb[:, rows, :], p)
        S = Sig[idx]
        V = torch.einsum("sra,rab->srb", p, S + S.transpose(-1, -2))
        g = qbar[:, rows, j].unsqueeze(-1) * V + mbar[:, rows, j].unsqueeze(-1) * Mu[idx].unsqueeze(0)
        if mode == MODE_U:
            Pbbar[:, rows, :] += torch.where(diag, g, torch.zeros_like(g))
            Pabar[:, rows, :] += torch.where(diag, torch.zeros_like(g), g)
        else:
            Pabar[:, rows, :] += g
    return Pabar, Pbbar


def weighted_gram(Pa, Pb, I, qbar, mbar, mode, SigBar, MuBar, seg=None):
    ns, B, Q = Pa.shape
    D = qbar.shape[-1]
    Il = I.long()
    for j in range(D):
        rows = torch.nonzero(Il >= j).view(-1)
        if rows.numel() == 0:
            continue
        idx = _pair_index(I[rows], j, mode, D)
        p = Pa[:, rows, :]
        if mode == MODE_U:
            p = torch.where((Il[rows] == j).view(1, -1, 1), Pb[:, rows, :], p)
        outer = torch.einsum("sr,sra,srb->rab", qbar[:, rows, j], p, p)
        SigBar.index_add_(0, idx, outer)
        MuBar.index_add_(0, idx, torch.einsum("sr,sra->ra", mbar[:, rows, j], p))


# ---- per-sample small stage -----------------------------------------------------------------
def sample_v_fwd(mu_v, Cv, zv):
    v = mu_v.unsqueeze(0) + zv @ Cv.t()
    return v, torch.exp(v)


def sample_v_bwd(ellzbar, vbar, ellz, zv, mu_v_bar, Cvbar):
    vb = vbar + ellzbar * ellz
    mu_v_bar += vb.sum(0)
    Cvbar += torch.tril(vb.t() @ zv)


# ---- row-wise element kernels ----------------------------------------------------------------
def ell_sd_fwd(c_ell, hyp):
    return torch.sqrt(hyp[H_S2_ELL] - c_ell + EPS)


def ell_sd_bwd(sdbar, sd, hyp, ghyp):
    vb = sdbar / (2.0 * sd)
    ghyp[H_S2_ELL] += vb.sum() * hyp[H_S2_ELL]
    return -vb


def ell_rows_fwd(Pell, v, zell, sdell):
    return torch.exp(v @ Pell.t() + zell * sdell.unsqueeze(0))


def ell_rows_bwd(ellxbar, ellx, Pell, v, zell, vbar, Pellbar, sdbar):
    tb = ellxbar * ellx                          # cotangent of tilde-ell  [ns,B]
    vbar += tb @ Pell
    Pellbar += tb.t() @ v
    sdbar += (tb * zell).sum(0)


def coef_sd_fwd(q, cL0, cL1, I, hyp):
    B, D = q.shape
    j = torch.arange(D).view(1, -1); Il = I.long().view(-1, 1)
    var = torch.where(j == Il, hyp[H_S2_L1] - cL1.view(-1, 1), hyp[H_S2_L0] - cL0.view(-1, 1)) + q
    return torch.where(j <= Il, torch.sqrt(var + EPS), torch.zeros_like(var))


def coef_sd_bwd(sdbar, sd, I, hyp, ghyp):
    B, D = sd.shape
    j = torch.arange(D).view(1, -1); Il = I.long().view(-1, 1)
    live = j <= Il
    qbar = torch.where(live, sdbar / (2.0 * torch.where(live, sd, torch.ones_like(sd))), torch.zeros_like(sd))
    dg = torch.where(j == Il, qbar, torch.zeros_like(qbar)).sum(1)
    off = qbar.sum(1) - dg
    ghyp[H_S2_L1] += dg.sum() * hyp[H_S2_L1]
    ghyp[H_S2_L0] += off.sum() * hyp[H_S2_L0]
    return qbar, -off, -dg


def _philox4x32_10(c, k):
    """c: uint32 array [...,4], k: uint32 array [...,2] (numpy, vectorised)."""
    import numpy as np
    c = [c[..., i].astype(np.uint64) for i in range(4)]
    k0 = k[..., 0].astype(np.uint64); k1 = k[..., 1].astype(np.uint64)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]; p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + W0) & MASK; k1 = (k1 + W1) & MASK
    return np.stack(c, -1).astype(np.uint32)


def noise_fill(ns, B, C, seed, stream_id, s0, gid, device=None, step_dev=None):
    """Counter-based N(0,1) noise of csrc/philox.cuh restated in numpy (float32 Box-Muller, widened to double)."""
    import numpy as np
    if step_dev is not None:                      # device step counter of the CUDA-graph path: stream |= step << 8
        stream_id = int(stream_id) | (int(step_dev) << 8)
    nq = (C + 3) // 4
    g = np.arange(B, dtype=np.uint64) if gid is None else np.asarray(gid.cpu() if torch.is_tensor(gid) else gid).astype(np.uint64)
    ctr = np.zeros((ns, B, nq, 4), dtype=np.uint32)
    ctr[..., 0] = (g & np.uint64(0xFFFFFFFF)).astype(np.uint32)[None, :, None]
    ctr[..., 1] = (g >> np.uint64(32)).astype(np.uint32)[None, :, None]
    ctr[..., 2] = (np.arange(ns, dtype=np.uint32) + np.uint32(s0))[:, None, None]
    ctr[..., 3] = np.arange(nq, dtype=np.uint32)[None, None, :]
    key = np.zeros((ns, B, nq, 2), dtype=np.uint32)
    key[..., 0] = np.uint32((seed ^ stream_id) & 0xFFFFFFFF)
    key[..., 1] = np.uint32(((seed >> 32) ^ (stream_id >> 32)) & 0xFFFFFFFF)
    u = _philox4x32_10(ctr, key)
    uf = ((u >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    z = np.empty((ns, B, nq, 4), dtype=np.float32)
    for h in range(2):
        r = np.sqrt(np.float32(-2.0) * np.log(uf[..., 2 * h]))
        ang = np.float32(6.28318530717958647692) * uf[..., 2 * h + 1]
        z[..., 2 * h] = r * np.cos(ang)
        z[..., 2 * h + 1] = r * np.sin(ang)
    return torch.from_numpy(z.reshape(ns, B, nq * 4)[..., :C].astype(np.float64)).contiguous()


def _explicit_noise(zL, noise, B, D):
    if zL is not None:
        return zL
    seed, stream_id, s0, ns, gid = noise[:5]
    return noise_fill(ns, B, D, seed, stream_id, s0, gid, step_dev=noise[5] if len(noise) > 5 else None)


def coef_sample_fwd(m, sd, zL, I, noise=None):
    D = m.shape[-1]
    zL = _explicit_noise(zL, noise, m.shape[0], D)
    j = torch.arange(D).view(1, 1, -1); Il = I.long().view(1, -1, 1)
    raw = m.unsqueeze(0) + zL * sd.unsqueeze(0)
    l = torch.where(j == Il, torch.exp(raw), raw)
    return torch.where(j <= Il, l, torch.zeros_like(l))


def coef_sample_bwd(lbar, l, zL, I, mbar, sdbar, noise=None):
    D = l.shape[-1]
    zL = _explicit_noise(zL, noise, l.shape[1], D)
    j = torch.arange(D).view(1, 1, -1); Il = I.long().view(1, -1, 1)
    rb = torch.where(j == Il, lbar * l, lbar)
    rb = torch.where(j <= Il, rb, torch.zeros_like(rb))
    mbar += rb.sum(0)
    sdbar += (rb * zL).sum(0)


def lik_rows(l, mg, qg, cG, y, I, hyp, scale, Rsum, ghyp):
    """Expected log-likelihood of a sample chunk and its cotangents (loss = -scale * sum_s R_s + ...)."""
    ns, B, D = l.shape
    j = torch.arange(D).view(1, 1, -1); Il = I.long().view(1, -1, 1)
    live = (j <= Il).to(F64)
    s2e = hyp[H_S2_ERR]
    s2g = (1.0 - cG.unsqueeze(-1) + qg) * live
    F = (l * mg * live).sum(-1)
    r = (y.view(1, -1) if y.dim() == 1 else y) - F          # y [B] shared, or [ns, B] one target vector per sample
    pen = (l * l * s2g).sum(-1)
    import math
    Rsum.copy_((-(r * r) / (2 * s2e) - 0.5 * torch.log(s2e) - math.log(math.sqrt(2 * math.pi))).sum(1) - 0.5 / s2e * pen.sum(1))
    ghyp[H_S2_ERR] += -scale * ((r * r / (2 * s2e) - 0.5).sum() + 0.5 / s2e * pen.sum())
    rr = (r / s2e).unsqueeze(-1)
    mgbar = -scale * rr * l * live
    qgbar = scale * (0.5 / s2e) * l * l * live
    cGbar = -qgbar.sum(-1)
    lbar = -scale * (rr * mg - (1.0 / s2e) * l * s2g) * live
    return lbar, mgbar, qgbar, cGbar


def pair_means(Pa, Pb, I, Mu, D, mode):
    ns, B, Q = Pa.shape
    m = torch.zeros(ns, B, D, dtype=F64)
    Il = I.long()
    for j in range(D):
        rows = torch.nonzero(Il >= j).view(-1)
        if rows.numel() == 0:
            continue
        idx = _pair_index(I[rows], j, mode, D)
        p = Pa[:, rows, :]
        if mode == MODE_U:
            p = torch.where((Il[rows] == j).view(1, -1, 1), Pb[:, rows, :], p)
        m[:, rows, j] = (p * Mu[idx].unsqueeze(0)).sum(-1)
    return m


def rowdot_live(l, g, I):
    D = l.shape[-1]
    live = (torch.arange(D).view(1, 1, -1) <= I.long().view(1, -1, 1)).to(F64)
    return (l * g * live).sum(-1)


def latent_fused(PG, cG, l, y, I, SigW, muW, hyp, scale, Rsum, ghyp, seg=None, rec=None):
    """= quadform_fwd(MODE_W) -> lik_rows -> quadform_bwd(MODE_W)."""
    D = l.shape[-1]
    qg, mg = quadform_fwd(PG, PG, I, SigW, muW, D, MODE_W)
    lbar, mgbar, qgbar, cGbar = lik_rows(l, mg, qg, cG, y, I, hyp, scale, Rsum, ghyp)
    PGbar, _ = quadform_bwd(PG, PG, I, SigW, muW, qgbar, mgbar, MODE_W)
    return lbar, mgbar, qgbar, cGbar, PGbar


# ---- SIM_code line ---------------------------------------------------------------------------------------
def nonstationary_cov(X1, sigma1, ell1, X2, sigma2, ell2, jitter, self_cov=False):
    n1, n2 = X1.shape[0], X2.shape[0]
    s1 = torch.ones(n1, dtype=F64) if sigma1 is None else sigma1
    s2 = torch.ones(n2, dtype=F64) if sigma2 is None else sigma2
    l1 = torch.ones(n1, dtype=F64) if ell1 is None else ell1
    l2 = torch.ones(n2, dtype=F64) if ell2 is None else ell2
    d = pairwise_dist(X1, X2)
    A = (l1 ** 2).view(-1, 1) + (l2 ** 2).view(1, -1)
    K = (s1.view(-1, 1) * s2.view(1, -1)) * torch.sqrt(2.0 * (l1.view(-1, 1) * l2.view(1, -1)) / A) * torch.exp(-d / A)
    return K + jitter * torch.eye(n1, n2, dtype=F64)


def nonstationary_cov_bwd(X1, sigma1, ell1, X2, sigma2, ell2, Kbar, want=(True, True, True, True)):
    """Autograd of the specification above (test infrastructure)."""
    args = [sigma1, ell1, sigma2, ell2]
    with torch.enable_grad():                      # may be called from inside a custom Function's backward
        leaves = [None if a is None else a.detach().clone().requires_grad_(True) for a in args]
        K = nonstationary_cov(X1, leaves[0], leaves[1], X2, leaves[2], leaves[3], 0.0)
        outs = []
        for a, w in zip(leaves, want):
            outs.append(torch.autograd.grad((K * Kbar).sum(), a, retain_graph=True)[0] if (w and a is not None) else None)
    return tuple(outs)


def hadamard_index_cov(Kx, Bf, indx1, indx2, diag=0.0):
    """out[i,j] = Kx[i,j] Bf[indx1[i], indx2[j]] (+ diag on i == j)."""
    Ki = Bf[indx1.long().view(-1, 1), indx2.long().view(1, -1)]
    return Kx * Ki + diag * torch.eye(Kx.shape[0], Kx.shape[1], dtype=F64)


def dense_loglik_bwd(Sinv, alpha, A, Bt, indx1, indx2, g):
    """Cotangents of -1/2 logdet S - 1/2 y^T S^-1 y, S = A o Bt[indx1, indx2] + sigma2 I (G = 1/2 (alpha alpha^T - S^-1))."""
    G = 0.5 * g.reshape(()) * (torch.outer(alpha, alpha) - Sinv)
    i1, i2 = indx1.long().view(-1, 1), indx2.long().view(1, -1)
    Abar = G * Bt[i1, i2]
    Btbar = torch.zeros_like(Bt)
    Btbar.index_put_((i1.expand_as(G), i2.expand_as(G)), G * A, accumulate=True)
    return Abar, Btbar, torch.diagonal(G).sum().reshape(1)


def sim_rbf_cov(X1, X2, alpha, beta, jitter, self_cov=False):
    d = pairwise_dist(X1 / beta, X2 / beta)
    return torch.exp(-0.5 * d) * alpha ** 2 + jitter * torch.eye(X1.shape[0], X2.shape[0], dtype=F64)


def pairwise_dist(X1, X2):
    return (X1 ** 2).sum(1).view(-1, 1) + (X2 ** 2).sum(1).view(1, -1) - 2.0 * X1 @ X2.t()


def gemm_nt(A, Bm, alpha=1.0, beta=0.0, C=None):
    out = alpha * (A @ Bm.t())
    if C is not None:
        C.copy_(out + beta * C)
        return C
    return out


def gemm_concurrent_mode(on):
    """No-op on the CPU specification."""
    return None


def potrf_big(A, info=None, slot=0, panel=0):
    L, bad = torch.linalg.cholesky_ex(A)
    if int(bad) != 0:
        if info is None:
            raise RuntimeError("cholesky: the leading minor of order %d is not positive-definite" % int(bad))
        info.fill_(int(bad))
    A.copy_(L)
    return A, L.diagonal().log().sum().reshape(1)


def build_augmented(K, r, alpha_dev, sigma2_dev, rnorm2_dev, out):
    T = K.shape[0]
    out[:T, :T] = alpha_dev.reshape(()) * K + sigma2_dev.reshape(()) * torch.eye(T, dtype=F64)
    out[T, :T] = r
    out[:T, T] = r
    out[T, T] = 1.0 + rnorm2_dev.reshape(()) / sigma2_dev.reshape(())
    return out


def augmented_results(A, T, hld_aug, hld_out, quad_out):
    quad_out.copy_((A[T, :T] ** 2).sum().reshape(quad_out.shape))
    hld_out.copy_((hld_aug.reshape(()) - torch.log(A[T, T])).reshape(hld_out.shape))


def tri_inv_block(L, out, scale=1.0):
    out.copy_(scale * torch.linalg.solve_triangular(torch.tril(L), torch.eye(L.shape[0], dtype=F64), upper=False))
    return out


def scale_add_diag_dev(K, alpha_dev, sigma2_dev, out=None):
    A = alpha_dev.reshape(()) * K + sigma2_dev.reshape(()) * torch.eye(K.shape[0], dtype=F64)
    if out is not None:
        out.copy_(A)
        return out
    return A


def potrs_vec(L, b):
    return torch.cholesky_solve(b.view(-1, 1), L).view(-1)


def scale_add_diag(K, alpha, sigma2):
    return alpha * K + sigma2 * torch.eye(K.shape[0], dtype=F64)


def kron_product(t1, t2):
    return torch.kron(t1, t2)


def eigh_small(A):
    return torch.linalg.eigh(A, UPLO="U")


def axpby_dev(x, y, a_dev, a_scale=1.0, b=1.0, out=None):
    res = (a_scale * a_dev.reshape(())) * x + b * y
    if out is not None:
        out.copy_(res)
        return out
    return res


def axpby(x, y, a, b):
    return a * x + b * y


def dot(x, y):
    return (x * y).sum().reshape(1)


def reparam_diag(mean, var, z):
    return mean + z * torch.sqrt(var + EPS)


def normal_logprob_sum(loc, scale, y):
    import math
    var = scale.reshape(()) ** 2
    return (-((y - loc) ** 2) / (2 * var) - torch.log(scale.reshape(())) - math.log(math.sqrt(2 * math.pi))).sum().reshape(1)


def sumsq_rows(x):
    return (x * x).sum(-1)


def lcorr(L):
    cov = L @ L.transpose(-1, -2)
    inv = torch.sqrt(torch.diag_embed(1.0 / torch.diagonal(cov, dim1=-2, dim2=-1)))
    return inv @ cov @ inv


def launch_count():
    """CPU specifications launch no kernels."""
    return 0


def lq_pad_records(Sig):
    """The CPU specifications read the covariances directly; the padded records are a device-side staging format."""
    return None
