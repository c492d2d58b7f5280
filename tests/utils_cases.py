"""Shared bodies: drop-in utils.* primitives (forward + autograd) against the oracle's torch-CPU restatement."""
import numpy as np
import torch

from oracle import nmgp_oracle as orc
from collaborative_nonstationary_multivariate_gaussian_process_b200 import utils

RTOL = 1e-9


def _rel(a, b):
    a = a.detach().cpu().double().reshape(-1); b = b.detach().cpu().double().reshape(-1)
    return float(torch.linalg.norm(a - b) / max(float(torch.linalg.norm(b)), 1e-300))


def _leaf(t, dev):
    return t.clone().to(dev).requires_grad_(True)


def run_all(dev):
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    N, M, D = 57, 12, 3
    X = torch.sort(torch.rand(N, generator=g, dtype=torch.float64))[0].view(-1, 1)
    Z = torch.linspace(0, 1, M, dtype=torch.float64).view(-1, 1)
    # ---- create_RBF / create_Gibbs with gradients --------------------------------------------------------
    s2c, lnc = torch.tensor(1.3, dtype=torch.float64, requires_grad=True), torch.tensor(0.4, dtype=torch.float64, requires_grad=True)
    Kc = orc.stationary_rbf(X, Z, s2c, lnc); W = rn(N, M); (Kc * W).sum().backward()
    s2d, lnd = _leaf(s2c.detach(), dev), _leaf(lnc.detach(), dev)
    Kd = utils.create_RBF(X.to(dev), Z.to(dev), scale2=s2d, length_scales=lnd); (Kd * W.to(dev)).sum().backward()
    assert _rel(Kd, Kc) < 1e-13 and _rel(s2d.grad, s2c.grad) < RTOL and _rel(lnd.grad, lnc.grad) < RTOL
    assert _rel(utils.create_RBF(Z.to(dev), scale2=0.7, length_scales=0.3), orc.stationary_rbf(Z, None, 0.7, 0.3)) < 1e-13
    exc = torch.exp(0.3 * rn(N)).requires_grad_(True); ezc = torch.exp(0.3 * rn(M)).requires_grad_(True)
    Gc = orc.gibbs_kernel(X, Z, exc, ezc); (Gc * W).sum().backward()
    exd, ezd = _leaf(exc.detach(), dev), _leaf(ezc.detach(), dev)
    Gd = utils.create_Gibbs(X.to(dev), Z.to(dev), exd, ezd); (Gd * W.to(dev)).sum().backward()
    assert _rel(Gd, Gc) < 1e-13 and _rel(exd.grad, exc.grad) < RTOL and _rel(ezd.grad, ezc.grad) < RTOL
    # ---- MGP_mu_sigma2 / MGP_mu with gradients w.r.t. everything ----------------------------------------------
    K12c = Gc.detach().clone().requires_grad_(True)
    K22c = orc.gibbs_kernel(Z, Z, ezc.detach(), ezc.detach()).clone().requires_grad_(True)
    muc = rn(D, M).requires_grad_(True)
    Lc = torch.tril(0.3 * rn(D, M, M)); Sigc = (Lc @ Lc.transpose(-1, -2)).requires_grad_(True)
    d11 = torch.ones(N, dtype=torch.float64)
    mc, sc = orc.marginal_stats(K12c, K22c, d11, muc, Sigc)
    W1, W2 = rn(D, N), rn(D, N)
    ((mc * W1).sum() + (sc * W2).sum()).backward()
    K12d, K22d, mud, Sigd = (_leaf(t.detach(), dev) for t in (K12c, K22c, muc, Sigc))
    md, sd = utils.MGP_mu_sigma2(K12d, K22d, d11.to(dev), mud, Sigd)
    ((md * W1.to(dev)).sum() + (sd * W2.to(dev)).sum()).backward()
    assert _rel(md, mc) < RTOL and _rel(sd, sc) < RTOL
    for a, b, n in ((K12d, K12c, "K12"), (K22d, K22c, "K22"), (mud, muc, "mu"), (Sigd, Sigc, "Sigma")):
        gb = b.grad if n != "K22" else 0.5 * (b.grad + b.grad.t())      # ours is the symmetric-matrix gradient
        ga = a.grad if n != "K22" else 0.5 * (a.grad + a.grad.t())
        assert _rel(ga, gb) < 1e-8, (n, _rel(ga, gb))
    assert _rel(utils.MGP_mu(K12d.detach(), K22d.detach(), mud.detach()), orc.marginal_mean(K12c.detach(), K22c.detach(), muc.detach())) < RTOL
    assert _rel(utils.MGP_mu(K12d.detach(), K22d.detach(), mud.detach()[1]), orc.marginal_mean(K12c.detach(), K22c.detach(), muc.detach()[1])) < RTOL
    # ---- samplers: same CPU random stream as the reference ---------------------------------------------------
    torch.manual_seed(5)
    sc_ = orc.marginal_sample(K12c.detach(), K22c.detach(), d11, muc.detach()[0], Sigc.detach()[0])
    torch.manual_seed(5)
    sd_ = utils.MGP_d(K12d.detach(), K22d.detach(), d11.to(dev), mud.detach()[0], Sigd.detach()[0])
    assert _rel(sd_, sc_) < RTOL
    torch.manual_seed(6)
    jc = orc.joint_sample(d11, K12c.detach(), K22c.detach(), muc.detach()[0], Sigc.detach()[0])
    torch.manual_seed(6)
    jd = utils.JGP_S(d11.to(dev), K12d.detach(), K22d.detach(), mud.detach()[0], Sigd.detach()[0])
    assert jd.shape == (N + M,) and _rel(jd, jc) < RTOL
    # ---- reparameterize -------------------------------------------------------------------------------------
    mean, var, z = rn(N), torch.rand(N, generator=g, dtype=torch.float64), rn(N)
    assert _rel(utils.reparameterize(mean.to(dev), var.to(dev), z.to(dev)), orc.reparam(mean, var, z)) < 1e-14
    vc = Sigc.detach()[0].clone().requires_grad_(True); mc_ = rn(M).requires_grad_(True); zz = rn(M)
    fc = orc.reparam(mc_, vc, zz, full_cov=True); (fc * rn(M).fill_(1.0)).sum().backward()
    vd, md_ = _leaf(vc.detach(), dev), _leaf(mc_.detach(), dev)
    fd = utils.reparameterize(md_, vd, zz.to(dev), full_cov=True); fd.sum().backward()
    assert _rel(fd, fc) < RTOL and _rel(md_.grad, mc_.grad) < RTOL
    assert _rel(0.5 * (vd.grad + vd.grad.t()), 0.5 * (vc.grad + vc.grad.t())) < 1e-8
    assert _rel(utils.reparameterize(muc.detach().to(dev), Sigc.detach().to(dev), rn(D, M).fill_(0.3).to(dev), full_cov=True),
                orc.reparam(muc.detach(), Sigc.detach(), torch.full((D, M), 0.3, dtype=torch.float64), full_cov=True)) < RTOL
    assert torch.equal(utils.mat2ltri(Sigc.detach().to(dev)).cpu(), orc.lower_part(Sigc.detach()))
    # ---- KL_Gaussian (reference-exact form) with gradients ------------------------------------------------------
    Xmu_c = rn(D, M).requires_grad_(True); XS_c = Sigc.detach().clone().requires_grad_(True)
    X2S_c = K22c.detach().clone().requires_grad_(True)
    klc = orc.kl_gaussian(Xmu_c, XS_c, torch.zeros(M, dtype=torch.float64), X2S_c); wv = rn(D); (klc * wv).sum().backward()
    Xmu_d, XS_d, X2S_d = _leaf(Xmu_c.detach(), dev), _leaf(XS_c.detach(), dev), _leaf(X2S_c.detach(), dev)
    kld = utils.KL_Gaussian(Xmu_d, XS_d, torch.zeros(M, dtype=torch.float64, device=dev), X2S_d); (kld * wv.to(dev)).sum().backward()
    assert _rel(kld, klc) < RTOL and _rel(Xmu_d.grad, Xmu_c.grad) < 1e-8
    sym = lambda t: 0.5 * (t + t.transpose(-1, -2))
    assert _rel(sym(XS_d.grad), sym(XS_c.grad)) < 1e-8 and _rel(sym(X2S_d.grad), sym(X2S_c.grad)) < 1e-8
    klv = utils.KL_Gaussian(Xmu_d.detach()[0], XS_d.detach()[0], torch.zeros(M, dtype=torch.float64, device=dev), X2S_d.detach())
    assert klv.dim() == 0 and abs(float(klv) - float(klc[0])) < RTOL * abs(float(klc[0]))
    # ---- small utilities ---------------------------------------------------------------------------------------
    loc, y = rn(N, 1), rn(N, 1); scl = torch.tensor(0.37, dtype=torch.float64)
    assert abs(float(utils.Normal_logprob(loc.to(dev), scl.to(dev), y.to(dev))) - float(orc.gauss_logprob_sum(loc, scl, y))) < 1e-11 * N
    Kpd = K22c.detach() + 1e-2 * torch.eye(M, dtype=torch.float64)
    assert _rel(utils.log_determinant_halfpower(torch.stack([Kpd, 2 * Kpd]).to(dev)), orc.half_logdet(torch.stack([Kpd, 2 * Kpd]))) < 1e-11
    bm = rn(4, 3, M, M)
    assert _rel(utils.batch_trace_XXT(bm.to(dev)), orc.frob2(bm)) < 1e-13
    Lf = torch.linalg.cholesky(Kpd)
    assert _rel(utils.batch_mahalanobis(Lf.to(dev), muc.detach().to(dev)), orc.mahalanobis(Lf, muc.detach())) < RTOL
