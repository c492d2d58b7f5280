"""Each CUDA kernel (through the C ABI wrappers in _ops) against its CPU specification
(oracle/kernel_specs.py) on seeded inputs.  Tolerance: FP64, 1e-11 relative (norm-wise) unless noted --
different summation order only."""
import numpy as np
import pytest
import torch

from oracle import kernel_specs as specs

pytestmark = pytest.mark.gpu

from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops  # noqa: E402

DEV = "cuda:0"
TOL = 1e-11


def g(t):
    if t is None:
        return None
    return t.to(DEV).contiguous()


def relerr(a, b):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    den = max(float(torch.linalg.norm(b)), 1e-300)
    return float(torch.linalg.norm(a - b)) / den


def check(a, b, tol=TOL, what=""):
    e = relerr(a, b)
    assert e <= tol, "%s rel err %.3e" % (what, e)


def spd(nb, Q, gen, scale=1.0):
    A = torch.randn(nb, Q, Q, generator=gen, dtype=torch.float64)
    return scale * (A @ A.transpose(-1, -2) / Q + 0.5 * torch.eye(Q, dtype=torch.float64))


def make_I(gen, B, D, empty=()):
    probs = torch.ones(D)
    for e in empty:
        probs[e] = 0
    I = torch.multinomial(probs, B, replacement=True, generator=gen)
    return torch.sort(I)[0].to(torch.int32)


CASES = [(20, 3, 37), (50, 8, 300), (100, 5, 131), (112, 2, 70)]
# 64 < Q <= 128: ring-pipelined DMMA quadratic forms, right-looking row solves, wide Gram / A^T B kernels
LQ_CASES = [(72, 6, 300), (100, 9, 700), (104, 3, 129), (128, 3, 200), (65, 12, 1030)]


@pytest.mark.parametrize("Q,nb", [(20, 5), (50, 70), (100, 3), (112, 2)])
def test_tril_syrk_potrf_and_adjoints(Q, nb):
    gen = torch.Generator().manual_seed(Q * 7 + nb)
    S = 0.3 * torch.randn(nb, Q, Q, generator=gen, dtype=torch.float64)
    Sig = ops.tril_syrk_fwd(g(S))
    check(Sig, specs.tril_syrk_fwd(S), what="tril_syrk_fwd")
    assert torch.equal(Sig, Sig.transpose(-1, -2)), "Sigma must be exactly symmetric"
    Gb = torch.randn(nb, Q, Q, generator=gen, dtype=torch.float64)
    check(ops.tril_syrk_bwd(g(S), g(Gb)), specs.tril_syrk_bwd(S, Gb), what="tril_syrk_bwd")
    A = specs.tril_syrk_fwd(S)
    C, hld = ops.potrf(g(A), 1e-4)
    Cr, hr = specs.potrf(A, 1e-4)
    check(C, Cr, 1e-10, "potrf C"); check(hld, hr, 1e-12, "potrf hld")
    assert float(torch.triu(C, 1).abs().max()) == 0.0
    Cb = torch.randn(nb, Q, Q, generator=gen, dtype=torch.float64)
    hb = torch.randn(nb, generator=gen, dtype=torch.float64)
    got = ops.potrf_bwd(g(Cr), g(Cb), g(hb))
    ref = specs.potrf_bwd(Cr, Cb, hb)
    check(got, ref, 1e-9, "potrf_bwd")      # conditioned by A: compare loosely, step-level test is the real bar


def test_potrf_raises_on_non_pd():
    A = torch.eye(8, dtype=torch.float64).repeat(3, 1, 1)
    A[1, 4, 4] = -1.0
    with pytest.raises(RuntimeError):
        ops.potrf(g(A), 0.0)


@pytest.mark.parametrize("Q,np_,nb", [(20, 1, 1), (50, 4, 70), (50, 1, 300), (100, 3, 5), (100, 1, 140), (128, 2, 9),
                                      (50, 32, 64)])
def test_kl(Q, np_, nb):
    gen = torch.Generator().manual_seed(Q + np_ + nb)
    CS, hS = specs.potrf(spd(nb, Q, gen, 0.05))
    R, hR = specs.potrf(spd(np_, Q, gen, 2.0))
    mu = torch.randn(nb, Q, generator=gen, dtype=torch.float64)
    kl, t = ops.kl_fwd(g(CS), g(hS), g(mu), g(R), g(hR))
    klr, tr = specs.kl_fwd(CS, hS, mu, R, hR)
    check(kl, klr, 1e-11, "kl")
    kb = torch.randn(np_, nb, generator=gen, dtype=torch.float64)
    got = ops.kl_bwd(g(kb), g(CS), g(mu), g(R), t)          # `t` is opaque: what this implementation saved in kl_fwd
    ref = specs.kl_bwd(kb, CS, mu, R, tr)
    for a, b, n in zip(got, ref, ("CSbar", "hldSbar", "mubar", "Rbar", "hldRbar")):
        check(a, b, 1e-10, "kl_bwd " + n)
    # the mathematically exact variant (explicit flag; not what the reference computes: quirk q10)
    kle, te = ops.kl_fwd(g(CS), g(hS), g(mu), g(R), g(hR), exact=True)
    kler, ter = specs.kl_fwd(CS, hS, mu, R, hR, exact=True)
    check(kle, kler, 1e-11, "kl exact")
    gote = ops.kl_bwd(g(kb), g(CS), g(mu), g(R), te, exact=True)
    refe = specs.kl_bwd(kb, CS, mu, R, ter, exact=True)
    for a, b, n in zip(gote, refe, ("CSbar", "hldSbar", "mubar", "Rbar", "hldRbar")):
        check(a, b, 1e-10, "kl_bwd exact " + n)


@pytest.mark.parametrize("Q,D,B", CASES)
def test_builds(Q, D, B):
    gen = torch.Generator().manual_seed(B)
    x = torch.rand(B, generator=gen, dtype=torch.float64) * 10
    z = torch.linspace(0, 10, Q, dtype=torch.float64)
    hyp = torch.tensor([1.3, 2.5, 0.7, 3.0, 1.1, 0.9, 0.05], dtype=torch.float64)
    K = ops.rbf_build_fwd(g(x), g(z), g(hyp), 2, 3, 0.0)
    check(K, specs.rbf_build_fwd(x, z, hyp, 2, 3, 0.0), 1e-14, "rbf fwd")
    Kzz = ops.rbf_build_fwd(g(z), g(z), g(hyp), 0, 1, 1e-4)
    check(Kzz, specs.rbf_build_fwd(z, z, hyp, 0, 1, 1e-4), 1e-14, "rbf zz")
    Kb = torch.randn(B, Q, generator=gen, dtype=torch.float64)
    gh = torch.zeros(7, dtype=torch.float64); ghd = g(gh.clone())
    specs.rbf_build_bwd(x, z, hyp, 2, 3, Kb, gh)
    ops.rbf_build_bwd(g(x), g(z), g(hyp), 2, 3, g(Kb), ghd)
    check(ghd, gh, 1e-11, "rbf bwd")
    ns = 3
    ellx = torch.exp(0.5 * torch.randn(ns, B, generator=gen, dtype=torch.float64))
    ellz = torch.exp(0.5 * torch.randn(ns, Q, generator=gen, dtype=torch.float64))
    KG = ops.gibbs_build_fwd(g(x), g(z), g(ellx), g(ellz), 0.0)
    check(KG, specs.gibbs_build_fwd(x, z, ellx, ellz, 0.0), 1e-14, "gibbs fwd")
    KGzz = ops.gibbs_build_fwd(g(z), g(z), g(ellz), g(ellz), 1e-4)
    check(KGzz, specs.gibbs_build_fwd(z, z, ellz, ellz, 1e-4), 1e-14, "gibbs zz")
    Kb = torch.randn(ns, B, Q, generator=gen, dtype=torch.float64)
    exb = torch.zeros(ns, B, dtype=torch.float64); ezb = torch.randn(ns, Q, generator=gen, dtype=torch.float64)
    ezb0 = ezb.clone()
    exd = g(torch.full_like(exb, 7.0)); ezd = g(ezb.clone())
    specs.gibbs_build_bwd(x, z, ellx, ellz, Kb, exb, ezb)
    ops.gibbs_build_bwd(g(x), g(z), g(ellx), g(ellz), g(Kb), exd, ezd)
    check(exd, exb, 1e-12, "gibbs bwd x"); check(ezd, ezb, 1e-11, "gibbs bwd z")
    exd2 = g(torch.full_like(exb, -3.0)); ezd2 = g(ezb0.clone())          # same with the forward values handed in
    ops.gibbs_build_bwd(g(x), g(z), g(ellx), g(ellz), g(Kb), exd2, ezd2, Kfwd=KG)
    check(exd2, exb, 1e-12, "gibbs bwd x (kept K)"); check(ezd2, ezb, 1e-11, "gibbs bwd z (kept K)")


@pytest.mark.parametrize("Q,D,B", CASES + LQ_CASES)
def test_solve_rows(Q, D, B):
    gen = torch.Generator().manual_seed(B + 1)
    ns = 2
    R, _ = specs.potrf(spd(ns, Q, gen))
    K = torch.randn(ns, B, Q, generator=gen, dtype=torch.float64)
    P, c = ops.solve_rows_fwd(g(K), g(R))
    Pr, cr = specs.solve_rows_fwd(K, R)
    check(P, Pr, 1e-11, "solve P"); check(c, cr, 1e-11, "solve c")
    Pb = torch.randn(ns, B, Q, generator=gen, dtype=torch.float64)
    cb = torch.randn(ns, B, generator=gen, dtype=torch.float64)
    Ab = torch.randn(ns, Q, Q, generator=gen, dtype=torch.float64); Abd = g(Ab.clone())
    Kbr = specs.solve_rows_bwd(Pb, cb, K, Pr, R, Ab)
    Kbd = ops.solve_rows_bwd(g(Pb), g(cb), g(K), g(Pr), g(R), Abd)
    check(Kbd, Kbr, 1e-11, "solve bwd K"); check(Abd, Ab, 1e-11, "solve bwd A")


@pytest.mark.parametrize("Q,D,B", CASES + [(50, 9, 64), (50, 4, 1)] + LQ_CASES + [(100, 4, 1)])
@pytest.mark.parametrize("mode", [0, 1])
def test_quadform_family(Q, D, B, mode):
    gen = torch.Generator().manual_seed(B * 3 + mode)
    ns = 2 if mode == 0 else 1
    I = make_I(gen, B, D, empty=(1,) if D > 2 else ())
    nm = D if mode == 0 else D * (D + 1) // 2
    Sig = spd(nm, Q, gen)
    Mu = torch.randn(nm, Q, generator=gen, dtype=torch.float64)
    Pa = torch.randn(ns, B, Q, generator=gen, dtype=torch.float64)
    Pb = torch.randn(ns, B, Q, generator=gen, dtype=torch.float64) if mode == 1 else Pa
    seg = ops.segment_offsets(g(I), D)
    assert torch.equal(seg.cpu(), specs.segment_offsets(I, D))
    q, m = ops.quadform_fwd(g(Pa), g(Pb), g(I), g(Sig), g(Mu), D, mode, seg=seg)
    qr, mr = specs.quadform_fwd(Pa, Pb, I, Sig, Mu, D, mode)
    check(q, qr, 1e-12, "quadform q"); check(m, mr, 1e-12, "quadform m")
    qb = torch.randn(ns, B, D, generator=gen, dtype=torch.float64)
    mb = torch.randn(ns, B, D, generator=gen, dtype=torch.float64)
    pa, pb = ops.quadform_bwd(g(Pa), g(Pb), g(I), g(Sig), g(Mu), g(qb), g(mb), mode, seg=seg)
    par, pbr = specs.quadform_bwd(Pa, Pb, I, Sig, Mu, qb, mb, mode)
    check(pa, par, 1e-12, "quadform bwd a")
    if mode == 1:
        check(pb, pbr, 1e-12, "quadform bwd b")
    SB = torch.randn(nm, Q, Q, generator=gen, dtype=torch.float64); MB = torch.randn(nm, Q, generator=gen, dtype=torch.float64)
    SBd, MBd = g(SB.clone()), g(MB.clone())
    specs.weighted_gram(Pa, Pb, I, qb, mb, mode, SB, MB)
    ops.weighted_gram(g(Pa), g(Pb), g(I), g(qb), g(mb), mode, SBd, MBd, seg=seg)
    check(SBd, SB, 1e-12, "gram Sig"); check(MBd, MB, 1e-12, "gram Mu")


@pytest.mark.parametrize("Q,D,B", CASES)
def test_row_elementwise(Q, D, B):
    gen = torch.Generator().manual_seed(B + 5)
    ns = 3
    rn = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    hyp = torch.tensor([1.3, 2.5, 0.7, 3.0, 1.1, 0.9, 0.05], dtype=torch.float64)
    I = make_I(gen, B, D)
    # sample_v
    mu_v = rn(Q); Cv = torch.tril(rn(Q, Q)) * 0.1; zv = rn(ns, Q)
    v, ez = ops.sample_v_fwd(g(mu_v), g(Cv), g(zv))
    vr, ezr = specs.sample_v_fwd(mu_v, Cv, zv)
    check(v, vr, 1e-13); check(ez, ezr, 1e-13)
    ezb = rn(ns, Q); vb = rn(ns, Q); mvb = rn(Q); Cvb = rn(Q, Q)
    mvd, Cvd = g(mvb.clone()), g(Cvb.clone())
    specs.sample_v_bwd(ezb, vb, ezr, zv, mvb, Cvb)
    ops.sample_v_bwd(g(ezb), g(vb), g(ezr), g(zv), mvd, Cvd)
    check(mvd, mvb, 1e-12); check(Cvd, Cvb, 1e-12)
    # ell sd / rows
    c = torch.rand(B, generator=gen, dtype=torch.float64)
    sd = ops.ell_sd_fwd(g(c), g(hyp)); sdr = specs.ell_sd_fwd(c, hyp); check(sd, sdr, 1e-14)
    sdb = rn(B); gh = torch.zeros(7, dtype=torch.float64); ghd = g(gh.clone())
    cbr = specs.ell_sd_bwd(sdb, sdr, hyp, gh); cbd = ops.ell_sd_bwd(g(sdb), g(sdr), g(hyp), ghd)
    check(cbd, cbr, 1e-13); check(ghd, gh, 1e-11)
    Pell = rn(B, Q) * 0.2; zell = rn(ns, B)
    ex = ops.ell_rows_fwd(g(Pell), g(vr * 0.1), g(zell), g(sdr)); exr = specs.ell_rows_fwd(Pell, vr * 0.1, zell, sdr)
    check(ex, exr, 1e-13, "ell_rows_fwd")
    exb = rn(ns, B); vbar = rn(ns, Q); Pb_ = rn(B, Q); sb = rn(B)
    vd, Pd, sdd = g(vbar.clone()), g(Pb_.clone()), g(sb.clone())
    specs.ell_rows_bwd(exb, exr, Pell, vr * 0.1, zell, vbar, Pb_, sb)
    ops.ell_rows_bwd(g(exb), g(exr), g(Pell), g(vr * 0.1), g(zell), vd, Pd, sdd)
    check(vd, vbar, 1e-11, "ell_rows_bwd v"); check(Pd, Pb_, 1e-12); check(sdd, sb, 1e-12)
    # coefficient sd / sample
    q = torch.rand(B, D, generator=gen, dtype=torch.float64); c0 = 0.5 * torch.rand(B, generator=gen, dtype=torch.float64)
    c1 = 0.5 * torch.rand(B, generator=gen, dtype=torch.float64)
    sdU = ops.coef_sd_fwd(g(q), g(c0), g(c1), g(I), g(hyp)); sdUr = specs.coef_sd_fwd(q, c0, c1, I, hyp)
    check(sdU, sdUr, 1e-14, "coef_sd_fwd")
    sdb = rn(B, D); gh = torch.zeros(7, dtype=torch.float64); ghd = g(gh.clone())
    ref = specs.coef_sd_bwd(sdb, sdUr, I, hyp, gh); got = ops.coef_sd_bwd(g(sdb), g(sdUr), g(I), g(hyp), ghd)
    for a, b in zip(got, ref):
        check(a, b, 1e-12, "coef_sd_bwd")
    check(ghd, gh, 1e-11)
    m = rn(B, D) * 0.3; zL = rn(ns, B, D)
    l = ops.coef_sample_fwd(g(m), g(sdUr), g(zL), g(I)); lr = specs.coef_sample_fwd(m, sdUr, zL, I)
    check(l, lr, 1e-14, "coef_sample_fwd")
    lb = rn(ns, B, D); mb = rn(B, D); sb2 = rn(B, D); mbd, sbd = g(mb.clone()), g(sb2.clone())
    specs.coef_sample_bwd(lb, lr, zL, I, mb, sb2); ops.coef_sample_bwd(g(lb), g(lr), g(zL), g(I), mbd, sbd)
    check(mbd, mb, 1e-12); check(sbd, sb2, 1e-12)
    # likelihood rows
    mg = rn(ns, B, D); qg = torch.rand(ns, B, D, generator=gen, dtype=torch.float64); cG = torch.rand(ns, B, generator=gen, dtype=torch.float64)
    y = rn(B)
    Rs = torch.zeros(ns, dtype=torch.float64); gh = torch.zeros(7, dtype=torch.float64)
    Rsd, ghd = g(Rs.clone()), g(gh.clone())
    ref = specs.lik_rows(lr, mg, qg, cG, y, I, hyp, 0.37, Rs, gh)
    got = ops.lik_rows(g(lr), g(mg), g(qg), g(cG), g(y), g(I), g(hyp), 0.37, Rsd, ghd)
    for a, b, n in zip(got, ref, ("lbar", "mgbar", "qgbar", "cGbar")):
        check(a, b, 1e-13, "lik " + n)
    check(Rsd, Rs, 1e-12, "lik Rsum"); check(ghd, gh, 1e-11, "lik ghyp")


def test_cpu_tensors_are_rejected():
    with pytest.raises(TypeError):
        ops.tril_syrk_fwd(torch.zeros(1, 4, 4, dtype=torch.float64))


@pytest.mark.parametrize("Q,D,B", [(20, 3, 37), (50, 8, 300), (50, 70, 517), (64, 5, 131), (100, 4, 90), (8, 2, 129),
                                   (21, 5, 150), (35, 66, 200)]
                         + LQ_CASES + [(100, 40, 517), (100, 2, 1)])
@pytest.mark.parametrize("per_sample_y", [False, True])
def test_latent_fused(Q, D, B, per_sample_y):
    """DMMA fused kernels (register-resident for Q <= 64, ring-pipelined for 64 < Q <= 128) against the composed
    specification; with one target vector shared by the samples or one per sample (subjects)."""
    gen = torch.Generator().manual_seed(B + 11)
    ns = 2
    rn = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    hyp = torch.tensor([1.3, 2.5, 0.7, 3.0, 1.1, 0.9, 0.05], dtype=torch.float64)
    I = make_I(gen, B, D, empty=(1,) if D > 2 else ())
    SigW = spd(D, Q, gen); muW = rn(D, Q)
    PG = rn(ns, B, Q) * 0.3; cG = torch.rand(ns, B, generator=gen, dtype=torch.float64)
    j = torch.arange(D).view(1, 1, -1)
    l = rn(ns, B, D) * (j <= I.long().view(1, -1, 1))
    y = rn(ns, B) if per_sample_y else rn(B)
    Rs = torch.zeros(ns, dtype=torch.float64); gh = torch.zeros(7, dtype=torch.float64)
    Rsd, ghd = g(Rs.clone()), g(gh.clone())
    ref = specs.latent_fused(PG, cG, l, y, I, SigW, muW, hyp, 0.37, Rs, gh)
    got = ops.latent_fused(g(PG), g(cG), g(l), g(y), g(I), g(SigW), g(muW), g(hyp), 0.37, Rsd, ghd)
    for a, b, n in zip(got, ref, ("lbar", "mgbar", "qgbar", "cGbar", "PGbar")):
        check(a, b, 1e-12, "latent_fused " + n)
    check(Rsd, Rs, 1e-12, "latent_fused Rsum"); check(ghd, gh, 1e-11, "latent_fused ghyp")


def test_counter_noise_kernel_and_in_kernel_sampling():
    """nmgp_noise_fill against its numpy restatement (float32 transcendental rounding: 1e-6), and the in-kernel noise
    of coef_sample_fwd/bwd against the explicit-noise path fed with the same generator's output (exact)."""
    gen = torch.Generator().manual_seed(3)
    B, D, ns = 333, 7, 3
    gid = torch.randperm(10 ** 6, generator=gen)[:B].to(torch.int64)
    seed, stream = 424242, (5 << 8) | 2
    z = ops.noise_fill(ns, B, D, seed, stream, 4, g(gid), DEV)
    zr = specs.noise_fill(ns, B, D, seed, stream, 4, gid)
    assert relerr(z, zr) < 1e-5
    assert abs(float(z.mean())) < 0.05 and abs(float(z.var()) - 1.0) < 0.05
    I = make_I(gen, B, D)
    m = torch.randn(B, D, generator=gen, dtype=torch.float64) * 0.3
    sd = torch.rand(B, D, generator=gen, dtype=torch.float64)
    noise = (seed, stream, 4, ns, g(gid))
    l_in = ops.coef_sample_fwd(g(m), g(sd), None, g(I), noise=noise)
    l_ex = ops.coef_sample_fwd(g(m), g(sd), z, g(I))
    assert torch.equal(l_in, l_ex)
    lb = torch.randn(ns, B, D, generator=gen, dtype=torch.float64)
    mb1, sb1 = g(torch.zeros(B, D, dtype=torch.float64)), g(torch.zeros(B, D, dtype=torch.float64))
    mb2, sb2 = mb1.clone(), sb1.clone()
    ops.coef_sample_bwd(g(lb), l_in, None, g(I), mb1, sb1, noise=noise)
    ops.coef_sample_bwd(g(lb), l_ex, z, g(I), mb2, sb2)
    assert torch.equal(mb1, mb2) and torch.equal(sb1, sb2)


@pytest.mark.parametrize("Q", [8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128, 13, 51, 99])
@pytest.mark.parametrize("mode", [0, 1])
def test_weighted_gram_every_block_count(Q, mode):
    """The Gram kernel's block-row roles are compile-time specialised per block count NB = ceil(Q/8) (odd / even dealing,
    narrow / wide CTAs): every NB from 1 to 16, both modes, ragged segments."""
    D, B, ns = 5, 203, 2 if mode == 0 else 1
    gen = torch.Generator().manual_seed(Q * 2 + mode)
    I = make_I(gen, B, D, empty=(1,))
    nm = D if mode == 0 else D * (D + 1) // 2
    Pa = torch.randn(ns, B, Q, generator=gen, dtype=torch.float64)
    Pb = torch.randn(ns, B, Q, generator=gen, dtype=torch.float64) if mode == 1 else Pa
    qb = torch.randn(ns, B, D, generator=gen, dtype=torch.float64)
    mb = torch.randn(ns, B, D, generator=gen, dtype=torch.float64)
    SB = torch.randn(nm, Q, Q, generator=gen, dtype=torch.float64); MB = torch.randn(nm, Q, generator=gen, dtype=torch.float64)
    SBd, MBd = g(SB.clone()), g(MB.clone())
    seg = ops.segment_offsets(g(I), D)
    specs.weighted_gram(Pa, Pb, I, qb, mb, mode, SB, MB)
    ops.weighted_gram(g(Pa), g(Pb), g(I), g(qb), g(mb), mode, SBd, MBd, seg=seg)
    check(SBd, SB, 1e-12, "gram Sig"); check(MBd, MB, 1e-12, "gram Mu")
