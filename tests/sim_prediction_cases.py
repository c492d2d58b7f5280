"""Shared body of the SIM_code prediction parity tests (CPU: kernel specs patched in; GPU: through the C ABI).
Golden outputs come from the unmodified reference (oracle/gen_golden_prediction.py)."""
import numpy as np
import torch

from tests import golden_util as gu

ORDER = ("mu_tilde_l", "alpha_tilde_l", "beta_tilde_l", "mu_tilde_sigma", "alpha_tilde_sigma", "beta_tilde_sigma")
# The reference solves the two GP-conditional systems Sigma + 1e-6 I (condition number ~1e8 for a smooth RBF on 60
# points) by LU and eigendecomposes K_x; we use Cholesky factors plus ONE step of FP64 iterative refinement on those
# systems (prediction._ConditionalGP.solve), which brings our side to rounding level: what is left is the reference's
# own LU error.  Measured (CPU specifications): <= 5.2e-10 everywhere except the SVC predictive variance (1.4e-9, a
# cancellation of two dense (NM)^2 solves).
RTOL = 1e-9
# Hadamard / SVC layouts: the same input appears once per output, so RBF_cov(x) + 1e-6 I has exactly repeated rows
# (condition number ~ N alpha^2 / 1e-6 ~ 1e8) and the reference goes through symeig(K) and an explicit inverse -- its own
# result carries ~cond * eps.  Bound below = measured worst case (1.4e-9) with margin.
HTOL = 5e-9


WORST = {"max": 0.0}          # largest relative deviation seen by the last run_all (reported by the tests)


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    e = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    WORST["max"] = max(WORST["max"], e)
    return e


def run_all(dev):
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import prediction
    g = gu.load("sim_prediction")
    d = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.float64)).to(dev)
    hyp = [torch.tensor(float(g[k]), dtype=torch.float64) for k in ORDER]
    M = g["Y"].shape[1]
    # helpers of SIM_code/Utility/utils.py
    Lv = prediction.uLvec2Lvec(d("uL_vec"), M)
    assert _rel(Lv.cpu().numpy(), g["L_vec"]) < 1e-15
    assert _rel(prediction.Lvec2uLvec(Lv, M).cpu().numpy(), g["uL_vec"]) < 1e-14
    Lm = prediction.vec2lowtriangle(Lv, M)
    assert float(torch.triu(Lm, 1).abs().max()) == 0.0
    assert _rel(prediction.lowtriangle2vec(Lm, M).cpu().numpy(), g["L_vec"]) < 1e-15
    ph = torch.arange(2 * (2 * 5 + 3 + 1), dtype=torch.float64).reshape(2, -1)
    a_, b_, c_, e_ = prediction.vec2pars(ph, 5, 2)
    assert a_.shape == (2, 5) and b_.shape == (2, 5) and c_.shape == (2, 3) and float(e_[1]) == float(ph[1, -1])
    yl = prediction.vec2list(torch.tensor([1., 2., 3., 4.]), torch.tensor([0, 1, 0, 1]))
    assert [t.tolist() for t in yl] == [[1., 3.], [2., 4.]]
    args = (d("tilde_l"), d("tilde_sigma"), d("uL_vec"), torch.tensor(float(g["tilde_s2"]), dtype=torch.float64).to(dev),
            d("Y"), d("x"))
    one = prediction.point_predmap(*args, d("grids")[2], *hyp)
    assert one.shape == (3, M)
    assert _rel(one.cpu().numpy(), g["point"]) < RTOL, _rel(one.cpu().numpy(), g["point"])
    allg = prediction.pointwise_predmap(*args, d("grids"), *hyp)
    assert allg.shape == g["pointwise"].shape
    assert _rel(allg.cpu().numpy(), g["pointwise"]) < RTOL, _rel(allg.cpu().numpy(), g["pointwise"])
    tst = prediction.test_predmap(*args, d("x_test"), *hyp)
    assert _rel(tst.cpu().numpy(), g["test"]) < RTOL, _rel(tst.cpu().numpy(), g["test"])
    # size-independent property: the band is symmetric around the mean and at least the noise level wide
    half = (allg[:, 2] - allg[:, 0]) / 2
    assert torch.allclose(allg[:, 1], (allg[:, 2] + allg[:, 0]) / 2, rtol=0, atol=1e-12)
    assert bool((half >= 1.96 * np.sqrt(np.exp(float(g["tilde_s2"]))) * (1 - 1e-9)).all())
    # sampling predictors: the global CPU generator is consumed exactly as by the reference (same seeds as the generator
    # script), so the draws themselves are reproduced
    hist = (d("tl_hist"), d("ts_hist"), d("uL_hist"), d("s2_hist"))
    torch.manual_seed(123)
    ps = prediction.point_predsample(*hist, d("Y"), d("x"), d("grids")[4], *hyp, 3)
    assert ps.shape == g["predsample_point"].shape
    assert _rel(ps.cpu().numpy(), g["predsample_point"]) < RTOL, _rel(ps.cpu().numpy(), g["predsample_point"])
    torch.manual_seed(321)
    pg = prediction.pointwise_predsample(*hist, d("Y"), d("x"), d("grids")[:3], *hyp, 3)
    assert isinstance(pg, np.ndarray) and pg.shape == g["predsample_grid"].shape
    assert _rel(pg, g["predsample_grid"]) < RTOL, _rel(pg, g["predsample_grid"])
    torch.manual_seed(77)
    mq, mm, ms = prediction.pointwise_predmap_sampling(6, *args, d("grids")[1:3], *hyp)
    assert mq.shape == g["mapsamp_q"].shape and mm.shape == g["mapsamp_mean"].shape
    for got, key in ((mq, "mapsamp_q"), (mm, "mapsamp_mean"), (ms, "mapsamp_std")):
        assert _rel(got, g[key]) < 10 * RTOL, (key, _rel(got, g[key]))        # std of 6 draws: difference of close numbers
    # stationary (_S) variants: the reference inverts the dense (N M) x (N M) matrix; here the eigen-block factors
    sc = lambda k: torch.tensor(float(g[k]), dtype=torch.float64).to(dev)
    s2t = torch.tensor(float(g["tilde_s2"]), dtype=torch.float64).to(dev)
    Sg = prediction.pointwise_predmap_S(sc("tl_S"), sc("ts_S"), d("uL_vec"), s2t, d("Y"), d("x"), d("grids"))
    assert _rel(Sg.cpu().numpy(), g["S_grid"]) < RTOL, _rel(Sg.cpu().numpy(), g["S_grid"])
    Sm, Ss = prediction.test_predmap_S(sc("tl_S"), sc("ts_S"), d("uL_vec"), s2t, d("Y"), d("x"), d("x_test"))
    assert _rel(Sm.cpu().numpy(), g["S_mean"]) < RTOL and _rel(Ss.cpu().numpy(), g["S_std"]) < RTOL
    np.random.seed(5)
    Sp = prediction.pointwise_predsample_S(d("tls_S"), d("tss_S"), d("uL_hist")[:3], d("s2_hist")[:3], d("Y"), d("x"),
                                           d("grids")[:4])
    assert Sp.shape == g["S_samp"].shape and _rel(Sp, g["S_samp"]) < RTOL, _rel(Sp, g["S_samp"])
    # Hadamard (irregular observations) MAP predictors: dense Cholesky of K_x * K_i + sigma2 I instead of symeig + inverse
    ih = torch.from_numpy(g["ih"]).to(dev)
    hargs = (d("tlh"), d("tsh"), d("L_vec_h"), s2t, d("xh"), ih, d("yh"))
    Hp = prediction.point_predmap_hadamard(*hargs, d("grids")[3], *hyp)
    assert _rel(Hp.cpu().numpy(), g["H_point"]) < HTOL, _rel(Hp.cpu().numpy(), g["H_point"])
    Hg = prediction.pointwise_predmap_hadmard(*hargs, d("grids")[:3], *hyp)
    assert _rel(Hg.cpu().numpy(), g["H_grid"]) < HTOL, _rel(Hg.cpu().numpy(), g["H_grid"])
    Hi = prediction.indexedpoint_predmap_hadamard(*hargs, d("grids")[5], torch.tensor(1), *hyp)
    assert Hi.shape == (3,) and _rel(Hi.cpu().numpy(), g["H_idx"]) < HTOL, _rel(Hi.cpu().numpy(), g["H_idx"])
    Ht = prediction.test_predmap_harmard(*hargs, d("xt_h"), torch.from_numpy(g["it_h"]), *hyp)
    assert _rel(Ht.cpu().numpy(), g["H_test"]) < HTOL, _rel(Ht.cpu().numpy(), g["H_test"])
    # stationary Hadamard (incl. the reference's quirk: the indexed variants take the prior variance of output 0)
    SHg = prediction.pointwise_predmap_S_hadamard(sc("tl_S"), sc("ts_S"), d("L_vec_h"), s2t, d("xh"), ih, d("yh"), d("grids")[:3])
    assert _rel(SHg.cpu().numpy(), g["SH_grid"]) < HTOL, _rel(SHg.cpu().numpy(), g["SH_grid"])
    SHm, SHs = prediction.test_predmap_S_hadamard(sc("tl_S"), sc("ts_S"), d("L_vec_h"), s2t, d("xh"), ih, d("yh"), d("xt_h"),
                                                  torch.from_numpy(g["it_h"]))
    assert _rel(SHm.cpu().numpy(), g["SH_mean"]) < HTOL and _rel(SHs.cpu().numpy(), g["SH_std"]) < HTOL
    # spatially varying coregionalisation: dense (N M) x (N M) system, one Cholesky for all grid points
    hyp_i = [torch.tensor(float(v), dtype=torch.float64) for v in g["hyp_i"]]
    INy, INL = prediction.pointwise_predmap_inhomogeneous(d("tli"), d("uLi"), s2t, d("Yi"), d("xi"), d("grids")[1:4], *hyp_i)
    assert INy.shape == g["IN_y"].shape and INL.shape == g["IN_L"].shape
    assert _rel(INy.cpu().numpy(), g["IN_y"]) < HTOL, _rel(INy.cpu().numpy(), g["IN_y"])
    assert _rel(INL.cpu().numpy(), g["IN_L"]) < HTOL, _rel(INL.cpu().numpy(), g["IN_L"])
    inh = (d("tli"), d("uLi"), s2t, d("Yi"), d("xi"), d("grids")[1:3])
    torch.manual_seed(91)
    q_, m_, s_ = prediction.pointwise_predmap_inhomogeneous_sampling(5, *inh, *hyp_i)
    for got, key in ((q_, "INS_q"), (m_, "INS_m"), (s_, "INS_s")):
        assert got.shape == g[key].shape and _rel(got, g[key]) < 10 * HTOL, (key, _rel(got, g[key]))
    torch.manual_seed(92)
    l_ = prediction.pointwise_predmap_inhomogeneous_sampling(4, *inh, *hyp_i, pred_smoothness=True)
    assert l_.shape == g["INS_l"].shape and _rel(l_, g["INS_l"]) < HTOL, _rel(l_, g["INS_l"])
    torch.manual_seed(93)
    L_ = prediction.pointwise_predmap_inhomogeneous_sampling(4, *inh, *hyp_i, pred_cov=True)
    assert L_.shape == g["INS_L"].shape and _rel(L_, g["INS_L"]) < HTOL, _rel(L_, g["INS_L"])
    torch.manual_seed(94)
    INP = prediction.pointwise_predsample_inhomogeneous(d("tli_h"), d("uLi_h"), d("s2i_h"), d("Yi"), d("xi"), d("grids")[1:3],
                                                        *hyp_i, 2)
    assert INP.shape == g["INP"].shape and _rel(INP, g["INP"]) < HTOL, _rel(INP, g["INP"])
    # SVC Hadamard (incl. the reference's return conventions for the indexed variants)
    svc = (d("tlh"), d("Lv_svc"), s2t, d("xh"), ih, d("yh"))
    SVg = prediction.pointwise_predmap_SVC_hadamard(*svc, d("grids")[2:5], *hyp_i)
    assert SVg.shape == g["SVC_grid"].shape and _rel(SVg.cpu().numpy(), g["SVC_grid"]) < HTOL, _rel(SVg.cpu().numpy(), g["SVC_grid"])
    SVm, SVv = prediction.test_predmap_SVC_hadamard(*svc, d("xt_h"), torch.from_numpy(g["it_h"]), *hyp_i)
    assert _rel(SVm.cpu().numpy(), g["SVC_m"]) < HTOL and _rel(SVv.cpu().numpy(), g["SVC_v"]) < HTOL
    SVi = prediction.indexedpoint_predmap_SVC_hadamard(*svc, d("grids")[1], torch.tensor(2), *hyp_i)
    assert SVi.shape == g["SVC_idx"].shape and _rel(SVi.cpu().numpy(), g["SVC_idx"]) < HTOL, _rel(SVi.cpu().numpy(), g["SVC_idx"])
    hh = (d("tlh_h"), d("tsh_h"), d("Lh_h"), d("s2h_h"), d("xh"), ih, d("yh"))
    torch.manual_seed(41)
    HSg = prediction.pointwise_predsample_hadamard(*hh, d("grids")[2:4], *hyp)
    assert HSg.shape == g["HS_grid"].shape and _rel(HSg.cpu().numpy(), g["HS_grid"]) < HTOL, _rel(HSg.cpu().numpy(), g["HS_grid"])
    torch.manual_seed(42)
    HSt = prediction.test_predsample_hadamard(*hh, d("xt_h")[:3], torch.from_numpy(g["it_h"])[:3], *hyp)
    assert HSt.shape == g["HS_test"].shape and _rel(HSt.cpu().numpy(), g["HS_test"]) < HTOL, _rel(HSt.cpu().numpy(), g["HS_test"])
    return {"point": _rel(one.cpu().numpy(), g["point"]), "pointwise": _rel(allg.cpu().numpy(), g["pointwise"]),
            "SVC_grid": _rel(SVg.cpu().numpy(), g["SVC_grid"]), "SVC_idx": _rel(SVi.cpu().numpy(), g["SVC_idx"]),
            "INP": _rel(INP, g["INP"]), "INS_m": _rel(m_, g["INS_m"]), "INS_l": _rel(l_, g["INS_l"]), "INS_L": _rel(L_, g["INS_L"]),
            "IN_y": _rel(INy.cpu().numpy(), g["IN_y"]), "IN_L": _rel(INL.cpu().numpy(), g["IN_L"]),
            "HS_grid": _rel(HSg.cpu().numpy(), g["HS_grid"]), "HS_test": _rel(HSt.cpu().numpy(), g["HS_test"]),
            "H_point": _rel(Hp.cpu().numpy(), g["H_point"]), "H_grid": _rel(Hg.cpu().numpy(), g["H_grid"]),
            "H_test": _rel(Ht.cpu().numpy(), g["H_test"]),
            "S_grid": _rel(Sg.cpu().numpy(), g["S_grid"]), "S_samp": _rel(Sp, g["S_samp"]),
            "mapsamp_mean": _rel(mm, g["mapsamp_mean"]), "mapsamp_std": _rel(ms, g["mapsamp_std"]),
            "predsample": _rel(ps.cpu().numpy(), g["predsample_point"]), "predsample_grid": _rel(pg, g["predsample_grid"])}
