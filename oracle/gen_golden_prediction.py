"""Golden vectors for the SIM_code predictive functions (code/SIM_code/Utility/prediction.py:337-458), produced by the
UNMODIFIED reference under the shim of oracle/gen_golden.py.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/gen_golden_prediction.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, install_shim  # noqa: E402


def main():
    install_shim()
    from Utility import prediction, utils as sim_utils
    torch.manual_seed(11)
    N, M, G = 60, 3, 7
    x = torch.sort(torch.rand(N).double())[0]
    tilde_l = (3 * (x - 1) ** 3 - 1.5) + 0.05 * torch.randn(N).double()        # sim.py:22-26 shape of log-ell
    tilde_sigma = 0.1 * torch.randn(N).double()
    Lm = torch.tril(torch.randn(M, M).double()); Lm[range(M), range(M)] = torch.exp(0.3 * torch.randn(M).double())
    uL_vec = sim_utils.Lvec2uLvec(sim_utils.lowtriangle2vec(Lm, M), M)
    tilde_s2 = torch.tensor(np.log(1e-2)).double()
    Y = torch.randn(N, M).double()
    grids = torch.linspace(0.03, 0.97, G).double()
    hyp = dict(mu_tilde_l=torch.tensor(-1.0).double(), alpha_tilde_l=torch.tensor(1.5).double(),
               beta_tilde_l=torch.tensor(0.3).double(), mu_tilde_sigma=torch.tensor(0.0).double(),
               alpha_tilde_sigma=torch.tensor(0.8).double(), beta_tilde_sigma=torch.tensor(0.4).double())
    order = ("mu_tilde_l", "alpha_tilde_l", "beta_tilde_l", "mu_tilde_sigma", "alpha_tilde_sigma", "beta_tilde_sigma")
    args = [hyp[k] for k in order]
    one = prediction.point_predmap(tilde_l, tilde_sigma, uL_vec, tilde_s2, Y, x, grids[2], *args)
    allg = prediction.pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_s2, Y, x, grids, *args)
    tst = prediction.test_predmap(tilde_l, tilde_sigma, uL_vec, tilde_s2, Y, x, x[::9] + 0.004, *args)
    # sampling predictors: a short parameter history (N_hist = 4, the last N_sample = 3 are used), seeded global RNG
    H = 4
    tl_h = torch.stack([tilde_l + 0.03 * torch.randn(N).double() for _ in range(H)])
    ts_h = torch.stack([tilde_sigma + 0.03 * torch.randn(N).double() for _ in range(H)])
    uL_h = torch.stack([uL_vec + 0.05 * torch.randn(uL_vec.numel()).double() for _ in range(H)])
    s2_h = tilde_s2 + 0.05 * torch.randn(H).double()
    import contextlib, io
    torch.manual_seed(123)
    ps_one = prediction.point_predsample(tl_h, ts_h, uL_h, s2_h, Y, x, grids[4], *args, 3)
    torch.manual_seed(321)
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints every grid point
        ps_grid = prediction.pointwise_predsample(tl_h, ts_h, uL_h, s2_h, Y, x, grids[:3], *args, 3)
    torch.manual_seed(77)
    with contextlib.redirect_stdout(io.StringIO()):
        mq, mm, ms = prediction.pointwise_predmap_sampling(6, tilde_l, tilde_sigma, uL_vec, tilde_s2, Y, x, grids[1:3], *args)
    # stationary (_S) variants: scalar log-ell / log-sigma
    tl_S, ts_S = torch.tensor(-1.8).double(), torch.tensor(0.2).double()
    S_grid = prediction.pointwise_predmap_S(tl_S, ts_S, uL_vec, tilde_s2, Y, x, grids)
    S_mean, S_std = prediction.test_predmap_S(tl_S, ts_S, uL_vec, tilde_s2, Y, x, x[::9] + 0.004)
    tls_S = tl_S + 0.05 * torch.randn(3).double(); tss_S = ts_S + 0.05 * torch.randn(3).double()
    np.random.seed(5)
    S_samp = prediction.pointwise_predsample_S(tls_S, tss_S, uL_h[:3], s2_h[:3], Y, x, grids[:4])
    # Hadamard (irregular observations): each output observed at its own subset of the inputs
    gen = torch.Generator().manual_seed(5)
    keep = torch.rand(N, M, generator=gen) < 0.6
    xh = torch.cat([x[keep[:, m]] for m in range(M)])
    ih = torch.cat([torch.full((int(keep[:, m].sum()),), m, dtype=torch.long) for m in range(M)])
    yh = torch.cat([Y[keep[:, m], m] for m in range(M)])
    tlh = (3 * (xh - 1) ** 3 - 1.5) + 0.05 * torch.randn(xh.numel()).double()
    tsh = 0.1 * torch.randn(xh.numel()).double()
    L_vec_h = sim_utils.uLvec2Lvec(uL_vec, M)
    tl_S0, ts_S0 = torch.tensor(-1.8).double(), torch.tensor(0.2).double()
    H_point = prediction.point_predmap_hadamard(tlh, tsh, L_vec_h, tilde_s2, xh, ih, yh, grids[3], *args)
    with contextlib.redirect_stdout(io.StringIO()):
        H_grid = prediction.pointwise_predmap_hadmard(tlh, tsh, L_vec_h, tilde_s2, xh, ih, yh, grids[:3], *args)
    H_idx = prediction.indexedpoint_predmap_hadamard(tlh, tsh, L_vec_h, tilde_s2, xh, ih, yh, grids[5], torch.tensor(1), *args)
    xt_h = x[::9][:5] + 0.004; it_h = torch.tensor([0, 2, 1, 1, 0])
    H_test = prediction.test_predmap_harmard(tlh, tsh, L_vec_h, tilde_s2, xh, ih, yh, xt_h, it_h, *args)
    SH_grid = prediction.pointwise_predmap_S_hadamard(tl_S0, ts_S0, L_vec_h, tilde_s2, xh, ih, yh, grids[:3])
    SH_mean, SH_std = prediction.test_predmap_S_hadamard(tl_S0, ts_S0, L_vec_h, tilde_s2, xh, ih, yh, xt_h, it_h)
    # spatially varying coregionalisation ("inhomogeneous"): a packed triangle per input
    Ni, Mi = 24, 2
    Pi = Mi * (Mi + 1) // 2
    gi = torch.Generator().manual_seed(17)
    xi = torch.sort(torch.rand(Ni, generator=gi).double())[0]
    tli = (3 * (xi - 1) ** 3 - 1.5) + 0.05 * torch.randn(Ni, generator=gi).double()
    uLi = (0.3 * torch.randn(Ni * Pi, generator=gi).double())
    Yi = torch.randn(Ni, Mi, generator=gi).double()
    hyp_i = [torch.tensor(v).double() for v in (-1.0, 1.5, 0.3, 0.1, 0.7, 0.35)]
    with contextlib.redirect_stdout(io.StringIO()):
        IN_y, IN_L = prediction.pointwise_predmap_inhomogeneous(tli, uLi, tilde_s2, Yi, xi, grids[1:4], *hyp_i)
    torch.manual_seed(91)
    with contextlib.redirect_stdout(io.StringIO()):
        INS_q, INS_m, INS_s = prediction.pointwise_predmap_inhomogeneous_sampling(5, tli, uLi, tilde_s2, Yi, xi, grids[1:3], *hyp_i)
    torch.manual_seed(92)
    with contextlib.redirect_stdout(io.StringIO()):
        INS_l = prediction.pointwise_predmap_inhomogeneous_sampling(4, tli, uLi, tilde_s2, Yi, xi, grids[1:3], *hyp_i, pred_smoothness=True)
    torch.manual_seed(93)
    with contextlib.redirect_stdout(io.StringIO()):
        INS_L = prediction.pointwise_predmap_inhomogeneous_sampling(4, tli, uLi, tilde_s2, Yi, xi, grids[1:3], *hyp_i, pred_cov=True)
    tli_h = torch.stack([tli + 0.03 * torch.randn(Ni, generator=gi).double() for _ in range(3)])
    uLi_h = torch.stack([uLi + 0.05 * torch.randn(Ni * Pi, generator=gi).double() for _ in range(3)])
    s2i_h = tilde_s2 + 0.05 * torch.randn(3, generator=gi).double()
    torch.manual_seed(94)
    with contextlib.redirect_stdout(io.StringIO()):
        INP = prediction.pointwise_predsample_inhomogeneous(tli_h, uLi_h, s2i_h, Yi, xi, grids[1:3], *hyp_i, 2)
    # SVC Hadamard: irregular observations with a packed triangle per observation (used raw, no exp)
    Nh = xh.numel()
    Lv_svc = 0.4 * torch.randn(Nh * (M * (M + 1) // 2), generator=gi).double() + 0.3
    with contextlib.redirect_stdout(io.StringIO()):
        SVC_grid = prediction.pointwise_predmap_SVC_hadamard(tlh, Lv_svc, tilde_s2, xh, ih, yh, grids[2:5], *hyp_i)
        SVC_m, SVC_v = prediction.test_predmap_SVC_hadamard(tlh, Lv_svc, tilde_s2, xh, ih, yh, xt_h, it_h, *hyp_i)
    SVC_idx = prediction.indexedpoint_predmap_SVC_hadamard(tlh, Lv_svc, tilde_s2, xh, ih, yh, grids[1], torch.tensor(2), *hyp_i)
    Hh = 3
    tlh_h = torch.stack([tlh + 0.03 * torch.randn(xh.numel()).double() for _ in range(Hh)])
    tsh_h = torch.stack([tsh + 0.03 * torch.randn(xh.numel()).double() for _ in range(Hh)])
    Lh_h = torch.stack([L_vec_h + 0.05 * torch.randn(L_vec_h.numel()).double() for _ in range(Hh)])
    s2h_h = tilde_s2 + 0.05 * torch.randn(Hh).double()
    torch.manual_seed(41)
    with contextlib.redirect_stdout(io.StringIO()):
        HS_grid = prediction.pointwise_predsample_hadamard(tlh_h, tsh_h, Lh_h, s2h_h, xh, ih, yh, grids[2:4], *args)
    torch.manual_seed(42)
    with contextlib.redirect_stdout(io.StringIO()):
        HS_test = prediction.test_predsample_hadamard(tlh_h, tsh_h, Lh_h, s2h_h, xh, ih, yh, xt_h[:3], it_h[:3], *args)
    np.savez_compressed(os.path.join(OUT, "sim_prediction.npz"), x=x.numpy(), tilde_l=tilde_l.numpy(),
                        tlh_h=tlh_h.numpy(), tsh_h=tsh_h.numpy(), Lh_h=Lh_h.numpy(), s2h_h=s2h_h.numpy(),
                        HS_grid=HS_grid.numpy(), HS_test=HS_test.numpy(), Lv_svc=Lv_svc.numpy(),
                        SVC_grid=SVC_grid.numpy(), SVC_m=SVC_m.numpy(), SVC_v=SVC_v.numpy(), SVC_idx=SVC_idx.numpy(),
                        xi=xi.numpy(), tli=tli.numpy(), uLi=uLi.numpy(), Yi=Yi.numpy(), hyp_i=np.array([float(v) for v in hyp_i]),
                        IN_y=IN_y.numpy(), IN_L=IN_L.numpy(), INS_q=INS_q, INS_m=INS_m, INS_s=INS_s, INS_l=INS_l, INS_L=INS_L,
                        tli_h=tli_h.numpy(), uLi_h=uLi_h.numpy(), s2i_h=s2i_h.numpy(), INP=INP,
                        SH_grid=SH_grid.numpy(), SH_mean=SH_mean.numpy(), SH_std=SH_std.numpy(),
                        xh=xh.numpy(), ih=ih.numpy(), yh=yh.numpy(), tlh=tlh.numpy(), tsh=tsh.numpy(), L_vec_h=L_vec_h.numpy(),
                        H_point=H_point.numpy(), H_grid=H_grid.numpy(), H_idx=H_idx.numpy(), H_test=H_test.numpy(),
                        xt_h=xt_h.numpy(), it_h=it_h.numpy(),
                        tl_S=float(tl_S), ts_S=float(ts_S), S_grid=S_grid.numpy(), S_mean=S_mean.numpy(), S_std=S_std.numpy(),
                        tls_S=tls_S.numpy(), tss_S=tss_S.numpy(), S_samp=S_samp,
                        mapsamp_q=mq, mapsamp_mean=mm, mapsamp_std=ms,
                        tl_hist=tl_h.numpy(), ts_hist=ts_h.numpy(), uL_hist=uL_h.numpy(), s2_hist=s2_h.numpy(),
                        predsample_point=ps_one.numpy(), predsample_grid=np.asarray(ps_grid),
                        tilde_sigma=tilde_sigma.numpy(), uL_vec=uL_vec.numpy(), tilde_s2=float(tilde_s2), Y=Y.numpy(),
                        grids=grids.numpy(), x_test=(x[::9] + 0.004).numpy(),
                        **{k: float(v) for k, v in hyp.items()}, point=one.numpy(), pointwise=allg.numpy(),
                        test=tst.numpy(), L_vec=sim_utils.uLvec2Lvec(uL_vec, M).numpy())
    print("point_predmap", one.numpy())


if __name__ == "__main__":
    main()
