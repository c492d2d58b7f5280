"""Shared bodies of the host-API parity tests (run on the GPU through the C ABI, and on the CPU with the
kernel specs patched in, to check the host logic without a device)."""
import numpy as np
import torch

from tests import golden_util as gu
from collaborative_nonstationary_multivariate_gaussian_process_b200 import nmgp_dsvi

RTOL = 1e-9
SIM_HYPER = {"sigma2_L0_log": 0., "length_scales_L0_log": 2., "sigma2_L1_log": 0., "length_scales_L1_log": 2.,
             "sigma2_tildeell_log": 0., "length_scales_tildeell_log": 0., "sigma2_err_log": -2.}


def split(vec, counts):
    return [a.reshape(-1, 1) for a in np.split(vec, np.cumsum(counts)[:-1])]


def predict_modelpt(device):
    g = gu.load("predict_modelpt")
    m = nmgp_dsvi.NMGP(200, 2, torch.from_numpy(g["Z"]).view(-1, 1), device=device)
    sd = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param_")}
    m.load_state_dict(sd)                                    # the reference checkpoint's keys load as-is
    pred = nmgp_dsvi.predict_Y(m, split(g["Xt"], g["nt_per_output"]))
    assert np.max(np.abs(pred - g["pred"])) <= RTOL * np.max(np.abs(g["pred"]))
    assert np.allclose(pred[:3], [3.46511351, -3.43782009, 2.70066831], atol=1e-8)
    rmse = np.sqrt(np.mean((pred[:, None] - g["Yt"][:, None]) ** 2))
    assert abs(rmse - 0.8258476499644105) < 1e-9


def forward_backward_reference_noise(device):
    """NMGP(seed) + global-generator noise in the reference's draw order reproduces the reference loss/grads."""
    g = gu.load("dsvi_sim_low")
    counts = [int((g["I"] == d).sum()) for d in range(2)]
    m = nmgp_dsvi.NMGP(200, 2, torch.from_numpy(g["Z"]).view(-1, 1), seed=int(g["seed"]), device=device)
    for k, v in SIM_HYPER.items():
        getattr(m, k).data.fill_(v)
    for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
        getattr(m, k).requires_grad = False
    Xl = [torch.from_numpy(a) for a in split(g["x"], counts)]
    Yl = [torch.from_numpy(a) for a in split(g["y"], counts)]
    torch.manual_seed(1000 + int(g["seed"]))
    loss = m(Xl, Yl)
    loss.backward()
    ref = float(g["loss"])
    assert abs(float(loss) - ref) <= RTOL * abs(ref)
    for k, prm in m.named_parameters():
        if prm.requires_grad:
            gu.check_grad(k, prm.grad.cpu().numpy(), g, RTOL)
        else:
            assert prm.grad is None


def unsorted_index_matches_sorted(device):
    """`index=` with outputs given out of order: same loss as the sorted call under the same noise."""
    g = gu.load("dsvi_ragged")
    D = int(g["D"])
    counts = [int((g["I"] == d).sum()) for d in range(D)]
    p = gu.case_params(g)
    m = nmgp_dsvi.NMGP(int(g["N"]), D, torch.from_numpy(g["Z"]).view(-1, 1), device=device)
    m.load_state_dict(p)
    Xl = [torch.from_numpy(a) for a in split(g["x"], counts)]
    Yl = [torch.from_numpy(a) for a in split(g["y"], counts)]
    noise = (torch.from_numpy(g["z_v"][:1]), torch.from_numpy(g["z_ell"][:1]), torch.from_numpy(g["z_L"][:1]))
    a = float(m(Xl, Yl, explicit_noise=noise))
    order = [2, 0, 3, 1]
    # explicit noise is indexed by sorted rows, so the permuted call must give the identical loss
    b = float(m([Xl[i] for i in order], [Yl[i] for i in order], index=order, explicit_noise=noise))
    assert abs(a - b) <= 1e-12 * abs(a)
    assert abs(a - float(g["losses"][0])) <= RTOL * abs(a)


def inference_trace(device):
    """inference() with the notebook settings: first five losses of the reference's seed-0 low_freq run and the
    posterior mean after those five Adam steps."""
    g = gu.load("inference_sim_low")
    X_list = split(g["X"], g["n_per_output"]); Y_list = split(g["Y"], g["n_per_output"])
    model, loss_list, time_list = nmgp_dsvi.inference(X_list, Y_list, np.linspace(0, 1, 20), 200, 2,
                                                      hyperpars=dict(SIM_HYPER), lr=0.005, itnum=5, seed=0,
                                                      show_ELBO=False, device=device)
    got = np.array([float(l) for l in loss_list])
    assert np.allclose(got[:3], [32497.476653962425, 30311.70398586406, 28249.774028662527], rtol=RTOL)
    assert np.max(np.abs(got - g["losses"]) / np.abs(g["losses"])) <= RTOL, got
    pred = nmgp_dsvi.predict_Y(model, split(g["Xt"], g["nt_per_output"]))
    assert np.max(np.abs(pred - g["pred_after"])) <= 1e-8 * np.max(np.abs(g["pred_after"]))
    assert len(time_list) == 5


def compute_elbo_replay(device):
    """compute_ELBO (quirk q5 included) with the reference's random stream: 3 draws on low_freq with model.pt."""
    g = gu.load("elbo_modelpt"); gm = gu.load("predict_modelpt")
    m = nmgp_dsvi.NMGP(200, 2, torch.from_numpy(gm["Z"]).view(-1, 1), device=device)
    m.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in gm.items() if k.startswith("param_")})
    Xl = [torch.from_numpy(a) for a in split(g["X"], g["n_per_output"])]
    Yl = [torch.from_numpy(a) for a in split(g["Y"], g["n_per_output"])]
    torch.manual_seed(77)
    e = float(m.compute_ELBO(Xl, Yl, n_sample=3))
    assert abs(e - float(g["elbo"])) <= RTOL * abs(float(g["elbo"])), (e, float(g["elbo"]))
    torch.manual_seed(77)
    e2 = float(m.compute_ELBO(Xl, Yl, n_sample=3, chunk=2))          # chunking must not change the random stream
    assert abs(e2 - e) <= 1e-12 * abs(e)


def posterior_sampling_replay(device):
    """sample_Y / sample_FY with the reference's random stream (2 draws each)."""
    g = gu.load("sample_Y_modelpt"); gm = gu.load("predict_modelpt")
    m = nmgp_dsvi.NMGP(200, 2, torch.from_numpy(gm["Z"]).view(-1, 1), device=device)
    m.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in gm.items() if k.startswith("param_")})
    torch.manual_seed(31)
    Ys, Ls, Gs, ells = nmgp_dsvi.sample_Y(m, split(g["X"], g["n_per_output"]), n_sample=2)
    for got, ref, name in ((Ys, g["Ys"], "Ys"), (Ls, g["Ls"], "Ls"), (Gs, g["Gs"], "Gs"), (ells, g["ells"], "ells")):
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        assert np.max(np.abs(got - ref)) <= RTOL * np.max(np.abs(ref)), (name, np.max(np.abs(got - ref)))
    g3 = gu.load("sample_FY_d3")
    m3 = nmgp_dsvi.NMGP(50, 3, torch.from_numpy(g3["Z"]).view(-1, 1), device=device)
    m3.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in g3.items() if k.startswith("param_")})
    torch.manual_seed(32)
    E, Y, C = nmgp_dsvi.sample_FY(m3, g3["grid"], n_sample=2)
    for got, ref, name in ((E, g3["ells"], "ells"), (Y, g3["Ys"], "Ys"), (C, g3["corrs"], "corrs")):
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        assert np.max(np.abs(got - ref)) <= RTOL * max(1.0, np.max(np.abs(ref))), (name, np.max(np.abs(got - ref)))
