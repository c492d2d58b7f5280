"""Pins the CPU oracle (oracle/nmgp_oracle.py) to the golden vectors produced by the
real reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import nmgp_oracle as orc
from tests import golden_util as gu

RTOL = 1e-12     # same library, same op order: the oracle should be at rounding level


@pytest.mark.parametrize("name", gu.DSVI_CASES)
def test_dsvi_step_matches_reference(name):
    g = gu.load(name)
    p = gu.case_params(g)
    Xl, Yl = gu.case_lists(g)
    Z = torch.from_numpy(g["Z"]).view(-1, 1)
    loss, grads = orc.step_loss_and_grads(p, Z, int(g["N"]), Xl, Yl, draws=gu.replay_draws(g),
                                          train_lengthscales=bool(int(g["train_len"])),
                                          outputs_lists=gu.case_target_lists(g))
    assert abs(float(loss) - float(g["loss"])) <= RTOL * abs(float(g["loss"]))
    for k in orc.PARAM_NAMES:
        if grads[k] is None:
            assert ("grad_" + k) not in g and ("grad_" + k + "__norm") not in g
            continue
        gu.check_grad(k, grads[k].numpy(), g, 1e-10)


def test_predict_modelpt_known_answer():
    g = gu.load("predict_modelpt")
    p = {k: torch.from_numpy(g["param_" + k]) for k in orc.PARAM_NAMES}
    nt = g["nt_per_output"]
    xs = np.split(g["Xt"], np.cumsum(nt)[:-1])
    Xl = [torch.from_numpy(x).view(-1, 1) for x in xs]
    pred = orc.posterior_mean(p, torch.from_numpy(g["Z"]).view(-1, 1), Xl).numpy()
    assert np.allclose(pred[:3], [3.46511351, -3.43782009, 2.70066831], atol=1e-8)
    assert np.max(np.abs(pred - g["pred"])) <= 1e-11 * np.max(np.abs(g["pred"]))
    rmse = np.sqrt(np.mean((pred[:, None] - g["Yt"][:, None]) ** 2))
    assert abs(rmse - 0.8258476499644105) < 1e-12


def test_mc_elbo_replay():
    g = gu.load("elbo_modelpt")
    gm = gu.load("predict_modelpt")
    p = {k: torch.from_numpy(gm["param_" + k]) for k in orc.PARAM_NAMES}
    n = g["n_per_output"]
    Xl = [torch.from_numpy(x).view(-1, 1) for x in np.split(g["X"], np.cumsum(n)[:-1])]
    Yl = [torch.from_numpy(x).view(-1, 1) for x in np.split(g["Y"], np.cumsum(n)[:-1])]
    noise = torch.from_numpy(g["noise"])
    pos = [0]

    def draw(shape):
        k = int(np.prod(shape))
        t = noise[pos[0]:pos[0] + k].reshape(shape)
        pos[0] += k
        return t
    e = orc.mc_elbo(p, torch.from_numpy(gm["Z"]).view(-1, 1), 200, Xl, Yl, n_sample=3, draw=draw)
    assert pos[0] == noise.numel()
    assert abs(float(e) - float(g["elbo"])) <= 1e-12 * abs(float(g["elbo"]))


def test_sim_code_kernels_and_kron():
    g = gu.load("sim_code")
    t = lambda k: torch.from_numpy(g[k])
    x1 = t("x1").view(-1, 1); x2 = t("x2").view(-1, 1)
    close = lambda a, b: np.max(np.abs(a.numpy() - b)) <= 1e-13 * max(1.0, np.max(np.abs(b)))
    assert close(orc.sim_nonstationary_cov(x1, t("sg1"), t("ell1")), g["K_self"])
    assert close(orc.sim_nonstationary_cov(x1, t("sg1"), t("ell1"), x2, t("sg2"), t("ell2")), g["K_cross"])
    assert close(orc.sim_nonstationary_cov(x1), g["K_def"])
    assert close(orc.sim_rbf_cov(x1, alpha=1.3, beta=0.2), g["R_self"])
    assert close(orc.sim_rbf_cov(x1, x2, alpha=0.7, beta=0.35), g["R_cross"])
    K = t("K_self"); Bf = t("Bf"); y = t("y"); mu = t("mu"); s2 = torch.tensor(float(g["s2"]), dtype=torch.float64)
    assert close(orc.kron_matvec(Bf, K, y), g["kron_mv"])
    assert close(orc.kron_dense(Bf, K[:5, :4]), g["kron_prod"])
    assert close(orc.kron_diag(torch.diagonal(Bf), torch.diagonal(K)), g["kron_diag"])
    assert abs(float(orc.kron_logdet(s2, Bf, K)) - float(g["kron_logdet"])) <= 1e-12 * abs(float(g["kron_logdet"]))
    inv = orc.kron_inverse(s2, Bf, K)
    assert np.allclose(torch.diagonal(inv).numpy(), g["kron_inv_diag"], rtol=1e-9)
    assert np.allclose(inv[7].numpy(), g["kron_inv_row7"], rtol=1e-8, atol=1e-9)
    lp0 = float(orc.mvn_logpdf_kron_eig(y, mu, Bf, K, s2))
    lp2 = float(orc.mvn_logpdf_dense(y, mu, Bf, K, s2))
    assert abs(lp0 - float(g["logpdf0"])) <= 1e-12 * abs(lp0)
    assert abs(lp2 - float(g["logpdf2"])) <= 1e-11 * abs(lp2)
    assert abs(lp0 - lp2) <= 1e-9 * abs(lp2)                       # distributions.py:163-169 identity
    rands = [t("rand_B"), t("rand_K")]
    lp1 = float(orc.mvn_logpdf_kron_eig_jittered(y, mu, Bf, K, s2, rand=lambda n: rands.pop(0)))
    assert abs(lp1 - float(g["logpdf1"])) <= 1e-11 * abs(lp1)
    # kron_mv identity eyeballed at kronecker_operation.py:112-115
    dense = torch.mv(orc.kron_dense(Bf, K), y)
    assert torch.max(torch.abs(dense - orc.kron_matvec(Bf, K, y))) < 1e-11


def test_sim_code_t200():
    g = gu.load("sim_code_t200")
    x1 = torch.from_numpy(g["x1"]).view(-1, 1)
    K = orc.sim_nonstationary_cov(x1, ell1=torch.from_numpy(g["ell1"]))
    assert np.allclose(K[17].numpy(), g["K_row17"], rtol=1e-13, atol=1e-15)
    y = torch.from_numpy(g["y"]); Bf = torch.from_numpy(g["Bf"])
    s2 = torch.tensor(float(g["s2"]), dtype=torch.float64)
    lp0 = float(orc.mvn_logpdf_kron_eig(y, torch.zeros_like(y), Bf, K, s2))
    assert abs(lp0 - float(g["logpdf0"])) <= 1e-11 * abs(lp0)
    assert abs(lp0 - float(g["logpdf2"])) <= 1e-9 * abs(lp0)
