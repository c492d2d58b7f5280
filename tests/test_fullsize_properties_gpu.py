"""Size-independent properties at BASELINE.json's full sizes -- ECoG (T=4096, D=64, Q=50, full batch B=262144), PM2.5
(T=2048, D=16, Q=100) and HCP (T=1200, D=15, Q=100, one target vector per subject) -- where the CPU oracle cannot run: (1) the row-sharded step sums to the unsharded step (linearity of every adjoint + rank-invariant
counter-based noise); (2) the hand-written gradient agrees with central finite differences of the loss along random
directions of the parameters (same noise); (3) the loss is finite and the zero-gradient blocks (quirk q8) are exact
zeros."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from collaborative_nonstationary_multivariate_gaussian_process_b200 import dsvi_step, nmgp_dsvi, parallel  # noqa: E402

DEV = "cuda:0"
S = 2                                 # S=2 keeps the test short; the per-sample work is what scales with S
# name -> (T, D, Q, driver hyper-parameters, per-subject targets)
SHAPES = {
    "ecog": (4096, 64, 50, {"length_scales_L0_log": 10., "length_scales_L1_log": 10., "length_scales_tildeell_log": 5.,
                            "sigma2_err_log": -5.}, False),
    "pm25": (2048, 16, 100, {"length_scales_L0_log": 10., "length_scales_L1_log": 10., "length_scales_tildeell_log": 10.}, False),
    "hcp": (1200, 15, 100, {"length_scales_L0_log": 5., "length_scales_L1_log": 5., "length_scales_tildeell_log": 5.}, True),
}
T, D, Q, HYPER, SUBJECTS = SHAPES["ecog"]


@pytest.fixture(params=sorted(SHAPES), autouse=True)
def shape(request):
    global T, D, Q, HYPER, SUBJECTS
    T, D, Q, HYPER, SUBJECTS = SHAPES[request.param]
    return request.param


def make_model():
    m = nmgp_dsvi.NMGP(T * D, D, torch.linspace(0, T - 1, Q, dtype=torch.float64).view(-1, 1), mu_v=np.ones(Q), seed=22,
                       device=DEV, noise="device")
    for k, v in HYPER.items():
        getattr(m, k).data.fill_(v)
    return m


def rows(rank, world):
    r = parallel.shard_rows_per_output([T] * D, rank, world)
    gid = parallel.global_row_ids([T] * D, r)
    g = torch.Generator().manual_seed(0)
    nsub = S if SUBJECTS else 1
    Y = torch.randn(nsub, D * T, generator=g, dtype=torch.float64)
    x = torch.from_numpy((gid % T).astype(np.float64))
    I = torch.from_numpy((gid // T).astype(np.int32))
    y = Y[:, torch.from_numpy(gid)]
    return x.to(DEV), (y if SUBJECTS else y[0]).contiguous().to(DEV), I.to(DEV), torch.from_numpy(gid).to(DEV)


def step(model, world=1, params=None, want_grads=True):
    p = {k: getattr(model, k).detach() for k in dsvi_step.PARAM_NAMES}
    if params is not None:
        p.update(params)
    zv, zell, _, key = model._device_noise(1, S)           # only the key/step matters below
    model._noise_step -= 1
    tot, grads = 0.0, None
    for rank in range(world):
        x, y, I, gid = rows(rank, world)
        model._noise_step = 7
        zv, zell, _, key = model._device_noise(x.shape[0], S, gid)
        loss, g = dsvi_step.dsvi_step(p, model.Z.reshape(-1), x, y, I, T * D, zv, zell, None, B_total=T * D,
                                      kl_weight=1.0 / world, kl_shard=(rank, world) if world > 1 else None,
                                      noise_key=key, row_gid=gid, want_grads=want_grads)
        tot += float(loss)
        if want_grads:
            grads = g if grads is None else {k: grads[k] + g[k] for k in g}
    return tot, grads


def test_sharded_step_sums_to_full_step_at_full_size():
    m = make_model()
    l1, g1 = step(m, 1)
    l4, g4 = step(m, 4)
    assert np.isfinite(l1)
    assert abs(l1 - l4) <= 1e-11 * abs(l1)
    for k in g1:
        den = max(float(torch.linalg.norm(g1[k])), 1e-300)
        assert float(torch.linalg.norm(g1[k] - g4[k])) / den <= 1e-9, k
    # quirk q8: blocks j > i of mu_U / sqrt_U never receive gradient
    iu = torch.triu_indices(D, D, offset=1)
    assert float(g1["mu_U"][iu[0], iu[1]].abs().max()) == 0.0
    assert float(g1["sqrt_U"][iu[0], iu[1]].abs().max()) == 0.0


def test_gradient_matches_finite_differences_at_full_size():
    m = make_model()
    _, g = step(m, 1)
    gen = torch.Generator().manual_seed(1)
    for names, h in ((("sigma2_err_log", "sigma2_tildeell_log", "sigma2_L0_log", "sigma2_L1_log"), 1e-5),
                     (("mu_v", "mu_W"), 1e-5), (("sqrt_v", "sqrt_W"), 1e-5), (("mu_U", "sqrt_U"), 1e-5)):
        dirs = {k: torch.randn(getattr(m, k).shape, generator=gen, dtype=torch.float64).to(DEV) for k in names}
        base = {k: getattr(m, k).detach() for k in names}
        lp, _ = step(m, 1, {k: base[k] + h * dirs[k] for k in names}, want_grads=False)
        lm, _ = step(m, 1, {k: base[k] - h * dirs[k] for k in names}, want_grads=False)
        fd = (lp - lm) / (2 * h)
        an = sum(float((g[k].reshape(-1) * dirs[k].reshape(-1)).sum()) for k in names)
        assert abs(fd - an) <= 2e-5 * max(abs(an), abs(fd)), (names, fd, an)
