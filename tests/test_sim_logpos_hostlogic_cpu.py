"""Forward values of the SIM_code log-posteriors / deviance (logpos.py) on the CPU against the reference's golden values
(C-ABI wrappers replaced by their specifications; the same kernels are exercised on the GPU by test_simcode_gpu and
test_sim_prediction_gpu)."""
import inspect

import numpy as np
import pytest
import torch

from oracle import kernel_specs as specs
from tests import golden_util as gu
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops


@pytest.fixture(autouse=True)
def spec_ops(monkeypatch):
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            monkeypatch.setattr(_ops, n, f)


def _close(got, ref, tol):
    got = np.array([float(v) for v in got]) if isinstance(got, (tuple, list)) else np.asarray(float(got))
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    assert np.all(err < tol), (got, ref, err)


def test_logpos_values_match_reference():
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import logpos
    g = gu.load("sim_logpos")
    d = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.float64))
    sc = lambda v: torch.tensor(float(v), dtype=torch.float64)
    hyp = [sc(v) for v in g["hyp"]]
    a, b, c = (float(v) for v in g["abc"])
    ts2 = sc(g["ts2"])
    TOL = 1e-9      # north_star tolerance on log-densities; measured 2e-14 with the kernel specifications
    _close(logpos.logpos(d("tilde_l"), d("tilde_sigma"), d("uL_vec"), ts2, d("Y"), d("x"), *hyp, a, b, c, verbose=True),
           g["logpos_verbose"], TOL)
    _close(logpos.logpos(d("tilde_l"), d("tilde_sigma"), d("uL_vec"), ts2, d("Y"), d("x"), *hyp, a, b, c, Prior=False),
           g["logpos_noprior"], 1e-9)
    pars = torch.cat([d("tilde_l"), d("tilde_sigma"), d("uL_vec"), ts2.view(1)])
    _close(logpos.nlogpos_obj(pars, d("Y"), d("x"), *[float(h) for h in hyp], a, b, c), g["nlogpos_obj"], TOL)
    _close(logpos.deviance(d("tilde_l"), d("tilde_sigma"), d("L_vec"), ts2, d("Y"), d("x")), g["deviance"], 1e-9)
    _close(logpos.deviance_obj(torch.cat([d("tilde_l"), d("tilde_sigma"), d("L_vec"), ts2.view(1)]), d("Y"), d("x")),
           g["deviance"], 1e-9)
    _close(logpos.logpos_S(sc(g["tlS"]), sc(g["tsS"]), d("uL_vec"), ts2, d("Y"), d("x"), sc(-1.0), sc(0.7), a, b, c, verbose=True),
           g["logpos_S_verbose"], 1e-9)
    ih = torch.from_numpy(g["ih"])
    _close(logpos.logpos_hadamard(d("tlh"), d("tsh"), d("L_vec"), ts2, d("xh"), ih, d("yh"), *hyp, a, b, c, verbose=True),
           g["logpos_hadamard_verbose"], TOL)
    _close(logpos.logpos_hadamard_S(sc(g["tlS"]), sc(g["tsS"]), d("L_vec"), ts2, d("xh"), ih, d("yh"), sc(-1.0), sc(0.7), a, b, c,
                                    verbose=True), g["logpos_hadamard_S_verbose"], 1e-9)
    parsH = torch.cat([d("tlh"), d("tsh"), d("L_vec"), ts2.view(1)])
    _close(logpos.nlogpos_obj_hadamard(parsH, d("xh"), ih, d("yh"), *[float(h) for h in hyp], a, b, c),
           g["nlogpos_obj_hadamard"], TOL)
    # spatially varying coregionalisation posteriors
    hyp_i = [sc(v) for v in g["hyp_i"]]
    _close(logpos.logpos_SVC(d("tli"), d("uLi"), ts2, d("Yi"), d("xi"), *hyp_i, a, b, verbose=True), g["logpos_SVC_verbose"], TOL)
    _close(logpos.nlogpos_obj_SVC(torch.cat([d("tli"), d("uLi"), ts2.view(1)]), d("Yi"), d("xi"), *[float(h) for h in hyp_i], a, b),
           g["nlogpos_obj_SVC"], TOL)
    _close(logpos.logpos_hadamard_SVC(d("tlh"), d("Lv_h"), ts2, d("xh"), ih, d("yh"), *hyp_i, a, b, verbose=True),
           g["logpos_hadamard_SVC_verbose"], TOL)
    Lf = [torch.tril(torch.arange(1., 5.).view(2, 2) + k) for k in range(3)]
    Kh = logpos.generate_K_index_SVC_hadamard(Lf, torch.tensor([0, 1, 1]))
    Lsel = torch.stack([Lf[0][0], Lf[1][1], Lf[2][1]])
    assert torch.allclose(Kh, Lsel @ Lsel.t(), rtol=0, atol=1e-12)
    assert logpos.generate_K_index_SVC(Lf).shape == (6, 6)
    # helpers
    i1, i2 = logpos.generate_vectorized_indexes(torch.tensor([0, 2]), torch.tensor([1, 0, 2]))
    assert i1.tolist() == [0, 0, 0, 2, 2, 2] and i2.tolist() == [1, 0, 2, 1, 0, 2]
    Bf = torch.arange(9, dtype=torch.float64).view(3, 3)
    assert torch.equal(logpos.generate_K_index(Bf, torch.tensor([2, 0])), torch.tensor([[8., 6.], [2., 0.]], dtype=torch.float64))


def test_objective_gradients_match_reference_autograd():
    """nlogpos_obj / deviance_obj / nlogpos_obj_S are differentiable drop-ins: values and gradients w.r.t. the parameter
    vector against the reference's own autograd (golden)."""
    from tests import sim_logpos_cases
    print(sim_logpos_cases.check_objective_gradients("cpu"))


def test_nan_retry_recovers_like_the_reference():
    """a block that is not positive definite must give NaN (as the reference's eigen route does), not an
    exception, so the jittered retry of logpos.py:267-268 can run."""
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import distributions
    T, D = 12, 2
    x = torch.linspace(0, 1, T, dtype=torch.float64).view(-1, 1)
    K = torch.exp(-0.5 * (x - x.t()) ** 2 / 0.09)
    B = torch.tensor([[1.0, 0.2], [0.2, 0.5]], dtype=torch.float64)
    y = torch.randn(T * D, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    bad = distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), B, K, torch.tensor(-0.5, dtype=torch.float64))
    assert bool(torch.isnan(bad))
    ok = distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), B, K, torch.tensor(0.1, dtype=torch.float64))
    assert bool(torch.isfinite(ok))
