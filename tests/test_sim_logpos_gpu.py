"""SIM_code log-posteriors on the GPU through the C ABI: values of every variant and the gradients of the Kronecker
family against the reference's golden values / autograd gradients (tolerances in tests/sim_logpos_cases.py: 1e-9)."""
import numpy as np
import pytest
import torch

from tests import golden_util as gu
from tests import sim_logpos_cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_objective_gradients_match_reference_autograd_gpu():
    print(sim_logpos_cases.check_objective_gradients(DEV))


def test_logpos_values_match_reference_gpu():
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import logpos
    g = gu.load("sim_logpos")
    d = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.float64)).to(DEV)
    sc = lambda v: torch.tensor(float(v), dtype=torch.float64, device=DEV)
    hyp = [sc(v) for v in g["hyp"]]
    a, b, c = (float(v) for v in g["abc"])
    ts2 = sc(g["ts2"])

    def close(got, ref, tol, what):
        got = np.array([float(v) for v in got]) if isinstance(got, (tuple, list)) else np.asarray(float(got))
        ref = np.asarray(ref, dtype=np.float64)
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        assert np.all(err < tol), (what, got, ref, err)
    TOL = sim_logpos_cases.VTOL
    close(logpos.logpos(d("tilde_l"), d("tilde_sigma"), d("uL_vec"), ts2, d("Y"), d("x"), *hyp, a, b, c, verbose=True),
          g["logpos_verbose"], TOL, "logpos")
    close(logpos.logpos_S(sc(g["tlS"]), sc(g["tsS"]), d("uL_vec"), ts2, d("Y"), d("x"), sc(-1.0), sc(0.7), a, b, c, verbose=True),
          g["logpos_S_verbose"], TOL, "logpos_S")
    ih = torch.from_numpy(g["ih"]).to(DEV)
    close(logpos.logpos_hadamard(d("tlh"), d("tsh"), d("L_vec"), ts2, d("xh"), ih, d("yh"), *hyp, a, b, c, verbose=True),
          g["logpos_hadamard_verbose"], TOL, "logpos_hadamard")
    close(logpos.logpos_hadamard_S(sc(g["tlS"]), sc(g["tsS"]), d("L_vec"), ts2, d("xh"), ih, d("yh"), sc(-1.0), sc(0.7), a, b, c,
                                   verbose=True), g["logpos_hadamard_S_verbose"], TOL, "logpos_hadamard_S")
    hyp_i = [sc(v) for v in g["hyp_i"]]
    close(logpos.logpos_SVC(d("tli"), d("uLi"), ts2, d("Yi"), d("xi"), *hyp_i, a, b, verbose=True), g["logpos_SVC_verbose"], TOL,
          "logpos_SVC")
    close(logpos.logpos_hadamard_SVC(d("tlh"), d("Lv_h"), ts2, d("xh"), ih, d("yh"), *hyp_i, a, b, verbose=True),
          g["logpos_hadamard_SVC_verbose"], TOL, "logpos_hadamard_SVC")
