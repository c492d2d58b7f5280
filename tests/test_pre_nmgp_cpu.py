"""pre_nmgp (optional NumPy/SciPy initialiser) against golden values of the reference's code/pre_nmgp.py
(oracle/gen_golden_pre.py).  Host-side helper: no CUDA involved."""
import numpy as np

from tests import golden_util as gu
from collaborative_nonstationary_multivariate_gaussian_process_b200 import pre_nmgp


def test_local_likelihood_matches_reference():
    g = gu.load("pre_nmgp")
    L = np.linalg.cholesky(g["Y"].T @ g["Y"] / (g["Y"].shape[0] - 1))
    xl, Yl = pre_nmgp.search_nearest_neighhood(g["x"], g["Y"], 0.4)
    assert np.array_equal(xl, g["xl"]) and np.array_equal(Yl, g["Yl"])
    for p, ref in zip(g["pars"], g["ll_part"]):
        got = pre_nmgp.compute_loglik_part(p, xl, Yl, L)
        assert abs(got - ref) <= 1e-12 * abs(ref)            # eigen route vs the reference's dense np.kron density
    assert abs(pre_nmgp.compute_loglik(g["pf"], xl, Yl) - float(g["ll_full"])) <= 1e-12 * abs(float(g["ll_full"]))


def test_pre_estimation_partial_matches_reference():
    g = gu.load("pre_nmgp")
    v, Lt, s2 = pre_nmgp.pre_estimation_partial(g["x"], g["Y"], g["z"])
    assert np.allclose(Lt, g["L_tensor"], rtol=1e-13, atol=0)
    # BFGS stops at its gradient tolerance (1e-5): the two optimisers walk rounding-different paths to the same optimum
    assert np.allclose(v, g["v"], atol=5e-5) and np.allclose(s2, g["s2log"], atol=5e-5)
    # and the reference's optimum is a stationary point of this implementation's objective
    xl, Yl = pre_nmgp.search_nearest_neighhood(g["x"], g["Y"], g["z"][0])
    f0 = pre_nmgp.objective_part(np.array([g["s2log"][0], g["v"][0]]), xl, Yl, Lt[:, :, 0])
    for d in ([1e-3, 0], [-1e-3, 0], [0, 1e-3], [0, -1e-3]):
        assert pre_nmgp.objective_part(np.array([g["s2log"][0] + d[0], g["v"][0] + d[1]]), xl, Yl, Lt[:, :, 0]) >= f0 - 1e-7
