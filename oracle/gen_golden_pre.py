"""Golden values of code/pre_nmgp.py (local ML initialiser) from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY:
    python oracle/gen_golden_pre.py"""
import importlib.util
import os

import numpy as np

REF = os.environ.get("NMGP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("ref_pre_nmgp", os.path.join(REF, "code", "pre_nmgp.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(3)
    x = np.sort(rng.random(60))
    Y = np.stack([np.sin(6 * x), np.cos(4 * x) + 0.3 * np.sin(6 * x), x ** 2], 1) + 0.05 * rng.standard_normal((60, 3))
    z = np.linspace(0.1, 0.9, 4)
    L = np.linalg.cholesky(Y.T @ Y / 59)
    xl, Yl = ref.search_nearest_neighhood(x, Y, 0.4)
    pars = np.array([[-6., -6.], [-3., -2.], [-1., 0.5]])
    ll_part = np.array([ref.compute_loglik_part(p, xl, Yl, L) for p in pars])
    pf = np.concatenate([[-3., -2.], L[np.tril_indices(3)]])
    v, Lt, s2 = ref.pre_estimation_partial(x, Y, z)
    np.savez_compressed(os.path.join(OUT, "pre_nmgp.npz"), x=x, Y=Y, z=z, xl=xl, Yl=Yl, pars=pars, ll_part=ll_part,
                        ll_full=ref.compute_loglik(pf, xl, Yl), pf=pf, v=v, L_tensor=Lt, s2log=s2)
    print(ll_part, v, s2)


if __name__ == "__main__":
    main()
