"""Golden vectors of the HCP-shaped step (BASELINE.json config 3): D=15 outputs, Q=100 inducing points, driver
hyper-parameters (code/NMGP_HCP.py:61-62: length scales e^5), two SUBJECTS observed on the same rows of the T=1200 grid
(different targets and different noise per forward); loss / gradients of the mean over the two reference forwards.

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference (read-only at /root/reference) under the shim of
oracle/gen_golden.py, in the build container:

    python oracle/gen_golden_hcp.py
"""
import numpy as np
import torch

from gen_golden import dsvi_case, install_shim


def main():
    install_shim()
    import nmgp_dsvi
    torch.set_num_threads(8)
    Tn, D, Q, B = 1200, 15, 100, 600
    rng = np.random.default_rng(21)
    grid = np.arange(Tn, dtype=np.float64)
    pick = np.sort(rng.choice(Tn * D, size=B, replace=False))
    Xs = [grid[pick[(pick // Tn) == d] % Tn] for d in range(D)]
    Y_lists = []
    for subj in range(2):
        Yfull = np.stack([np.sin(grid / 60.0 * (1 + 0.05 * d) + subj) + 0.3 * rng.standard_normal(Tn) for d in range(D)])
        Y_lists.append([Yfull[d][(pick[(pick // Tn) == d] % Tn)] for d in range(D)])
    hyper = {"length_scales_L0_log": 5., "length_scales_L1_log": 5., "length_scales_tildeell_log": 5.}
    dsvi_case(nmgp_dsvi, "dsvi_hcp_like", Xs, Y_lists[0], np.linspace(0, Tn - 1, Q), 2 * Tn * D, hyper, seed=22,
              init={"mu_v": np.ones(Q)}, store_params=False, n_forward=2, Y_lists=Y_lists)


if __name__ == "__main__":
    main()
