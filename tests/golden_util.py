"""Helpers to load tests/golden/*.npz (written by oracle/gen_golden.py)."""
import os
import zlib

import numpy as np
import torch

from oracle import nmgp_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DSVI_CASES = ("dsvi_sim_low", "dsvi_sim_high", "dsvi_sim_varying", "dsvi_ragged", "dsvi_ecog_like", "dsvi_pm25_like",
              "dsvi_hcp_like")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def case_params(g):
    """Parameters of a golden DSVI case: stored in full, or re-created from the seed
    and pinned by the stored (sum, sum of squares) checksums."""
    D, Q = int(g["D"]), int(g["Q"])
    if int(g["store_params"]):
        return {k: torch.from_numpy(g["param_" + k]).to(torch.float64) for k in orc.PARAM_NAMES}
    init = {k[5:]: g[k] for k in g if k.startswith("init_")}
    p = orc.init_params(D, Q, seed=int(g["seed"]), **init)
    for k in orc.PARAM_NAMES:
        if ("param_" + k) in g:
            p[k] = torch.from_numpy(g["param_" + k]).to(torch.float64).reshape(p[k].shape)
        s = g["paramsum_" + k]
        a = p[k].numpy()
        assert np.allclose([a.sum(), (a ** 2).sum()], s, rtol=1e-13, atol=1e-13), k
    return p


def case_lists(g):
    I = g["I"]
    D = int(g["D"])
    x = torch.from_numpy(g["x"]); y = torch.from_numpy(g["y"])
    Xl = [x[torch.from_numpy(I == d)].view(-1, 1) for d in range(D)]
    Yl = [y[torch.from_numpy(I == d)].view(-1, 1) for d in range(D)]
    return Xl, Yl


def case_targets(g):
    """y [B], or ys [S, B] when every forward of the case has its own targets (subjects, dsvi_hcp_like)."""
    return g["ys"] if "ys" in g else g["y"]


def case_target_lists(g):
    """None, or one outputs_list per forward for the oracle (cases with per-forward targets)."""
    if "ys" not in g:
        return None
    I = g["I"]; D = int(g["D"])
    return [[torch.from_numpy(ys[I == d]).view(-1, 1) for d in range(D)] for ys in g["ys"]]


def replay_draws(g):
    """One ReplayDraw per recorded forward, expanding the gathered z_L back to the
    per-pair (B,) vectors the oracle consumes (entries the reference discards are 0)."""
    D = int(g["D"]); I = g["I"]
    draws = []
    for s in range(int(g["n_forward"])):
        seq = [torch.from_numpy(g["z_v"][s]), torch.from_numpy(g["z_ell"][s])]
        zL = g["z_L"][s]
        for i in range(D):
            for j in range(i + 1):
                z = np.zeros(I.shape[0])
                z[I == i] = zL[I == i, j]
                seq.append(torch.from_numpy(z))
        draws.append(orc.ReplayDraw(seq))
    return draws


def check_grad(name, got, g, rtol):
    """Norm-wise comparison of one gradient tensor against the golden record (full or summarised)."""
    key = "grad_" + name
    got = np.asarray(got, dtype=np.float64)
    if key in g:
        ref = g[key]
        denom = max(np.linalg.norm(ref), 1e-300)
        err = np.linalg.norm(got.reshape(ref.shape) - ref) / denom
        assert err <= rtol, (name, err)
        return err
    flat = got.reshape(-1)
    nrm = float(g[key + "__norm"])
    rng = np.random.default_rng(zlib.crc32(key.encode()))
    proj = np.array([flat @ rng.standard_normal(flat.size) for _ in range(4)])
    idx = rng.choice(flat.size, size=512, replace=False)
    assert np.array_equal(idx, g[key + "__idx"])
    e1 = abs(np.linalg.norm(flat) - nrm) / nrm
    e2 = np.max(np.abs(proj - g[key + "__proj"])) / nrm
    e3 = np.linalg.norm(flat[idx] - g[key + "__val"]) / max(np.linalg.norm(g[key + "__val"]), 1e-300)
    assert max(e1, e2 / 10, e3) <= rtol, (name, e1, e2, e3)
    return max(e1, e3)
