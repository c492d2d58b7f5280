"""No-gradient paths built from the same kernels: deterministic posterior mean
(``predict_Y``, code/nmgp_dsvi.py:666-722) and the Monte-Carlo ELBO (``compute_ELBO``,
code/nmgp_dsvi.py:303-404)."""
from __future__ import annotations

import numpy as np
import torch

from . import _ops as ops
from . import dsvi_step as _step
from .dsvi_step import (EPS, HYPER_ORDER, H_LEN_ELL, H_LEN_L0, H_LEN_L1, H_S2_ELL, H_S2_L0, H_S2_L1, MODE_U, MODE_W,
                        packed_pair_index)

F64 = torch.float64


def _stationary_systems(x, Z, hyp, want_c=False):
    B, Q = x.shape[0], Z.shape[0]
    out = {}
    for name, is2, ilen in (("ell", H_S2_ELL, H_LEN_ELL), ("L0", H_S2_L0, H_LEN_L0), ("L1", H_S2_L1, H_LEN_L1)):
        A = ops.rbf_build_fwd(Z, Z, hyp, is2, ilen, EPS).reshape(1, Q, Q)
        R, hldR = ops.potrf(A, 0.0)
        K12 = ops.rbf_build_fwd(x, Z, hyp, is2, ilen, 0.0).reshape(1, B, Q)
        P, c = ops.solve_rows_fwd(K12, R)
        out[name] = dict(R=R, hldR=hldR, K12=K12, P=P, c=c)
    return out


def posterior_mean(p, Z, x, I):
    """Y*[n] = sum_{j<=I[n]} Lhat[I[n],j](x_n) Ghat[j](x_n) with Lhat/Ghat the conditional means through the
    inducing points (MGP_mu, code/utils.py:149-157); only the (row, own-output) entries the reference gathers
    at :721-722 are evaluated.  x sorted by output id, I int32."""
    D, Q = p["mu_W"].shape
    B = x.shape[0]
    dev = x.device
    zeros = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
    hyp = ops.hyper_exp(torch.stack([p[k].detach().reshape(()) for k in HYPER_ORDER]))
    sysm = _stationary_systems(x, Z, hyp)
    mu_v = p["mu_v"].detach().contiguous()
    v, ellZ = ops.sample_v_fwd(mu_v, zeros(Q, Q), zeros(1, Q))               # v = mu_v, ellZ = exp(mu_v)
    ellx = ops.ell_rows_fwd(sysm["ell"]["P"][0], v, zeros(1, B), zeros(B))     # exp(P mu_v)
    flat = packed_pair_index(D, dev)
    muU = p["mu_U"].detach().reshape(D * D, Q).index_select(0, flat)
    mU = ops.pair_means(sysm["L0"]["P"], sysm["L1"]["P"], I, muU, D, MODE_U)
    lhat = ops.coef_sample_fwd(mU[0], zeros(B, D), zeros(1, B, D), I)          # exp on the diagonal entry
    A_G = ops.gibbs_build_fwd(Z, Z, ellZ, ellZ, EPS)
    R_G, _ = ops.potrf(A_G, 0.0)
    KG = ops.gibbs_build_fwd(x, Z, ellx, ellZ, 0.0)
    PG, _ = ops.solve_rows_fwd(KG, R_G)
    ghat = ops.pair_means(PG, PG, I, p["mu_W"].detach().contiguous(), D, MODE_W)
    return ops.rowdot_live(lhat, ghat, I)[0]


def mc_elbo(model, inputs_list, outputs_list, index=None, n_sample=1000, verbose=False):
    raise NotImplementedError("compute_ELBO (SURVEY.md 8f row 2) is scheduled after the training hot path")
