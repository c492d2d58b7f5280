"""CPU oracle for the NMGP DSVI hot path -- TEST INFRASTRUCTURE ONLY.

This file is a float64 torch-CPU *restatement* of the reference algorithm
(Corleno/Collaborative_Nonstationary_Multivariate_Gaussian_Process).  It is the
checker for the CUDA path, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under
``collaborative_nonstationary_multivariate_gaussian_process_b200/`` imports it.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real reference
(shimmed for torch>=2: ``torch.solve``/``torch.symeig``/matplotlib stubs, see
SURVEY.md 8c) in the build container, runs it on seeded inputs with recorded
noise, and stores inputs/outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function here against those files.

The statement deliberately keeps the reference's *operation order* (one LU solve
per marginal call, the D(D+1)/2 pair loop, four Cholesky calls per KL, the
``upper=True`` triangular-solve quirk) so that (a) results agree with the
reference to rounding and (b) timing it is a fair "port" CPU baseline.

Gradients come from torch autograd exactly as in the reference
(``loss.backward`` at code/nmgp_dsvi.py:847).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

F64 = torch.float64
JITTER = 1e-4          # code/utils.py:7  (tridiagonal_jitter)
SIM_JITTER = 1e-6      # code/SIM_code/Utility/settings.py:3

PARAM_NAMES = (
    "mu_W", "sqrt_W", "mu_v", "sqrt_v", "mu_U", "sqrt_U",
    "sigma2_tildeell_log", "length_scales_tildeell_log",
    "sigma2_L0_log", "length_scales_L0_log",
    "sigma2_L1_log", "length_scales_L1_log",
    "sigma2_err_log",
)


# --------------------------------------------------------------------------
# noise sources
# --------------------------------------------------------------------------
def reference_draw(shape) -> torch.Tensor:
    """float32 standard normals cast to float64 from the global CPU generator
    (code/utils.py:123,226,234 -- quirk q2)."""
    return torch.randn(shape).to(F64)


class ReplayDraw:
    """Feeds pre-recorded noise tensors back in call order."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        self._t = [torch.as_tensor(t, dtype=F64) for t in tensors]
        self._i = 0

    def __call__(self, shape):
        t = self._t[self._i]
        self._i += 1
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t


class RecordDraw:
    def __init__(self, inner: Callable = reference_draw):
        self.inner = inner
        self.log: List[torch.Tensor] = []

    def __call__(self, shape):
        t = self.inner(shape)
        self.log.append(t.clone())
        return t


# --------------------------------------------------------------------------
# primitives (code/utils.py)
# --------------------------------------------------------------------------
def lower_part(S: torch.Tensor) -> torch.Tensor:
    """code/utils.py:68-72 mat2ltri: copy with the strict upper triangle zeroed."""
    out = S.clone()
    r, c = np.triu_indices(S.shape[-2], k=1, m=S.shape[-1])
    out[..., r, c] = 0
    return out


def stationary_rbf(X, X2=None, scale2=1.0, length_scales=1.0):
    """code/utils.py:75-94: explicit-difference squared distance on inputs divided
    by the length-scale first, then scale2*exp(-r2/2)."""
    A = X / length_scales
    Bm = A if X2 is None else X2 / length_scales
    diff = A.unsqueeze(1) - Bm.unsqueeze(0)
    r2 = (diff * diff).sum(-1)
    return scale2 * torch.exp(-0.5 * r2)


def gibbs_kernel(X, X2, ell_X, ell_X2, scale2=1.0):
    """code/utils.py:97-103."""
    diff = X.unsqueeze(1) - X2.unsqueeze(0)
    r2 = (diff * diff).sum(-1)
    denom = (ell_X ** 2).unsqueeze(1) + (ell_X2 ** 2).unsqueeze(0)
    pref = torch.sqrt(2 * (ell_X.unsqueeze(1) * ell_X2.unsqueeze(0)) / denom)
    return scale2 * pref * torch.exp(-r2 / denom)


def reparam(mean, var, z, full_cov=False, use_std=False):
    """code/utils.py:15-65."""
    if var is None:
        return mean
    if not full_cov:
        return mean + z * (var + JITTER) ** 0.5
    n = mean.shape[-1]
    if use_std:
        chol = var
    else:
        chol = torch.linalg.cholesky(var + JITTER * torch.eye(n, dtype=F64))
    return mean + torch.matmul(chol, z.unsqueeze(-1))[..., 0]


def _through_inducing(K12, K22):
    """P = K12 (K22 + eps I)^-1 through an LU solve (code/utils.py:117-119)."""
    A = K22 + torch.eye(K22.shape[0], dtype=F64) * JITTER
    return torch.linalg.solve(A, K12.t()).t()


def marginal_sample(K12, K22, d11, mu, Sigma, draw=reference_draw):
    """code/utils.py:106-125 MGP_d."""
    P = _through_inducing(K12, K22)
    m = torch.matmul(P, mu.unsqueeze(-1))[..., 0]
    s2 = d11 - (P * K12).sum(-1) + (P.matmul(Sigma) * P).sum(-1)
    z = draw(m.shape)
    return reparam(m, s2, z)


def marginal_stats(K12, K22, d11, mu, Sigma):
    """code/utils.py:128-146 MGP_mu_sigma2."""
    P = _through_inducing(K12, K22)
    m = torch.matmul(P, mu.unsqueeze(-1))[..., 0]
    s2 = d11 - (P * K12).sum(-1) + (P.matmul(Sigma) * P).sum(-1)
    return m, s2


def marginal_mean(K12, K22, mu):
    """code/utils.py:149-157 MGP_mu."""
    P = _through_inducing(K12, K22)
    return torch.matmul(P, mu.unsqueeze(-1))[..., 0]


def joint_sample(K11_diag, K12, K22, mu, Sigma, draw=reference_draw):
    """code/utils.py:216-237 JGP_S: (f(X), u) with f(X)_i independent given u."""
    z_u = draw(mu.shape)
    u = reparam(mu, Sigma, z_u, full_cov=True)
    P = _through_inducing(K12, K22)
    m = torch.matmul(P, u.unsqueeze(-1))[..., 0]
    s2 = K11_diag - (P * K12).sum(1)
    z = draw(m.shape)
    f = reparam(m, s2, z)
    return torch.cat([f, u])


def gauss_logprob_sum(loc, scale, y):
    """code/utils.py:268-272 Normal_logprob."""
    var = scale ** 2
    return (-((y - loc) ** 2) / (2 * var) - torch.log(scale)
            - math.log(math.sqrt(2 * math.pi))).sum()


def half_logdet(K):
    """code/utils.py:275-277."""
    return torch.linalg.cholesky(K).diagonal(dim1=-2, dim2=-1).log().sum(-1)


def frob2(bmat):
    """code/utils.py:280-287 batch_trace_XXT."""
    return bmat.reshape(*bmat.shape[:-2], -1).pow(2).sum(-1)


def mahalanobis(L, bx):
    """code/utils.py:290-329 for a single (n,n) factor L and bx of shape (...,n):
    ||L^-1 x||^2 (explicit upper=False at :321)."""
    flat = bx.reshape(-1, bx.shape[-1])
    sol = torch.linalg.solve_triangular(L, flat.t(), upper=False)
    return sol.pow(2).sum(0).reshape(bx.shape[:-1])


def kl_gaussian(X_mu, X_Sigma, X2_mu, X2_Sigma, exact=False):
    """code/utils.py:332-351 KL_Gaussian.

    exact=False reproduces the reference bit-for-bit in exact arithmetic: the
    trace term is formed with ``triangular_solve(..., upper=True)`` on the *lower*
    prior factor, i.e. a row scaling by 1/diag (quirk q10).  exact=True is the
    true KL (behind a flag in the product as well)."""
    n = X_mu.shape[-1]
    eye = torch.eye(n, dtype=F64) * JITTER
    S1 = X_Sigma + eye
    S2 = X2_Sigma + eye
    half_term1 = half_logdet(S2) - half_logdet(S1)
    C1 = torch.linalg.cholesky(S1)
    C2 = torch.linalg.cholesky(S2)
    if exact:
        sol = torch.linalg.solve_triangular(C2, C1, upper=False)
    else:
        sol = C1 / C2.diagonal().unsqueeze(-1)
    term2 = frob2(sol)
    term3 = mahalanobis(C2, X2_mu - X_mu)
    return half_term1 + 0.5 * (term2 + term3 - n)


# --------------------------------------------------------------------------
# model (code/nmgp_dsvi.py)
# --------------------------------------------------------------------------
def init_params(D: int, Q: int, seed: int = 22, mu_v=None, mu_W=None, mu_U=None,
                sqrt_v=None, sqrt_W=None, sqrt_U=None) -> Dict[str, torch.Tensor]:
    """Parameter creation in the draw order of code/nmgp_dsvi.py:114-155."""
    torch.random.manual_seed(seed)
    s = 0.1
    p: Dict[str, torch.Tensor] = {}

    def given(a):
        return torch.from_numpy(np.asarray(a)).to(F64)
    p["mu_W"] = (0.1 * torch.randn(D, Q).to(F64)) if mu_W is None else given(mu_W)
    p["sqrt_W"] = (s * torch.randn(D, Q, Q).to(F64)) if sqrt_W is None else given(sqrt_W)
    p["mu_v"] = (-4 * torch.ones(Q, dtype=F64)) if mu_v is None else given(mu_v)
    p["sqrt_v"] = (s * torch.randn(Q, Q).to(F64)) if sqrt_v is None else given(sqrt_v)
    p["mu_U"] = (0.1 * torch.randn(D, D, Q).to(F64)) if mu_U is None else given(mu_U)
    p["sqrt_U"] = (s * torch.randn(D, D, Q, Q).to(F64)) if sqrt_U is None else given(sqrt_U)
    for name, val in (("sigma2_tildeell_log", 0.), ("length_scales_tildeell_log", -4.),
                      ("sigma2_L0_log", 0.), ("length_scales_L0_log", -4.),
                      ("sigma2_L1_log", 0.), ("length_scales_L1_log", -4.),
                      ("sigma2_err_log", -2.)):
        p[name] = torch.tensor(val, dtype=F64)
    return p


def _row_index(inputs_list, D, index=None):
    ids = range(D) if index is None else index
    I = np.hstack([np.repeat(j, x.shape[0]) for x, j in zip(inputs_list, ids)])
    return I.astype(np.int64)


def neg_selbo(p: Dict[str, torch.Tensor], Z: torch.Tensor, N: int,
              inputs_list, outputs_list, index=None, draw=reference_draw,
              exact_kl: bool = False, parts: Optional[dict] = None) -> torch.Tensor:
    """code/nmgp_dsvi.py:157-301 NMGP.forward: one reparameterised MC estimate of -ELBO."""
    D, Q = p["mu_W"].shape
    I = _row_index(inputs_list, D, index)
    rows = torch.arange(I.shape[0])
    cols = torch.from_numpy(I)
    x = torch.cat(list(inputs_list)).view(-1, 1)
    y = torch.cat(list(outputs_list)).view(-1, 1)
    B = x.shape[0]

    LW = lower_part(p["sqrt_W"]); Sigma_W = LW @ LW.transpose(-1, -2)
    Lv = lower_part(p["sqrt_v"]); Sigma_v = Lv @ Lv.t()
    LU = lower_part(p["sqrt_U"]); Sigma_U = LU @ LU.transpose(-1, -2)
    s2_ell = torch.exp(p["sigma2_tildeell_log"]); len_ell = torch.exp(p["length_scales_tildeell_log"])
    s2_L0 = torch.exp(p["sigma2_L0_log"]); len_L0 = torch.exp(p["length_scales_L0_log"])
    s2_L1 = torch.exp(p["sigma2_L1_log"]); len_L1 = torch.exp(p["length_scales_L1_log"])
    s2_err = torch.exp(p["sigma2_err_log"])

    # log length-scale GP, sampled jointly at X and Z
    d_ell = torch.ones(B, dtype=F64) * s2_ell
    K12_ell = stationary_rbf(x, Z, s2_ell, len_ell)
    K22_ell = stationary_rbf(Z, None, s2_ell, len_ell)
    joint = joint_sample(d_ell, K12_ell, K22_ell, p["mu_v"], Sigma_v, draw)
    ell_X = torch.exp(joint[:B]); ell_Z = torch.exp(joint[B:])

    # mixing coefficients: the D(D+1)/2 loop of :228-237
    d_L0 = torch.ones(B, dtype=F64) * s2_L0
    K12_L0 = stationary_rbf(x, Z, s2_L0, len_L0); K22_L0 = stationary_rbf(Z, None, s2_L0, len_L0)
    d_L1 = torch.ones(B, dtype=F64) * s2_L1
    K12_L1 = stationary_rbf(x, Z, s2_L1, len_L1); K22_L1 = stationary_rbf(Z, None, s2_L1, len_L1)
    Lfull = torch.zeros(D, D, B, dtype=F64)
    for i in range(D):
        for j in range(i + 1):
            if i == j:
                Lfull[i, j, :] = torch.exp(marginal_sample(K12_L1, K22_L1, d_L1, p["mu_U"][i, j, :],
                                                           Sigma_U[i, j, :, :], draw))
            else:
                Lfull[i, j, :] = marginal_sample(K12_L0, K22_L0, d_L0, p["mu_U"][i, j, :],
                                                 Sigma_U[i, j, :, :], draw)
    l_rows = Lfull.permute(2, 0, 1)[rows, cols]          # (B, D): row I[n] of L(x_n)

    # latent functions through the Gibbs kernel, q(W) marginalised analytically
    d_G = torch.ones(B, dtype=F64) * 1.0
    K12_G = gibbs_kernel(x, Z, ell_X, ell_Z, 1.0)
    K22_G = gibbs_kernel(Z, Z, ell_Z, ell_Z, 1.0)
    mu_g, s2_g = marginal_stats(K12_G, K22_G, d_G, p["mu_W"], Sigma_W)

    F = (l_rows * mu_g.t()).sum(1).view(-1, 1)
    R = gauss_logprob_sum(F, torch.sqrt(s2_err), y)
    R = R - 0.5 / s2_err * (l_rows ** 2 * s2_g.t()).sum()

    zeros = torch.zeros(Q, dtype=F64)
    KL_W = kl_gaussian(p["mu_W"], Sigma_W, zeros, K22_G, exact_kl).sum()
    KL_v = kl_gaussian(p["mu_v"], Sigma_v, zeros, K22_ell, exact_kl)
    mu1, S1, mu0, S0 = [], [], [], []
    for i in range(D):
        mu1.append(p["mu_U"][i, i, :]); S1.append(Sigma_U[i, i, :, :])
        if i > 0:
            mu0.append(p["mu_U"][i, :i, :].view(i, Q)); S0.append(Sigma_U[i, :i, :, :].view(i, Q, Q))
    KL_U = kl_gaussian(torch.stack(mu1), torch.stack(S1), zeros, K22_L1, exact_kl).sum()
    KL_U = KL_U + kl_gaussian(torch.cat(mu0), torch.cat(S0), zeros, K22_L0, exact_kl).sum()
    if parts is not None:
        parts.update(R=R.detach(), KL_W=KL_W.detach(), KL_v=KL_v.detach(), KL_U=KL_U.detach(),
                     ell_X=ell_X.detach(), ell_Z=ell_Z.detach(), l_rows=l_rows.detach(),
                     mu_g=mu_g.detach(), s2_g=s2_g.detach(), F=F.detach())
    return -(N / B * R - KL_W - KL_v - KL_U)


def step_loss_and_grads(p, Z, N, inputs_list, outputs_list, index=None, draws: Sequence = (reference_draw,),
                        exact_kl=False, train_lengthscales=False, outputs_lists=None):
    """S-sample estimate = mean of S consecutive reference forwards on unchanged
    parameters, plus autograd gradients for the 13 parameters.
    ``outputs_lists`` (one outputs_list per draw): every forward sees its own targets -- the subjects of an HCP-shaped
    step (BASELINE.json config 3)."""
    leaves = {}
    for k in PARAM_NAMES:
        t = p[k].detach().clone()
        is_len = k.startswith("length_scales")
        t.requires_grad_(train_lengthscales or not is_len)
        leaves[k] = t
    total = 0.0
    for s_, d in enumerate(draws):
        Yl_s = outputs_list if outputs_lists is None else outputs_lists[s_]
        total = total + neg_selbo(leaves, Z, N, inputs_list, Yl_s, index, d, exact_kl)
    loss = total / len(draws)
    loss.backward()
    grads = {k: (leaves[k].grad.clone() if leaves[k].grad is not None else None) for k in PARAM_NAMES}
    return loss.detach(), grads


def posterior_mean(p, Z, inputs_list, index=None):
    """code/nmgp_dsvi.py:666-722 predict_Y (deterministic, no noise)."""
    D, Q = p["mu_W"].shape
    I = _row_index(inputs_list, D, index)
    rows = torch.arange(I.shape[0]); cols = torch.from_numpy(I)
    x = torch.cat(list(inputs_list)).view(-1, 1)
    B = x.shape[0]
    q = {k: v.detach() for k, v in p.items()}
    s2_ell = torch.exp(q["sigma2_tildeell_log"]); len_ell = torch.exp(q["length_scales_tildeell_log"])
    s2_L0 = torch.exp(q["sigma2_L0_log"]); len_L0 = torch.exp(q["length_scales_L0_log"])
    s2_L1 = torch.exp(q["sigma2_L1_log"]); len_L1 = torch.exp(q["length_scales_L1_log"])
    K12 = stationary_rbf(x, Z, s2_ell, len_ell); K22 = stationary_rbf(Z, None, s2_ell, len_ell)
    ell_X = torch.exp(marginal_mean(K12, K22, q["mu_v"])); ell_Z = torch.exp(q["mu_v"])
    K12_L0 = stationary_rbf(x, Z, s2_L0, len_L0); K22_L0 = stationary_rbf(Z, None, s2_L0, len_L0)
    K12_L1 = stationary_rbf(x, Z, s2_L1, len_L1); K22_L1 = stationary_rbf(Z, None, s2_L1, len_L1)
    Lhat = torch.zeros(D, D, B, dtype=F64)
    for i in range(D):
        for j in range(i + 1):
            if i == j:
                Lhat[i, j, :] = torch.exp(marginal_mean(K12_L1, K22_L1, q["mu_U"][i, j, :]))
            else:
                Lhat[i, j, :] = marginal_mean(K12_L0, K22_L0, q["mu_U"][i, j, :])
    KG12 = gibbs_kernel(x, Z, ell_X, ell_Z); KG22 = gibbs_kernel(Z, Z, ell_Z, ell_Z)
    Ghat = marginal_mean(KG12, KG22, q["mu_W"])                    # (D, B)
    Yhat = torch.matmul(Lhat.permute(2, 0, 1), Ghat.t().unsqueeze(2))[:, :, 0]
    return Yhat[rows, cols]


def mc_elbo(p, Z, N, inputs_list, outputs_list, index=None, n_sample=1000, draw=reference_draw):
    """code/nmgp_dsvi.py:303-404 compute_ELBO including quirk q5: the coefficient
    gather uses the transposed layout (permute(2,1,0)), no N/B factor, KL_W from
    the last draw's Gibbs K22."""
    D, Q = p["mu_W"].shape
    q = {k: v.detach() for k, v in p.items()}
    I = _row_index(inputs_list, D, index)
    rows = torch.arange(I.shape[0]); cols = torch.from_numpy(I)
    x = torch.cat(list(inputs_list)).view(-1, 1); y = torch.cat(list(outputs_list)).view(-1, 1)
    B = x.shape[0]
    LW = lower_part(q["sqrt_W"]); Sigma_W = LW @ LW.transpose(-1, -2)
    Lv = lower_part(q["sqrt_v"]); Sigma_v = Lv @ Lv.t()
    LU = lower_part(q["sqrt_U"]); Sigma_U = LU @ LU.transpose(-1, -2)
    s2_ell = torch.exp(q["sigma2_tildeell_log"]); len_ell = torch.exp(q["length_scales_tildeell_log"])
    s2_L0 = torch.exp(q["sigma2_L0_log"]); len_L0 = torch.exp(q["length_scales_L0_log"])
    s2_L1 = torch.exp(q["sigma2_L1_log"]); len_L1 = torch.exp(q["length_scales_L1_log"])
    s2_err = torch.exp(q["sigma2_err_log"])
    vals = []
    for _ in range(n_sample):
        d_ell = torch.ones(N, dtype=F64) * s2_ell       # sized with self.N (:334); requires B == N
        K12_ell = stationary_rbf(x, Z, s2_ell, len_ell); K22_ell = stationary_rbf(Z, None, s2_ell, len_ell)
        joint = joint_sample(d_ell, K12_ell, K22_ell, q["mu_v"], Sigma_v, draw)
        ell_X = torch.exp(joint[:B]); ell_Z = torch.exp(joint[B:])
        d_L0 = torch.ones(B, dtype=F64) * s2_L0; d_L1 = torch.ones(B, dtype=F64) * s2_L1
        K12_L0 = stationary_rbf(x, Z, s2_L0, len_L0); K22_L0 = stationary_rbf(Z, None, s2_L0, len_L0)
        K12_L1 = stationary_rbf(x, Z, s2_L1, len_L1); K22_L1 = stationary_rbf(Z, None, s2_L1, len_L1)
        Lfull = torch.zeros(D, D, B, dtype=F64)
        for i in range(D):
            for j in range(i + 1):
                if i == j:
                    Lfull[i, j, :] = torch.exp(marginal_sample(K12_L1, K22_L1, d_L1, q["mu_U"][i, j, :],
                                                               Sigma_U[i, j, :, :], draw))
                else:
                    Lfull[i, j, :] = marginal_sample(K12_L0, K22_L0, d_L0, q["mu_U"][i, j, :],
                                                     Sigma_U[i, j, :, :], draw)
        l_rows = Lfull.permute(2, 1, 0)[rows, cols]
        K12_G = gibbs_kernel(x, Z, ell_X, ell_Z); K22_G = gibbs_kernel(Z, Z, ell_Z, ell_Z)
        mu_g, s2_g = marginal_stats(K12_G, K22_G, torch.ones(B, dtype=F64), q["mu_W"], Sigma_W)
        F = (l_rows * mu_g.t()).sum(1).view(-1, 1)
        R = gauss_logprob_sum(F, torch.sqrt(s2_err), y) - 0.5 / s2_err * (l_rows ** 2 * s2_g.t()).sum()
        vals.append(R)
    zeros = torch.zeros(Q, dtype=F64)
    KL_W = kl_gaussian(q["mu_W"], Sigma_W, zeros, K22_G).sum()
    KL_v = kl_gaussian(q["mu_v"], Sigma_v, zeros, K22_ell)
    mu1, S1, mu0, S0 = [], [], [], []
    for i in range(D):
        mu1.append(q["mu_U"][i, i, :]); S1.append(Sigma_U[i, i])
        if i > 0:
            mu0.append(q["mu_U"][i, :i, :].reshape(i, Q)); S0.append(Sigma_U[i, :i].reshape(i, Q, Q))
    KL_U = kl_gaussian(torch.stack(mu1), torch.stack(S1), zeros, K22_L1).sum() \
        + kl_gaussian(torch.cat(mu0), torch.cat(S0), zeros, K22_L0).sum()
    return torch.stack(vals).mean() - KL_W - KL_v - KL_U


def reference_noise_for_step(B: int, D: int, Q: int, I: np.ndarray, generator_draw=reference_draw):
    """Draw one forward's noise in the reference order (z_v(Q), z_ell(B),
    then z_ij(B) for i=0..D-1, j=0..i) and return both the replay list and the
    gathered arrays the C-ABI consumes: z_v (Q), z_ell (B), z_L (B,D) with
    z_L[n,j] = z_{I[n],j}[n] (zero for j > I[n])."""
    seq = [generator_draw((Q,)), generator_draw((B,))]
    zL = torch.zeros(B, D, dtype=F64)
    rows = torch.arange(B)
    It = torch.from_numpy(np.asarray(I, dtype=np.int64))
    for i in range(D):
        for j in range(i + 1):
            z = generator_draw((B,))
            seq.append(z)
            sel = It == i
            zL[rows[sel], j] = z[sel]
    return seq, seq[0], seq[1], zL


# --------------------------------------------------------------------------
# SIM_code line (code/SIM_code/Utility/*.py)
# --------------------------------------------------------------------------
def sq_dists_gemm(x, y=None):
    """kernels.py:5-21 pairwise_distances: ||x||^2 + ||y||^2 - 2 x y^T (not clamped)."""
    xn = (x ** 2).sum(1).view(-1, 1)
    if y is None:
        y = x
        yn = xn.view(1, -1)
    else:
        yn = (y ** 2).sum(1).view(1, -1)
    return xn + yn - 2.0 * torch.mm(x, y.t())


def sim_rbf_cov(X1, X2=None, alpha=1.0, beta=1.0):
    """kernels.py:24-43 RBF_cov (alpha is a standard deviation here)."""
    if X2 is None:
        X2 = X1
        base = torch.eye(X1.shape[0], dtype=F64) * SIM_JITTER
    else:
        base = torch.zeros(X1.shape[0], X2.shape[0], dtype=F64)
    d = sq_dists_gemm(X1 / beta, X2 / beta)
    return base + torch.exp(-0.5 * d) * alpha ** 2


def sim_nonstationary_cov(X1, sigma1=None, ell1=None, X2=None, sigma2=None, ell2=None):
    """kernels.py:46-73 Nonstationary_RBF_cov."""
    n1 = X1.shape[0]
    if sigma1 is None:
        sigma1 = torch.ones(n1, dtype=F64)
    if ell1 is None:
        ell1 = torch.ones(n1, dtype=F64)
    if X2 is None:
        X2, sigma2, ell2 = X1, sigma1, ell1
        base = torch.eye(n1, dtype=F64) * SIM_JITTER
    else:
        base = torch.zeros(n1, X2.shape[0], dtype=F64)
    d = sq_dists_gemm(X1, X2)
    A = (ell1 ** 2).view(-1, 1) + (ell2 ** 2).view(1, -1)
    Bm = ell1.view(-1, 1) * ell2.view(1, -1)
    C = sigma1.view(-1, 1) * sigma2.view(1, -1)
    return base + C * torch.sqrt(2.0 * Bm / A) * torch.exp(-d / A)


def kron_dense(t1, t2):
    """kronecker_operation.py:5-22."""
    return torch.kron(t1, t2)


def kron_diag(d1, d2):
    """kronecker_operation.py:25-33."""
    return (d1.view(-1, 1) * d2.view(1, -1)).reshape(-1)


def _eigh_U(A):
    # torch.symeig(A, eigenvectors=True) defaulted to upper=True (reads the upper triangle)
    return torch.linalg.eigh(A, UPLO="U")


def kron_inverse(sigma2, Bm, K):
    """kronecker_operation.py:36-54 kron_inv: dense (sigma2 I + B (x) K)^-1 via per-factor eigh."""
    wB, vB = _eigh_U(Bm); wK, vK = _eigh_U(K)
    U = kron_dense(vB, vK)
    t = kron_diag(wB, wK)
    return (U * (1.0 / (t + sigma2)).unsqueeze(0)) @ U.t()


def kron_logdet(sigma2, Bm, K):
    """kronecker_operation.py:57-69."""
    wB, _ = _eigh_U(Bm); wK, _ = _eigh_U(K)
    return torch.log(kron_diag(wB, wK) + sigma2).sum()


def kron_matvec(Bm, K, y):
    """kronecker_operation.py:72-85 kron_mv: (B (x) K) y with y output-major."""
    M = Bm.shape[1]; N = K.shape[1]
    Y = y.view(M, N).t()
    A = K @ Y @ Bm.t()
    return A.t().contiguous().view(-1)


def mvn_logpdf_kron_eig(y, mu, Bm, K, sigma2):
    """distributions.py:26-52 multivariate_normal_logpdf0 (unnormalised)."""
    wB, vB = _eigh_U(Bm); wK, vK = _eigh_U(K)
    a = kron_matvec(vB.t(), vK.t(), y - mu)
    t = kron_diag(wB, wK)
    return -0.5 * torch.log(t + sigma2).sum() - 0.5 * torch.dot(a * (1.0 / (sigma2 + t)), a)


def mvn_logpdf_kron_eig_jittered(y, mu, Bm, K, sigma2, precision=1e-6, rand=torch.rand):
    """distributions.py:55-96 multivariate_normal_logpdf1: same as above after adding a
    random U(0,1)*1e-6 diagonal to both factors (B first, then K -- draw order)."""
    Bm = Bm + torch.diag(rand(Bm.shape[0]).to(F64) * precision)
    K = K + torch.diag(rand(K.shape[0]).to(F64) * precision)
    return mvn_logpdf_kron_eig(y, mu, Bm, K, sigma2)


def mvn_logpdf_dense(y, mu, Bm, K, sigma2):
    """distributions.py:99-113 multivariate_normal_logpdf2 (+ :10-23, which drops the
    -N/2 log 2pi constant, quirk q7)."""
    S = kron_dense(Bm, K) + sigma2 * torch.eye(Bm.shape[0] * K.shape[0], dtype=F64)
    r = y - mu
    return -0.5 * torch.logdet(S) - 0.5 * torch.dot(r, torch.mv(torch.inverse(S), r))
