"""Log joint posteriors of code/SIM_code/Utility/logpos.py (Kronecker NMGP, its stationary variant, and the Hadamard /
irregular-observation variants), the deviance, and the index helpers next to them.

The Kronecker family -- ``logpos`` / ``nlogpos_obj``, ``logpos_S`` / ``nlogpos_obj_S``, ``deviance`` / ``deviance_obj`` --
is DIFFERENTIABLE w.r.t. the parameter vector, like the reference's (its drivers hand these objectives to gradient-based
optimisers / HMC through autograd): hand-written adjoints of the Gibbs build (nmgp_nonstationary_cov_bwd), of the
eigen-block Kronecker log-density (Cholesky-based, distributions._KronLogpdf0) and of the fixed-covariance GP priors.
The dense (Hadamard / spatially-varying-coregionalisation) variants return forward values.  Every heavy step runs on
the C-ABI kernels: covariance builds, the eigen-block Kronecker log-density (no symeig of K_x), dense blocked Cholesky
in place of torch.inverse + torch.logdet and of MultivariateNormal's own factorisation.
"""
import math

import torch

from . import _ops as ops
from . import distributions, kernels, kronecker_operation
from .prediction import uLvec2Lvec, uLvecs2Lvecs, vec2lowtriangle


# ---- parameter-vector helpers (logpos.py:17-72): slicing only ---------------------------------------------------------
def vec2pars(pars, N, M):
    P = M * (M + 1) // 2
    return pars[:N], pars[N:2 * N], pars[2 * N:2 * N + P], pars[-1]


def vec2pars_SVC(pars, N, M):
    P = M * (M + 1) // 2
    return pars[:N], pars[N:N + N * P], pars[-1]


def vec2pars_S(pars, M):
    P = M * (M + 1) // 2
    return pars[0], pars[1], pars[2:2 + P], pars[-1]


def vec2pars_hadamard_SVC(pars, N, M):
    return vec2pars_SVC(pars, N, M)


def generate_vectorized_indexes(indx1, indx2):
    """logpos.py:74-84."""
    N1, N2 = indx1.shape[0], indx2.shape[0]
    return indx1.view(-1, 1).repeat(1, N2).view(-1).long(), indx2.repeat(N1).long()


def generate_K_index(B_f, indx):
    """logpos.py:87-98: K_i[n, n'] = B_f[indx_n, indx_n'] (a gather)."""
    i = indx.long()
    return B_f[i.view(-1, 1), i.view(1, -1)]


# ---- shared pieces -----------------------------------------------------------------------------------------------------
class _MVNLogprobFixedCov(torch.autograd.Function):
    """log N(value | mean 1, Sigma) for a covariance that carries no gradient (the GP priors of logpos.py:271-283 use
    RBF_cov with fixed hyper-parameters): d/d value = -Sigma^-1 (value - mean), from the same Cholesky solve."""

    @staticmethod
    def forward(ctx, value, mean, Sigma):
        r = (value.detach() - mean).contiguous()
        S = Sigma.detach()
        L, hld = ops.potrf_big(S.clone())
        # one step of FP64 iterative refinement: these prior covariances are RBF + 1e-6 I (condition number ~1e8), a plain
        # Cholesky solve carries ~cond * eps = 1e-8, above the 1e-9 this path is compared at
        a = ops.potrs_vec(L, r)
        res = ops.axpby(r, ops.gemm_nt(a.view(1, -1), S.contiguous()).view(-1), 1.0, -1.0)       # r - Sigma a
        a = ops.axpby(a, ops.potrs_vec(L, res), 1.0, 1.0)
        ctx.save_for_backward(a)
        quad = ops.dot(r, a).reshape(())
        return -0.5 * quad - hld.reshape(()) - 0.5 * r.numel() * math.log(2.0 * math.pi)

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        return -g * a, None, None


def _mvn_logprob(value, mean, Sigma):
    """torch.distributions.MultivariateNormal(mean 1, covariance_matrix=Sigma).log_prob(value) through one blocked Cholesky."""
    return _MVNLogprobFixedCov.apply(value, mean, Sigma)


class _GemmNT(torch.autograd.Function):
    """C = A B^T on the tensor-core GEMM with its adjoint (Abar = Cbar B, Bbar = Cbar^T A)."""

    @staticmethod
    def forward(ctx, A, Bm):
        ctx.save_for_backward(A.detach(), Bm.detach())
        return ops.gemm_nt(A.detach().contiguous(), Bm.detach().contiguous())

    @staticmethod
    def backward(ctx, g):
        A, Bm = ctx.saved_tensors
        g = g.contiguous()
        return ops.gemm_nt(g, Bm.t().contiguous()), ops.gemm_nt(g.t().contiguous(), A.t().contiguous())


class _NormalLogprobSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, loc, s_eff, shift):
        vc = v.detach().contiguous().view(-1)
        ctx.save_for_backward(vc)
        ctx.loc, ctx.var, ctx.shape = loc, s_eff * s_eff, v.shape
        base = ops.normal_logprob_sum(torch.full_like(vc, loc), torch.full((1,), s_eff, dtype=vc.dtype, device=vc.device),
                                      vc).reshape(())
        return base + shift

    @staticmethod
    def backward(ctx, g):
        (vc,) = ctx.saved_tensors
        return (ops.axpby(vc, torch.full_like(vc, ctx.loc), -1.0 / ctx.var, 1.0 / ctx.var) * g).reshape(ctx.shape), None, None, None


def _normal_logprob_sum(v, loc, scale):
    """sum torch.distributions.Normal(loc, scale).log_prob(v), including its dtype behaviour: python numbers become
    tensors of the default dtype (float32), so Normal(0, c) carries float32 roundings of c^2 and log(c) -- reproduced."""
    def as_param(p_, like):
        if torch.is_tensor(p_):
            return p_.detach().cpu()
        return torch.tensor(float(p_), dtype=like.dtype if like is not None else torch.get_default_dtype())
    first = next((t for t in (loc, scale) if torch.is_tensor(t)), None)
    loc_t, scale_t = as_param(loc, first), as_param(scale, first)
    var = float(scale_t ** 2)
    log_scale = float(scale_t.log())
    s_eff = math.sqrt(var)
    return _NormalLogprobSum.apply(v, float(loc_t), s_eff, v.numel() * (math.log(s_eff) - log_scale))


def _coregionalisation(L_vec, M):
    L = vec2lowtriangle(L_vec, M).contiguous()
    return _GemmNT.apply(L, L)


class _DenseIndexedLoglik(torch.autograd.Function):
    """-1/2 logdet S - 1/2 y^T S^-1 y for S = A o Bt[i1, i2] + sigma2 I (the torch.inverse + torch.logdet likelihoods of
    logpos.py:350-352, 521-526, 611-616, 690-694) through one blocked Cholesky, with the adjoint the reference gets from
    autograd: dloglik/dS = 1/2 (alpha alpha^T - S^-1), alpha = S^-1 y; S^-1 from the factor by tensor-core GEMMs
    (kronecker_operation.chol_inverse), the chain to A, the indexed table Bt and sigma2 in one kernel."""

    @staticmethod
    def forward(ctx, A, Bt, i1, i2, sigma2, y):
        Ad, Btd = A.detach().contiguous(), Bt.detach().contiguous()
        L, hld = ops.potrf_big(ops.hadamard_index_cov(Ad, Btd, i1, i2, float(sigma2)))
        yc = y.detach().contiguous()
        alpha = ops.potrs_vec(L, yc)
        ctx.save_for_backward(Ad, Btd, L, alpha)
        ctx.idx = (i1, i2)
        return (-hld - 0.5 * ops.dot(yc, alpha)).reshape(())

    @staticmethod
    def backward(ctx, g):
        Ad, Btd, L, alpha = ctx.saved_tensors
        i1, i2 = ctx.idx
        Sinv = kronecker_operation.chol_inverse(L)
        Abar, Btbar, s2bar = ops.dense_loglik_bwd(Sinv, alpha, Ad, Btd, i1, i2, g.detach().reshape(1).contiguous())
        return Abar, Btbar, None, None, s2bar.reshape(()), None


def _dense_loglik(K_x, B_f, indx, sigma2_err, y):
    """-1/2 logdet(S) - 1/2 y^T S^-1 y for S = K_x * K_i + sigma2 I (logpos.py:521-526), dense Cholesky."""
    ii = indx.to(torch.int32).contiguous()
    return _DenseIndexedLoglik.apply(K_x, B_f, ii, ii, sigma2_err, y)


class _StationaryRBF(torch.autograd.Function):
    """kernels.RBF_cov(x, alpha, beta) for scalar tensors alpha, beta that carry gradient (logpos_hadamard_S,
    logpos.py:685): alpha^2 exp(-|x_i - x_j|^2 / (2 beta^2)) is the Gibbs kernel with constant sigma = alpha, ell = beta,
    so the adjoint is the per-point Gibbs adjoint summed over the points."""

    @staticmethod
    def forward(ctx, xc, alpha, beta):
        ctx.save_for_backward(xc, alpha.detach(), beta.detach())
        return kernels.RBF_cov(xc, alpha=float(alpha), beta=float(beta))

    @staticmethod
    def backward(ctx, Kbar):
        xc, alpha, beta = ctx.saved_tensors
        one = torch.ones(xc.shape[0], dtype=torch.float64, device=xc.device)
        s, l = (alpha * one).contiguous(), (beta * one).contiguous()
        g1, gl1, g2, gl2 = ops.nonstationary_cov_bwd(xc, s, l, xc, s, l, Kbar.contiguous())
        return None, ops.dot(one, ops.axpby(g1, g2, 1.0, 1.0)).reshape(alpha.shape), \
            ops.dot(one, ops.axpby(gl1, gl2, 1.0, 1.0)).reshape(beta.shape)


# ---- deviance (logpos.py:176-213) --------------------------------------------------------------------------------------
def deviance(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, Y, x):
    """-2 log-likelihood of the Kronecker model (same value as the reference's dense-inverse route)."""
    N, M = Y.shape
    y = Y.t().contiguous().view(-1)
    B_f = _coregionalisation(L_vec, M)
    K_x = kernels.Nonstationary_RBF_cov(x.view(-1, 1), sigma1=torch.exp(tilde_sigma), ell1=torch.exp(tilde_l))
    return -2.0 * distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), B_f, K_x, torch.exp(tilde_sigma2_err))


def deviance_obj(pars, Y, x):
    N, M = Y.shape
    return deviance(*vec2pars(pars, N, M), Y, x)


# ---- Kronecker NMGP posterior (logpos.py:216-296) ----------------------------------------------------------------------
def logpos(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma,
           alpha_tilde_sigma, beta_tilde_sigma, a, b, c, verbose=False, Prior=True):
    N, M = Y.shape
    y = Y.t().contiguous().view(-1)
    B_f = _coregionalisation(uLvec2Lvec(uL_vec, M), M)
    sigma2_err = torch.exp(tilde_sigma2_err)
    xc = x.contiguous().view(-1, 1)
    K_x = kernels.Nonstationary_RBF_cov(xc, sigma1=torch.exp(tilde_sigma), ell1=torch.exp(tilde_l))
    loglik = distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), B_f, K_x, sigma2_err)
    while bool(loglik != loglik):                         # the reference's NaN retry with the jittered variant
        loglik = distributions.multivariate_normal_logpdf1(y, torch.zeros_like(y), B_f, K_x, sigma2_err)
    lp_l = _mvn_logprob(tilde_l, float(mu_tilde_l), kernels.RBF_cov(xc, alpha=float(alpha_tilde_l), beta=float(beta_tilde_l)))
    lp_s = _mvn_logprob(tilde_sigma, float(mu_tilde_sigma),
                        kernels.RBF_cov(xc, alpha=float(alpha_tilde_sigma), beta=float(beta_tilde_sigma)))
    lp_L = _normal_logprob_sum(uL_vec, 0.0, c)
    lp_e = distributions.inverse_gamma_logpdf(sigma2_err, alpha=a, beta=b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_s + lp_L + lp_e + tilde_sigma2_err
    return (res, loglik, lp_l, lp_s, lp_L, lp_e) if verbose else res


def nlogpos_obj(pars, Y, x, mu_tilde_l=0., alpha_tilde_l=1., beta_tilde_l=1., mu_tilde_sigma=0., alpha_tilde_sigma=1.,
                beta_tilde_sigma=1., a=1, b=1, c=10, verbose=False, Prior=True):
    N, M = Y.shape
    out = logpos(*vec2pars(pars, N, M), Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
                 beta_tilde_sigma, a, b, c, verbose, Prior)
    return (-out[0],) + tuple(out[1:]) if verbose else -out


# ---- stationary variant (logpos.py:383-462) ----------------------------------------------------------------------------
def logpos_S(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, sigma_tilde_l, a, b, c, verbose=False,
             Prior=True):
    N, M = Y.shape
    y = Y.t().contiguous().view(-1)
    B_f = _coregionalisation(uLvec2Lvec(uL_vec, M), M)
    sigma2_err = torch.exp(tilde_sigma2_err)
    one = torch.ones(N, dtype=torch.float64, device=Y.device)
    K_x = kernels.Nonstationary_RBF_cov(x.contiguous().view(-1, 1), sigma1=torch.exp(tilde_sigma * one),
                                        ell1=torch.exp(tilde_l * one))
    loglik = distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), B_f, K_x, sigma2_err)
    while bool(loglik != loglik):
        loglik = distributions.multivariate_normal_logpdf1(y, torch.zeros_like(y), B_f, K_x, sigma2_err)
    lp_l = _normal_logprob_sum(tilde_l.reshape(1), mu_tilde_l, sigma_tilde_l)
    lp_L = _normal_logprob_sum(uL_vec, 0.0, c)
    lp_e = distributions.inverse_gamma_logpdf(sigma2_err, alpha=a, beta=b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_L + lp_e + tilde_sigma2_err
    return (res, loglik, lp_l, lp_L, lp_e) if verbose else res


def nlogpos_obj_S(pars, Y, x, mu_tilde_l, sigma_tilde_l, a=1, b=1, c=10, verbose=False, Prior=True):
    M = Y.shape[1]
    out = logpos_S(*vec2pars_S(pars, M), Y, x, mu_tilde_l, sigma_tilde_l, a, b, c, verbose, Prior)
    return (-out[0],) + tuple(out[1:]) if verbose else -out


# ---- Hadamard (irregular observations) variants (logpos.py:465-563, 662-716) ------------------------------------------
def logpos_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                    mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, a, b, c, verbose=False, Prior=True):
    M = int(torch.unique(indx).numel())
    B_f = _coregionalisation(L_vec, M)
    sigma2_err = torch.exp(tilde_sigma2_err)
    xc = x.contiguous().view(-1, 1)
    K_x = kernels.Nonstationary_RBF_cov(xc, sigma1=torch.exp(tilde_sigma), ell1=torch.exp(tilde_l))
    loglik = _dense_loglik(K_x, B_f, indx, sigma2_err, y)
    lp_l = _mvn_logprob(tilde_l, float(mu_tilde_l), kernels.RBF_cov(xc, alpha=float(alpha_tilde_l), beta=float(beta_tilde_l)))
    lp_s = _mvn_logprob(tilde_sigma, float(mu_tilde_sigma),
                        kernels.RBF_cov(xc, alpha=float(alpha_tilde_sigma), beta=float(beta_tilde_sigma)))
    lp_L = _normal_logprob_sum(L_vec, 0.0, c)
    lp_e = distributions.inverse_gamma_logpdf_u(sigma2_err, alpha=a, beta=b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_s + lp_L + lp_e + tilde_sigma2_err
    return (res, loglik, lp_l, lp_s, lp_L, lp_e) if verbose else res


def nlogpos_obj_hadamard(pars, x, indx, y, mu_tilde_l=0., alpha_tilde_l=1., beta_tilde_l=1., mu_tilde_sigma=0.,
                         alpha_tilde_sigma=1., beta_tilde_sigma=1., a=1, b=1, c=10, verbose=False, Prior=True):
    N = y.shape[0]
    M = int(torch.unique(indx).numel())
    out = logpos_hadamard(*vec2pars(pars, N, M), x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma,
                          alpha_tilde_sigma, beta_tilde_sigma, a, b, c, verbose, Prior)
    return (-out[0],) + tuple(out[1:]) if verbose else -out


def logpos_hadamard_S(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, mu_tilde_l, sigma_tilde_l, a, b, c,
                      verbose=False, Prior=True):
    M = int(torch.unique(indx).numel())
    B_f = _coregionalisation(L_vec, M)
    sigma2_err = torch.exp(tilde_sigma2_err)
    K_x = _StationaryRBF.apply(x.contiguous().view(-1, 1), torch.exp(tilde_sigma), torch.exp(tilde_l))
    loglik = _dense_loglik(K_x, B_f, indx, sigma2_err, y)
    lp_l = _normal_logprob_sum(tilde_l.reshape(1), mu_tilde_l, sigma_tilde_l)
    lp_L = _normal_logprob_sum(L_vec, 0.0, c)
    lp_e = distributions.inverse_gamma_logpdf_u(sigma2_err, alpha=a, beta=b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_L + lp_e + tilde_sigma2_err
    return (res, loglik, lp_l, lp_L, lp_e) if verbose else res


def nlogpos_obj_hadamard_S(pars, x, indx, y, mu_tilde_l, sigma_tilde_l, a=1, b=1, c=10, verbose=False, Prior=True):
    M = int(torch.unique(indx).numel())
    out = logpos_hadamard_S(*vec2pars_S(pars, M), x, indx, y, mu_tilde_l, sigma_tilde_l, a, b, c, verbose, Prior)
    return (-out[0],) + tuple(out[1:]) if verbose else -out


# ---- spatially varying coregionalisation posteriors (logpos.py:299-380, 566-660) ---------------------------------------
def generate_K_index_SVC(L_f_list):
    """logpos.py:111-118 (row-indexed (n, m) order, as the reference)."""
    L = torch.cat(list(L_f_list), dim=0).contiguous()
    return ops.gemm_nt(L, L)


def generate_K_index_SVC_hadamard0(L_f_list, indexes):
    """logpos.py:121-124."""
    L = torch.stack([L_f[int(i), :] for L_f, i in zip(L_f_list, indexes)]).contiguous()
    return ops.gemm_nt(L, L)


def generate_K_index_SVC_hadamard(L_f_list, indexes):
    """logpos.py:127-137: the same N x N table as generate_K_index_SVC_hadamard0 (the reference builds it entry by entry)."""
    return generate_K_index_SVC_hadamard0(L_f_list, indexes)


def show_covs(pars, Y, x):
    """logpos.py:140-157: prints the coregionalisation matrix, the time kernel and the noise variance."""
    N, M = Y.shape
    tilde_l, tilde_sigma, L_vec, tilde_sigma2_err = vec2pars(pars, N, M)
    print("B_f: {}".format(_coregionalisation(L_vec, M)))
    print("K_x: {}".format(kernels.Nonstationary_RBF_cov(x.view(-1, 1), sigma1=torch.exp(tilde_sigma), ell1=torch.exp(tilde_l))))
    print("sigma2_err: {}".format(torch.exp(tilde_sigma2_err)))


def show_covs_hadamard(pars, x, indx):
    """logpos.py:160-173."""
    N = x.shape[0]
    M = int(torch.unique(indx).numel())
    _, _, L_vec, tilde_sigma2_err = vec2pars(pars, N, M)
    print("B_f: {}".format(_coregionalisation(L_vec, M)))
    print("sigma2_err: {}".format(torch.exp(tilde_sigma2_err)))


def _triangles(L_vecs, N, M):
    P = M * (M + 1) // 2
    Lmat = torch.zeros(N, M, M, dtype=torch.float64, device=L_vecs.device)
    idx = torch.tril_indices(M, M, device=L_vecs.device)
    Lmat[:, idx[0], idx[1]] = L_vecs.reshape(N, P)
    return Lmat


class _GPPriorColumns(torch.autograd.Function):
    """sum over the columns of V [N, P] of MultivariateNormal(mu 1, Sigma).log_prob(V[:, p]) for a fixed Sigma: one
    factorisation, P solves; d/dV[:, p] = -Sigma^-1 (V[:, p] - mu) from the same solves."""

    @staticmethod
    def forward(ctx, V, mu, Sigma):
        Sd = Sigma.detach().contiguous()
        L, hld = ops.potrf_big(Sd.clone())
        N, P = V.shape
        total = -P * (hld.reshape(()) + 0.5 * N * math.log(2.0 * math.pi))
        sols = []
        for p_ in range(P):
            r = (V.detach()[:, p_] - mu).contiguous()
            a = ops.potrs_vec(L, r)
            res = ops.axpby(r, ops.gemm_nt(a.view(1, -1), Sd).view(-1), 1.0, -1.0)       # one refinement step (cond ~1e8)
            a = ops.axpby(a, ops.potrs_vec(L, res), 1.0, 1.0)
            sols.append(a)
            total = total - 0.5 * ops.dot(r, a).reshape(())
        ctx.save_for_backward(torch.stack(sols, dim=1))
        return total

    @staticmethod
    def backward(ctx, g):
        (sol,) = ctx.saved_tensors
        return -g * sol, None, None


def _gp_prior_entries(V, mu, Sigma):
    return _GPPriorColumns.apply(V, mu, Sigma)


def logpos_SVC(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, a, b,
               verbose=False, Prior=True):
    N, M = Y.shape
    P = M * (M + 1) // 2
    y = Y.t().contiguous().view(-1)                         # output-major
    Lmat = _triangles(uLvecs2Lvecs(uL_vecs, N, M), N, M)
    Lrow = Lmat.permute(1, 0, 2).reshape(M * N, M).contiguous()                 # row (m, n) = L_n[m, :]
    nidx = torch.arange(N, dtype=torch.int32, device=Y.device).repeat(M).contiguous()
    sigma2_err = torch.exp(tilde_sigma2_err)
    xc = x.contiguous().view(-1, 1)
    K_x = kernels.Nonstationary_RBF_cov(xc, ell1=torch.exp(tilde_l))
    loglik = _DenseIndexedLoglik.apply(_GemmNT.apply(Lrow, Lrow), K_x, nidx, nidx, sigma2_err, y)
    lp_l = _mvn_logprob(tilde_l, float(mu_tilde_l), kernels.RBF_cov(xc, alpha=float(alpha_tilde_l), beta=float(beta_tilde_l)))
    lp_L = _gp_prior_entries(uL_vecs.reshape(N, P), float(mu_L), kernels.RBF_cov(xc, alpha=float(alpha_L), beta=float(beta_L)))
    lp_e = distributions.inverse_gamma_logpdf(sigma2_err, alpha=a, beta=b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_L + lp_e + tilde_sigma2_err
    return (res, loglik, lp_l, lp_L, lp_e) if verbose else res


def nlogpos_obj_SVC(pars, Y, x, mu_tilde_l=0., alpha_tilde_l=5., beta_tilde_l=1., mu_L=0., alpha_L=5., beta_L=1., a=1, b=1,
                    verbose=False, Prior=True):
    N, M = Y.shape
    out = logpos_SVC(*vec2pars_SVC(pars, N, M), Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, a, b,
                     verbose, Prior)
    return (-out[0],) + tuple(out[1:]) if verbose else -out


def logpos_hadamard_SVC(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L,
                        beta_L, a, b, verbose=False, Prior=True):
    N = y.shape[0]
    M = int(torch.unique(indx).numel())
    P = M * (M + 1) // 2
    Lmat = _triangles(L_vecs, N, M)                         # used as given (no exp of the diagonals)
    Lsel = Lmat[torch.arange(N, device=y.device), indx.long().to(y.device)].contiguous()
    ident = torch.arange(N, dtype=torch.int32, device=y.device)
    sigma2_err = torch.exp(tilde_sigma2_err)
    xc = x.contiguous().view(-1, 1)
    K_x = kernels.Nonstationary_RBF_cov(xc, ell1=torch.exp(tilde_l))
    loglik = _DenseIndexedLoglik.apply(_GemmNT.apply(Lsel, Lsel), K_x, ident, ident, sigma2_err, y)
    lp_l = _mvn_logprob(tilde_l, float(mu_tilde_l), kernels.RBF_cov(xc, alpha=float(alpha_tilde_l), beta=float(beta_tilde_l)))
    lp_L = _gp_prior_entries(L_vecs.reshape(N, P), float(mu_L), kernels.RBF_cov(xc, alpha=float(alpha_L), beta=float(beta_L)))
    lp_e = distributions.inverse_gamma_logpdf_u(sigma2_err, alpha=a, beta=b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_L + lp_e + tilde_sigma2_err
    return (res, loglik, lp_l, lp_L, lp_e) if verbose else res


def nlogpos_obj_hadamard_SVC(pars, x, indx, y, mu_tilde_l=0., alpha_tilde_l=1., beta_tilde_l=1., mu_L=0., alpha_L=1., beta_L=1.,
                             a=1, b=1, verbose=False, Prior=True):
    N = y.shape[0]
    M = int(torch.unique(indx).numel())
    out = logpos_hadamard_SVC(*vec2pars_hadamard_SVC(pars, N, M), x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L,
                              alpha_L, beta_L, a, b, verbose, Prior)
    return (-out[0],) + tuple(out[1:]) if verbose else -out

