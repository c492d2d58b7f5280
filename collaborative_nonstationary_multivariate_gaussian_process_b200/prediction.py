"""Drop-in for the MAP predictive functions of code/SIM_code/Utility/prediction.py (point_predmap 337-408,
pointwise_predmap 410-430, test_predmap 432-458) and the vector <-> lower-triangle helpers of
code/SIM_code/Utility/utils.py:10-88 they use.

The reference redoes, for EVERY test point, two T x T LU solves and torch.symeig(K_x).  Here everything that does not
depend on the test point is factorised once (`KroneckerPosterior`): the two GP-conditional systems, the eigen-blocks
sigma2 I + lambda_k K_x of sigma2 I + B_f (x) K_x (one blocked Cholesky per block, no symeig of K_x), and
alpha_k = A_k^-1 (V^T Y^T)_k.  Because k_f = B_f (x) k_x for a single test input, in the eigenbasis of B_f

    mean_m     = sum_k lambda_k V[m,k] * (k_x . alpha_k)
    k_f,m^T S^-1 k_f,m = sum_k (lambda_k V[m,k])^2 * (k_x . A_k^-1 k_x)

so one test point costs D vector solves and 2D dot products.  Results equal the reference's to rounding (tests).
"""
import numpy as np
import torch

from . import _ops as ops
from . import kernels, kronecker_operation, settings


# ---- code/SIM_code/Utility/utils.py:10-88 (index plumbing on tiny vectors) -------------------------------------------
def _diag_positions(M):
    return list(np.cumsum(np.arange(1, M + 1)) - 1)


def uLvec2Lvec(uL_vec, M):
    """utils.py:10-22: exponentiate the diagonal entries of the packed lower triangle."""
    on = _diag_positions(M)
    L_vec = uL_vec.clone()
    L_vec[on] = torch.exp(uL_vec[on])
    return L_vec


def Lvec2uLvec(L_vec, M):
    """utils.py:24-36."""
    on = _diag_positions(M)
    uL_vec = L_vec.clone()
    uL_vec[on] = torch.log(L_vec[on])
    return uL_vec


def vec2lowtriangle(x, N=None):
    """utils.py:56-74."""
    if N * (N + 1) / 2 != x.shape[0]:
        raise ValueError("check the dimension size!")
    mat = torch.zeros(N, N, dtype=torch.float64, device=x.device)      # settings.torchType = DoubleTensor
    idx = torch.tril_indices(N, N, device=x.device)
    mat[idx[0], idx[1]] = x
    return mat


def lowtriangle2vec(L, N=None):
    """utils.py:77-88."""
    idx = torch.tril_indices(N, N, device=L.device)
    return L[idx[0], idx[1]]


# ---------------------------------------------------------------------------------------------------------------------
class KroneckerPosterior:
    """Everything of point_predmap that does not depend on x_star (prediction.py:350-381)."""

    def __init__(self, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                 mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma):
        N, M = Y.shape
        self.N, self.M = N, M
        dev = Y.device
        self.x = x.contiguous().view(-1, 1)
        f = lambda v: float(v)
        self.hyp_l = (f(mu_tilde_l), f(alpha_tilde_l), f(beta_tilde_l))
        self.hyp_s = (f(mu_tilde_sigma), f(alpha_tilde_sigma), f(beta_tilde_sigma))
        # GP conditionals of log-ell and log-sigma: mu + k^T Sigma^-1 (tilde - mu); Sigma^-1 (tilde - mu) is hoisted
        self.beta = []
        for tilde, (mu, alpha, beta) in ((tilde_l, self.hyp_l), (tilde_sigma, self.hyp_s)):
            Sigma = kernels.RBF_cov(self.x, alpha=alpha, beta=beta)
            Lc, _ = ops.potrf_big(Sigma)
            self.beta.append(ops.potrs_vec(Lc, (tilde - mu).contiguous()))
        self.sigma2_err = torch.exp(tilde_sigma2_err)
        self.l = torch.exp(tilde_l).contiguous()
        self.sigma = torch.exp(tilde_sigma).contiguous()
        L = vec2lowtriangle(uLvec2Lvec(uL_vec, M), M)
        self.B_f = ops.gemm_nt(L.contiguous(), L.contiguous())
        K_x = kernels.Nonstationary_RBF_cov(self.x, sigma1=self.sigma, ell1=self.l)
        y = Y.t().contiguous().view(-1)
        self.blocks = []          # (lambda_k, chol(sigma2 I + lambda_k K_x), alpha_k)
        Rt = None
        for k, lam_k, Lk, hld, V in kronecker_operation._factor_blocks(self.sigma2_err, self.B_f, K_x):
            if Rt is None:
                self.V = V
                Rt = ops.gemm_nt(V.t().contiguous(), y.view(M, N).t().contiguous())        # rows: (V^T (x) I) y
            self.blocks.append((lam_k, Lk, ops.potrs_vec(Lk, Rt[k].contiguous())))
        self.lam = torch.tensor([b[0] for b in self.blocks], dtype=torch.float64, device=dev)

    def point(self, x_star):
        """prediction.py:353-408 for one test input: [3, M] = (mean - 1.96 sd, mean, mean + 1.96 sd)."""
        xs = x_star.reshape(1, 1).to(self.x.dtype)
        est = []
        for b, (mu, alpha, beta) in zip(self.beta, (self.hyp_l, self.hyp_s)):
            k = kernels.RBF_cov(self.x, xs, alpha=alpha, beta=beta).view(-1)
            est.append(mu + ops.dot(k.contiguous(), b).reshape(()))
        l_star, sigma_star = torch.exp(est[0]).view(1), torch.exp(est[1]).view(1)
        k_x = kernels.Nonstationary_RBF_cov(X1=self.x, sigma1=self.sigma, ell1=self.l, X2=xs, sigma2=sigma_star,
                                            ell2=l_star).view(-1).contiguous()
        k_ss = kernels.Nonstationary_RBF_cov(X1=xs, sigma1=sigma_star, ell1=l_star).view(())      # incl. the 1e-6 jitter
        dots = torch.stack([torch.stack((ops.dot(k_x, a_k).reshape(()),
                                         ops.dot(k_x, ops.potrs_vec(Lk, k_x)).reshape(())))
                            for _, Lk, a_k in self.blocks])                                        # [D, 2]
        c = self.V * self.lam.view(1, -1)                       # c[m, k] = lambda_k V[m, k]
        mu_f = c @ dots[:, 0]
        sigma2_f = torch.diagonal(self.B_f) * k_ss - (c * c) @ dots[:, 1]
        sigma2_y = sigma2_f + self.sigma2_err
        sigma2_y = torch.where(sigma2_y <= 0, torch.full_like(sigma2_y, settings.precision), sigma2_y)
        sd = torch.sqrt(sigma2_y)
        return torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd])


def point_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_star, mu_tilde_l, alpha_tilde_l,
                  beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:337-408."""
    post = KroneckerPosterior(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l,
                              beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
    return post.point(x_star)


def pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l,
                      beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:410-430: [N_grid, 3, M]; the factorisations are shared by all grid points."""
    post = KroneckerPosterior(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l,
                              beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
    return torch.stack([post.point(g) for g in grids])


def test_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                 beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:432-458: same loop over held-out inputs."""
    return pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                             beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)


test_predmap.__test__ = False      # not a pytest test
