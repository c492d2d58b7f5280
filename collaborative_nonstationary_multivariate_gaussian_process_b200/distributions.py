"""Drop-in for code/SIM_code/Utility/distributions.py (un-normalised multivariate normal log-densities)."""
import numpy as np
import torch

from . import _ops as ops
from . import kronecker_operation, settings


def multivariate_normal_logpdf(y, mu, logdetSigma, invSigma):
    """distributions.py:10-23 (the -N/2 log 2pi constant is dropped there, quirk q7)."""
    r = ops.axpby(y.contiguous(), mu.contiguous(), 1.0, -1.0)
    Sr = ops.gemm_nt(invSigma.contiguous(), r.view(1, -1)).view(-1)
    return (-0.5 * logdetSigma - 0.5 * ops.dot(r, Sr)).reshape(())


def multivariate_normal_logpdf0(y, mu, B, K, sigma2, shard=None):
    """distributions.py:26-52: -1/2 logdet(S) - 1/2 r^T S^-1 r, S = B (x) K + sigma2 I, via the eigen-blocks of B and
    one blocked Cholesky per block (see kronecker_operation).  shard = (rank, world) returns this rank's partial sum
    over its eigen-blocks (parallel.kron_logpdf0_sharded adds the partials with one all-reduce)."""
    D, T = B.shape[0], K.shape[0]
    r = ops.axpby(y.contiguous(), mu.contiguous(), 1.0, -1.0)
    half_logdet, quad, Rt = None, None, None
    for m, lam_m, L, hld, V in kronecker_operation._factor_blocks(sigma2, B, K, shard):
        if Rt is None:
            # rows of Rt: (V^T (x) I) r  ->  Rt = V^T R with R = r.view(D, T)
            Rt = ops.gemm_nt(V.t().contiguous(), r.view(D, T).t().contiguous())      # [D, T]
        rm = Rt[m].contiguous()
        xm = ops.potrs_vec(L, rm)
        q = ops.dot(rm, xm)
        quad = q if quad is None else quad + q
        half_logdet = hld if half_logdet is None else half_logdet + hld
    if half_logdet is None:                       # a rank that owns no block (world > D)
        return torch.zeros((), dtype=torch.float64, device=K.device)
    return (-half_logdet - 0.5 * quad).reshape(())


def multivariate_normal_logpdf1(y, mu, B, K, sigma2):
    """distributions.py:55-96: same after adding U(0,1)*1e-6 to both diagonals (B first, then K -- the global CPU
    generator is consumed in that order)."""
    dB = (torch.rand(B.size(0)).type(settings.torchType) * settings.precision).to(B.device)
    dK = (torch.rand(K.size(0)).type(settings.torchType) * settings.precision).to(K.device)
    return multivariate_normal_logpdf0(y, mu, B + torch.diag(dB), K + torch.diag(dK), sigma2)


def multivariate_normal_logpdf2(y, mu, B, K, sigma2):
    """distributions.py:99-113: dense reference path; Cholesky replaces torch.logdet + torch.inverse."""
    S = kronecker_operation.kronecker_product(B, K)
    A = ops.scale_add_diag(S, 1.0, float(sigma2))
    L, hld = ops.potrf_big(A)
    r = ops.axpby(y.contiguous(), mu.contiguous(), 1.0, -1.0)
    x = ops.potrs_vec(L, r)
    return (-hld - 0.5 * ops.dot(r, x)).reshape(())


# scalar priors (distributions.py:116-137): plain scalar arithmetic, kept for API completeness
def inverse_gamma_logpdf_u(x, alpha=1., beta=1.):
    return (-alpha - 1) * torch.log(x) - beta / x


def inverse_gamma_logpdf(x, alpha=1., beta=1.):
    from math import lgamma
    return (-alpha - 1) * torch.log(x) - beta / x + alpha * np.log(beta) - lgamma(alpha)


def gamma_logpdf(x, alpha=1., beta=1.):
    from math import lgamma
    return (alpha - 1) * torch.log(x) - beta * x + alpha * np.log(beta) - lgamma(alpha)
