"""Drop-in for code/SIM_code/Utility/kernels.py: covariance builds of the exact/Kronecker line on the GPU.
CUDA float64 tensors in, CUDA float64 tensors out (forward only in this round; the reference's SIM_code ships
no optimiser that would differentiate them)."""
import torch

from . import _ops as ops
from . import settings


def _c(t):
    return t.contiguous()


def pairwise_distances(x, y=None):
    """kernels.py:5-21: dist[i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j (GEMM form, may be slightly negative)."""
    x = _c(x)
    return ops.pairwise_dist(x, x if y is None else _c(y))


def RBF_cov(X1, X2=None, alpha=1., beta=1.):
    """kernels.py:24-43: alpha^2 exp(-dist(X1/beta, X2/beta)/2), + jitter*I when X2 is None."""
    X1 = _c(X1)
    self_cov = X2 is None
    return ops.sim_rbf_cov(X1, X1 if self_cov else _c(X2), float(alpha), float(beta), settings.jitter if self_cov else 0.0)


def Nonstationary_RBF_cov(X1, sigma1=None, ell1=None, X2=None, sigma2=None, ell2=None):
    """kernels.py:46-73: sigma_i sigma_j sqrt(2 l_i l_j/(l_i^2+l_j^2)) exp(-dist/(l_i^2+l_j^2)), + 1e-6 I on the
    self-covariance.  Missing sigma/ell default to ones as in the reference."""
    X1 = _c(X1)
    opt = lambda t: None if t is None else _c(t)
    if X2 is None:
        return ops.nonstationary_cov(X1, opt(sigma1), opt(ell1), X1, opt(sigma1), opt(ell1), settings.jitter)
    return ops.nonstationary_cov(X1, opt(sigma1), opt(ell1), _c(X2), opt(sigma2), opt(ell2), 0.0)
