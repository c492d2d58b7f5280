"""Drop-in for code/SIM_code/Utility/kernels.py: covariance builds of the exact/Kronecker line on the GPU.
CUDA float64 tensors in, CUDA float64 tensors out.  ``Nonstationary_RBF_cov`` is differentiable w.r.t. its per-point
``sigma`` / ``ell`` arguments (hand-written adjoint kernel), which is what the reference's autograd provides when the
log-posteriors of logpos.py are used as optimisation objectives."""
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
    return ops.sim_rbf_cov(X1, X1 if self_cov else _c(X2), float(alpha), float(beta),
                           settings.jitter if self_cov else 0.0, self_cov=self_cov)


class _NonstatCov(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X1, sigma1, ell1, X2, sigma2, ell2, self_cov):
        d = lambda t: None if t is None else _c(t.detach())
        X1d, X2d = d(X1), (d(X1) if self_cov else d(X2))
        s1, l1 = d(sigma1), d(ell1)
        s2, l2 = (s1, l1) if self_cov else (d(sigma2), d(ell2))
        ctx.args = (X1d, s1, l1, X2d, s2, l2)
        ctx.self_cov = self_cov
        return ops.nonstationary_cov(X1d, s1, l1, X2d, s2, l2, settings.jitter if self_cov else 0.0, self_cov=self_cov)

    @staticmethod
    def backward(ctx, Kbar):
        X1d, s1, l1, X2d, s2, l2 = ctx.args
        need = ctx.needs_input_grad
        if ctx.self_cov:
            want = (need[1], need[2], need[1], need[2])
        else:
            want = (need[1], need[2], need[4], need[5])
        g1, gl1, g2, gl2 = ops.nonstationary_cov_bwd(X1d, s1, l1, X2d, s2, l2, _c(Kbar), want)
        if ctx.self_cov:
            gs = None if g1 is None else ops.axpby(g1, g2, 1.0, 1.0)
            gl = None if gl1 is None else ops.axpby(gl1, gl2, 1.0, 1.0)
            return None, gs, gl, None, None, None, None
        return None, g1, gl1, None, g2, gl2, None


def Nonstationary_RBF_cov(X1, sigma1=None, ell1=None, X2=None, sigma2=None, ell2=None):
    """kernels.py:46-73: sigma_i sigma_j sqrt(2 l_i l_j/(l_i^2+l_j^2)) exp(-dist/(l_i^2+l_j^2)), + 1e-6 I on the
    self-covariance.  Missing sigma/ell default to ones as in the reference."""
    self_cov = X2 is None
    return _NonstatCov.apply(X1, sigma1, ell1, None if self_cov else X2, sigma2, ell2, self_cov)
