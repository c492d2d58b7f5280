"""Drop-in for the GP primitives of the reference (code/utils.py): same names, argument order and return shapes,
CUDA float64 tensors, differentiable through hand-written adjoint kernels (torch.autograd.Function wrappers over the
C ABI).  ``NMGP.forward`` does not go through these wrappers (it runs the fused step of ``dsvi_step.py``); they exist
so that code written against ``utils.py`` keeps working, and they are tested against the same oracle.

Small O(N) glue (adding two cotangents, broadcasting a scalar) uses torch tensor ops; every O(N*M), O(M^3) or
transcendental computation is a kernel of ``libnmgp_b200.so``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.autograd import Function

from . import _ops as ops

TensorType = torch.DoubleTensor          # code/utils.py:6
tridiagonal_jitter = 1e-4                # code/utils.py:7
F64 = torch.float64
MODE_W = 0


def _flat1(t):
    return t.reshape(-1).contiguous()


def _as_dev_scalar(v, ref):
    if torch.is_tensor(v):
        return v.detach().to(ref.device, F64).reshape(())
    return torch.tensor(float(v), dtype=F64, device=ref.device)


# ------------------------------------------------------------------------------------------------------------
class _RBFBuild(Function):
    """scale2 * exp(-0.5 ((x - x2)/len)^2), code/utils.py:75-94 (1-D inputs, quirk q11)."""

    @staticmethod
    def forward(ctx, x, z, scale2, length):
        hyp = torch.stack([scale2.detach().reshape(()), length.detach().reshape(())]).contiguous()
        K = ops.rbf_build_fwd(x, z, hyp, 0, 1, 0.0)
        ctx.save_for_backward(x, z, hyp)
        return K

    @staticmethod
    def backward(ctx, Kbar):
        x, z, hyp = ctx.saved_tensors
        g = torch.zeros(2, dtype=F64, device=x.device)
        ops.rbf_build_bwd(x, z, hyp, 0, 1, Kbar.contiguous(), g)      # gradients w.r.t. the logs
        return None, None, g[0] / hyp[0], g[1] / hyp[1]


def squared_distance(X, X2):
    """code/utils.py:75-81 (explicit differences; O(N*M) host-side helper kept for API completeness)."""
    d = X.unsqueeze(1) - (X if X2 is None else X2).unsqueeze(0)
    return (d * d).sum(-1)


def squared_dist(X, X2, length_scales):
    """code/utils.py:84-88."""
    return squared_distance(X / length_scales, (X if X2 is None else X2) / length_scales)


def create_RBF(X, X2=None, scale2=1., length_scales=1.):
    """code/utils.py:91-94."""
    if X.dim() == 2 and X.shape[1] != 1:
        raise ValueError("inputs are one-dimensional on this path (inputs.view(-1,1), code/nmgp_dsvi.py:168)")
    x = _flat1(X)
    z = x if X2 is None else _flat1(X2)
    s2 = scale2 if torch.is_tensor(scale2) else _as_dev_scalar(scale2, X)
    ln = length_scales if torch.is_tensor(length_scales) else _as_dev_scalar(length_scales, X)
    return _RBFBuild.apply(x, z, s2.to(X.device), ln.to(X.device))


class _GibbsBuild(Function):
    """sqrt(2ab/(a^2+b^2)) exp(-(x-z)^2/(a^2+b^2)), code/utils.py:97-103."""

    @staticmethod
    def forward(ctx, x, z, ellx, ellz):
        K = ops.gibbs_build_fwd(x, z, ellx.reshape(1, -1).contiguous(), ellz.reshape(1, -1).contiguous(), 0.0)[0]
        ctx.save_for_backward(x, z, ellx, ellz)
        return K

    @staticmethod
    def backward(ctx, Kbar):
        x, z, ellx, ellz = ctx.saved_tensors
        ex = torch.empty(1, x.numel(), dtype=F64, device=x.device)
        ez = torch.zeros(1, z.numel(), dtype=F64, device=x.device)
        ops.gibbs_build_bwd(x, z, ellx.reshape(1, -1).contiguous(), ellz.reshape(1, -1).contiguous(),
                            Kbar.reshape(1, x.numel(), z.numel()).contiguous(), ex, ez)
        return None, None, ex[0].reshape(ellx.shape), ez[0].reshape(ellz.shape)


def create_Gibbs(X, X2, ell_X, ell_X2, scale2=1.):
    """code/utils.py:97-103."""
    K = _GibbsBuild.apply(_flat1(X), _flat1(X2), ell_X.contiguous(), ell_X2.contiguous())
    if torch.is_tensor(scale2) or scale2 != 1.:
        K = scale2 * K
    return K


# ------------------------------------------------------------------------------------------------------------
class _InducingSolve(Function):
    """P = K12 (K22 + eps I)^-1 and c = rowsum(P o K12) (code/utils.py:117-122): Cholesky + two triangular sweeps
    instead of the reference's LU solve."""

    @staticmethod
    def forward(ctx, K12, K22):
        N, M = K12.shape
        R, _ = ops.potrf(K22.reshape(1, M, M).contiguous(), tridiagonal_jitter)
        K = K12.reshape(1, N, M).contiguous()
        P, c = ops.solve_rows_fwd(K, R)
        ctx.save_for_backward(K, P, R)
        return P[0], c[0]

    @staticmethod
    def backward(ctx, Pbar, cbar):
        K, P, R = ctx.saved_tensors
        _, N, M = K.shape
        Abar = torch.zeros(1, M, M, dtype=F64, device=K.device)
        Kbar = ops.solve_rows_bwd(Pbar.reshape(1, N, M).contiguous(), cbar.reshape(1, N).contiguous(), K, P, R, Abar)
        return Kbar[0], Abar[0]


class _QuadMeans(Function):
    """m[d,n] = P[n] . mu[d],  q[d,n] = P[n] Sigma[d] P[n]^T  (code/utils.py:120-122,143-144)."""

    @staticmethod
    def forward(ctx, P, mu, Sigma):
        N, M = P.shape
        D = mu.shape[0]
        I = torch.full((N,), D - 1, dtype=torch.int32, device=P.device)
        P3 = P.reshape(1, N, M).contiguous()
        q, m = ops.quadform_fwd(P3, P3, I, Sigma.contiguous(), mu.contiguous(), D, MODE_W)
        ctx.save_for_backward(P3, mu, Sigma, I)
        return m[0].t(), q[0].t()

    @staticmethod
    def backward(ctx, mbar, qbar):
        P3, mu, Sigma, I = ctx.saved_tensors
        _, N, M = P3.shape
        D = mu.shape[0]
        qb = qbar.t().reshape(1, N, D).contiguous()
        mb = mbar.t().reshape(1, N, D).contiguous()
        Pbar, _ = ops.quadform_bwd(P3, P3, I, Sigma.contiguous(), mu.contiguous(), qb, mb, MODE_W)
        Sb = torch.zeros_like(Sigma)
        Mb = torch.zeros_like(mu)
        ops.weighted_gram(P3, P3, I, qb, mb, MODE_W, Sb, Mb)
        return Pbar[0], Mb, Sb


def _marginal_moments(K12, K22, mu, Sigma):
    P, c = _InducingSolve.apply(K12, K22)
    batched = mu.dim() == 2
    mu2 = mu if batched else mu.unsqueeze(0)
    if Sigma is None:
        Sig2 = torch.zeros(mu2.shape[0], mu2.shape[1], mu2.shape[1], dtype=F64, device=mu.device)
    else:
        Sig2 = Sigma if batched else Sigma.unsqueeze(0)
    m, q = _QuadMeans.apply(P, mu2, Sig2)
    if not batched:
        m, q = m[0], q[0]
    return m, q, c


def MGP_mu_sigma2(K12, K22, d11, mu, Sigma):
    """code/utils.py:128-146: mu_Y (...,N) and sigma2_Y (...,N)."""
    m, q, c = _marginal_moments(K12, K22, mu, Sigma)
    return m, d11 - c + q


def MGP_d(K12, K22, d11, mu, Sigma):
    """code/utils.py:106-125: sample of the marginalised element-wise GP (noise from the global CPU generator, float32
    cast to float64 -- quirk q2)."""
    m, s2 = MGP_mu_sigma2(K12, K22, d11, mu, Sigma)
    z = torch.randn(m.size()).type(TensorType).to(m.device)
    return reparameterize(m, s2, z, full_cov=False)


def MGP_mu(K12, K22, mu, device0=None):
    """code/utils.py:149-157."""
    m, _, _ = _marginal_moments(K12, K22, mu, None)
    return m


# ------------------------------------------------------------------------------------------------------------
class _ReparamDiag(Function):
    @staticmethod
    def forward(ctx, mean, var, z):
        ctx.save_for_backward(var, z)
        return ops.reparam_diag(mean.contiguous(), var.contiguous(), z.contiguous())

    @staticmethod
    def backward(ctx, g):
        var, z = ctx.saved_tensors
        sd = torch.sqrt(var + tridiagonal_jitter)
        return g, g * z / (2.0 * sd), g * sd


class _ReparamFull(Function):
    """mean + chol(var + eps I) z for one (N,N) covariance (code/utils.py:34-48)."""

    @staticmethod
    def forward(ctx, mean, var, z):
        N = mean.shape[0]
        C, _ = ops.potrf(var.reshape(1, N, N).contiguous(), tridiagonal_jitter)
        f, _ = ops.sample_v_fwd(mean.contiguous(), C[0], z.reshape(1, N).contiguous())
        ctx.save_for_backward(C, z)
        return f[0]

    @staticmethod
    def backward(ctx, g):
        C, z = ctx.saved_tensors
        N = z.numel()
        mub = torch.zeros(N, dtype=F64, device=g.device)
        Cb = torch.zeros(N, N, dtype=F64, device=g.device)
        zero = torch.zeros(1, N, dtype=F64, device=g.device)
        ops.sample_v_bwd(zero, g.reshape(1, N).contiguous(), zero, z.reshape(1, N).contiguous(), mub, Cb)
        Sb = ops.potrf_bwd(C, Cb.reshape(1, N, N), torch.zeros(1, dtype=F64, device=g.device))
        return mub, Sb[0], None


def reparameterize(mean, var, z, full_cov=False, use_std=False):
    """code/utils.py:15-65."""
    if var is None:
        return mean
    if full_cov is False:
        return _ReparamDiag.apply(mean, var, z)
    if use_std:
        return mean + torch.matmul(var, z.unsqueeze(-1))[..., 0]
    if mean.dim() == 1:
        return _ReparamFull.apply(mean, var, z)
    lead = mean.shape[:-1]
    n = mean.shape[-1]
    outs = [_ReparamFull.apply(m, v, zz) for m, v, zz in zip(mean.reshape(-1, n), var.reshape(-1, n, n), z.reshape(-1, n))]
    return torch.stack(outs).reshape(*lead, n)


def mat2ltri(X):
    """code/utils.py:68-72: copy with the strict upper triangle zeroed."""
    return torch.tril(X)


def JGP_S(K11_diag, K12, K22, mu, Sigma):
    """code/utils.py:216-237: joint sample (f(X), u), f(X)_i independent given u; returns cat([f, u])."""
    z_v = torch.randn(mu.size()).type(TensorType).to(mu.device)
    sampled_v = reparameterize(mu, Sigma, z_v, full_cov=True)
    m, _, c = _marginal_moments(K12, K22, sampled_v, None)
    z = torch.randn(m.size()).type(TensorType).to(mu.device)
    f = reparameterize(m, K11_diag - c, z, full_cov=False)
    return torch.cat([f, sampled_v])


# ------------------------------------------------------------------------------------------------------------
class _NormalLogprob(Function):
    @staticmethod
    def forward(ctx, loc, scale, y):
        ctx.save_for_backward(loc, scale, y)
        return ops.normal_logprob_sum(loc.reshape(-1).contiguous(), scale.reshape(1).contiguous(),
                                      y.reshape(-1).contiguous()).reshape(())

    @staticmethod
    def backward(ctx, g):
        loc, scale, y = ctx.saved_tensors
        r = y - loc
        gl = g * r / scale ** 2
        gs = g * ((r * r).sum() / scale ** 3 - r.numel() / scale)
        return gl, gs.reshape(scale.shape), -gl


def Normal_logprob(loc, scale, y):
    """code/utils.py:268-272."""
    return _NormalLogprob.apply(loc, scale, y)


class _HalfLogdet(Function):
    @staticmethod
    def forward(ctx, K):
        M = K.shape[-1]
        C, hld = ops.potrf(K.reshape(-1, M, M).contiguous(), 0.0)
        ctx.save_for_backward(C)
        ctx.shape = K.shape
        return hld.reshape(K.shape[:-2])

    @staticmethod
    def backward(ctx, g):
        (C,) = ctx.saved_tensors
        Ab = ops.potrf_bwd(C, torch.zeros_like(C), g.reshape(-1).contiguous())
        return Ab.reshape(ctx.shape)


def log_determinant_halfpower(K):
    """code/utils.py:275-277: sum(log(diag(chol(K))))."""
    return _HalfLogdet.apply(K)


class _SumSq(Function):
    @staticmethod
    def forward(ctx, x2d):
        ctx.save_for_backward(x2d)
        return ops.sumsq_rows(x2d.contiguous())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return 2.0 * g.unsqueeze(-1) * x


def batch_trace_XXT(bmat):
    """code/utils.py:280-287."""
    n, m = bmat.size(-1), bmat.size(-2)
    return _SumSq.apply(bmat.reshape(-1, m * n)).reshape(bmat.shape[:-2])


class _Mahalanobis(Function):
    """x^T (L L^T)^-1 x per row of x for one lower factor L (code/utils.py:290-329, upper=False at :321)."""

    @staticmethod
    def forward(ctx, L, x2d):
        M = L.shape[-1]
        R = torch.tril(L).reshape(1, M, M).contiguous()
        K = x2d.reshape(1, -1, M).contiguous()
        P, c = ops.solve_rows_fwd(K, R)
        ctx.save_for_backward(K, P, R)
        return c[0]

    @staticmethod
    def backward(ctx, g):
        K, P, R = ctx.saved_tensors
        _, N, M = K.shape
        Abar = torch.zeros(1, M, M, dtype=F64, device=K.device)
        Kbar = ops.solve_rows_bwd(torch.zeros_like(K), g.reshape(1, N).contiguous(), K, P, R, Abar)
        Lbar = ops.tril_syrk_bwd(R, Abar)                 # A = L L^T  ->  Lbar = tril((Abar + Abar^T) L)
        return Lbar[0], Kbar[0]


def batch_mahalanobis(bL, bx):
    """code/utils.py:290-329 for a single factor (bL of shape (n,n) or (1,n,n)), bx of shape (...,n) -- the only form the
    model uses."""
    n = bx.size(-1)
    if bL.numel() != n * n:
        raise NotImplementedError("batched factors are not used on the reference's hot path")
    return _Mahalanobis.apply(bL.reshape(n, n), bx.reshape(-1, n)).reshape(bx.shape[:-1])


class _KLGaussian(Function):
    @staticmethod
    def forward(ctx, X_mu, X_Sigma, X2_mu, X2_Sigma, exact=False):
        M = X_mu.shape[-1]
        CS, hS = ops.potrf(X_Sigma.reshape(-1, M, M).contiguous(), tridiagonal_jitter)
        R, hR = ops.potrf(X2_Sigma.reshape(1, M, M).contiguous(), tridiagonal_jitter)
        delta = (X2_mu.reshape(1, M) - X_mu.reshape(-1, M)).contiguous()
        kl, t = ops.kl_fwd(CS, hS, delta, R, hR, exact=exact)
        ctx.save_for_backward(CS, delta, R)
        ctx.kl_saved, ctx.exact = t, exact                # opaque state of the KL kernels (tensors not tracked by autograd)
        ctx.shapes = (X_mu.shape, X_Sigma.shape, X2_mu.shape, X2_Sigma.shape)
        return kl[0].reshape(X_mu.shape[:-1])

    @staticmethod
    def backward(ctx, g):
        CS, delta, R = ctx.saved_tensors
        nb = CS.shape[0]
        CSb, hSb, db, Rb, hRb = ops.kl_bwd(g.reshape(1, nb).contiguous(), CS, delta, R, ctx.kl_saved, exact=ctx.exact)
        SigX = ops.potrf_bwd(CS, CSb, hSb)
        SigX2 = ops.potrf_bwd(R, Rb, hRb)
        s = ctx.shapes
        return (-db).reshape(s[0]), SigX.reshape(s[1]), db.sum(0).reshape(s[2]), SigX2.reshape(s[3]), None


def KL_Gaussian(X_mu, X_Sigma, X2_mu, X2_Sigma, device0=None, exact=False):
    """code/utils.py:332-351, including the diagonal-only trace term produced by ``triangular_solve(..., upper=True)``
    on a lower factor (quirk q10): every ELBO the reference has printed contains it.  ``exact=True`` (not in the
    reference) evaluates the mathematically correct KL(N(X_mu, X_Sigma + eps I) || N(X2_mu, X2_Sigma + eps I))."""
    return _KLGaussian.apply(X_mu, X_Sigma, X2_mu, X2_Sigma, bool(exact))
