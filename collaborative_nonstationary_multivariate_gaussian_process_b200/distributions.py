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


def _rotate(V, r, D, T):
    """rows of the result: (V^T (x) I) r  ->  Rt = V^T R with R = r.view(D, T)."""
    return ops.gemm_nt(V.t().contiguous(), r.view(D, T).t().contiguous())      # [D, T]


class _KronLogpdf0(torch.autograd.Function):
    """-1/2 sum_m [logdet A_m + r_m^T A_m^-1 r_m],  A_m = sigma2 I + lam_m K,  r_m = row m of V^T (y - mu).view(D, T).

    Adjoint (what autograd through the two symeig calls of distributions.py:37-51 gives the reference): with
    alpha_m = A_m^-1 r_m and G_m = -1/2 A_m^-1 + 1/2 alpha_m alpha_m^T,
        Kbar = sum_m lam_m G_m,   lam_bar_m = <G_m, K>,   sigma2_bar = sum_m tr G_m,   Rt_bar[m] = -alpha_m,
    B receives V (diag(lam_bar) + F o (V^T Vbar)) V^T with F_ij = 1/(lam_j - lam_i) and Vbar = R Rt_bar^T (the standard
    symmetric-eigendecomposition adjoint; like the reference's it needs distinct eigenvalues).  A_m^-1 comes from the
    block's Cholesky factor (kronecker_operation.chol_inverse: tensor-core GEMMs), block by block on the pipeline's
    streams with one K-bar accumulator per stream."""

    @staticmethod
    def forward(ctx, y, mu, B, K, sigma2, shard):
        D, T = B.shape[0], K.shape[0]
        r = ops.axpby(y.detach().contiguous(), mu.detach().contiguous(), 1.0, -1.0)
        lam, V = ops.eigh_small(B.detach().contiguous())
        Rt = _rotate(V, r, D, T)
        needs_adjoint = any(torch.is_tensor(t) and t.requires_grad for t in (y, mu, B, K, sigma2))
        res = kronecker_operation.block_pipeline(sigma2, B, K, Rt=Rt, shard=shard, want_alpha=needs_adjoint)
        val = -(res["hld"].sum()) - 0.5 * res["quad"].sum()
        s2 = torch.as_tensor(sigma2, dtype=torch.float64).detach().reshape(1).to(K.device)
        alpha = res["alpha"] if res["alpha"] is not None else torch.empty(0, dtype=torch.float64, device=K.device)
        ctx.save_for_backward(r, B.detach(), K.detach(), s2, alpha, lam, V)
        ctx.shard = shard
        ctx.s2_shape = sigma2.shape if torch.is_tensor(sigma2) else None
        return kronecker_operation.nan_if_not_pd(val, res["info"]).reshape(())

    @staticmethod
    def backward(ctx, g):
        r, B, K, s2, alpha, lam, V = ctx.saved_tensors
        D, T = B.shape[0], K.shape[0]
        dev = K.device
        need_y, need_mu, need_B, need_K, need_s2 = ctx.needs_input_grad[:5]
        gy = gmu = gB = gK = gs2 = None
        if need_B or need_K or need_s2:
            Kc = K.contiguous()
            nslot = kronecker_operation.NSLOT
            Kbar = [torch.zeros_like(Kc) for _ in range(nslot)] if need_K else None
            lam_bar = torch.zeros(D, dtype=torch.float64, device=dev)
            tr_bar = torch.zeros(D, dtype=torch.float64, device=dev)

            def per_block(m, L, slot):                       # on the block's stream, L = chol(A_m) still in its buffer
                Ainv = kronecker_operation.chol_inverse(L)
                am = alpha[m].contiguous()
                if need_B:
                    Ka = ops.gemm_nt(am.view(1, T), Kc).view(-1)                    # K alpha_m (K symmetric)
                    lam_bar[m:m + 1] = -0.5 * ops.dot(Ainv.view(-1), Kc.view(-1)) + 0.5 * ops.dot(am, Ka)
                if need_s2:
                    tr_bar[m:m + 1] = -0.5 * ops.dot(torch.diagonal(Ainv).contiguous(), torch.ones_like(am)) + 0.5 * ops.dot(am, am)
                if need_K:
                    ops.axpby_dev(Ainv.view(-1), Kbar[slot].view(-1), lam[m:m + 1], -0.5, 1.0, out=Kbar[slot].view(-1))
                    u = ops.axpby_dev(am, am, lam[m:m + 1], 1.0, 0.0)               # lam_m alpha_m
                    ops.gemm_nt(u.view(T, 1), am.view(T, 1), alpha=0.5, beta=1.0, C=Kbar[slot])
            kronecker_operation.block_pipeline(s2, B, K, Rt=None, shard=ctx.shard, per_block=per_block)
            if need_K:
                tot = Kbar[0]
                for extra in Kbar[1:]:
                    tot = ops.axpby(tot.view(-1), extra.view(-1), 1.0, 1.0).view(T, T)
                gK = tot * g
            if need_s2:
                gs2 = (tr_bar.sum() * g).reshape(ctx.s2_shape if ctx.s2_shape is not None else ())
        if need_y or need_mu or need_B:
            Rt_bar = -alpha * g                                                      # [D, T]
            if need_y or need_mu:
                gy_full = ops.gemm_nt(V.contiguous(), Rt_bar.t().contiguous()).reshape(-1)      # (V Rt_bar) flattened
                gy = gy_full if need_y else None
                gmu = -gy_full if need_mu else None
            if need_B:
                # D x D algebra of the symmetric eigendecomposition adjoint (tiny; torch on the device)
                Vbar = r.view(D, T) @ Rt_bar.t()
                dl = lam.view(1, D) - lam.view(D, 1)
                F = torch.where(dl != 0, 1.0 / torch.where(dl != 0, dl, torch.ones_like(dl)), torch.zeros_like(dl))
                inner = torch.diag(lam_bar * g) + F * (V.t() @ Vbar)
                gB = V @ inner @ V.t()
                gB = 0.5 * (gB + gB.t())
        return gy, gmu, gB, gK, gs2, None


def multivariate_normal_logpdf0(y, mu, B, K, sigma2, shard=None):
    """distributions.py:26-52: -1/2 logdet(S) - 1/2 r^T S^-1 r, S = B (x) K + sigma2 I, via the eigen-blocks of B and
    one blocked Cholesky per block (see kronecker_operation).  shard = (rank, world) returns this rank's partial sum
    over its eigen-blocks (parallel.kron_logpdf0_sharded adds the partials with one all-reduce).  Differentiable
    w.r.t. y, mu, B, K and sigma2.  NaN (not an exception) when a block is not positive definite, like the reference's
    eigen route, so the callers' jittered retry (logpos.py:267-268) works."""
    return _KronLogpdf0.apply(y, mu, B, K, sigma2, shard)


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
