"""Drop-in for code/SIM_code/Utility/kronecker_operation.py on the GPU.

The structured operations never need an eigen-decomposition of the T x T factor K: with B = V diag(lam) V^T
(D x D, Jacobi in one CTA),  sigma2 I + B (x) K = (V (x) I) blkdiag_m(sigma2 I + lam_m K) (V^T (x) I), so the
log-determinant and solves reduce to D independent T x T blocked Cholesky factorisations whose trailing updates run
on the FP64 tensor cores.  The reference's route is two torch.symeig calls
(kronecker_operation.py:45-47, 66-67); results agree to rounding times the conditioning of the blocks."""
import torch

from . import _ops as ops


def kronecker_product(t1, t2):
    """kronecker_operation.py:5-22."""
    return ops.kron_product(t1.contiguous(), t2.contiguous())


def kronecker_product_diag(d1, d2):
    """kronecker_operation.py:25-33."""
    return ops.kron_product(d1.contiguous().view(-1, 1), d2.contiguous().view(-1, 1)).view(-1)


def _factor_blocks(sigma2, B, K, shard=None):
    """Yields (m, lam_m, L_m, half_logdet_m, V) with L_m = chol(sigma2 I + lam_m K); V from the Jacobi eigensolver.
    shard = (rank, world): only the eigen-blocks m = rank, rank + world, ... (the blocks are independent)."""
    lam, V = ops.eigh_small(B.contiguous())
    lam_host = lam.cpu()
    s2 = float(sigma2)
    Kc = K.contiguous()
    first, step = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
    for m in range(first, B.shape[0], step):
        A = ops.scale_add_diag(Kc, float(lam_host[m]), s2)
        L, hld = ops.potrf_big(A)
        yield m, float(lam_host[m]), L, hld, V


def kron_logdet(sigma2, B, K):
    """kronecker_operation.py:57-69: log det(sigma2 I + B (x) K)."""
    total = None
    for m, lam_m, L, hld, V in _factor_blocks(sigma2, B, K):
        total = 2.0 * hld if total is None else total + 2.0 * hld
    return total.reshape(())


def kron_inv(sigma2, B, K):
    """kronecker_operation.py:36-54: dense (sigma2 I + B (x) K)^-1 (only sensible for small D*T, as in the reference)."""
    T = K.shape[0]
    D = B.shape[0]
    out = torch.zeros(D * T, D * T, dtype=torch.float64, device=K.device)
    eye = torch.eye(T, dtype=torch.float64, device=K.device)
    for m, lam_m, L, hld, V in _factor_blocks(sigma2, B, K):
        Minv = torch.stack([ops.potrs_vec(L, eye[c].contiguous()) for c in range(T)], dim=1).contiguous()
        vm = V[:, m].contiguous()
        outer = ops.gemm_nt(vm.view(-1, 1).contiguous(), vm.view(-1, 1).contiguous())
        out = ops.axpby(out.view(-1), ops.kron_product(outer, Minv).view(-1), 1.0, 1.0).view(D * T, D * T)
    return out


def kron_mv(B, K, y):
    """kronecker_operation.py:72-85: (B (x) K) y for output-major y, as K Y B^T without forming the product:
    two tensor-core GEMMs, the second written directly in the output-major layout."""
    M = B.shape[1]
    N = K.shape[1]
    Yt = y.contiguous().view(M, N)                        # Y^T  (M x N)
    KY = ops.gemm_nt(K.contiguous(), Yt)                  # K Y        [N1, M]
    return ops.gemm_nt(B.contiguous(), KY).view(-1)       # B (K Y)^T  [M1, N1] -> flattened output-major
