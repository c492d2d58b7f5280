"""Drop-in for code/SIM_code/Utility/kronecker_operation.py on the GPU.

The structured operations never need an eigen-decomposition of the T x T factor K: with B = V diag(lam) V^T
(D x D, Jacobi in one CTA),  sigma2 I + B (x) K = (V (x) I) blkdiag_m(sigma2 I + lam_m K) (V^T (x) I), so the
log-determinant and solves reduce to D independent T x T blocked Cholesky factorisations whose trailing updates run
on the FP64 tensor cores.  The reference's route is two torch.symeig calls
(kronecker_operation.py:45-47, 66-67); results agree to rounding times the conditioning of the blocks.

The D factorisations are independent: ``block_pipeline`` deals them round-robin to a few CUDA streams (each with its own
look-ahead scratch slot in the library) so that the latency-bound panel phase of one block overlaps the tensor-core
trailing updates of the others, reads the eigenvalue of each block on the device (no host round trip per block) and
defers the positive-definiteness check to one flag per block."""
import torch

from . import _ops as ops

import os

# concurrent eigen-block factorisations (streams / library scratch slots / T x T buffers); the library has 8 slots
NSLOT = max(1, min(8, int(os.environ.get("NMGP_KRON_SLOTS", "4"))))
_streams = {}


def _slot_streams(dev):
    key = str(dev)
    if key not in _streams:
        _streams[key] = [torch.cuda.Stream(device=dev) for _ in range(NSLOT)]
    return _streams[key]


def kronecker_product(t1, t2):
    """kronecker_operation.py:5-22."""
    return ops.kron_product(t1.contiguous(), t2.contiguous())


def kronecker_product_diag(d1, d2):
    """kronecker_operation.py:25-33."""
    return ops.kron_product(d1.contiguous().view(-1, 1), d2.contiguous().view(-1, 1)).view(-1)


def _dev_scalar(v, dev):
    return torch.as_tensor(v, dtype=torch.float64).detach().reshape(1).to(dev)


def block_pipeline(sigma2, B, K, Rt=None, shard=None, per_block=None, want_alpha=True):
    """Factorises A_m = sigma2 I + lam_m K for the eigen-blocks m of B (all, or the shard (rank, world): m = rank,
    rank + world, ...) -- NSLOT at a time on NSLOT streams -- and per block computes hld_m = 1/2 logdet A_m and, when
    ``Rt`` [D, T] is given, alpha_m = A_m^-1 Rt[m] and quad_m = Rt[m] . alpha_m.  ``per_block(m, L_m, slot)`` (optional)
    runs on the block's stream right after its factorisation (the adjoint uses it).  Nothing is read back to the host.
    With ``want_alpha=False`` (values only) the quadratic form comes out of the factorisation itself: the block is
    factorised as the augmented (T+1) x (T+1) system [[A_m, r_m], [r_m^T, c]] whose factor carries (L^-1 r_m)^T in its last
    row -- no triangular-solve launches at all.
    Returns dict(lam, V, hld [D], quad [D], alpha [D, T] or None, info [D] int32, blocks)."""
    D, T = B.shape[0], K.shape[0]
    dev = K.device
    lam, V = ops.eigh_small(B.detach().contiguous())
    Kc = K.detach().contiguous()
    s2 = _dev_scalar(sigma2, dev)
    first, step = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
    blocks = list(range(first, D, step))
    hld = torch.zeros(D, dtype=torch.float64, device=dev)
    quad = torch.zeros(D, dtype=torch.float64, device=dev)
    info = torch.zeros(D, dtype=torch.int32, device=dev)
    augmented = Rt is not None and not want_alpha and per_block is None
    alpha = torch.zeros(D, T, dtype=torch.float64, device=dev) if (Rt is not None and not augmented) else None
    on_gpu = Kc.is_cuda
    nslot = min(NSLOT, max(len(blocks), 1))
    # several blocks in flight hide the panel latency of each other, so wider panels (fewer, larger trailing GEMMs) pay
    # earlier than for a single factorisation: measured 1047 -> 990 ms at T=8192, D=128 (profiles/README.md)
    panel = 512 if (T >= 6144 and nslot > 1) else 0
    if augmented:
        lda = T + 8 - (T % 8) if T % 8 else T + 8                    # even, 64-byte aligned rows: 16-byte vector paths stay on
        bufs = [torch.empty(T + 1, lda, dtype=torch.float64, device=dev) for _ in range(nslot)]
        rn2 = (Rt * Rt).sum(1).contiguous()                          # |r_m|^2 per block (device)
    else:
        bufs = [torch.empty_like(Kc) for _ in range(nslot)]
    streams = _slot_streams(dev)[:nslot] if on_gpu else [None] * nslot
    main = torch.cuda.current_stream(dev) if on_gpu else None
    for st in streams:
        if st is not None:
            st.wait_stream(main)
    # several factorisations in flight: the library stages its GEMM operands with cp.async meanwhile (the TMA-fed kernel was
    # not bit-reproducible with four blocks in flight at T = 12288, profiles/README.md); also covers per_block callbacks
    concurrent = on_gpu and nslot > 1
    if concurrent:
        ops.gemm_concurrent_mode(True)
    try:
        _run_blocks(blocks, nslot, streams, on_gpu, augmented, bufs, Kc, Rt, lam, s2, rn2 if augmented else None, info, hld,
                    quad, alpha, panel, per_block, T)
    finally:
        if concurrent:
            ops.gemm_concurrent_mode(False)
    for st in streams:
        if st is not None:
            main.wait_stream(st)
    return dict(lam=lam, V=V, hld=hld, quad=quad, alpha=alpha, info=info, blocks=blocks, bufs=bufs)


def _run_blocks(blocks, nslot, streams, on_gpu, augmented, bufs, Kc, Rt, lam, s2, rn2, info, hld, quad, alpha, panel,
                per_block, T):
    for idx, m in enumerate(blocks):
        slot = idx % nslot
        ctx = torch.cuda.stream(streams[slot]) if on_gpu else _Null()
        with ctx:
            if augmented:
                Aug = bufs[slot][:, :T + 1]
                ops.build_augmented(Kc, Rt[m].contiguous(), lam[m:m + 1], s2, rn2[m:m + 1], out=bufs[slot])
                _, h = ops.potrf_big(Aug, info=info[m:m + 1], slot=slot, panel=panel)
                ops.augmented_results(bufs[slot], T, h, hld[m:m + 1], quad[m:m + 1])
                continue
            A = ops.scale_add_diag_dev(Kc, lam[m:m + 1], s2, out=bufs[slot])
            L, h = ops.potrf_big(A, info=info[m:m + 1], slot=slot, panel=panel)
            hld[m:m + 1] = h
            if Rt is not None:
                rm = Rt[m].contiguous()
                xm = ops.potrs_vec(L, rm)
                alpha[m] = xm
                quad[m:m + 1] = ops.dot(rm, xm)
            if per_block is not None:
                per_block(m, L, slot)


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def nan_if_not_pd(value, info):
    """The reference's eigen route yields NaN when sigma2 + lam w <= 0 somewhere (log of a non-positive number); the
    callers in logpos.py retry with the jittered log-density then (logpos.py:267-268).  Same contract here, on the
    device: NaN if any block failed to factorise."""
    bad = (info != 0).any()
    return torch.where(bad, torch.full_like(value, float("nan")), value)


def _factor_blocks(sigma2, B, K, shard=None):
    """Yields (m, lam_m, L_m, half_logdet_m, V) with L_m = chol(sigma2 I + lam_m K), one block at a time, every factor in
    its own buffer (for callers that keep the factors: prediction.py).  shard = (rank, world): only the eigen-blocks
    m = rank, rank + world, ... (the blocks are independent)."""
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
    res = block_pipeline(sigma2, B, K)
    return nan_if_not_pd(2.0 * res["hld"].sum(), res["info"]).reshape(())


def chol_inverse(L):
    """(L L^T)^-1 from the lower Cholesky factor L (T x T) with tensor-core GEMMs only: U = L^-T is built block column
    by block column (128-wide: the diagonal blocks are inverted in one CTA each, the rest is two GEMMs per block), then
    A^-1 = U U^T.  Used by the adjoints of the Cholesky-based log-densities."""
    T = L.shape[0]
    nb = 128
    U = torch.zeros_like(L)                                  # upper triangular: U[0:i, i] blocks + diagonal blocks
    Dinv = torch.empty(nb, nb, dtype=L.dtype, device=L.device)
    for k0 in range(0, T, nb):
        h = min(nb, T - k0)
        Lkk = L[k0:k0 + h, k0:k0 + h]
        ops.tri_inv_block(Lkk, Dinv[:h, :h])                 # Dinv = inv(L_kk) (lower)
        U[k0:k0 + h, k0:k0 + h] = Dinv[:h, :h].t()
        if k0 > 0:
            # W^T = U[0:k0, 0:k0] L[k, 0:k0]^T  (k0 x h);   U[0:k0, k] = -W^T Dinv^T
            Wt = ops.gemm_nt(U[0:k0, 0:k0], L[k0:k0 + h, 0:k0])
            ops.gemm_nt(Wt, Dinv[:h, :h], alpha=-1.0, beta=0.0, C=U[0:k0, k0:k0 + h])
    return ops.gemm_nt(U, U)


def kron_inv(sigma2, B, K):
    """kronecker_operation.py:36-54: dense (sigma2 I + B (x) K)^-1 (only sensible for small D*T, as in the reference):
    sum_m (v_m v_m^T) (x) A_m^-1 with A_m^-1 from the block's Cholesky factor."""
    T = K.shape[0]
    D = B.shape[0]
    out = torch.zeros(D * T, D * T, dtype=torch.float64, device=K.device)
    for m, lam_m, L, hld, V in _factor_blocks(sigma2, B, K):
        Minv = chol_inverse(L)
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
