"""Tensor-level wrappers over the C ABI (include/nmgp_b200.h).

Each function validates its arguments (CUDA, float64/int32, contiguous), allocates
outputs with torch (device memory is torch's job; arithmetic is not) and launches the
sm_100a kernel on torch's current stream.  CPU tensors are rejected: there is no
fallback path.  ``oracle/kernel_specs.py`` holds a same-named CPU specification of
every function for the tests.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_int, c_int64, c_void_p

import torch

from ._lib import check, lib

F64 = torch.float64
LQ_MIN_Q = 65          # the register-resident DMMA kernels cover Q <= 64, the ring-pipelined / right-looking ones 65..128
MODE_W, MODE_U = 0, 1
MAX_Q = 128


def _same_device(t: torch.Tensor) -> None:
    """Kernels are enqueued on the CURRENT device's current stream (and the library-owned scratch lives there), so a
    tensor of another device would be dereferenced on the wrong GPU: refuse it.  Use ``torch.cuda.device(model.device)``
    (``nmgp_dsvi`` does) when a model lives on a GPU that is not the current one."""
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError("tensor on %s but the current CUDA device is cuda:%d; wrap the call in "
                           "torch.cuda.device(...)" % (t.device, torch.cuda.current_device()))


def _d(t: torch.Tensor) -> c_void_p:
    if not (t.is_cuda and t.dtype == F64 and t.is_contiguous()):
        raise TypeError("expected a contiguous CUDA float64 tensor, got %s %s contiguous=%s (no CPU fallback)"
                        % (t.device, t.dtype, t.is_contiguous()))
    _same_device(t)
    return c_void_p(t.data_ptr())


def _i(t: torch.Tensor) -> c_void_p:
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise TypeError("expected a contiguous CUDA int32 tensor, got %s %s" % (t.device, t.dtype))
    _same_device(t)
    return c_void_p(t.data_ptr())


def _optd(t):
    return c_void_p(0) if t is None else _d(t)


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _empty(ref, *shape):
    return torch.empty(*shape, dtype=F64, device=ref.device)


def _zeros(ref, *shape):
    return torch.zeros(*shape, dtype=F64, device=ref.device)


def _checkQ(Q):
    if Q > MAX_Q:
        raise ValueError("number of inducing points Q=%d exceeds the supported maximum %d" % (Q, MAX_Q))


# ---------------------------------------------------------------------------------------------
def hyper_exp(logs):
    out = torch.empty_like(logs)
    check(lib().nmgp_hyper_exp(_d(logs), _d(out), c_int(logs.numel()), _stream()), "nmgp_hyper_exp")
    return out


def segment_offsets(I, D):
    seg = torch.empty(D + 1, dtype=torch.int32, device=I.device)
    check(lib().nmgp_segment_offsets(_i(I), _i(seg), c_int64(I.numel()), c_int(D), _stream()), "nmgp_segment_offsets")
    return seg


def tril_syrk_fwd(S):
    nb, Q, _ = S.shape
    _checkQ(Q)
    out = torch.empty_like(S)
    check(lib().nmgp_tril_syrk_fwd(_d(S), _d(out), c_int(nb), c_int(Q), _stream()), "nmgp_tril_syrk_fwd")
    return out


def tril_syrk_bwd(S, SigBar):
    nb, Q, _ = S.shape
    out = torch.empty_like(S)
    check(lib().nmgp_tril_syrk_bwd(_d(S), _d(SigBar), _d(out), c_int(nb), c_int(Q), _stream()), "nmgp_tril_syrk_bwd")
    return out


def potrf(A, jitter=0.0, info=None):
    """Batched lower Cholesky of A + jitter*I with hld = sum(log(diag)).  Raises RuntimeError (like torch.cholesky in
    the reference) when a matrix is not positive definite.  With ``info`` (an int32 device scalar that accumulates
    1 + index of a failing matrix) the check is deferred to the caller -- no host synchronisation here."""
    nb, Q, _ = A.shape
    _checkQ(Q)
    C = torch.empty_like(A)
    hld = _empty(A, nb)
    deferred = info is not None
    if not deferred:
        info = torch.zeros(1, dtype=torch.int32, device=A.device)
    check(lib().nmgp_potrf_batched(_d(A), c_double(jitter), _d(C), _d(hld), _i(info), c_int(nb), c_int(Q), _stream()),
          "nmgp_potrf_batched")
    if not deferred:
        raise_if_not_pd(info)
    return C, hld


def raise_if_not_pd(info):
    bad = int(info.item())
    if bad != 0:
        raise RuntimeError("cholesky: matrix %d of a batch is not positive-definite" % (bad - 1))


def potrf_bwd(C, Cbar, hldbar):
    nb, Q, _ = C.shape
    out = torch.empty_like(C)
    if Q >= LQ_MIN_Q:                       # tensor-core formulation (A^T B reduction + two DMMA right solves)
        w1, w2 = torch.empty_like(C), torch.empty_like(C)
        check(lib().nmgp_potrf_bwd_batched_lq(_d(C), _d(Cbar), _d(hldbar), _d(out), _d(w1), _d(w2), c_int(nb), c_int(Q),
                                              _stream()), "nmgp_potrf_bwd_batched_lq")
        return out
    check(lib().nmgp_potrf_bwd_batched(_d(C), _d(Cbar), _d(hldbar), _d(out), c_int(nb), c_int(Q), _stream()),
          "nmgp_potrf_bwd_batched")
    return out


def atb(A, Bm, C, sign=1.0):
    """C[s] += sign * A[s]^T Bm[s]  (A, Bm [ns,B,Q], C [ns,Q,Q]) on the FP64 tensor cores."""
    ns, B, Q = A.shape
    check(lib().nmgp_atb(_d(A), _d(Bm), _d(C), c_double(sign), c_int(ns), c_int64(B), c_int(Q), _stream()), "nmgp_atb")
    return C


def kl_rbar(R, G, Rbar):
    """Rbar[p] += -tril(G[p] R[p])."""
    np_, Q, _ = R.shape
    check(lib().nmgp_kl_rbar(_d(R), _d(G), _d(Rbar), c_int(np_), c_int(Q), _stream()), "nmgp_kl_rbar")
    return Rbar


def _kl_exact_rows(CS, mu, np_):
    """[np, nb (Q+1), Q]: per b the row mu_b followed by the Q columns of CS_b (as rows), repeated for every prior p."""
    nb, Q, _ = CS.shape
    rows = torch.cat([mu.reshape(nb, 1, Q), CS.transpose(1, 2)], dim=1).reshape(1, nb * (Q + 1), Q)
    return rows.expand(np_, nb * (Q + 1), Q).contiguous()


def kl_fwd(CS, hldS, mu, R, hldR, exact=False):
    """kl[p,b] = KL(N(mu_b, CS_b CS_b^T) || N(0, R_p R_p^T)).  Default: the reference's form (quirk q10: the trace term
    only sees diag(R_p), code/utils.py:349).  ``exact=True``: the true trace term tr((R_p R_p^T)^-1 CS_b CS_b^T),
    obtained from the same DMMA row solve with the columns of CS_b as extra rows.
    Returns (kl [np,nb], saved); ``saved`` is what kl_bwd needs."""
    nb, Q, _ = CS.shape
    _checkQ(Q)
    np_ = R.shape[0]
    if exact:
        K = _kl_exact_rows(CS, mu, np_)
        P, c = solve_rows_fwd(K, R)
        kl = hldR.reshape(np_, 1) - hldS.reshape(1, nb) + 0.5 * (c.reshape(np_, nb, Q + 1).sum(-1) - Q)
        return kl, (P,)
    kl = _empty(CS, np_, nb)
    t = _empty(CS, np_, nb, Q)
    rs = _empty(CS, nb, Q)
    work = _empty(CS, np_, nb, Q)
    check(lib().nmgp_kl_fwd(_d(CS), _d(hldS), _d(mu), _d(R), _d(hldR), _d(kl), _d(t), _d(rs), _d(work),
                            c_int(np_), c_int(nb), c_int(Q), _stream()), "nmgp_kl_fwd")
    return kl, (t, rs)


def kl_bwd(klbar, CS, mu, R, saved, exact=False):
    nb, Q, _ = CS.shape
    np_ = R.shape[0]
    Rbar = _zeros(CS, np_, Q, Q)
    if exact:
        (P,) = saved                                              # [np, nb (Q+1), Q] = rows (R R^T)^-1
        Wr = (klbar.reshape(np_, nb, 1, 1) * P.reshape(np_, nb, Q + 1, Q))
        rows_bar = Wr.sum(0)                                      # [nb, Q+1, Q]: cotangent of (mu_b ; columns of CS_b)
        mubar = rows_bar[:, 0].contiguous()
        CSbar = torch.tril(rows_bar[:, 1:].transpose(1, 2)).contiguous()
        G = _zeros(CS, np_, Q, Q)
        atb(Wr.reshape(np_, nb * (Q + 1), Q).contiguous(), P, G, 1.0)
        kl_rbar(R, G, Rbar)
        return CSbar, -klbar.sum(0), mubar, Rbar, klbar.sum(1)
    t, rs = saved
    CSbar = torch.empty_like(CS)
    hldSbar = _empty(CS, nb)
    mubar = _empty(CS, nb, Q)
    hldRbar = _empty(CS, np_)
    work = _empty(CS, np_, nb, Q)
    rsb = _empty(CS, nb, Q)
    G = _empty(CS, np_, Q, Q)
    check(lib().nmgp_kl_bwd(_d(klbar), _d(CS), _d(R), _d(t), _d(rs), _d(CSbar), _d(hldSbar), _d(mubar), _d(Rbar),
                            _d(hldRbar), _d(work), _d(rsb), _d(G), c_int(np_), c_int(nb), c_int(Q), _stream()),
          "nmgp_kl_bwd")
    return CSbar, hldSbar, mubar, Rbar, hldRbar


def rbf_build_fwd(x, z, hyp, is2, ilen, jitter=0.0):
    B, Q = x.numel(), z.numel()
    K = _empty(x, B, Q)
    check(lib().nmgp_rbf_build_fwd(_d(x), _d(z), _d(hyp), c_int(is2), c_int(ilen), c_double(jitter), _d(K),
                                   c_int64(B), c_int(Q), _stream()), "nmgp_rbf_build_fwd")
    return K


def rbf_build_bwd(x, z, hyp, is2, ilen, Kbar, ghyp):
    B, Q = x.numel(), z.numel()
    check(lib().nmgp_rbf_build_bwd(_d(x), _d(z), _d(hyp), c_int(is2), c_int(ilen), _d(Kbar), _d(ghyp),
                                   c_int64(B), c_int(Q), _stream()), "nmgp_rbf_build_bwd")


def gibbs_build_fwd(x, z, ellx, ellz, jitter=0.0):
    ns, B = ellx.shape
    Q = z.numel()
    K = _empty(x, ns, B, Q)
    check(lib().nmgp_gibbs_build_fwd(_d(x), _d(z), _d(ellx), _d(ellz), c_double(jitter), _d(K),
                                     c_int(ns), c_int64(B), c_int(Q), _stream()), "nmgp_gibbs_build_fwd")
    return K


def gibbs_build_bwd(x, z, ellx, ellz, Kbar, ellxbar, ellzbar, Kfwd=None):
    ns, B = ellx.shape
    Q = z.numel()
    check(lib().nmgp_gibbs_build_bwd(_d(x), _d(z), _d(ellx), _d(ellz), _d(Kbar), _optd(Kfwd), _d(ellxbar), _d(ellzbar),
                                     c_int(ns), c_int64(B), c_int(Q), _stream()), "nmgp_gibbs_build_bwd")


def solve_rows_fwd(K, R):
    ns, B, Q = K.shape
    _checkQ(Q)
    P = torch.empty_like(K)
    c = _empty(K, ns, B)
    check(lib().nmgp_solve_rows_fwd(_d(K), _d(R), _d(P), _d(c), c_int(ns), c_int64(B), c_int(Q), _stream()),
          "nmgp_solve_rows_fwd")
    return P, c


def solve_rows_bwd(Pbar, cbar, K, P, R, Abar):
    ns, B, Q = K.shape
    Kbar = torch.empty_like(K)
    work_t = torch.empty_like(K)                           # kept alive until the launch below has been enqueued
    work = _optd(work_t)
    check(lib().nmgp_solve_rows_bwd(_d(Pbar), _d(cbar), _d(K), _d(P), _d(R), _d(Kbar), _d(Abar), work,
                                    c_int(ns), c_int64(B), c_int(Q), _stream()), "nmgp_solve_rows_bwd")
    return Kbar


def lq_pad_records(Sig):
    """Padded, half-split copies of a batch of Q x Q covariances (64 < Q <= 128) in the order the large-Q DMMA kernels
    stream them (csrc/nmgp_quadform_lq.cu); build once per step and pass as ``rec``."""
    n, Q, _ = Sig.shape
    lib().nmgp_lq_record_doubles.restype = ctypes.c_longlong
    per = int(lib().nmgp_lq_record_doubles(c_int(Q)))
    if per <= 0:
        raise ValueError("lq_pad_records: Q=%d outside 65..128" % Q)
    rec = torch.empty(n * per, dtype=F64, device=Sig.device)
    check(lib().nmgp_lq_pad_records(_d(Sig), _d(rec), c_int(n), c_int(Q), _stream()), "nmgp_lq_pad_records")
    return rec




def quadform_fwd(Pa, Pb, I, Sig, Mu, D, mode, seg=None, rec=None):
    ns, B, Q = Pa.shape
    seg = segment_offsets(I, D) if seg is None else seg
    q = _zeros(Pa, ns, B, D)
    m = _zeros(Pa, ns, B, D)
    if mode == MODE_U and Q >= LQ_MIN_Q:
        rec = lq_pad_records(Sig) if rec is None else rec
        check(lib().nmgp_lq_coef_quadform(c_int(0), _d(Pa), _d(Pb), _i(I), _d(rec), _d(Mu), _d(q), _d(m), c_void_p(0),
                                          c_void_p(0), c_void_p(0), c_void_p(0), c_int(ns), c_int64(B), c_int(Q),
                                          c_int(D), _stream()), "nmgp_lq_coef_quadform")
        return q, m
    check(lib().nmgp_quadform_fwd(_d(Pa), _d(Pb), _i(I), _i(seg), _d(Sig), _d(Mu), _d(q), _d(m),
                                  c_int(ns), c_int64(B), c_int(Q), c_int(D), c_int(mode), _stream()), "nmgp_quadform_fwd")
    return q, m


def quadform_bwd(Pa, Pb, I, Sig, Mu, qbar, mbar, mode, seg=None, rec=None):
    ns, B, Q = Pa.shape
    D = qbar.shape[-1]
    seg = segment_offsets(I, D) if seg is None else seg
    Pabar = torch.empty_like(Pa)
    Pbbar = torch.empty_like(Pa) if mode == MODE_U else Pabar
    if mode == MODE_U and Q >= LQ_MIN_Q:
        rec = lq_pad_records(Sig) if rec is None else rec
        check(lib().nmgp_lq_coef_quadform(c_int(1), _d(Pa), _d(Pb), _i(I), _d(rec), _d(Mu), c_void_p(0), c_void_p(0),
                                          _d(qbar), _d(mbar), _d(Pabar), _d(Pbbar), c_int(ns), c_int64(B), c_int(Q),
                                          c_int(D), _stream()), "nmgp_lq_coef_quadform")
        return Pabar, Pbbar
    check(lib().nmgp_quadform_bwd(_d(Pa), _d(Pb), _i(I), _i(seg), _d(Sig), _d(Mu), _d(qbar), _d(mbar),
                                  _d(Pabar), _d(Pbbar), c_int(ns), c_int64(B), c_int(Q), c_int(D), c_int(mode),
                                  _stream()), "nmgp_quadform_bwd")
    return Pabar, (Pbbar if mode == MODE_U else None)


def weighted_gram(Pa, Pb, I, qbar, mbar, mode, SigBar, MuBar, seg=None):
    ns, B, Q = Pa.shape
    D = qbar.shape[-1]
    seg = segment_offsets(I, D) if seg is None else seg
    check(lib().nmgp_weighted_gram(_d(Pa), _d(Pb), _i(I), _i(seg), _d(qbar), _d(mbar), _d(SigBar), _d(MuBar),
                                   c_int(ns), c_int64(B), c_int(Q), c_int(D), c_int(mode), _stream()), "nmgp_weighted_gram")


def sample_v_fwd(mu_v, Cv, zv):
    S, Q = zv.shape
    v = torch.empty_like(zv)
    ellz = torch.empty_like(zv)
    check(lib().nmgp_sample_v_fwd(_d(mu_v), _d(Cv), _d(zv), _d(v), _d(ellz), c_int(S), c_int(Q), _stream()),
          "nmgp_sample_v_fwd")
    return v, ellz


def sample_v_bwd(ellzbar, vbar, ellz, zv, mu_v_bar, Cvbar):
    S, Q = zv.shape
    check(lib().nmgp_sample_v_bwd(_d(ellzbar), _d(vbar), _d(ellz), _d(zv), _d(mu_v_bar), _d(Cvbar),
                                  c_int(S), c_int(Q), _stream()), "nmgp_sample_v_bwd")


def ell_sd_fwd(c_ell, hyp):
    sd = torch.empty_like(c_ell)
    check(lib().nmgp_ell_sd_fwd(_d(c_ell), _d(hyp), _d(sd), c_int64(c_ell.numel()), _stream()), "nmgp_ell_sd_fwd")
    return sd


def ell_sd_bwd(sdbar, sd, hyp, ghyp):
    cbar = torch.empty_like(sd)
    check(lib().nmgp_ell_sd_bwd(_d(sdbar), _d(sd), _d(hyp), _d(ghyp), _d(cbar), c_int64(sd.numel()), _stream()),
          "nmgp_ell_sd_bwd")
    return cbar


def ell_rows_fwd(Pell, v, zell, sdell):
    ns, Q = v.shape
    B = Pell.shape[0]
    ellx = _empty(Pell, ns, B)
    check(lib().nmgp_ell_rows_fwd(_d(Pell), _d(v), _d(zell), _d(sdell), _d(ellx), c_int(ns), c_int64(B), c_int(Q),
                                  _stream()), "nmgp_ell_rows_fwd")
    return ellx


def ell_rows_bwd(ellxbar, ellx, Pell, v, zell, vbar, Pellbar, sdbar):
    ns, Q = v.shape
    B = Pell.shape[0]
    check(lib().nmgp_ell_rows_bwd(_d(ellxbar), _d(ellx), _d(Pell), _d(v), _d(zell), _d(vbar), _d(Pellbar), _d(sdbar),
                                  c_int(ns), c_int64(B), c_int(Q), _stream()), "nmgp_ell_rows_bwd")


def coef_sd_fwd(q, cL0, cL1, I, hyp):
    B, D = q.shape
    sd = torch.empty_like(q)
    check(lib().nmgp_coef_sd_fwd(_d(q), _d(cL0), _d(cL1), _i(I), _d(hyp), _d(sd), c_int64(B), c_int(D), _stream()),
          "nmgp_coef_sd_fwd")
    return sd


def coef_sd_bwd(sdbar, sd, I, hyp, ghyp):
    B, D = sd.shape
    qbar = torch.empty_like(sd)
    cL0bar = _empty(sd, B)
    cL1bar = _empty(sd, B)
    check(lib().nmgp_coef_sd_bwd(_d(sdbar), _d(sd), _i(I), _d(hyp), _d(ghyp), _d(qbar), _d(cL0bar), _d(cL1bar),
                                 c_int64(B), c_int(D), _stream()), "nmgp_coef_sd_bwd")
    return qbar, cL0bar, cL1bar


def _l(t):
    if t is None:
        return c_void_p(0)
    if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()):
        raise TypeError("expected a contiguous CUDA int64 tensor")
    return c_void_p(t.data_ptr())


def _noise_args(noise):
    """noise = (seed, stream_id, s0, ns, gid[, step_dev]) for in-kernel generation; None when explicit draws are passed.
    step_dev: int64 device scalar holding the step counter (CUDA-graph replay), see nmgp_noise_fill."""
    if noise is None:
        return ctypes.c_uint64(0), ctypes.c_uint64(0), c_int(0), c_void_p(0), c_void_p(0)
    seed, stream_id, s0, _, gid = noise[:5]
    step_dev = noise[5] if len(noise) > 5 else None
    return ctypes.c_uint64(seed), ctypes.c_uint64(stream_id), c_int(s0), _l(gid), _l(step_dev)


def coef_sample_fwd(m, sd, zL, I, noise=None):
    B, D = m.shape
    ns = zL.shape[0] if zL is not None else noise[3]
    l = _empty(m, ns, B, D)
    a = _noise_args(noise)
    check(lib().nmgp_coef_sample_fwd(_d(m), _d(sd), _optd(zL), _i(I), _d(l), c_int(ns), c_int64(B), c_int(D),
                                     a[0], a[1], a[2], a[3], a[4], _stream()), "nmgp_coef_sample_fwd")
    return l


def coef_sample_bwd(lbar, l, zL, I, mbar, sdbar, noise=None):
    ns, B, D = l.shape
    a = _noise_args(noise)
    check(lib().nmgp_coef_sample_bwd(_d(lbar), _d(l), _optd(zL), _i(I), _d(mbar), _d(sdbar), c_int(ns), c_int64(B),
                                     c_int(D), a[0], a[1], a[2], a[3], a[4], _stream()), "nmgp_coef_sample_bwd")


def noise_fill(ns, B, C, seed, stream_id, s0, gid, device, step_dev=None):
    out = torch.empty(ns, B, C, dtype=F64, device=device)
    check(lib().nmgp_noise_fill(_d(out), c_int(ns), c_int64(B), c_int(C), ctypes.c_uint64(seed),
                                ctypes.c_uint64(stream_id), c_int(s0), _l(gid), _l(step_dev), _stream()), "nmgp_noise_fill")
    return out


def _ystride(y, ns, B):
    """y [B]: one target vector shared by the samples of the chunk; y [ns, B]: one per sample (subjects)."""
    if y.dim() == 1:
        return 0
    if y.shape[0] != ns or y.shape[1] != B:
        raise ValueError("per-sample targets must be [ns, B] = [%d, %d], got %s" % (ns, B, tuple(y.shape)))
    return B


def lik_rows(l, mg, qg, cG, y, I, hyp, scale, Rsum, ghyp):
    ns, B, D = l.shape
    lbar = torch.empty_like(l); mgbar = torch.empty_like(l); qgbar = torch.empty_like(l)
    cGbar = _empty(l, ns, B)
    Rsum.zero_()
    check(lib().nmgp_lik_rows(_d(l), _d(mg), _d(qg), _d(cG), _d(y), _i(I), _d(hyp), c_double(scale), _d(Rsum), _d(ghyp),
                              _d(lbar), _d(mgbar), _d(qgbar), _d(cGbar), c_int(ns), c_int64(B), c_int(D),
                              c_int64(_ystride(y, ns, B)), _stream()),
          "nmgp_lik_rows")
    return lbar, mgbar, qgbar, cGbar


def pair_means(Pa, Pb, I, Mu, D, mode):
    ns, B, Q = Pa.shape
    m = _empty(Pa, ns, B, D)
    check(lib().nmgp_pair_means(_d(Pa), _d(Pb), _i(I), _d(Mu), _d(m), c_int(ns), c_int64(B), c_int(Q), c_int(D),
                                c_int(mode), _stream()), "nmgp_pair_means")
    return m


def rowdot_live(l, g, I):
    ns, B, D = l.shape
    F = _empty(l, ns, B)
    check(lib().nmgp_rowdot_live(_d(l), _d(g), _i(I), _d(F), c_int(ns), c_int64(B), c_int(D), _stream()),
          "nmgp_rowdot_live")
    return F


def latent_fused(PG, cG, l, y, I, SigW, muW, hyp, scale, Rsum, ghyp, seg=None, rec=None):
    """Latent-function statistics, expected log-likelihood and every row cotangent of one sample chunk in a
    single DMMA kernel (register-resident for Q <= 64, ring-pipelined for 64 < Q <= 128; ``rec`` = lq_pad_records(SigW)
    may be passed to reuse the padded records over the sample chunks of a step).
    Returns (lbar, mgbar, qgbar, cGbar, PGbar); Rsum (=) and ghyp (+=) in place."""
    ns, B, Q = PG.shape
    D = l.shape[-1]
    lbar = torch.empty_like(l); mgbar = torch.empty_like(l); qgbar = torch.empty_like(l)
    cGbar = _empty(l, ns, B)
    PGbar = torch.empty_like(PG)
    Rsum.zero_()
    if Q >= LQ_MIN_Q:
        rec = lq_pad_records(SigW) if rec is None else rec
        check(lib().nmgp_lq_latent_fused(_d(PG), _d(cG), _d(l), _d(y), _i(I), _d(rec), _d(muW), _d(hyp), c_double(scale),
                                         _d(Rsum), _d(ghyp), _d(lbar), _d(mgbar), _d(qgbar), _d(cGbar), _d(PGbar),
                                         c_int(ns), c_int64(B), c_int(Q), c_int(D), c_int64(_ystride(y, ns, B)),
                                         _stream()), "nmgp_lq_latent_fused")
        return lbar, mgbar, qgbar, cGbar, PGbar
    seg = segment_offsets(I, D) if seg is None else seg
    pq = pm = c_void_p(0)
    check(lib().nmgp_latent_fused(_d(PG), _d(cG), _d(l), _d(y), _i(I), _i(seg), _d(SigW), _d(muW), _d(hyp),
                                  c_double(scale), _d(Rsum), _d(ghyp), _d(lbar), _d(mgbar), _d(qgbar), _d(cGbar),
                                  _d(PGbar), pq, pm, c_int(ns), c_int64(B), c_int(Q), c_int(D),
                                  c_int64(_ystride(y, ns, B)), _stream()),
          "nmgp_latent_fused")
    return lbar, mgbar, qgbar, cGbar, PGbar


def reparam_diag(mean, var, z):
    out = torch.empty_like(mean)
    check(lib().nmgp_reparam_diag(_d(mean), _d(var), _d(z), _d(out), c_int64(mean.numel()), _stream()), "nmgp_reparam_diag")
    return out


def normal_logprob_sum(loc, scale, y):
    out = _zeros(loc, 1)
    check(lib().nmgp_normal_logprob_sum(_d(loc), _d(scale), _d(y), _d(out), c_int64(loc.numel()), _stream()),
          "nmgp_normal_logprob_sum")
    return out


def sumsq_rows(x):
    rows, cols = x.shape
    out = _empty(x, rows)
    check(lib().nmgp_sumsq_rows(_d(x), _d(out), c_int64(rows), c_int64(cols), _stream()), "nmgp_sumsq_rows")
    return out


def lcorr(L):
    nmat, D = L.shape[0], L.shape[-1]
    out = torch.empty_like(L)
    check(lib().nmgp_lcorr(_d(L), _d(out), c_int64(nmat), c_int(D), _stream()), "nmgp_lcorr")
    return out


# ---- SIM_code (exact / Kronecker) line --------------------------------------------------------------
def nonstationary_cov(X1, sigma1, ell1, X2, sigma2, ell2, jitter, self_cov=False):
    """Nonstationary_RBF_cov.  ``self_cov``: X2/sigma2/ell2 ARE X1/sigma1/ell1 (the reference's X2=None call): only the
    tiles on and below the diagonal are evaluated and mirrored, jitter is added on the diagonal."""
    T1, dx = X1.shape
    T2 = X2.shape[0]
    K = _empty(X1, T1, T2)
    check(lib().nmgp_nonstationary_cov(_d(X1), _optd(sigma1), _optd(ell1), _d(X2), _optd(sigma2), _optd(ell2),
                                       c_double(jitter), _d(K), c_int64(T1), c_int64(T2), c_int(dx),
                                       c_int(1 if self_cov else 0), _stream()), "nmgp_nonstationary_cov")
    return K


def nonstationary_cov_bwd(X1, sigma1, ell1, X2, sigma2, ell2, Kbar, want=(True, True, True, True)):
    """Adjoint of nonstationary_cov w.r.t. (sigma1, ell1, sigma2, ell2) for 1-D inputs; entries of ``want`` switch the
    four outputs (None where not wanted or where the argument itself was None).  For a self-covariance the caller adds the
    row-side and the column-side gradients."""
    T1, dx = X1.shape
    T2 = X2.shape[0]
    outs = []
    for arg, w, n in ((sigma1, want[0], T1), (ell1, want[1], T1), (sigma2, want[2], T2), (ell2, want[3], T2)):
        outs.append(_zeros(X1, n) if (w and arg is not None) else None)
    check(lib().nmgp_nonstationary_cov_bwd(_d(X1), _optd(sigma1), _optd(ell1), _d(X2), _optd(sigma2), _optd(ell2),
                                           _d(Kbar), _optd(outs[0]), _optd(outs[1]), _optd(outs[2]), _optd(outs[3]),
                                           c_int64(T1), c_int64(T2), c_int(dx), _stream()), "nmgp_nonstationary_cov_bwd")
    return tuple(outs)


def hadamard_index_cov(Kx, Bf, indx1, indx2, diag=0.0):
    N1, N2 = Kx.shape
    out = _empty(Kx, N1, N2)
    check(lib().nmgp_hadamard_index_cov(_d(Kx), _d(Bf), _i(indx1), _i(indx2), c_double(diag), _d(out), c_int64(N1),
                                        c_int64(N2), c_int(Bf.shape[1]), _stream()), "nmgp_hadamard_index_cov")
    return out


def dense_loglik_bwd(Sinv, alpha, A, Bt, indx1, indx2, g):
    """Cotangents (Abar [N,N], Btbar like Bt, s2bar [1]) of -1/2 logdet S - 1/2 y^T S^-1 y with S = A o Bt[indx1, indx2] +
    sigma2 I, from Sinv = S^-1, alpha = S^-1 y and the upstream cotangent g (device scalar)."""
    N = A.shape[0]
    Abar = _empty(A, N, N)
    Btbar = _zeros(Bt, *Bt.shape)
    s2bar = _zeros(A, 1)
    check(lib().nmgp_dense_loglik_bwd(_d(Sinv), _d(alpha), _d(A), _d(Bt), _i(indx1), _i(indx2), _d(g), _d(Abar), _d(Btbar),
                                      _d(s2bar), c_int64(N), c_int(Bt.shape[0]), c_int(Bt.shape[1]), _stream()),
          "nmgp_dense_loglik_bwd")
    return Abar, Btbar, s2bar


def sim_rbf_cov(X1, X2, alpha, beta, jitter, self_cov=False):
    T1, dx = X1.shape
    T2 = X2.shape[0]
    K = _empty(X1, T1, T2)
    check(lib().nmgp_sim_rbf_cov(_d(X1), _d(X2), c_double(alpha), c_double(beta), c_double(jitter), _d(K),
                                 c_int64(T1), c_int64(T2), c_int(dx), c_int(1 if self_cov else 0), _stream()),
          "nmgp_sim_rbf_cov")
    return K


def pairwise_dist(X1, X2):
    T1, dx = X1.shape
    T2 = X2.shape[0]
    out = _empty(X1, T1, T2)
    check(lib().nmgp_pairwise_dist(_d(X1), _d(X2), _d(out), c_int64(T1), c_int64(T2), c_int(dx), _stream()),
          "nmgp_pairwise_dist")
    return out


def gemm_nt(A, Bm, alpha=1.0, beta=0.0, C=None):
    """C = alpha * A @ Bm.T + beta * C  (A [M,K], Bm [N,K] row-major) on the FP64 tensor cores.  Row-strided 2-D views
    (unit column stride, e.g. blocks of a larger matrix) are accepted for A, Bm and C."""
    M, K = A.shape
    N = Bm.shape[0]
    if C is None:
        C = _empty(A, M, N)
        beta = 0.0
    for t in (A, Bm, C):
        if not (t.is_cuda and t.dtype == F64 and t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1)):
            raise TypeError("gemm_nt expects CUDA float64 2-D tensors with unit column stride")
        _same_device(t)
    ld = lambda t: int(t.stride(0)) if t.shape[0] > 1 else max(int(t.shape[1]), 1)
    if M == 0 or N == 0:
        return C
    check(lib().nmgp_gemm_nt(c_void_p(A.data_ptr()), c_void_p(Bm.data_ptr()), c_void_p(C.data_ptr()), c_int64(M),
                             c_int64(N), c_int64(K), c_int64(max(ld(A), K)), c_int64(max(ld(Bm), K)), c_int64(max(ld(C), N)),
                             c_double(alpha), c_double(beta), _stream()), "nmgp_gemm_nt")
    return C


def build_augmented(K, r, alpha_dev, sigma2_dev, rnorm2_dev, out):
    """out [T+1 rows, leading dimension out.stride(0)] = [[alpha K + sigma2 I, .], [r^T, 1 + |r|^2 / sigma2]]."""
    T = K.shape[0]
    check(lib().nmgp_build_augmented(_d(K), _d(r), c_void_p(out.data_ptr()), c_int64(T), c_int64(out.stride(0)),
                                     c_void_p(alpha_dev.data_ptr()), c_void_p(sigma2_dev.data_ptr()),
                                     c_void_p(rnorm2_dev.data_ptr()), _stream()), "nmgp_build_augmented")
    return out


def augmented_results(A, T, hld_aug, hld_out, quad_out):
    check(lib().nmgp_augmented_results(c_void_p(A.data_ptr()), c_int64(T), c_int64(A.stride(0)), _d(hld_aug),
                                       c_void_p(hld_out.data_ptr()), c_void_p(quad_out.data_ptr()), _stream()),
          "nmgp_augmented_results")


def potrf_big(A, info=None, slot=0, panel=0):
    """In-place blocked lower Cholesky of the square matrix A; returns (A, sum(log(diag))).  Raises RuntimeError on a
    non-positive pivot like torch.cholesky -- unless ``info`` (int32 device scalar, receives the order of the failing
    leading minor, 0 if none) is given: then nothing is read back here (no host synchronisation).  ``slot`` (0..3)
    selects the library's look-ahead scratch; factorisations running concurrently on different streams need different
    slots."""
    T = A.shape[0]
    hld = _empty(A, 1)
    deferred = info is not None
    if not deferred:
        info = torch.zeros(1, dtype=torch.int32, device=A.device)
    if not (A.is_cuda and A.dtype == F64 and A.dim() == 2 and A.stride(1) == 1):
        raise TypeError("potrf_big expects a CUDA float64 matrix with unit column stride")
    _same_device(A)
    check(lib().nmgp_potrf_big_slot(c_void_p(A.data_ptr()), c_int64(T), c_int64(A.stride(0) if T > 1 else 1), _d(hld),
                                    _i(info), c_int(slot), c_int(panel), _stream()), "nmgp_potrf_big")
    if not deferred:
        bad = int(info.item())
        if bad != 0:
            raise RuntimeError("cholesky: the leading minor of order %d is not positive-definite" % bad)
    return A, hld


def gemm_concurrent_mode(on):
    """Tell the library that several of its GEMMs / factorisations are about to run concurrently on different streams (its
    GEMMs then stage operands with cp.async, see nmgp_gemm_concurrent_mode)."""
    lib().nmgp_gemm_concurrent_mode(c_int(1 if on else 0))


def tri_inv_block(L, out, scale=1.0):
    """out = scale * inverse of the lower-triangular block L (n <= 128; 2-D views with unit column stride)."""
    n = L.shape[0]
    check(lib().nmgp_tri_inv_block(c_void_p(L.data_ptr()), c_int64(L.stride(0) if n > 1 else n), c_int(n),
                                   c_void_p(out.data_ptr()), c_int64(out.stride(0) if n > 1 else n), c_double(scale),
                                   _stream()), "nmgp_tri_inv_block")
    return out


def potrs_vec(L, b):
    x = b.clone()
    check(lib().nmgp_potrs_vec(_d(L), c_int64(L.shape[0]), c_int64(L.shape[0]), _d(x), _stream()), "nmgp_potrs_vec")
    return x


def scale_add_diag(K, alpha, sigma2):
    A = torch.empty_like(K)
    check(lib().nmgp_scale_add_diag(_d(K), _d(A), c_int64(K.shape[0]), c_double(alpha), c_double(sigma2), _stream()),
          "nmgp_scale_add_diag")
    return A


def scale_add_diag_dev(K, alpha_dev, sigma2_dev, out=None):
    """A = alpha K + sigma2 I with both scalars read on the device (1-element float64 tensors / views)."""
    A = torch.empty_like(K) if out is None else out
    check(lib().nmgp_scale_add_diag_dev(_d(K), _d(A), c_int64(K.shape[0]), c_void_p(alpha_dev.data_ptr()),
                                        c_void_p(sigma2_dev.data_ptr()), _stream()), "nmgp_scale_add_diag_dev")
    return A


def kron_product(t1, t2):
    h1, w1 = t1.shape
    h2, w2 = t2.shape
    out = _empty(t1, h1 * h2, w1 * w2)
    check(lib().nmgp_kron_product(_d(t1), _d(t2), _d(out), c_int(h1), c_int(w1), c_int64(h2), c_int64(w2), _stream()),
          "nmgp_kron_product")
    return out


def eigh_small(A):
    n = A.shape[0]
    lib().nmgp_eigh_small_work.restype = ctypes.c_longlong
    w = _empty(A, n); V = _empty(A, n, n); work = _empty(A, int(lib().nmgp_eigh_small_work(c_int(n))))
    check(lib().nmgp_eigh_small(_d(A), _d(w), _d(V), _d(work), c_int(n), _stream()), "nmgp_eigh_small")
    return w, V


def axpby(x, y, a, b):
    out = torch.empty_like(x)
    check(lib().nmgp_axpby(_d(x), _d(y), _d(out), c_int64(x.numel()), c_double(a), c_double(b), _stream()), "nmgp_axpby")
    return out


def axpby_dev(x, y, a_dev, a_scale=1.0, b=1.0, out=None):
    """out = (a_scale * a_dev[0]) * x + b * y with a_dev a 1-element device tensor (no host read)."""
    out = torch.empty_like(x) if out is None else out
    check(lib().nmgp_axpby_dev(_d(x), _d(y), _d(out), c_int64(x.numel()), c_void_p(a_dev.data_ptr()), c_double(a_scale),
                               c_double(b), _stream()), "nmgp_axpby_dev")
    return out


def dot(x, y):
    out = _zeros(x, 1)
    check(lib().nmgp_dot(_d(x), _d(y), _d(out), c_int64(x.numel()), _stream()), "nmgp_dot")
    return out


# ---- optional per-call CUDA-event timing (used by bench.py; off by default) -------------------------
import functools as _functools

_PROFILE = None
_CALLS = [0]


def launch_count() -> int:
    """Kernel launches issued by libnmgp_b200.so since it was loaded (+ launches replayed from captured CUDA graphs)."""
    lib().nmgp_launch_count.restype = ctypes.c_uint64
    return int(lib().nmgp_launch_count()) + _GRAPH_LAUNCHES[0]


_GRAPH_LAUNCHES = [0]


def _profile_begin():
    global _PROFILE
    _PROFILE = []
    _CALLS[0] = 0


def _profile_end():
    """Returns {op name: (calls, total milliseconds)} and the number of C-ABI calls since _profile_begin()."""
    global _PROFILE
    rec, _PROFILE = _PROFILE or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in rec:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + e0.elapsed_time(e1))
    return out, _CALLS[0]


def _instrument(fn):
    @_functools.wraps(fn)
    def wrapped(*a, **k):
        _CALLS[0] += 1
        if _PROFILE is None:
            return fn(*a, **k)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        _PROFILE.append((fn.__name__, e0, e1))
        return r
    return wrapped


for _name, _fn in list(globals().items()):
    if callable(_fn) and not _name.startswith("_") and getattr(_fn, "__module__", None) == __name__ \
            and _name not in ("check", "lib", "launch_count"):
        globals()[_name] = _instrument(_fn)


