"""One DSVI iteration (S-sample reparameterised -ELBO and its hand-written gradient).

This is the B200 restatement of ``NMGP.forward`` + ``loss.backward`` of the reference
(code/nmgp_dsvi.py:157-301 and :847), de-duplicated as described in DESIGN.md:

* the four distinct inducing systems (tilde-ell, L0, L1, Gibbs) are factored once
  (the reference re-solves them D(D+1)/2 + 2 times, code/utils.py:119,142,230);
* the coefficient statistics are evaluated only for the (row n, j <= I[n]) pairs the
  reference keeps at code/nmgp_dsvi.py:238;
* the latent statistics only for j <= I[n] (the others are multiplied by zero at :255-258);
* every adjoint is written out -- no autograd graph exists.

Host code here only sequences kernels from ``_ops`` (the C-ABI wrappers) and moves
buffers; all arithmetic on B-, Q^2- or S-sized data happens in the CUDA kernels.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, Optional

import torch

from . import _ops as ops

EPS = 1e-4                                   # tridiagonal_jitter, code/utils.py:7
HYPER_ORDER = ("sigma2_tildeell_log", "length_scales_tildeell_log", "sigma2_L0_log",
               "length_scales_L0_log", "sigma2_L1_log", "length_scales_L1_log", "sigma2_err_log")
H_S2_ELL, H_LEN_ELL, H_S2_L0, H_LEN_L0, H_S2_L1, H_LEN_L1, H_S2_ERR = range(7)
MODE_W, MODE_U = 0, 1
PARAM_NAMES = ("mu_W", "sqrt_W", "mu_v", "sqrt_v", "mu_U", "sqrt_U") + HYPER_ORDER

_pair_cache: Dict[tuple, torch.Tensor] = {}
_side_streams: Dict[str, "torch.cuda.Stream"] = {}


def _side_stream(dev):
    key = str(dev)
    if key not in _side_streams:
        # highest priority: the few-CTA, latency-bound KL kernels must get SM slots ahead of the queued row-kernel CTAs,
        # otherwise they only start when a row kernel drains and the overlap is lost
        _side_streams[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _side_streams[key]


def packed_pair_index(D: int, device) -> torch.Tensor:
    """Flat indices i*D+j of the live coefficient pairs, packed as: the D diagonal
    pairs (i,i) first, then the strictly-lower pairs (i,j<i) row-major
    (slot D + i(i-1)/2 + j).  Pairs with j > i never receive gradient (quirk q8)."""
    key = (D, str(device))
    if key not in _pair_cache:
        idx = [i * D + i for i in range(D)]
        idx += [i * D + j for i in range(D) for j in range(i)]
        _pair_cache[key] = torch.tensor(idx, dtype=torch.int64, device=device)
    return _pair_cache[key]


def _sample_budget_bytes() -> float:
    """Bytes the [ns,B,*] temporaries of one pass may take: NMGP_SAMPLE_BUDGET_GB, default 6 GB.  (Larger launches do not
    pay: 4 samples per launch already fill 55 waves at the ECoG shape; 131.3 / 131.4 / 130.5 ms per step at 6 / 13 / 50 GB,
    profiles/README.md.)"""
    import os
    try:
        return float(os.environ.get("NMGP_SAMPLE_BUDGET_GB", "6")) * 1e9
    except ValueError:
        return 6e9


def default_sample_chunk(S: int, B: int, Q: int, D: int, budget_bytes: Optional[float] = None) -> int:
    """Samples processed per pass so the [ns,B,*] temporaries stay within a budget."""
    per_sample = 8.0 * B * (5 * Q + 7 * D + 8)
    budget = _sample_budget_bytes() if budget_bytes is None else budget_bytes
    return max(1, min(S, int(budget // max(per_sample, 1.0))))


def dsvi_step(p: Dict[str, torch.Tensor], Z: torch.Tensor, x: torch.Tensor, y: torch.Tensor,
              I: torch.Tensor, N: int, z_v: torch.Tensor, z_ell: torch.Tensor, z_L: torch.Tensor, *,
              B_total: Optional[int] = None, kl_weight: float = 1.0,
              sample_chunk: Optional[int] = None, want_grads: bool = True, pair_index: Optional[torch.Tensor] = None,
              latent_order: Optional[torch.Tensor] = None, aux: Optional[dict] = None,
              kl_shard: Optional[tuple] = None, noise_key: Optional[tuple] = None,
              row_gid: Optional[torch.Tensor] = None, defer_pd_check: bool = False,
              S_total: Optional[int] = None, sample_offset: int = 0, exact_kl: bool = False):
    """Returns (loss, grads) for rows (x, y, I) -- I sorted ascending, int32 -- and noise
    z_v [S,Q], z_ell [S,B], z_L [S,B,D] (z_L[s,n,j] is the draw for pair (I[n], j)).

    Under row sharding (several ranks each holding a slice of the minibatch) pass the
    global row count as ``B_total`` and ``kl_weight = 1/world_size``; the sum over
    ranks of the returned loss/grads is then the full-batch value.

    ``z_L`` may be None: the coefficient noise is then generated inside the sampling kernels from
    ``noise_key = (seed, stream_id)`` and ``row_gid`` (int64 global row ids; default arange(B)), see nmgp_noise_fill.

    ``kl_shard = (rank, world)`` additionally splits the batched small-matrix work of the KL terms (Cholesky of the
    P coefficient covariances, KL forward/backward) over the ranks instead of replicating it: each KL_U pair and each
    sample's KL_W is then evaluated by exactly one rank with weight 1 (KL_v stays replicated with ``kl_weight``).

    ``y`` may be [S, B]: one target vector per sample -- the S "samples" are then S subjects observed on the same rows
    (HCP-shaped step: the mean over subjects of the reference's one-draw forward on that subject's data).

    ``S_total`` / ``sample_offset`` (sample sharding: each rank holds ALL rows but only samples sample_offset ..
    sample_offset + S - 1 of S_total): the estimate is scaled by 1/S_total, every local sample's KL_W is evaluated
    here, and the counter-based noise is keyed by the global sample index.

    ``exact_kl``: evaluate the mathematically correct KL terms instead of the reference's (quirk q10, SURVEY 8a: the
    reference's trace term only uses diag(chol(K)); every ELBO it prints contains that).  Not the default.

    ``defer_pd_check`` (set under multi-rank sharding): a failed Cholesky is not raised here -- a rank that raised alone
    would leave the others waiting in the gradient all-reduce -- but returned as ``aux["pd_info"]`` (int32 device scalar,
    1 + index of the failing matrix); ``parallel.allreduce_loss_and_grads`` reduces it and raises on every rank.

    ``pair_index`` (flat i*D+j per packed pair slot) and ``latent_order`` (permutation of the D latent functions)
    override the default packing; compute_ELBO uses them to evaluate the reference's transposed coefficient gather
    (quirk q5) with the same kernels.  ``aux`` (a dict) receives the per-sample pieces of the estimate.
    """
    D, Q = p["mu_W"].shape
    B = x.shape[0]
    S = z_v.shape[0]
    if y.dim() == 2 and tuple(y.shape) != (S, B):
        raise ValueError("per-sample targets must be [S, B] = [%d, %d], got %s" % (S, B, tuple(y.shape)))
    dev = x.device
    f64 = torch.float64
    S_tot = S if S_total is None else int(S_total)
    scale = float(N) / float((B if B_total is None else B_total) * S_tot)
    zeros = lambda *s: torch.zeros(*s, dtype=f64, device=dev)
    ns_max = sample_chunk or default_sample_chunk(S, B, Q, D)

    pd_info = torch.zeros(1, dtype=torch.int32, device=dev)     # one deferred positive-definiteness check per step
    hyp = ops.hyper_exp(torch.stack([p[k].detach().reshape(()) for k in HYPER_ORDER]))
    ghyp = zeros(7)

    # ---- variational covariances: Sigma = tril(S) tril(S)^T, C = chol(Sigma + eps I) -------
    flat = packed_pair_index(D, dev) if pair_index is None else pair_index
    npair = flat.shape[0]
    SU = p["sqrt_U"].detach().reshape(D * D, Q, Q).index_select(0, flat)
    muU = p["mu_U"].detach().reshape(D * D, Q).index_select(0, flat)
    sqrt_v = p["sqrt_v"].detach().reshape(1, Q, Q)
    sqrt_W = p["sqrt_W"].detach()
    mu_W = p["mu_W"].detach()
    if latent_order is not None:
        assert not want_grads
        sqrt_W = sqrt_W.index_select(0, latent_order).contiguous()
        mu_W = mu_W.index_select(0, latent_order).contiguous()
    mu_v = p["mu_v"].detach()
    Sig_v = ops.tril_syrk_fwd(sqrt_v)
    Sig_W = ops.tril_syrk_fwd(sqrt_W)
    Sig_U = ops.tril_syrk_fwd(SU)
    C_v, hld_v = ops.potrf(Sig_v, EPS, info=pd_info)
    C_W, hld_W = ops.potrf(Sig_W, EPS, info=pd_info)
    rank_, world_ = kl_shard if kl_shard is not None else (0, 1)
    part = lambda n: slice((n * rank_) // world_, (n * (rank_ + 1)) // world_)
    sl1 = part(D)                                             # diagonal pairs handled here
    sl0 = slice(D + part(npair - D).start, D + part(npair - D).stop)   # strictly-lower pairs handled here
    slS = part(S) if S_total is None else slice(0, S)         # samples whose KL_W is evaluated here
    w_sh = kl_weight if kl_shard is None else 1.0
    n1, n0, nS = sl1.stop - sl1.start, sl0.stop - sl0.start, slS.stop - slS.start

    # ---- the three stationary inducing systems ------------------------------------------------
    sysm = {}
    systems = (("ell", H_S2_ELL, H_LEN_ELL), ("L0", H_S2_L0, H_LEN_L0), ("L1", H_S2_L1, H_LEN_L1))
    # one batched factorisation for the three Q x Q systems (each launch of a single small matrix is pure latency)
    A3 = torch.stack([ops.rbf_build_fwd(Z, Z, hyp, is2, ilen, EPS) for _, is2, ilen in systems])
    R3, hldR3 = ops.potrf(A3, 0.0, info=pd_info)
    for k_sys, (name, is2, ilen) in enumerate(systems):
        R, hldR = R3[k_sys:k_sys + 1], hldR3[k_sys:k_sys + 1]
        K12 = ops.rbf_build_fwd(x, Z, hyp, is2, ilen, 0.0).reshape(1, B, Q)
        P, c = ops.solve_rows_fwd(K12, R)
        sysm[name] = dict(R=R, hldR=hldR, K12=K12, P=P, c=c, is2=is2, ilen=ilen)
    sd_ell = ops.ell_sd_fwd(sysm["ell"]["c"][0], hyp)
    seg = ops.segment_offsets(I, D)
    # 64 < Q <= 128: padded records of the covariances for the ring-pipelined DMMA kernels, built once per step
    recU = ops.lq_pad_records(Sig_U) if Q >= ops.LQ_MIN_Q else None
    recW = ops.lq_pad_records(Sig_W) if Q >= ops.LQ_MIN_Q else None
    qU, mU = ops.quadform_fwd(sysm["L0"]["P"], sysm["L1"]["P"], I, Sig_U, muU, D, MODE_U, seg=seg, rec=recU)
    sdU = ops.coef_sd_fwd(qU[0], sysm["L0"]["c"][0], sysm["L1"]["c"][0], I, hyp)

    # ---- per-sample inducing draws, Gibbs K22 factor -----------------------------------------
    v, ellZ = ops.sample_v_fwd(mu_v, C_v[0], z_v)
    A_G = ops.gibbs_build_fwd(Z, Z, ellZ, ellZ, EPS)
    R_G, hld_G = ops.potrf(A_G, 0.0, info=pd_info)

    # The KL terms touch only Q x Q matrices (n_K of them): latency-bound work that is independent of the row kernels
    # of the sample loop, so it runs on a side stream and overlaps them (it is the serial fraction under row sharding).
    side = _side_stream(dev) if x.is_cuda else None
    main = torch.cuda.current_stream(dev) if x.is_cuda else None
    if side is not None:
        side.wait_stream(main)
    with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
        # ---- KL terms (reference-exact form, quirk q10) and their cotangents ----------------------
        if n1:
            C_U1, hld_U1 = ops.potrf(Sig_U[sl1], EPS, info=pd_info)
        if n0:
            C_U0, hld_U0 = ops.potrf(Sig_U[sl0], EPS, info=pd_info)
        kl_v, t_v = ops.kl_fwd(C_v, hld_v, mu_v.reshape(1, Q), sysm["ell"]["R"], sysm["ell"]["hldR"], exact=exact_kl)
        loss_kl = kl_weight * kl_v.sum()
        kl_W = None
        if nS:
            kl_W, t_W = ops.kl_fwd(C_W, hld_W, mu_W, R_G[slS], hld_G[slS], exact=exact_kl)
            loss_kl = loss_kl + w_sh * kl_W.sum() / S_tot
        klU_sum = zeros(())
        if n1:
            kl_U1, t_U1 = ops.kl_fwd(C_U1, hld_U1, muU[sl1], sysm["L1"]["R"], sysm["L1"]["hldR"], exact=exact_kl)
            klU_sum = klU_sum + kl_U1.sum()
        if n0:
            kl_U0, t_U0 = ops.kl_fwd(C_U0, hld_U0, muU[sl0], sysm["L0"]["R"], sysm["L0"]["hldR"], exact=exact_kl)
            klU_sum = klU_sum + kl_U0.sum()
        loss_kl = loss_kl + w_sh * klU_sum

        full = lambda shape, val: torch.full(shape, val, dtype=f64, device=dev)
        if want_grads:
            CWbar = zeros(D, Q, Q); hldWbar = zeros(D); muWbar = zeros(D, Q)
            RGbar = zeros(S, Q, Q); hldGbar = zeros(S)
            if nS:
                a, b, c_, rg, hg = ops.kl_bwd(full((nS, D), w_sh / S_tot), C_W, mu_W, R_G[slS], t_W, exact=exact_kl)
                CWbar, hldWbar, muWbar = a, b, c_
                RGbar[slS] = rg; hldGbar[slS] = hg
            Cvbar, hldvbar, muvbar, Rellbar, hldRellbar = ops.kl_bwd(full((1, 1), kl_weight), C_v, mu_v.reshape(1, Q),
                                                                     sysm["ell"]["R"], t_v, exact=exact_kl)
            CUbar1 = hldUbar1 = CUbar0 = hldUbar0 = None
            muUbar = zeros(npair, Q)
            RL1bar = zeros(1, Q, Q); hldRL1bar = zeros(1); RL0bar = zeros(1, Q, Q); hldRL0bar = zeros(1)
            if n1:
                CUbar1, hldUbar1, c_, RL1bar, hldRL1bar = ops.kl_bwd(full((1, n1), w_sh), C_U1, muU[sl1], sysm["L1"]["R"], t_U1, exact=exact_kl)
                muUbar[sl1] = c_
            if n0:
                CUbar0, hldUbar0, c_, RL0bar, hldRL0bar = ops.kl_bwd(full((1, n0), w_sh), C_U0, muU[sl0], sysm["L0"]["R"], t_U0, exact=exact_kl)
                muUbar[sl0] = c_
            muvbar = muvbar.reshape(Q).clone()
            # Cholesky adjoints that depend on the KL terms only
            AG_kl = ops.potrf_bwd(R_G, RGbar, hldGbar)
            SigW_kl = ops.potrf_bwd(C_W, CWbar, hldWbar)
            SigU_kl1 = ops.potrf_bwd(C_U1, CUbar1, hldUbar1) if n1 else None
            SigU_kl0 = ops.potrf_bwd(C_U0, CUbar0, hldUbar0) if n0 else None

    # ---- accumulators filled by the sample loop ------------------------------------------------
    SigWbar = zeros(D, Q, Q); muWrows = zeros(D, Q)     # row-kernel parts; the KL parts join after the side stream
    Pellbar = zeros(B, Q); sdellbar = zeros(B)
    mUbar = zeros(B, D); sdUbar = zeros(B, D)
    AGbar = zeros(S, Q, Q); ellZbar = zeros(S, Q); vbar = zeros(S, Q)
    Rsum = zeros(S)
    P_ell = sysm["ell"]["P"][0]

    for s0 in range(0, S, ns_max):
        sl = slice(s0, min(S, s0 + ns_max))
        ellx = ops.ell_rows_fwd(P_ell, v[sl], z_ell[sl], sd_ell)
        if z_L is not None:
            zl_, nz_ = z_L[sl], None
        else:
            zl_, nz_ = None, (int(noise_key[0]), int(noise_key[1]), sl.start + int(sample_offset), sl.stop - sl.start, row_gid,
                          noise_key[2] if len(noise_key) > 2 else None)
        l = ops.coef_sample_fwd(mU[0], sdU, zl_, I, noise=nz_)
        KG = ops.gibbs_build_fwd(x, Z, ellx, ellZ[sl], 0.0)
        PG, cG = ops.solve_rows_fwd(KG, R_G[sl])
        lbar, mgbar, qgbar, cGbar, PGbar = ops.latent_fused(PG, cG, l, y if y.dim() == 1 else y[sl], I, Sig_W, mu_W, hyp,
                                                            scale, Rsum[sl], ghyp, seg=seg, rec=recW)
        if not want_grads:
            continue
        ops.weighted_gram(PG, PG, I, qgbar, mgbar, MODE_W, SigWbar, muWrows, seg=seg)
        KGbar = ops.solve_rows_bwd(PGbar, cGbar, KG, PG, R_G[sl], AGbar[sl])
        ellxbar = torch.empty_like(ellx)
        ops.gibbs_build_bwd(x, Z, ellx, ellZ[sl], KGbar, ellxbar, ellZbar[sl], Kfwd=KG)
        ops.ell_rows_bwd(ellxbar, ellx, P_ell, v[sl], z_ell[sl], vbar[sl], Pellbar, sdellbar)
        ops.coef_sample_bwd(lbar, l, zl_, I, mUbar, sdUbar, noise=nz_)

    if side is not None:
        main.wait_stream(side)
        for t_ in (loss_kl, kl_v, kl_W):
            if t_ is not None:
                t_.record_stream(main)
    loss = loss_kl - scale * Rsum.sum()
    if aux is not None:
        aux.update(Rsum=Rsum, kl_W=kl_W, kl_v=kl_v.sum(), kl_U=klU_sum)
    if not want_grads:
        if defer_pd_check and aux is not None:
            aux["pd_info"] = pd_info
        else:
            ops.raise_if_not_pd(pd_info)
        return loss, None

    # ---- backward of the per-sample small stage ------------------------------------------------
    if side is not None:
        for t_ in (AG_kl, SigW_kl, SigU_kl1, SigU_kl0, muWbar, muUbar, muvbar, Cvbar, hldvbar, Rellbar, hldRellbar,
                   RL1bar, hldRL1bar, RL0bar, hldRL0bar):
            if t_ is not None:
                t_.record_stream(main)
    AGbar += AG_kl
    muWbar = muWbar + muWrows
    tmp = torch.empty_like(ellZbar)
    ops.gibbs_build_bwd(Z, Z, ellZ, ellZ, AGbar, tmp, ellZbar)
    ellZbar += tmp
    Cvbar = Cvbar.reshape(Q, Q).clone()
    ops.sample_v_bwd(ellZbar, vbar, ellZ, z_v, muvbar, Cvbar)

    # ---- backward of the sample-independent coefficient statistics ------------------------------
    qUbar, cL0bar, cL1bar = ops.coef_sd_bwd(sdUbar, sdU, I, hyp, ghyp)
    PL0bar, PL1bar = ops.quadform_bwd(sysm["L0"]["P"], sysm["L1"]["P"], I, Sig_U, muU,
                                      qUbar.reshape(1, B, D), mUbar.reshape(1, B, D), MODE_U, seg=seg, rec=recU)
    SigUbar = zeros(npair, Q, Q)
    ops.weighted_gram(sysm["L0"]["P"], sysm["L1"]["P"], I, qUbar.reshape(1, B, D), mUbar.reshape(1, B, D),
                      MODE_U, SigUbar, muUbar, seg=seg)
    cellbar = ops.ell_sd_bwd(sdellbar, sd_ell, hyp, ghyp)

    # Cholesky adjoints of the three stationary factors and of C_v in one batched launch
    chol_adj = ops.potrf_bwd(torch.cat([sysm["ell"]["R"], sysm["L0"]["R"], sysm["L1"]["R"], C_v.reshape(1, Q, Q)]),
                             torch.cat([Rellbar.reshape(1, Q, Q), RL0bar.reshape(1, Q, Q), RL1bar.reshape(1, Q, Q),
                                        Cvbar.reshape(1, Q, Q)]),
                             torch.cat([hldRellbar.reshape(1), hldRL0bar.reshape(1), hldRL1bar.reshape(1),
                                        hldvbar.reshape(1)]))
    for k_sys, (name, Pbar, cbar) in enumerate((("ell", Pellbar.reshape(1, B, Q), cellbar.reshape(1, B)),
                                                ("L0", PL0bar, cL0bar.reshape(1, B)),
                                                ("L1", PL1bar, cL1bar.reshape(1, B)))):
        sy = sysm[name]
        Abar = zeros(1, Q, Q)
        K12bar = ops.solve_rows_bwd(Pbar, cbar, sy["K12"], sy["P"], sy["R"], Abar)
        Abar += chol_adj[k_sys:k_sys + 1]
        ops.rbf_build_bwd(x, Z, hyp, sy["is2"], sy["ilen"], K12bar[0], ghyp)
        ops.rbf_build_bwd(Z, Z, hyp, sy["is2"], sy["ilen"], Abar[0], ghyp)

    # ---- Cholesky / LL^T adjoints back to the sqrt parameters -----------------------------------
    g_sqrt_v = ops.tril_syrk_bwd(sqrt_v, chol_adj[3:4]).reshape(Q, Q)
    SigWbar += SigW_kl
    g_sqrt_W = ops.tril_syrk_bwd(sqrt_W, SigWbar)
    if n1:
        SigUbar[sl1] += SigU_kl1
    if n0:
        SigUbar[sl0] += SigU_kl0
    g_SU = ops.tril_syrk_bwd(SU, SigUbar)
    g_sqrt_U = zeros(D * D, Q, Q).index_copy_(0, flat, g_SU).reshape(D, D, Q, Q)
    g_mu_U = zeros(D * D, Q).index_copy_(0, flat, muUbar).reshape(D, D, Q)

    grads = {"mu_W": muWbar, "sqrt_W": g_sqrt_W, "mu_v": muvbar, "sqrt_v": g_sqrt_v,
             "mu_U": g_mu_U, "sqrt_U": g_sqrt_U}
    for i, k in enumerate(HYPER_ORDER):
        grads[k] = ghyp[i]
    # the only host synchronisation of the step, after everything is enqueued (torch.cholesky's RuntimeError)
    if defer_pd_check and aux is not None:
        aux["pd_info"] = pd_info
    else:
        ops.raise_if_not_pd(pd_info)
    return loss, grads
