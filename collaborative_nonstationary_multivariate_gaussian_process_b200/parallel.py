"""Multi-GPU plumbing for the DSVI step: one process per GPU, rows of the minibatch sharded across ranks,
ONE all-reduce (NCCL over NVLink/NVSwitch, sum, float64) of the packed gradient + loss per step.

The reference has no distributed code; the decomposition follows SURVEY.md 8e: every
adjoint is linear in its cotangent, so each rank can run the complete backward on its row shard with the
KL terms weighted 1/world_size, and the sum over ranks is the full-batch gradient.  Parameters stay
replicated; every rank applies the identical Adam update.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_rows_per_output(counts: Sequence[int], rank: int, world: int) -> List[np.ndarray]:
    """Row indices (within each output) owned by ``rank``: a stride-``world`` slice of every output, so each
    rank sees every output and the per-row cost (proportional to I[n]+1) is balanced."""
    return [np.arange(rank, int(c), world) for c in counts]


def shard_samples(n_samples: int, rank: int, world: int):
    """Contiguous block of samples (MC draws, or the subjects of an HCP-style step) owned by ``rank``."""
    return (n_samples * rank) // world, (n_samples * (rank + 1)) // world


def configure_model_for_sharding(model, total_rows: int, rank: int, world: int, shard: str = "rows",
                                 n_samples_total: int = None):
    """shard="rows" (default): rows of the minibatch are dealt to the ranks -- global row count for the N/B scaling, KL
    terms split over the ranks.  shard="samples": every rank keeps all rows but only its block of the ``n_samples_total``
    samples / subjects (``shard_samples``); the sample-independent coefficient statistics are then replicated (1/S_local
    of a rank's work).  Either way the device noise is counter-based and keyed by global row and sample ids, so the
    estimate does not depend on the number of ranks."""
    model.step_options = dict(B_total=int(total_rows), kl_weight=1.0 / world,
                              kl_shard=(rank, world) if world > 1 else None,
                              defer_pd_check=world > 1)      # a Cholesky failure is raised collectively, after the all-reduce
    if shard == "samples":
        lo, hi = shard_samples(int(n_samples_total), rank, world)
        model.step_options.update(S_total=int(n_samples_total), sample_offset=lo)
    elif shard != "rows":
        raise ValueError("shard must be 'rows' or 'samples'")
    for name in ("mu_U", "sqrt_U"):      # [D, D, ...] with exact-zero gradient blocks for j > i: reduced in packed form
        prm = getattr(model, name, None)
        if prm is not None and prm.dim() >= 2 and prm.shape[0] == prm.shape[1]:
            prm._nmgp_pair_D = int(prm.shape[0])


def global_row_ids(counts: Sequence[int], rows_per_output: Sequence[np.ndarray]) -> np.ndarray:
    """Global ids (position in the unsharded, output-sorted minibatch) of a rank's rows."""
    starts = np.cumsum([0] + [int(c) for c in counts[:-1]])
    return np.concatenate([starts[d] + np.asarray(r, dtype=np.int64) for d, r in enumerate(rows_per_output)]).astype(np.int64)


_pending_pd = []          # reduced PD flags not yet looked at (check="defer")


def raise_if_pending_not_pd():
    """Host check of the PD flags that ``allreduce_loss_and_grads(..., check="defer")`` left on the device."""
    flags, _pending_pd[:] = list(_pending_pd), []
    for f in flags:
        if float(f) != 0.0:
            raise RuntimeError("cholesky: a matrix of the step is not positive-definite on %d rank(s)" % int(round(float(f))))


def allreduce_loss_and_grads(loss: torch.Tensor, params: Sequence[torch.nn.Parameter], group=None,
                             pd_info=None, check: str = "sync") -> torch.Tensor:
    """Sum loss and all gradients over ranks with a single collective on one flat float64 buffer.

    ``pd_info`` (the step's deferred positive-definiteness flag, ``model._last_pd_info``) rides in the same buffer:
    under ``kl_shard`` every rank factorises only its slice of the coefficient covariances, so a failure is seen by
    one rank only; raising it there before the collective would leave the other ranks hanging in the all-reduce.
    After the reduction every rank sees a non-zero slot and raises the same RuntimeError the reference's
    torch.cholesky would.  ``check="defer"`` skips the host read (one synchronisation per step) and keeps the reduced
    flag for ``raise_if_pending_not_pd()``.

    Parameters tagged by `configure_model_for_sharding` as coefficient-pair tensors (`mu_U`, `sqrt_U`: [D, D, ...]) only
    travel as their live (i, j <= i) blocks -- the other blocks are exact zeros on every rank (reference quirk q8) --
    which cuts the buffer from 84.9 MB to 22.6 MB at D = 64, Q = 50."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return loss.detach()
    from .dsvi_step import packed_pair_index
    live = [p for p in params if p.grad is not None]
    pd_slot = (pd_info != 0).to(loss.dtype).reshape(1) if pd_info is not None else torch.zeros_like(loss.detach().reshape(1))
    chunks, plan = [loss.detach().reshape(1), pd_slot], []
    for p in live:
        Dp = getattr(p, "_nmgp_pair_D", None)
        if Dp is not None and p.grad.dim() >= 2 and p.grad.shape[0] == Dp and p.grad.shape[1] == Dp:
            idx = packed_pair_index(Dp, p.grad.device)
            rows = p.grad.view(Dp * Dp, -1).index_select(0, idx)
            plan.append((p, idx, rows.shape))
            chunks.append(rows.reshape(-1))
        else:
            plan.append((p, None, None))
            chunks.append(p.grad.reshape(-1))
    flat = torch.cat(chunks)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 2
    for p, idx, shp in plan:
        if idx is None:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        else:
            n = shp[0] * shp[1]
            p.grad.view(p.grad.shape[0] * p.grad.shape[1], -1).index_copy_(0, idx, flat[off:off + n].view(shp))
        off += n
    if pd_info is not None:
        _pending_pd.append(flat[1])
        if check == "sync":                                   # same decision on every rank (host read after the collective)
            raise_if_pending_not_pd()
    return flat[0]


def kron_logpdf0_sharded(y, mu, B, K, sigma2, group=None):
    """multivariate_normal_logpdf0 with the D eigen-blocks (sigma2 I + lambda_m K) dealt round-robin to the ranks: every
    rank factorises its blocks with the blocked Cholesky and the partial log-densities (-1/2 logdet_m - 1/2 quad_m) are
    summed with one all-reduce of a single double (scale sweep).  A single T x T factorisation is not split."""
    import torch.distributed as dist
    from . import distributions
    if not (dist.is_available() and dist.is_initialized()):
        return distributions.multivariate_normal_logpdf0(y, mu, B, K, sigma2)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    part = distributions.multivariate_normal_logpdf0(y, mu, B, K, sigma2, shard=(rank, world)).reshape(1).clone()
    dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part.reshape(())
