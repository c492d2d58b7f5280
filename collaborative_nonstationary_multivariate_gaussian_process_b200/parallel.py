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


def configure_model_for_sharding(model, total_rows: int, rank: int, world: int):
    """Row sharding: global row count for the N/B scaling, KL terms split over the ranks.  The device noise is
    counter-based and keyed by global row ids (pass ``row_gid`` to ``forward_rows``), so no per-rank generator state."""
    model.step_options = dict(B_total=int(total_rows), kl_weight=1.0 / world,
                              kl_shard=(rank, world) if world > 1 else None)


def global_row_ids(counts: Sequence[int], rows_per_output: Sequence[np.ndarray]) -> np.ndarray:
    """Global ids (position in the unsharded, output-sorted minibatch) of a rank's rows."""
    starts = np.cumsum([0] + [int(c) for c in counts[:-1]])
    return np.concatenate([starts[d] + np.asarray(r, dtype=np.int64) for d, r in enumerate(rows_per_output)]).astype(np.int64)


def allreduce_loss_and_grads(loss: torch.Tensor, params: Sequence[torch.nn.Parameter], group=None) -> torch.Tensor:
    """Sum loss and all gradients over ranks with a single collective on one flat float64 buffer."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return loss.detach()
    live = [p for p in params if p.grad is not None]
    flat = torch.cat([loss.detach().reshape(1)] + [p.grad.reshape(-1) for p in live])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 1
    for p in live:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
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
