"""world_size-2 gloo run of the row-sharded step on the CPU (kernel specs patched in for the C-ABI wrappers):
sum over ranks of (loss, gradients) after the single all-reduce == the single-process full-batch step."""
import inspect
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import golden_util as gu


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import kernel_specs as specs
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, nmgp_dsvi, parallel
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            setattr(_ops, n, f)
    g = gu.load("dsvi_ragged")
    D = int(g["D"])
    p = gu.case_params(g)
    model = nmgp_dsvi.NMGP(int(g["N"]), D, torch.from_numpy(g["Z"]).view(-1, 1), device="cpu")
    model.load_state_dict(p)
    counts = [int((g["I"] == d).sum()) for d in range(D)]
    rows = parallel.shard_rows_per_output(counts, rank, world)
    starts = np.cumsum([0] + counts[:-1])
    sel = np.concatenate([starts[d] + rows[d] for d in range(D)]).astype(np.int64)
    B = g["x"].shape[0]
    parallel.configure_model_for_sharding(model, B, rank, world)
    x = torch.from_numpy(g["x"][sel]); y = torch.from_numpy(g["y"][sel])
    I = torch.from_numpy(g["I"][sel].astype(np.int32))
    noise = (torch.from_numpy(g["z_v"]), torch.from_numpy(g["z_ell"][:, sel]), torch.from_numpy(g["z_L"][:, sel]))
    loss = model.forward_rows(x, y, I, explicit_noise=noise)
    loss.backward()
    tot = parallel.allreduce_loss_and_grads(loss, list(model.parameters()))
    if rank == 0:
        res = {"loss": float(tot)}
        for k, prm in model.named_parameters():
            res[k] = prm.grad.detach().numpy().copy()
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_row_sharding_matches_reference(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    g = gu.load("dsvi_ragged")
    assert abs(res["loss"] - float(g["loss"])) <= 1e-9 * abs(float(g["loss"]))
    for k in ("mu_W", "sqrt_W", "mu_v", "sqrt_v", "mu_U", "sqrt_U", "sigma2_err_log", "length_scales_L0_log"):
        gu.check_grad(k, res[k], g, 1e-9)


def test_shard_rows_cover_every_row_once():
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import parallel
    counts = [5, 0, 13, 1]
    for world in (1, 2, 3, 8):
        seen = [np.concatenate([parallel.shard_rows_per_output(counts, r, world)[d] for r in range(world)]) for d in range(4)]
        for d, c in enumerate(counts):
            assert sorted(seen[d].tolist()) == list(range(c))
