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


def _nonpd_worker(rank, world, port, outdir):
    """One coefficient covariance that only rank 1 factorises (kl_shard) is poisoned: BOTH ranks must raise after the
    all-reduce (a rank raising alone before it would leave the other one hanging in the collective)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import datetime
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=60))
    from oracle import kernel_specs as specs
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, nmgp_dsvi, parallel
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            setattr(_ops, n, f)
    g = gu.load("dsvi_ragged")
    D = int(g["D"])
    model = nmgp_dsvi.NMGP(int(g["N"]), D, torch.from_numpy(g["Z"]).view(-1, 1), device="cpu")
    model.load_state_dict(gu.case_params(g))
    with torch.no_grad():
        model.sqrt_U[D - 1, D - 2].fill_(float("nan"))        # last strictly-lower pair: rank 1's slice
    counts = [int((g["I"] == d).sum()) for d in range(D)]
    rows = parallel.shard_rows_per_output(counts, rank, world)
    starts = np.cumsum([0] + counts[:-1])
    sel = np.concatenate([starts[d] + rows[d] for d in range(D)]).astype(np.int64)
    parallel.configure_model_for_sharding(model, g["x"].shape[0], rank, world)
    noise = (torch.from_numpy(g["z_v"]), torch.from_numpy(g["z_ell"][:, sel]), torch.from_numpy(g["z_L"][:, sel]))
    loss = model.forward_rows(torch.from_numpy(g["x"][sel]), torch.from_numpy(g["y"][sel]),
                              torch.from_numpy(g["I"][sel].astype(np.int32)), explicit_noise=noise)
    local_flag = int(model._last_pd_info)
    loss.backward()
    raised = False
    try:
        parallel.allreduce_loss_and_grads(loss, list(model.parameters()), pd_info=model._last_pd_info)
    except RuntimeError as e:
        raised = "positive-definite" in str(e)
    torch.save({"raised": raised, "local_flag": local_flag}, os.path.join(outdir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_nonpd_is_raised_on_every_rank(tmp_path):
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_nonpd_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(str(tmp_path / ("r%d.pt" % r)), weights_only=False) for r in range(2)]
    assert res[0]["raised"] and res[1]["raised"]
    assert res[0]["local_flag"] == 0 and res[1]["local_flag"] != 0      # only one rank saw it locally


def test_shard_rows_cover_every_row_once():
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import parallel
    counts = [5, 0, 13, 1]
    for world in (1, 2, 3, 8):
        seen = [np.concatenate([parallel.shard_rows_per_output(counts, r, world)[d] for r in range(world)]) for d in range(4)]
        for d, c in enumerate(counts):
            assert sorted(seen[d].tolist()) == list(range(c))


def test_device_noise_is_independent_of_the_sharding(monkeypatch):
    """Counter-based noise keyed by global row ids: the sum over two row shards equals the unsharded step exactly
    (same draws), which is what makes 1/2/4/8-GPU runs agree."""
    from oracle import kernel_specs as specs
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, nmgp_dsvi, parallel
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            monkeypatch.setattr(_ops, n, f)
    g = gu.load("dsvi_ragged")
    D = int(g["D"]); B = g["x"].shape[0]
    counts = [int((g["I"] == d).sum()) for d in range(D)]

    def run(world):
        tot, grads = 0.0, None
        for rank in range(world):
            model = nmgp_dsvi.NMGP(int(g["N"]), D, torch.from_numpy(g["Z"]).view(-1, 1), device="cpu", noise="device")
            model.load_state_dict(gu.case_params(g))
            model.noise_seed = 1234
            rows = parallel.shard_rows_per_output(counts, rank, world)
            gid = parallel.global_row_ids(counts, rows)
            if world > 1:
                parallel.configure_model_for_sharding(model, B, rank, world)
            loss = model.forward_rows(torch.from_numpy(g["x"][gid]), torch.from_numpy(g["y"][gid]),
                                      torch.from_numpy(g["I"][gid].astype(np.int32)), n_mc=3, row_gid=torch.from_numpy(gid))
            loss.backward()
            tot += float(loss)
            gr = {k: p.grad.clone() for k, p in model.named_parameters()}
            grads = gr if grads is None else {k: grads[k] + gr[k] for k in gr}
        return tot, grads
    l1, g1 = run(1)
    l2, g2 = run(2)
    assert abs(l1 - l2) <= 1e-12 * abs(l1)
    for k in g1:
        den = max(float(torch.linalg.norm(g1[k])), 1e-300)
        assert float(torch.linalg.norm(g1[k] - g2[k])) / den <= 1e-10, k


def test_counter_noise_statistics():
    from oracle import kernel_specs as specs
    z = specs.noise_fill(4, 5000, 7, seed=99, stream_id=(3 << 8) | 2, s0=10, gid=None)
    assert z.shape == (4, 5000, 7)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.var()) - 1.0) < 0.02
    z2 = specs.noise_fill(2, 5000, 7, seed=99, stream_id=(3 << 8) | 2, s0=12, gid=None)
    assert torch.equal(z[2:], z2)                      # sample offset only shifts the counter
    assert float(z.to(torch.float32).to(torch.float64).sub(z).abs().max()) == 0.0   # float32-valued (quirk q2)


def _kron_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import kernel_specs as specs
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, parallel
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            setattr(_ops, n, f)
    g = gu.load("sim_code")
    val = parallel.kron_logpdf0_sharded(torch.from_numpy(g["y"]), torch.from_numpy(g["mu"]), torch.from_numpy(g["Bf"]),
                                        torch.from_numpy(g["K_self"]), torch.tensor(float(g["s2"]), dtype=torch.float64))
    if rank == 0:
        torch.save({"logpdf0": float(val)}, out)
    dist.destroy_process_group()


def test_two_rank_kronecker_logpdf_matches_reference(tmp_path):
    """eigen-blocks of B dealt to two ranks, one all-reduce: equals the reference's multivariate_normal_logpdf0"""
    out = str(tmp_path / "kron.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_kron_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    g = gu.load("sim_code")
    assert abs(res["logpdf0"] - float(g["logpdf0"])) <= 1e-9 * abs(float(g["logpdf0"]))


def _subject_worker(rank, world, port, out):
    """HCP-style step: two subjects observed on the same rows, one Monte-Carlo draw each; rank r keeps all rows and
    subject r only (shard="samples").  Rank 0 also evaluates the unsharded two-subject step for comparison."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import kernel_specs as specs
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, nmgp_dsvi, parallel
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            setattr(_ops, n, f)
    g = gu.load("dsvi_ragged")
    D, B = int(g["D"]), g["x"].shape[0]
    S = 2
    gen = torch.Generator().manual_seed(9)
    x = torch.from_numpy(g["x"]); I = torch.from_numpy(g["I"].astype(np.int32))
    Y = torch.stack([torch.from_numpy(g["y"]), torch.from_numpy(g["y"]) + 0.3 * torch.randn(B, generator=gen, dtype=torch.float64)])
    Q = g["Z"].shape[0]
    zv = torch.randn(S, Q, generator=gen, dtype=torch.float64)
    zell = torch.randn(S, B, generator=gen, dtype=torch.float64)
    zL = torch.randn(S, B, D, generator=gen, dtype=torch.float64)

    def fresh():
        m = nmgp_dsvi.NMGP(int(g["N"]), D, torch.from_numpy(g["Z"]).view(-1, 1), device="cpu")
        m.load_state_dict(gu.case_params(g))
        return m
    model = fresh()
    parallel.configure_model_for_sharding(model, B, rank, world, shard="samples", n_samples_total=S)
    lo, hi = parallel.shard_samples(S, rank, world)
    loss = model.forward_rows(x, Y[lo:hi], I, explicit_noise=(zv[lo:hi], zell[lo:hi], zL[lo:hi]))
    loss.backward()
    tot = parallel.allreduce_loss_and_grads(loss, list(model.parameters()), pd_info=model._last_pd_info)
    if rank == 0:
        ref = fresh()
        lref = ref.forward_rows(x, Y, I, explicit_noise=(zv, zell, zL))
        lref.backward()
        res = {"loss": float(tot), "loss_ref": float(lref)}
        for (k, prm), (_, pr) in zip(model.named_parameters(), ref.named_parameters()):
            res[k] = prm.grad.detach().numpy().copy()
            res["ref_" + k] = pr.grad.detach().numpy().copy()
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_subject_sharding_matches_unsharded(tmp_path):
    out = str(tmp_path / "res_subjects.pt")
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_subject_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert abs(res["loss"] - res["loss_ref"]) <= 1e-12 * abs(res["loss_ref"])
    for k in ("mu_W", "sqrt_W", "mu_v", "sqrt_v", "mu_U", "sqrt_U", "sigma2_err_log", "sigma2_L0_log"):
        a, b = res[k].reshape(-1), res["ref_" + k].reshape(-1)
        assert np.linalg.norm(a - b) <= 1e-11 * max(np.linalg.norm(b), 1e-300), k
