"""Diagnostic: host time to enqueue one resident DSVI step (no host sync inside) vs its GPU time."""
import sys, time, json
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import numpy as np, torch
import bench
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, nmgp_dsvi, parallel, dsvi_step
frac = int(sys.argv[1]) if len(sys.argv) > 1 else 1      # emulate a 1/frac row shard
w = dict(bench.WORKLOADS["ecog"])
T, D, Q, S = w["T"], w["D"], w["Q"], w["S"]
X_list, Y_list, z = bench.synthetic_problem(w)
dev = torch.device("cuda:0")
model = nmgp_dsvi.NMGP(T * D, D, torch.from_numpy(z).view(-1, 1), mu_v=np.ones(Q), seed=22, device=dev, noise="device")
for k, v in w["hyper"].items():
    getattr(model, k).data.fill_(v)
for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
    getattr(model, k).requires_grad = False
opt = torch.optim.Adam(model.parameters(), lr=0.005)
parallel.configure_model_for_sharding(model, T * D, 0, frac)
rows = parallel.shard_rows_per_output([T] * D, 0, frac)
xd = torch.cat([X_list[d][rows[d]] for d in range(D)]).to(dev); yd = torch.cat([Y_list[d][rows[d]] for d in range(D)]).to(dev)
Id = torch.from_numpy(np.repeat(np.arange(D, dtype=np.int32), [len(r) for r in rows])).to(dev)
gid = torch.from_numpy(parallel.global_row_ids([T] * D, rows)).to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = model.forward_rows(xd, yd, Id, n_mc=S, row_gid=gid)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
orig = _ops.raise_if_not_pd
for name, fn in (("with_sync", orig), ("no_sync", lambda info: None)):
    _ops.raise_if_not_pd = fn
    dsvi_step.ops.raise_if_not_pd = fn
    torch.cuda.synchronize()
    hs = []; t_all0 = time.perf_counter()
    for _ in range(5):
        t0 = time.perf_counter(); step(); hs.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t_all0) * 1e3 / 5
    print(json.dumps({"shard": "1/%d" % frac, "mode": name, "host_ms_per_step": [round(h, 2) for h in hs], "wall_ms_per_step": round(tot, 2)}))
