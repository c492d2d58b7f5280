#!/usr/bin/env python
"""Benchmark of the NMGP DSVI hot path (BASELINE.json metric: DSVI iterations/s, ELBO + gradient + Adam,
at the ECoG shape T=4096, D("M")=64, S=32, Q=50, full batch B = T*D = 262144; FP64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the reference algorithm on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DSVI iters/sec (ELBO+grad) at T=4096,M=64,S=32"
UNIT = "iters/s"
WORKLOADS = {
    # name: T, D, Q, S, hyper (driver settings)
    "ecog": dict(T=4096, D=64, Q=50, S=32, hyper={"length_scales_L0_log": 10., "length_scales_L1_log": 10.,
                                                    "length_scales_tildeell_log": 5., "sigma2_err_log": -5.}),
    "pm25": dict(T=2048, D=16, Q=100, S=8, hyper={"length_scales_L0_log": 10., "length_scales_L1_log": 10.,
                                                   "length_scales_tildeell_log": 10.}),
    "tiny": dict(T=256, D=4, Q=20, S=2, hyper={"length_scales_L0_log": 3., "length_scales_L1_log": 3.,
                                                "length_scales_tildeell_log": 2., "sigma2_err_log": -2.}),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ecog", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-baseline", default="auto", choices=["auto", "skip"])
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def synthetic_problem(w, seed=0):
    """Shared grid X_d = arange(T) (as in the ECoG/HCP drivers, NMGP_ECoG_full.py:108-110), Y = N(0,1) draws,
    Z = linspace(0, T-1, Q); float64, generated on the CPU with a fixed seed."""
    g = torch.Generator().manual_seed(seed)
    T, D, Q = w["T"], w["D"], w["Q"]
    X_list = [torch.arange(T, dtype=torch.float64) for _ in range(D)]
    Y_list = [torch.randn(T, generator=g, dtype=torch.float64) for _ in range(D)]
    z = np.linspace(0, T - 1, Q)
    return X_list, Y_list, z


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
def reference_step_fn(w, Bs, seed=0):
    """One iteration (forward + autograd backward + Adam, S=1) of the CPU port of the reference algorithm
    (oracle/nmgp_oracle.py: same operation order as code/nmgp_dsvi.py:157-301,847,854, including the D(D+1)/2-call
    MGP_d loop) on Bs random rows of the workload's T x D grid, all D channels, the workload's Q and hyper-parameters."""
    from oracle import nmgp_oracle as orc
    T, D, Q = w["T"], w["D"], w["Q"]
    Bs = min(Bs, T * D)
    rng = np.random.default_rng(seed)
    p = orc.init_params(D, Q, seed=22, mu_v=np.ones(Q))
    for k, v in w["hyper"].items():
        p[k] = torch.tensor(float(v), dtype=torch.float64)
    pick = np.sort(rng.choice(T * D, size=Bs, replace=False))
    Xl = [torch.from_numpy((pick[(pick // T) == d] % T).astype(np.float64)).view(-1, 1) for d in range(D)]
    Yl = [torch.from_numpy(rng.standard_normal(x.shape[0])).view(-1, 1) for x in Xl]
    Z = torch.linspace(0, T - 1, Q, dtype=torch.float64).view(-1, 1)
    opt_state = {}

    def step():
        loss, grads = orc.step_loss_and_grads(p, Z, T * D, Xl, Yl)
        for k, gk in grads.items():                      # Adam as in code/nmgp_dsvi.py:854 (negligible cost)
            if gk is None:
                continue
            m, v, t = opt_state.get(k, (torch.zeros_like(gk), torch.zeros_like(gk), 0))
            t += 1
            m = 0.9 * m + 0.1 * gk; v = 0.999 * v + 0.001 * gk * gk
            p[k] = p[k] - 0.005 * (m / (1 - 0.9 ** t)) / ((v / (1 - 0.999 ** t)).sqrt() + 1e-8)
            opt_state[k] = (m, v, t)
        return float(loss)
    return step, Bs


def reference_estimate(w, reps_small=1, warm_small=0, B1=512, B2=2048):
    """Reference iterations/s on the full workload from a bounded sample: the reference's step time is
    t(B) = a + b*B (a: per-call overhead of its 2080-pair loop and autograd bookkeeping, independent of B; b: per-row
    arithmetic), measured at B1 (the ECoG driver's minibatch, NMGP_ECoG_full.py:288) and B2 rows; the full workload is
    S sequential forwards on B = T*D rows, so t_full = S * (a + b*T*D).  The reference cannot run the full batch
    itself (8.6 GB (D,D,B) tensor, ~45 s of autograd bookkeeping per call: SURVEY.md 6)."""
    T, D, S = w["T"], w["D"], w["S"]
    f1, B1 = reference_step_fn(w, B1)
    for _ in range(warm_small):
        f1()
    t1s = []
    for _ in range(max(1, reps_small)):
        t0 = time.perf_counter(); f1(); t1s.append(time.perf_counter() - t0)
    t1 = float(np.median(t1s))
    f2, B2 = reference_step_fn(w, B2, seed=1)
    t0 = time.perf_counter(); f2(); t2 = time.perf_counter() - t0
    b = max((t2 - t1) / max(B2 - B1, 1), 0.0)
    a = max(t1 - b * B1, 0.0)
    t_full = S * (a + b * T * D)
    desc = ("oracle port of the reference step (forward + autograd backward + Adam, S=1) on the T=%d x D=%d grid, Q=%d: "
            "measured %.2f s at B=%d rows and %.2f s at B=%d rows; model t(B)=a+b*B with a=%.2f s, b=%.3g s/row; full "
            "workload = S=%d forwards on B=%d rows -> %.0f s per iteration (extrapolated; the measured sample is %.4f "
            "iters/s at B=%d, S=1)" % (T, D, w["Q"], t1, B1, t2, B2, a, b, S, T * D, t_full, 1.0 / t1, B1))
    return {"value": 1.0 / t_full, "t_small": t1, "t_small_all": t1s, "t_large": t2, "B1": B1, "B2": B2, "a": a, "b": b,
            "t_full": t_full, "desc": desc}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # one reference call costs ~45 s at D=64 whatever the batch, so the timed region is clamped to keep
    # the run within a few minutes; the clamped counts are what the JSON reports
    k_eff, w_eff = 1, 0
    est = reference_estimate(w, reps_small=k_eff, warm_small=w_eff)
    line = {"impl": "reference", "metric": METRIC, "value": est["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": k_eff, "warmup": w_eff, "ms_per_step": est["t_full"] * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, w), "requested_steps": args.steps, "requested_warmup": args.warmup,
                       "sample_ms_per_step": est["t_small"] * 1e3},
            "cpu_baseline": {"value": est["value"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": est["desc"]},
            "e2e": {"value": est["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(args, w):
    return "%s-shaped DSVI step: T=%d, D=%d outputs, Q=%d inducing, S=%d MC samples, full batch B=%d, shared grid" % (
        args.workload, w["T"], w["D"], w["Q"], w["S"], w["T"] * w["D"])


# ------------------------------------------------------------------------------------------------------
def fp64_yardstick(dev, n=8192, reps=3):
    """cuBLAS DGEMM rate measured in this run: the FP64 roofline denominator (MEASURED_PEAKS.json has none)."""
    a = torch.randn(n, n, device=dev, dtype=torch.float64); b = torch.randn(n, n, device=dev, dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_b200(args, w):
    import torch.distributed as dist
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, nmgp_dsvi, parallel
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, D, Q, S = w["T"], w["D"], w["Q"], w["S"]
    X_list, Y_list, z = synthetic_problem(w)
    Btot = T * D
    model = nmgp_dsvi.NMGP(Btot, D, torch.from_numpy(z).view(-1, 1), mu_v=np.ones(Q), seed=22, device=dev, noise="device")
    for k, v in w["hyper"].items():
        getattr(model, k).data.fill_(v)
    for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
        getattr(model, k).requires_grad = False                       # fix_hyperpars=True, the drivers' setting
    opt = torch.optim.Adam(model.parameters(), lr=0.005)
    parallel.configure_model_for_sharding(model, Btot, rank, world)
    rows = parallel.shard_rows_per_output([T] * D, rank, world)
    Xh = [X_list[d][rows[d]].contiguous().pin_memory() for d in range(D)]
    Yh = [Y_list[d][rows[d]].contiguous().pin_memory() for d in range(D)]
    Bloc = sum(int(x.shape[0]) for x in Xh)
    xd = torch.cat(Xh).to(dev); yd = torch.cat(Yh).to(dev)
    Id = torch.from_numpy(np.repeat(np.arange(D, dtype=np.int32), [int(x.shape[0]) for x in Xh])).to(dev)
    gid = torch.from_numpy(parallel.global_row_ids([T] * D, rows)).to(dev)
    params = list(model.parameters())

    def step_resident():
        opt.zero_grad(set_to_none=True)
        loss = model.forward_rows(xd, yd, Id, n_mc=S, row_gid=gid)
        loss.backward()
        tot = parallel.allreduce_loss_and_grads(loss, params)
        opt.step()
        return tot

    def step_e2e():
        opt.zero_grad(set_to_none=True)
        loss = model(Xh, Yh, n_mc=S, noise="device", row_gid=gid)                   # host lists -> H2D inside
        loss.backward()
        tot = parallel.allreduce_loss_and_grads(loss, params)
        opt.step()
        return float(tot.cpu())                                       # D2H read of the step's result

    def timed(fn, k, prof=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if prof:
            _ops._profile_begin()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        profd = _ops._profile_end() if prof else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, profd

    peak = fp64_yardstick(dev) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    ms, last, (prof, ncalls) = timed(step_resident, args.steps, prof=True)
    clocks = clk.stop() if rank == 0 else None
    ms_step = ms / args.steps
    e2e = None
    if not args.no_e2e:
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        ms2, _, _ = timed(step_e2e, args.steps)
        e2e = {"value": 1e3 * args.steps / ms2, "unit": UNIT, "h2d_bytes_per_step": int(Bloc * (8 + 8 + 4)) * world,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": ms2 / args.steps}

    if rank == 0:
        # roofline of the dominant kernels: dense convention, 2 Q^2 flop per (row, used pair) quadratic form
        pairs_per_sample = float(sum((d + 1) * int(Xh[d].shape[0]) for d in range(D)))
        kern = {}
        # dense convention: 2 Q^2 flop per (row, used pair) for each of V = P Sigma, its adjoint and
        # the Gram accumulation; the fused latent kernel covers the first two for the S samples, the coefficient
        # (U) side runs once per step.
        for name, per_pair, nsamp in (("quadform_fwd", 2.0 * Q * Q, 1), ("quadform_bwd", 2.0 * Q * Q, 1),
                                      ("weighted_gram", 2.0 * Q * Q, S + 1), ("latent_fused", 4.0 * Q * Q, S)):
            if name in prof:
                calls, tms = prof[name]
                flops = args.steps * nsamp * pairs_per_sample * per_pair
                kern[name] = {"calls": calls, "ms_total": tms, "share_of_step": tms / ms,
                              "tflops": flops / (tms * 1e-3) / 1e12}
        others = {k: {"calls": c, "ms_total": t_, "share_of_step": t_ / ms} for k, (c, t_) in prof.items() if k not in kern}
        dom = max(kern, key=lambda k: kern[k]["ms_total"]) if kern else None
        roof = None
        if dom:
            # DRAM traffic per launch of the dominant kernel from the committed ncu --set full capture (profiles/)
            traffic = None
            try:
                summ = json.load(open(os.path.join(ROOT, "profiles", "ncu_r1_full_summary.json")))
                key = {"latent_fused": "k_latent_fused", "weighted_gram": "k_gram_mma"}.get(dom)
                for kname, rec in summ.items():
                    if key and kname.startswith(key):
                        scale_u = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                        ur = scale_u.get(rec["units"]["dram__bytes_read.sum"], 1.0)
                        uw = scale_u.get(rec["units"]["dram__bytes_write.sum"], 1.0)
                        traffic = rec["dram__bytes_read.sum"] * ur + rec["dram__bytes_write.sum"] * uw
            except Exception:
                traffic = None
            # flops the kernel really issues per (row, pair): one padded V = P Sigma GEMM (latent_fused) / the lower
            # 8x8 blocks of the Gram matrix (weighted_gram) -- the dense convention counts 4 Q^2 / 2 Q^2
            KSp, NBp = (Q + 3) // 4, (Q + 7) // 8
            executed_per_pair = {"latent_fused": 2.0 * (4 * KSp) * (8 * NBp),
                                 "weighted_gram": 2.0 * 64 * (NBp * (NBp + 1) // 2)}.get(dom)
            roof = {"kernel": dom, "bound": "tensor", "achieved": kern[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                    "frac": kern[dom]["tflops"] / peak, "traffic": traffic,
                    "peak_source": "FP64 DGEMM (torch.matmul, cuBLAS) 8192^3 measured in this run; "
                                   "MEASURED_PEAKS.json holds no FP64 figure",
                    "convention": "achieved = dense-convention flops of SURVEY.md 8d (2 Q^2 per quadratic form and per "
                                  "adjoint); the kernel re-uses V = P Sigma for the adjoint, see executed_tflops",
                    "avg_launch_ms": kern[dom]["ms_total"] / kern[dom]["calls"],
                    "share_of_step": kern[dom]["share_of_step"]}
            if executed_per_pair:
                nsamp = S if dom == "latent_fused" else S + 1
                roof["executed_tflops"] = args.steps * nsamp * pairs_per_sample * executed_per_pair / (kern[dom]["ms_total"] * 1e-3) / 1e12
                roof["executed_frac"] = roof["executed_tflops"] / peak
        F_step = 3.0 * ((S + 1) * pairs_per_sample * (2.0 * Q * Q + 2.0 * Q) + S * 2.0 * Bloc * Q * Q) * world
        line = {"metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, w), "rows_per_gpu": Bloc, "sharding": "rows strided over ranks",
                           "l2": "per-step working set (>= 5 B*Q doubles per sample chunk, >1 GB) exceeds the 126 MB L2",
                           "noise": "device (counter-based, in-kernel)", "optimizer": "Adam lr=0.005"},
                "step_tflops_fp64": F_step / (ms_step * 1e-3) / 1e12,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(ncalls),
                "roofline": roof, "kernels": kern, "other_ops": others, "loss": float(last)}
        if args.cpu_baseline == "auto" and world == 1:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            est = reference_estimate(w)
            line["cpu_baseline"] = {"value": est["value"], "unit": UNIT, "cores": torch.get_num_threads(),
                                    "kind": "port", "sample": est["desc"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
