#!/usr/bin/env python
"""Scale-sweep measurements of the SIM_code line (BASELINE.json config 5): batched Gibbs-kernel build (HBM roofline),
DMMA GEMM and blocked Cholesky (FP64 tensor roofline), Kronecker log-density.  One JSON line per size.
Not the headline metric (bench.py is); evidence for DESIGN.md / profiles/."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops  # noqa: E402
from collaborative_nonstationary_multivariate_gaussian_process_b200 import distributions, kernels  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def kron_sharded(args):
    """Kronecker log-density with the D eigen-blocks dealt over the ranks (scale sweep): launched under
    torchrun like bench.py; time = max over ranks (CUDA events between barriers), rank 0 prints one line per T."""
    import torch.distributed as dist
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import parallel
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D = args.kron_D
    for T in [int(s) for s in args.sizes.split(",")]:
        g = torch.Generator().manual_seed(T)
        x = torch.sort(torch.rand(T, generator=g, dtype=torch.float64))[0].view(-1, 1).to(dev)
        ell = torch.exp(3 * (x.view(-1) - 1) ** 3 - 3.0)                     # sim.py:22-26
        K = kernels.Nonstationary_RBF_cov(x, ell1=ell)
        Lb = torch.tril(torch.randn(D, D, generator=g, dtype=torch.float64)); Bf = (Lb @ Lb.t() / D).to(dev)
        y = torch.randn(D * T, generator=g, dtype=torch.float64).to(dev)
        mu = torch.zeros_like(y)
        s2 = torch.tensor(1e-2, dtype=torch.float64)
        f = lambda: parallel.kron_logpdf0_sharded(y, mu, Bf, K, s2)
        lp = f()                                                             # warm-up (scratch, streams)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); lp = f(); e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            flops = D * T ** 3 / 3.0
            print(json.dumps({"T": T, "D": D, "n_gpus": world, "kron_logpdf0_ms": float(ms.item()),
                              "TFs_total": flops / (float(ms.item()) * 1e-3) / 1e12, "logpdf0": float(lp.cpu()),
                              "blocks_per_rank": (D + world - 1) // world}), flush=True)
        del K, y
    if world > 1:
        dist.destroy_process_group()


def sweep_problem(T, D, dev, seed=0):
    """BASELINE config 5 inputs: x = sort(U(0,1)), log-ell(x) = 3 (x-1)^3 - 3, sigma = 1, sigma2_err = 1e-2,
    B_f = L L^T / D with L = tril(randn(D, D)), y ~ N(0, I)."""
    g = torch.Generator().manual_seed(seed + T)
    x = torch.sort(torch.rand(T, generator=g, dtype=torch.float64))[0].view(-1, 1)
    ell = torch.exp(3 * (x.view(-1) - 1) ** 3 - 3.0)
    Lb = torch.tril(torch.randn(D, D, generator=g, dtype=torch.float64))
    Bf = Lb @ Lb.t() / D
    y = torch.randn(D * T, generator=g, dtype=torch.float64)
    mv = lambda t: t if dev is None else t.to(dev)
    return mv(x), mv(ell), mv(Bf), mv(y), torch.tensor(1e-2, dtype=torch.float64)


def run_reference_sweep(args):
    """The UNMODIFIED reference (oracle/_ref: kernels.Nonstationary_RBF_cov + distributions.multivariate_normal_logpdf0,
    i.e. two symeig calls) on the host cores, at a bounded size (its eigen-decomposition of K is O(10 T^3)): one
    evaluation = kernel build + log-density."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ref_runner
    torch.set_num_threads(os.cpu_count() or 1)
    Ts, Ds = min(args.sweep_T, 2048), min(args.sweep_D, 16)
    x, ell, Bf, y, s2 = sweep_problem(Ts, Ds, None)
    if ref_runner.available():
        ker, kro, dis = ref_runner.sim_modules()
        kind = "reference"

        def one():
            K = ker.Nonstationary_RBF_cov(x, ell1=ell)
            return float(dis.multivariate_normal_logpdf0(y, torch.zeros_like(y), Bf, K, s2))
    else:
        from oracle import nmgp_oracle as orc
        kind = "port"

        def one():
            K = orc.sim_nonstationary_cov(x, ell1=ell)
            return float(orc.mvn_logpdf_kron_eig(y, torch.zeros_like(y), Bf, K, s2))
    one()
    times = []
    t_start = time.perf_counter()
    while len(times) < max(args.steps, 1) and (len(times) < 1 or time.perf_counter() - t_start < args.ref_budget):
        t0 = time.perf_counter(); one(); times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    cfg = sweep_config(args)
    print(json.dumps({"impl": "reference", "metric": SWEEP_METRIC % (args.sweep_T, args.sweep_D), "value": 1.0 / med,
                      "unit": "evals/s", "n_gpus": args.gpus, "steps": len(times), "warmup": 1, "ms_per_step": med * 1e3,
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": cfg, "value_kind": "measured at the bounded size T=%d, D=%d (NOT the configuration's size: "
                      "the reference's symeig(K) is O(10 T^3); value is evaluations/s of that smaller problem)" % (Ts, Ds),
                      "cpu_baseline": {"value": 1.0 / med, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": kind,
                                       "sample": "kernel build + multivariate_normal_logpdf0 at T=%d, D=%d: median %.3f s "
                                                 "of %d calls" % (Ts, Ds, med, len(times))},
                      "e2e": {"value": 1.0 / med, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
          flush=True)


SWEEP_METRIC = "Kronecker log-density evaluations/sec (Gibbs-kernel build + eigen-block Cholesky) at T=%d,M=%d"


def sweep_config(args):
    return {"workload": "scale sweep: T=%d time points x D=%d outputs, nonstationary kernel build + "
                        "multivariate_normal_logpdf0 (D eigen-block Cholesky factorisations dealt to the ranks)"
                        % (args.sweep_T, args.sweep_D), "T": args.sweep_T, "D": args.sweep_D}


def quick_line(T, D, rank, world, dev, timed, peak):
    """One short measured line of the scale-sweep configuration for bench.py's `other_configs` (same step as `run`)."""
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import parallel
    xh, ellh, Bfh, yh, s2 = sweep_problem(T, D, None)
    xh, ellh, yh, Bfh = xh.pin_memory(), ellh.pin_memory(), yh.pin_memory(), Bfh.pin_memory()
    x, ell, Bf, y = xh.to(dev), ellh.to(dev), Bfh.to(dev), yh.to(dev)
    mu = torch.zeros_like(y)

    def step():
        return parallel.kron_logpdf0_sharded(y, mu, Bf, kernels.Nonstationary_RBF_cov(x, ell1=ell), s2)

    def step_e2e():
        xd, elld, Bd, yd = (t.to(dev, non_blocking=True) for t in (xh, ellh, Bfh, yh))
        return float(parallel.kron_logpdf0_sharded(yd, torch.zeros_like(yd), Bd, kernels.Nonstationary_RBF_cov(xd, ell1=elld), s2).cpu())
    for _ in range(3):
        step()
    k = 2
    ms, lp, _, nl = timed(step, k)
    ms2, _, _, _ = timed(step_e2e, k)
    blocks_local = (D + world - 1) // world
    tf = blocks_local * T ** 3 / 3.0 / (ms / k * 1e-3) / 1e12
    return {"metric": SWEEP_METRIC % (T, D), "config": {"workload": "scale sweep: T=%d x D=%d, nonstationary kernel build + "
                                                        "multivariate_normal_logpdf0 (eigen-blocks dealt to the ranks)" % (T, D),
                                                        "T": T, "D": D},
            "n_gpus": world, "steps": k, "warmup": 3, "value": 1e3 * k / ms, "unit": "evals/s", "ms_per_step": ms / k,
            "gpu_launches": int(nl), "logpdf0": float(lp),
            "e2e": {"value": 1e3 * k / ms2, "unit": "evals/s", "ms_per_step": ms2 / k,
                    "h2d_bytes_per_step": int(8 * (2 * T + D * D + D * T)) * world, "d2h_bytes_per_step": 8 * world},
            "roofline": {"kernel": "nmgp_potrf_big (blocked Cholesky, %d eigen-blocks per GPU in flight on streams)" % blocks_local,
                         "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                         "frac": (tf / peak) if peak else None, "algorithmic": "T^3/3 flop per block (build included in the time)"}}


def run(args):
    """bench.py --workload sweep: one step = build K (T x T, HBM roofline) + Kronecker log-density (D blocked Cholesky
    factorisations on the FP64 tensor cores, sharded over the ranks, one all-reduce of a double)."""
    if args.impl == "reference":
        return run_reference_sweep(args)
    import torch.distributed as dist
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import parallel
    import bench
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, D = args.sweep_T, args.sweep_D
    xh, ellh, Bfh, yh, s2 = sweep_problem(T, D, None)
    xh, ellh, yh, Bfh = xh.pin_memory(), ellh.pin_memory(), yh.pin_memory(), Bfh.pin_memory()
    x, ell, Bf, y = xh.to(dev), ellh.to(dev), Bfh.to(dev), yh.to(dev)
    mu = torch.zeros_like(y)
    t_build = [0.0]

    def step():
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        K = kernels.Nonstationary_RBF_cov(x, ell1=ell)
        e1.record()
        lp = parallel.kron_logpdf0_sharded(y, mu, Bf, K, s2)
        step.ev.append((e0, e1))
        return lp

    def step_e2e():
        xd, elld, Bd, yd = (t.to(dev, non_blocking=True) for t in (xh, ellh, Bfh, yh))
        K = kernels.Nonstationary_RBF_cov(xd, ell1=elld)
        return float(parallel.kron_logpdf0_sharded(yd, torch.zeros_like(yd), Bd, K, s2).cpu())
    step.ev = []
    peak = bench.fp64_yardstick(dev) if rank == 0 else None
    hbm = 6650.0
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", hbm)
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_src = "fallback"

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, ops.launch_count() - n0
    for _ in range(max(args.warmup, 1)):
        step()
    step.ev = []
    clk = bench.ClockSampler(local)
    if rank == 0:
        clk.start()
    ms, lp, nl = timed(step, args.steps)
    clocks = clk.stop() if rank == 0 else None
    build_ms = float(np.mean([a.elapsed_time(b) for a, b in step.ev]))
    ms2, _, _ = timed(step_e2e, args.steps)
    if rank == 0:
        ms_step = ms / args.steps
        blocks_local = (D + world - 1) // world
        chol_ms = ms_step - build_ms
        tf = blocks_local * T ** 3 / 3.0 / (chol_ms * 1e-3) / 1e12
        line = {"metric": SWEEP_METRIC % (T, D), "value": 1e3 / ms_step, "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": sweep_config(args),
                "setup": {"blocks_per_rank": blocks_local, "l2": "every T x T factorisation (%.2f GB) exceeds the 126 MB L2" % (8.0 * T * T / 1e9),
                          "concurrent_blocks": "3 streams / scratch slots per GPU"},
                "clocks": clocks, "gpu_launches": int(nl), "logpdf0": float(lp),
                "e2e": {"value": 1e3 * args.steps / ms2, "unit": "evals/s", "ms_per_step": ms2 / args.steps,
                        "h2d_bytes_per_step": int(8 * (2 * T + D * D + D * T)) * world, "d2h_bytes_per_step": 8 * world},
                "roofline": {"kernel": "nmgp_potrf_big (blocked Cholesky, %d eigen-blocks per GPU, 3 in flight)" % blocks_local,
                             "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                             "traffic": None, "algorithmic": "T^3/3 flop per block",
                             "peak_source": "FP64 DGEMM (torch.matmul, cuBLAS) 8192^3 measured in this run",
                             "avg_launch_ms": chol_ms, "share_of_step": chol_ms / ms_step},
                "build": {"kernel": "k_simcov (Nonstationary_RBF_cov)", "bound": "hbm", "ms": build_ms,
                          "achieved": 8.0 * T * T / (build_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                          "frac": 8.0 * T * T / (build_ms * 1e-3) / 1e9 / hbm, "peak_source": hbm_src,
                          "algorithmic": "8 T^2 bytes written (full matrix)"}}
        if args.cpu_baseline == "auto" and world == 1:
            from oracle import ref_runner
            torch.set_num_threads(os.cpu_count() or 1)
            Ts, Ds = min(T, 2048), min(D, 16)
            xs, es, Bs, ys, s2s = sweep_problem(Ts, Ds, None)
            if ref_runner.available():
                ker, kro, dis = ref_runner.sim_modules()
                kind = "reference"
                one = lambda: float(dis.multivariate_normal_logpdf0(ys, torch.zeros_like(ys), Bs, ker.Nonstationary_RBF_cov(xs, ell1=es), s2s))
            else:
                from oracle import nmgp_oracle as orc
                kind = "port"
                one = lambda: float(orc.mvn_logpdf_kron_eig(ys, torch.zeros_like(ys), Bs, orc.sim_nonstationary_cov(xs, ell1=es), s2s))
            one()
            t0 = time.perf_counter(); one(); tc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / tc, "unit": "evals/s (of the bounded sample)", "cores": torch.get_num_threads(),
                                    "kind": kind, "sample": "kernel build + multivariate_normal_logpdf0 at the bounded size "
                                    "T=%d, D=%d: %.3f s per evaluation (after one warm-up)" % (Ts, Ds, tc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2048,4096,8192,16384")
    ap.add_argument("--D", type=int, default=4)
    ap.add_argument("--kron-D", type=int, default=0, help="run only the sharded Kronecker log-density with this many outputs")
    args = ap.parse_args()
    if args.kron_D > 0:
        return kron_sharded(args)
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    dg = timed(lambda: torch.matmul(a, a))
    dgemm_peak = 2.0 * n ** 3 / (dg * 1e-3) / 1e12
    del a
    for T in [int(s) for s in args.sizes.split(",")]:
        g = torch.Generator().manual_seed(T)
        x = torch.sort(torch.rand(T, generator=g, dtype=torch.float64))[0].view(-1, 1).to(dev)
        ell = torch.exp(3 * (x.view(-1) - 1) ** 3 - 3.0)                     # sim.py:22-26
        K = kernels.Nonstationary_RBF_cov(x, ell1=ell)
        t_build = timed(lambda: kernels.Nonstationary_RBF_cov(x, ell1=ell))
        A = ops.scale_add_diag(K, 1.0, 1e-2)
        work = torch.empty_like(A)

        def chol():
            work.copy_(A)
            ops.potrf_big(work)
        t_copy = timed(lambda: work.copy_(A))
        t_chol = timed(chol) - t_copy
        Bm = torch.randn(T, T, device=dev, dtype=torch.float64)
        t_gemm = timed(lambda: ops.gemm_nt(A, Bm))
        t_ref = timed(lambda: torch.linalg.cholesky(A))                      # cuSOLVER yardstick, not on the product path
        line = {"T": T, "build_ms": t_build, "build_GBs": 8.0 * T * T / (t_build * 1e-3) / 1e9, "build_frac_hbm": 8.0 * T * T / (t_build * 1e-3) / 1e9 / hbm,
                "potrf_ms": t_chol, "potrf_TFs": T ** 3 / 3.0 / (t_chol * 1e-3) / 1e12, "potrf_frac": T ** 3 / 3.0 / (t_chol * 1e-3) / 1e12 / dgemm_peak,
                "gemm_nt_ms": t_gemm, "gemm_nt_TFs": 2.0 * T ** 3 / (t_gemm * 1e-3) / 1e12, "gemm_nt_frac": 2.0 * T ** 3 / (t_gemm * 1e-3) / 1e12 / dgemm_peak,
                "cusolver_potrf_ms": t_ref, "dgemm_peak_TFs": dgemm_peak, "hbm_peak_GBs": hbm}
        if T <= 8192:
            D = args.D
            Lb = torch.tril(torch.randn(D, D, generator=g, dtype=torch.float64)); Bf = (Lb @ Lb.t()).to(dev)
            y = torch.randn(D * T, generator=g, dtype=torch.float64).to(dev)
            s2 = torch.tensor(1e-2, dtype=torch.float64)
            t0 = time.perf_counter()
            lp = distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), Bf, K, s2)
            torch.cuda.synchronize()
            line["logpdf0_D%d_ms" % D] = (time.perf_counter() - t0) * 1e3
            line["logpdf0"] = float(lp.cpu())
        print(json.dumps(line), flush=True)
        del K, A, work, Bm


if __name__ == "__main__":
    main()
