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
