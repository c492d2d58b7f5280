#!/usr/bin/env python
"""Benchmark of the NMGP DSVI hot path (BASELINE.json metric: DSVI iterations/s, ELBO + gradient + Adam,
at the ECoG shape T=4096, D("M")=64, S=32, Q=50, full batch B = T*D = 262144; FP64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (default workload: ecog)
    python bench.py --impl reference ...                      # the UNMODIFIED reference (oracle/_ref) on the host cores
    python bench.py --workload {ecog,pm25,hcp,sim,tiny} [--rows R] [--S s]
    python bench.py --workload sweep                          # BASELINE config 5 (kernel build + Kronecker/Cholesky)

One JSON line on stdout (rank 0).  DESIGN.md "Measurement" documents every field.
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

UNIT = "iters/s"
_E10 = {"length_scales_L0_log": 10., "length_scales_L1_log": 10.}
WORKLOADS = {
    # T, D, Q, S and driver hyper-parameters; ref_rows = the reference driver's own minibatch size
    "ecog": dict(T=4096, D=64, Q=50, S=32, lr=0.005, ref_rows=512,          # NMGP_ECoG_full.py:285-302
                 hyper=dict(_E10, length_scales_tildeell_log=5., sigma2_err_log=-5.),
                 metric="DSVI iters/sec (ELBO+grad) at T=4096,M=64,S=32"),
    "pm25": dict(T=2048, D=16, Q=100, S=8, lr=0.01, ref_rows=1000,          # NMGP_PM25.py:63-64,219-241
                 hyper=dict(_E10, length_scales_tildeell_log=10.),
                 metric="DSVI iters/sec (ELBO+grad) at T=2048,M=16,S=8 (PM2.5-shaped)"),
    "hcp": dict(T=1200, D=15, Q=100, S=64, lr=0.01, ref_rows=1000, subjects=True,   # NMGP_HCP.py:61-62,210-233
                hyper=dict(length_scales_L0_log=5., length_scales_L1_log=5., length_scales_tildeell_log=5.),
                metric="DSVI iters/sec (ELBO+grad) at T=1200,M=15, 64 subjects per step (HCP-shaped)"),
    "sim": dict(T=100, D=2, Q=20, S=1, lr=0.005, ref_rows=200, unit_grid=True,      # NMGP_SIM.ipynb cell 2
                hyper=dict(sigma2_L0_log=0., length_scales_L0_log=2., sigma2_L1_log=0., length_scales_L1_log=2.,
                           sigma2_tildeell_log=0., length_scales_tildeell_log=0., sigma2_err_log=-2.),
                metric="DSVI iters/sec (ELBO+grad) at T=100,M=2,S=1 (shipped simulation shape)"),
    "tiny": dict(T=256, D=4, Q=20, S=2, lr=0.005, ref_rows=1024,
                 hyper=dict(length_scales_L0_log=3., length_scales_L1_log=3., length_scales_tildeell_log=2.,
                            sigma2_err_log=-2.),
                 metric="DSVI iters/sec (ELBO+grad) at T=256,M=4,S=2 (smoke shape)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ecog", choices=sorted(WORKLOADS) + ["sweep"])
    ap.add_argument("--rows", type=int, default=0, help="random row subset of the T x D grid (0 = full batch)")
    ap.add_argument("--S", type=int, default=0, help="Monte-Carlo samples per iteration (0 = the workload's)")
    ap.add_argument("--cpu-baseline", default="auto", choices=["auto", "skip"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step from a CUDA graph")
    ap.add_argument("--others", default="auto", choices=["auto", "skip"],
                    help="auto: the default (ecog, full batch) run also measures short lines of BASELINE's other configs "
                         "(pm25, hcp, sim, scale sweep) and attaches them as `other_configs`")
    ap.add_argument("--ref-budget", type=float, default=float(os.environ.get("NMGP_REF_BUDGET_S", "300")),
                    help="wall-clock budget (s) of the reference arm's calls; it stops early when the requested "
                         "steps do not fit and reports the steps it timed")
    ap.add_argument("--sweep-T", type=int, default=8192)
    ap.add_argument("--sweep-D", type=int, default=128)
    return ap.parse_args()


def resolve(args):
    w = dict(WORKLOADS[args.workload])
    if args.S > 0:
        w["S"] = args.S
    w["rows"] = min(args.rows, w["T"] * w["D"]) if args.rows > 0 else w["T"] * w["D"]
    w["name"] = args.workload
    return w


def config_of(w):
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    full = w["rows"] == w["T"] * w["D"]
    what = ("%d subjects per step, one MC draw each, B=%d rows per subject" % (w["S"], w["rows"])) if w.get("subjects") \
        else "S=%d MC samples, %s B=%d" % (w["S"], "full batch" if full else "random row subset", w["rows"])
    return {"workload": "%s-shaped DSVI step: T=%d, D=%d outputs, Q=%d inducing, %s, shared grid"
                        % (w["name"], w["T"], w["D"], w["Q"], what),
            "T": w["T"], "D": w["D"], "Q": w["Q"], "S": w["S"], "rows": w["rows"]}


def grid_inputs(w):
    T = w["T"]
    return (np.arange(T, dtype=np.float64) / T) if w.get("unit_grid") else np.arange(T, dtype=np.float64)


def synthetic_problem(w, seed=0):
    """Shared grid X_d = arange(T) (as in the ECoG/HCP drivers, NMGP_ECoG_full.py:108-110), Y = N(0,1) draws
    ([S, T] per output for the subject workload), Z = linspace over the grid; float64, CPU, fixed seed.  With
    rows < T*D a random subset of the grid rows (the drivers' minibatch, seeded) is kept."""
    g = torch.Generator().manual_seed(seed)
    T, D, Q = w["T"], w["D"], w["Q"]
    xg = torch.from_numpy(grid_inputs(w))
    nsub = w["S"] if w.get("subjects") else 1
    Y = torch.randn(D, nsub, T, generator=g, dtype=torch.float64)
    z = np.linspace(float(xg[0]), float(xg[-1]), Q)
    keep = [np.arange(T) for _ in range(D)]
    if w["rows"] < T * D:
        pick = np.sort(np.random.default_rng(seed).choice(T * D, size=w["rows"], replace=False))
        keep = [pick[(pick // T) == d] % T for d in range(D)]
    return xg, Y, z, keep


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
# Reference side: the unmodified reference (oracle/_ref, vendored by oracle/build_ref.py) or, where that directory is
# absent, the oracle port (oracle/nmgp_oracle.py, pinned to the reference by the golden vectors).
def reference_stepper(w):
    """Returns (kind, make) where make(rows, seed) -> closure running ONE reference call (zero_grad, forward,
    backward(retain_graph=True), Adam.step; S=1) on `rows` random rows of the workload's grid."""
    from oracle import ref_runner
    T, D, Q = w["T"], w["D"], w["Q"]
    if ref_runner.available():
        ref = ref_runner.reference_module()
        xg = grid_inputs(w)
        Z = torch.linspace(float(xg[0]), float(xg[-1]), Q, dtype=torch.float64).view(-1, 1)
        model = ref.NMGP(T * D, D, Z, mu_v=np.ones(Q), seed=22)
        for k, v in w["hyper"].items():
            getattr(model, k).data.fill_(float(v))
        for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
            getattr(model, k).requires_grad = False                      # fix_hyperpars=True (nmgp_dsvi.py:795-814)
        opt = torch.optim.Adam(model.parameters(), lr=w["lr"])

        def make(rows, seed=0):
            rng = np.random.default_rng(seed)
            rows = min(int(rows), T * D)
            pick = np.sort(rng.choice(T * D, size=rows, replace=False)) if rows < T * D else np.arange(T * D)
            Xl = [torch.from_numpy(xg[pick[(pick // T) == d] % T]).view(-1, 1) for d in range(D)]
            Yl = [torch.from_numpy(rng.standard_normal(x.shape[0])).view(-1, 1) for x in Xl]

            def call():                                                 # code/nmgp_dsvi.py:832-854
                opt.zero_grad()
                loss = model(Xl, Yl)
                loss.backward(retain_graph=True)
                opt.step()
                return float(loss.detach())
            return call
        return "reference", make

    from oracle import nmgp_oracle as orc

    def make(rows, seed=0):
        rng = np.random.default_rng(seed)
        rows = min(int(rows), T * D)
        p = orc.init_params(D, Q, seed=22, mu_v=np.ones(Q))
        for k, v in w["hyper"].items():
            p[k] = torch.tensor(float(v), dtype=torch.float64)
        xg = grid_inputs(w)
        pick = np.sort(rng.choice(T * D, size=rows, replace=False)) if rows < T * D else np.arange(T * D)
        Xl = [torch.from_numpy(xg[pick[(pick // T) == d] % T]).view(-1, 1) for d in range(D)]
        Yl = [torch.from_numpy(rng.standard_normal(x.shape[0])).view(-1, 1) for x in Xl]
        Z = torch.linspace(float(xg[0]), float(xg[-1]), Q, dtype=torch.float64).view(-1, 1)
        st = {}

        def call():
            loss, grads = orc.step_loss_and_grads(p, Z, T * D, Xl, Yl)
            for k, gk in grads.items():                      # Adam as in code/nmgp_dsvi.py:854 (negligible cost)
                if gk is None:
                    continue
                m, v, t = st.get(k, (torch.zeros_like(gk), torch.zeros_like(gk), 0))
                t += 1
                m = 0.9 * m + 0.1 * gk; v = 0.999 * v + 0.001 * gk * gk
                p[k] = p[k] - w["lr"] * (m / (1 - 0.9 ** t)) / ((v / (1 - 0.999 ** t)).sqrt() + 1e-8)
                st[k] = (m, v, t)
            return float(loss)
        return call
    return "port", make


def reference_measure(w, steps, warmup, budget_s, fit=True):
    """Times reference calls inside a wall-clock budget.

    exact (the configuration fits the reference: --rows given, or a small workload): one step = S calls on the
    configuration's rows; value = 1 / median step time -- no extrapolation.

    bounded sample (full batch of the big workloads: the reference needs an 8.6 GB (D,D,B) tensor and minutes per call
    there): calls at the reference driver's own minibatch B1 = ref_rows (S=1) and, for the fit, at 2.5 B1 and 4 B1.
    `value` is the MEASURED BOUND 1 / (S * median t(B1)): one iteration of the workload is S forward/backward calls on
    >= B1 rows each, and a call on more rows cannot be faster, so the reference cannot exceed it.  The least-squares
    model t(B) = a + b B over all timed calls gives the extrapolated full-batch time, reported beside it with its
    residual (`extrapolated`), never as `value`."""
    T, D, S = w["T"], w["D"], w["S"]
    t_start = time.perf_counter()
    kind, make = reference_stepper(w)
    full = T * D
    exact = w["rows"] < full or w["rows"] * S <= 4096
    samples = []                                            # (rows, seconds)
    if exact:
        call = make(w["rows"], seed=0)

        def one_step():
            t0 = time.perf_counter()
            for _ in range(S):
                call()
            return time.perf_counter() - t0
        t_first = one_step()
        n_warm = 1
        while n_warm < warmup and (time.perf_counter() - t_start) + t_first < 0.25 * budget_s:
            one_step(); n_warm += 1
        times = []
        while len(times) < steps and (len(times) < 1 or (time.perf_counter() - t_start) + np.median(times) < budget_s):
            times.append(one_step())
        med = float(np.median(times))
        return {"kind": kind, "exact": True, "value": 1.0 / med, "steps": len(times), "warmup": n_warm,
                "mean_call_s": float(np.mean(times)), "median_step_s": med, "times_s": times,
                "sample": "%s: %d timed iterations (after %d warm-up) of the exact configuration: S=%d sequential "
                          "forward+backward+Adam calls on B=%d rows; median %.4f s per iteration, min %.4f, max %.4f"
                          % ("unmodified reference (oracle/_ref)" if kind == "reference" else "oracle port", len(times),
                             n_warm, S, w["rows"], med, min(times), max(times))}
    B1 = min(int(w["ref_rows"]), full)
    B2, B3 = min(int(2.5 * B1), full), min(4 * B1, full)
    calls = {b: make(b, seed=i) for i, b in enumerate((B1, B2, B3))}
    calls[B1]()                                             # one warm-up call (allocator, thread pools, lazy init)
    n_warm = 1
    order = [B1, B3, B1, B2] if fit else [B1]
    need = 3 if fit else 1
    i = 0
    while len(samples) < max(steps, need):
        b = order[i % len(order)]; i += 1
        est = max([t for bb, t in samples if bb == b], default=max([t for _, t in samples], default=0.0))
        if len(samples) >= need and (time.perf_counter() - t_start) + est > budget_s:
            break
        t0 = time.perf_counter(); calls[b](); samples.append((b, time.perf_counter() - t0))
    t1 = [t for b, t in samples if b == B1]
    med1 = float(np.median(t1))
    out = {"kind": kind, "exact": False, "value": 1.0 / (S * med1), "steps": len(samples), "warmup": n_warm,
           "mean_call_s": float(np.mean([t for _, t in samples])), "median_call_B1_s": med1, "B1": B1,
           "samples": [[int(b), float(t)] for b, t in samples]}
    bs = np.array([b for b, _ in samples], dtype=np.float64); ts = np.array([t for _, t in samples])
    ext = None
    if len(set(bs.tolist())) >= 2:
        A = np.stack([np.ones_like(bs), bs], 1)
        (a, b), *_ = np.linalg.lstsq(A, ts, rcond=None)
        resid = float(np.sqrt(np.mean((A @ np.array([a, b]) - ts) ** 2)))
        b = max(float(b), 0.0)
        t_full = S * (float(a) + b * full)
        ext = {"a_s": float(a), "b_s_per_row": b, "rms_residual_s": resid, "points": sorted(set(int(x) for x in bs)),
               "t_iteration_s": t_full, "value": 1.0 / t_full if t_full > 0 else None,
               "note": "least-squares t(B) = a + b*B over all timed calls, iteration = S*(a + b*T*D); extrapolation, "
                       "not used as `value`"}
    out["extrapolated"] = ext
    out["sample"] = ("%s: %d timed calls after 1 warm-up (forward + backward(retain_graph) + Adam, S=1) on random rows of "
                     "the T=%d x D=%d grid, Q=%d: median %.2f s at B=%d rows (%d calls: min %.2f, max %.2f)%s; value = "
                     "1/(S*median) with S=%d: a measured upper bound of the reference's iterations/s (the full batch has "
                     "%dx more rows per call)"
                     % ("unmodified reference (oracle/_ref)" if kind == "reference" else "oracle port of the reference",
                        len(samples), T, D, w["Q"], med1, B1, len(t1), min(t1), max(t1),
                        "" if ext is None else "; fit a=%.2f s, b=%.3g s/row (rms residual %.2f s) -> %.0f s per "
                        "full-batch iteration (extrapolated)" % (ext["a_s"], ext["b_s_per_row"], ext["rms_residual_s"],
                                                                 ext["t_iteration_s"]),
                        S, full // B1))
    return out


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = reference_measure(w, args.steps, args.warmup, args.ref_budget)
    small = None
    if not m["exact"]:
        small = {"config": "B=%d random grid rows, S=1 (the reference driver's minibatch): no extrapolation" % m["B1"],
                 "ms_per_step": m["median_call_B1_s"] * 1e3, "value": 1.0 / m["median_call_B1_s"], "unit": UNIT}
    line = {"impl": "reference", "metric": w["metric"], "value": m["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": m["steps"], "warmup": m["warmup"],
            # mean wall time of one TIMED call/iteration of the sample (timed region = steps * ms_per_step)
            "ms_per_step": m["mean_call_s"] * 1e3 * (w["S"] if m["exact"] else 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w),
            "value_kind": "measured (exact configuration)" if m["exact"] else
                          "measured bound: 1/(S * median call time at the driver minibatch); see cpu_baseline.sample",
            "requested": {"steps": args.steps, "warmup": args.warmup, "budget_s": args.ref_budget},
            "extrapolated": m.get("extrapolated"), "same_config_pair": small,
            "cpu_baseline": {"value": m["value"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": m["kind"],
                             "sample": m["sample"]},
            "e2e": {"value": m["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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


def ncu_traffic(dom, Q):
    """DRAM bytes per launch of the dominant kernel from this round's committed `ncu --set full` captures
    (profiles/r2/ncu_r2_*_full_summary.json, written by profiles/scripts/ncu_summarize.py).  Static: the capture is a
    separate profiler run of the same bench command at the shape named in the returned source string."""
    if Q <= 64:
        fname, shape = "profiles/r2/ncu_r2_ecog_full_summary.json", "ECoG shape, 4 samples per launch"
        prefix = {"latent_fused": "void k_latent_fused", "weighted_gram": "void k_gram"}.get(dom)
    else:
        fname, shape = "profiles/r2/ncu_r2_pm25_full_summary.json", "PM2.5 shape, 8 samples per launch"
        prefix = {"latent_fused": "void <unnamed>::k_lq<13, 0>", "weighted_gram": "void k_gram"}.get(dom)
    try:
        summ = json.load(open(os.path.join(ROOT, fname)))
    except Exception:
        return None, None
    for kname, rec in summ.items():
        if prefix and kname.startswith(prefix) and "dram_bytes_per_launch" in rec:
            return rec["dram_bytes_per_launch"], "%s (%s)" % (fname, shape)
    return None, None


class Problem:
    """Model, optimiser and device/host copies of one rank's share of a DSVI workload."""

    def __init__(self, w, rank, world, dev, noise="device", graph=True):
        from collaborative_nonstationary_multivariate_gaussian_process_b200 import nmgp_dsvi, parallel
        self.w, self.rank, self.world, self.dev = w, rank, world, dev
        T, D, Q, S = w["T"], w["D"], w["Q"], w["S"]
        xg, Y, z, keep = synthetic_problem(w)
        self.N = T * D * (S if w.get("subjects") else 1)         # observations behind the N/B factor
        model = nmgp_dsvi.NMGP(self.N, D, torch.from_numpy(z).view(-1, 1), mu_v=np.ones(Q), seed=22, device=dev,
                               noise=noise)
        for k, v in w["hyper"].items():
            getattr(model, k).data.fill_(v)
        for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
            getattr(model, k).requires_grad = False                   # fix_hyperpars=True, the drivers' setting
        self.model = model
        # fused=True: one multi-tensor kernel for the 13 parameter tensors (85 MB at the ECoG shape, replicated on every rank)
        self.opt = torch.optim.Adam(model.parameters(), lr=w["lr"], capturable=bool(graph), fused=True)
        self.params = list(model.parameters())
        counts = [int(k.shape[0]) for k in keep]
        Btot = int(sum(counts))
        if w.get("subjects"):
            # subjects (one MC draw each) are dealt to the ranks; every rank keeps all rows of its subjects
            self.subj = list(range(*parallel.shard_samples(S, rank, world)))
            rows = [np.arange(c) for c in counts]
            parallel.configure_model_for_sharding(model, Btot, rank, world, shard="samples", n_samples_total=S)
            self.n_mc = len(self.subj)
        else:
            self.subj = None
            rows = parallel.shard_rows_per_output(counts, rank, world)
            parallel.configure_model_for_sharding(model, Btot, rank, world)
            self.n_mc = S
        self.Xh = [xg[keep[d]][rows[d]].contiguous().pin_memory() for d in range(D)]
        if self.subj is None:
            self.Yh = [Y[d, 0][keep[d]][rows[d]].contiguous().pin_memory() for d in range(D)]
        else:
            self.Yh = [Y[d][self.subj][:, keep[d]][:, rows[d]].contiguous().pin_memory() for d in range(D)]  # [S_loc, T_d]
        self.Bloc = sum(int(x.shape[0]) for x in self.Xh)
        self.xd = torch.cat(self.Xh).to(dev)
        self.yd = torch.cat(self.Yh, dim=-1).to(dev)
        self.Id = torch.from_numpy(np.repeat(np.arange(D, dtype=np.int32), [int(x.shape[0]) for x in self.Xh])).to(dev)
        self.gid = torch.from_numpy(parallel.global_row_ids(counts, rows)).to(dev)
        self.parallel = parallel
        self.h2d_bytes = int(self.Bloc * (8 + 4) + self.yd.numel() * 8)
        self.graphed = None
        if graph:
            from collaborative_nonstationary_multivariate_gaussian_process_b200.graph_step import GraphedStep
            self.graphed = GraphedStep(model, self.opt, self.xd, self.yd, self.Id, n_mc=self.n_mc, row_gid=self.gid,
                                       distributed=world > 1)
            # pinned staging buffers of the end-to-end path (host lists -> pinned -> static device buffers)
            self.x_pin = torch.empty(self.xd.shape, dtype=torch.float64).pin_memory()
            self.y_pin = torch.empty(self.yd.shape, dtype=torch.float64).pin_memory()
            self.I_pin = self.Id.cpu().pin_memory()

    def check(self):
        if self.graphed is not None:
            self.graphed.check()
        self.parallel.raise_if_pending_not_pd()

    def step_resident(self):
        if self.graphed is not None:
            return self.graphed.step()
        self.opt.zero_grad(set_to_none=True)
        loss = self.model.forward_rows(self.xd, self.yd, self.Id, n_mc=self.n_mc, row_gid=self.gid)
        loss.backward()
        tot = self.parallel.allreduce_loss_and_grads(loss, self.params, pd_info=self.model._last_pd_info, check="defer")
        self.opt.step()
        return tot

    def step_e2e(self):
        if self.graphed is not None:
            # the caller's per-output host lists are packed into pinned memory, copied to the device, the captured
            # iteration is replayed on them and the loss is read back
            torch.cat([t.reshape(-1) for t in self.Xh], out=self.x_pin)
            torch.cat(self.Yh, dim=-1, out=self.y_pin)
            self.graphed.load_rows(self.x_pin, self.y_pin, self.I_pin)
            return float(self.graphed.step().cpu())
        self.opt.zero_grad(set_to_none=True)
        loss = self.model(self.Xh, self.Yh, n_mc=self.n_mc, noise="device", row_gid=self.gid,
                          subjects=self.subj is not None)                      # host lists -> H2D inside
        loss.backward()
        tot = self.parallel.allreduce_loss_and_grads(loss, self.params, pd_info=self.model._last_pd_info, check="defer")
        self.opt.step()
        return float(tot.cpu())                                       # D2H read of the step's result


def other_configs(args, rank, world, dev, timed, peak):
    """Short measured lines of BASELINE.json's other configurations, taken in the same process right after the headline
    workload so that the driver's record holds them too: PM2.5-shaped and simulation-shaped steps (single GPU only: they
    are one-GPU configurations), the HCP-shaped subject-sharded step and the scale-sweep evaluation (every N).  Same
    timing rules as the headline (warm-up >= 3, CUDA events, max over ranks); `e2e` from pinned host lists."""
    out = {}
    k2 = max(args.steps, 5)
    names = ["hcp"] + (["pm25", "sim"] if world == 1 else [])
    for name in names:
        w2 = dict(WORKLOADS[name]); w2["rows"] = w2["T"] * w2["D"]; w2["name"] = name
        pb2 = Problem(w2, rank, world, dev, graph=not args.no_graph)
        for _ in range(3):
            pb2.step_resident()
        pb2.check()
        ms, last, _, nl = timed(pb2.step_resident, k2)
        for _ in range(2):
            pb2.step_e2e()
        ms2, _, _, _ = timed(pb2.step_e2e, k2)
        pb2.check()
        out[name] = {"metric": w2["metric"], "config": config_of(w2), "n_gpus": world, "steps": k2, "warmup": 3,
                     "value": 1e3 * k2 / ms, "unit": UNIT, "ms_per_step": ms / k2, "gpu_launches": int(nl),
                     "e2e": {"value": 1e3 * k2 / ms2, "unit": UNIT, "ms_per_step": ms2 / k2,
                             "h2d_bytes_per_step": pb2.h2d_bytes * world, "d2h_bytes_per_step": 8 * world},
                     "loss": float(last)}
        del pb2
        torch.cuda.empty_cache()
    import bench_sweep
    out["sweep"] = bench_sweep.quick_line(args.sweep_T, args.sweep_D, rank, world, dev, timed, peak)
    return out


def run_b200(args, w):
    import torch.distributed as dist
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, D, Q, S = w["T"], w["D"], w["Q"], w["S"]
    pb = Problem(w, rank, world, dev, graph=not args.no_graph)

    def timed(fn, k, prof=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if prof:
            _ops._profile_begin()
        n0 = _ops.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        nl = _ops.launch_count() - n0
        profd = _ops._profile_end() if prof else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, profd, nl

    peak = fp64_yardstick(dev) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        pb.step_resident()
    pb.check()
    # per-kernel breakdown: eager steps with CUDA events around every C-ABI call (a graph replay cannot be instrumented)
    graphed, pb.graphed = pb.graphed, None
    ms_inst, _, (prof, ncalls), _ = timed(lambda: (pb.step_resident(), pb.model.advance_noise_step())[0], args.steps,
                                          prof=True)
    pb.graphed = graphed
    pb.check()
    # the timed region of `value`: K steps (graph replays unless --no-graph), clocks sampled meanwhile
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    ms, last, _, nlaunch = timed(pb.step_resident, args.steps)
    clocks = clk.stop() if rank == 0 else None
    pb.check()
    ms_step = ms / args.steps
    e2e = None
    if not args.no_e2e:
        for _ in range(max(1, min(args.warmup, 2))):
            pb.step_e2e()
        ms2, _, _, _ = timed(pb.step_e2e, args.steps)
        e2e = {"value": 1e3 * args.steps / ms2, "unit": UNIT, "h2d_bytes_per_step": pb.h2d_bytes * world,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": ms2 / args.steps}
    pb.check()

    small = None
    if rank == 0 and world == 1 and w["name"] == "ecog" and w["rows"] == T * D and not args.no_e2e:
        # the un-extrapolated pair of BASELINE.md 3.4: the reference driver's own minibatch (B=512 random rows, S=1),
        # end to end from host lists; `bench.py --impl reference` reports the same configuration as same_config_pair
        w2 = dict(w, rows=w["ref_rows"], S=1)
        pb2 = Problem(w2, 0, 1, dev, graph=not args.no_graph)
        for _ in range(3):
            pb2.step_e2e()
        ks = max(args.steps, 10)
        ms3, _, _, _ = timed(pb2.step_e2e, ks)
        small = {"config": "B=%d random grid rows, S=1 (the reference driver's minibatch): no extrapolation" % w2["rows"],
                 "ms_per_step": ms3 / ks, "value": 1e3 * ks / ms3, "unit": UNIT, "path": "e2e (host lists in, loss out)"}
        del pb2

    if rank == 0:
        Xh = pb.Xh
        n_mc = pb.n_mc
        # roofline of the dominant kernels: dense convention, 2 Q^2 flop per (row, used pair) quadratic form
        pairs_per_sample = float(sum((d + 1) * int(Xh[d].shape[0]) for d in range(D)))
        kern = {}
        # dense convention: 2 Q^2 flop per (row, used pair) for each of V = P Sigma, its adjoint and
        # the Gram accumulation; the fused latent kernel covers the first two for the S samples, the coefficient
        # (U) side runs once per step.
        for name, per_pair, nsamp in (("quadform_fwd", 2.0 * Q * Q, 1), ("quadform_bwd", 2.0 * Q * Q, 1),
                                      ("weighted_gram", 2.0 * Q * Q, n_mc + 1), ("latent_fused", 4.0 * Q * Q, n_mc)):
            if name in prof:
                calls, tms = prof[name]
                flops = args.steps * nsamp * pairs_per_sample * per_pair
                # share of the TIMED (graph-replayed) step: the op's device time per step over ms_per_step; ops that run on
                # the side stream (Cholesky / KL chain) overlap the row kernels, so the shares need not sum to one
                kern[name] = {"calls": calls, "ms_total": tms, "share_of_step": tms / ms,
                              "tflops": flops / (tms * 1e-3) / 1e12}
        others = {k: {"calls": c, "ms_total": t_, "share_of_step": t_ / ms} for k, (c, t_) in prof.items() if k not in kern}
        dom = max(kern, key=lambda k: kern[k]["ms_total"]) if kern else None
        roof = None
        if dom:
            traffic, tsrc = ncu_traffic(dom, Q)
            # algorithmic DRAM bytes of one launch of the dominant kernel: the fused kernel reads P [B,Q],
            # l [B,D], y, c and writes lbar, qbar, mbar [B,D], Pbar [B,Q], cbar per sample; the Gram kernel reads P and
            # the live halves of qbar / mbar
            ns_launch = max(1, round(n_mc * args.steps / max(kern[dom]["calls"], 1))) if dom == "latent_fused" else None
            alg_bytes = None
            if dom == "latent_fused":
                alg_bytes = 8.0 * ns_launch * pb.Bloc * (2 * Q + 4 * D + 3)
            elif dom == "weighted_gram":
                alg_bytes = 8.0 * (n_mc * args.steps / max(kern[dom]["calls"] - args.steps, 1)) * pb.Bloc * (Q + D)
            # flops the kernel really issues per (row, pair): one padded V = P Sigma GEMM (latent_fused) / the lower
            # 8x8 blocks of the Gram matrix (weighted_gram) -- the dense convention counts 4 Q^2 / 2 Q^2
            KSp, NBp = (Q + 3) // 4, (Q + 7) // 8
            executed_per_pair = {"latent_fused": 2.0 * (4 * KSp) * (8 * NBp),
                                 "weighted_gram": 2.0 * 64 * (NBp * (NBp + 1) // 2)}.get(dom)
            roof = {"kernel": dom, "bound": "tensor", "achieved": kern[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                    "frac": kern[dom]["tflops"] / peak, "traffic": traffic,
                    "traffic_source": None if traffic is None else
                    "static: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full "
                    "capture %s; a separate profiler run of this command, not measured in this run" % tsrc,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "FP64 DGEMM (torch.matmul, cuBLAS) 8192^3 measured in this run; "
                                   "MEASURED_PEAKS.json holds no FP64 figure",
                    "convention": "achieved = dense-convention flops of SURVEY.md 8d (2 Q^2 per quadratic form and per "
                                  "adjoint); the kernel re-uses V = P Sigma for the adjoint, see executed_tflops",
                    "avg_launch_ms": kern[dom]["ms_total"] / kern[dom]["calls"],
                    "share_of_step": kern[dom]["share_of_step"]}
            if executed_per_pair:
                nsamp = n_mc if dom == "latent_fused" else n_mc + 1
                roof["executed_tflops"] = args.steps * nsamp * pairs_per_sample * executed_per_pair / (kern[dom]["ms_total"] * 1e-3) / 1e12
                roof["executed_frac"] = roof["executed_tflops"] / peak
        F_step = 3.0 * ((n_mc + 1) * pairs_per_sample * (2.0 * Q * Q + 2.0 * Q) + n_mc * 2.0 * pb.Bloc * Q * Q) * world
        line = {"metric": w["metric"], "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_of(w),
                "setup": {"rows_per_gpu": pb.Bloc, "sharding": "subjects dealt to ranks" if w.get("subjects") else
                          "rows strided over ranks",
                          "l2": "per-step working set (>= 5 B*Q doubles per sample chunk) exceeds the 126 MB L2"
                          if pb.Bloc * Q * 8 * 5 > 126e6 else "working set fits L2: a 256 MB buffer is not flushed between "
                          "steps because every step rewrites all its intermediates (>= L2 at the named full-batch shapes)",
                          "noise": "device (counter-based, in-kernel)", "optimizer": "Adam lr=%g" % w["lr"],
                          "cuda_graph": "whole iteration replayed from one CUDA graph" if pb.graphed is not None else "off",
                          "ms_per_step_eager_instrumented": ms_inst / args.steps,
                          "instrumented_note": "per-op times come from an eager pass with CUDA events around every C-ABI call; "
                                               "ops issued on the side stream (Cholesky / KL of the coefficient covariances) "
                                               "include their queueing behind the row kernels there (0.4 .. 60 ms run to run); "
                                               "`value` is the graph-replayed step"},
                "step_tflops_fp64": F_step / (ms_step * 1e-3) / 1e12,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(nlaunch), "abi_calls": int(ncalls),
                "roofline": roof, "kernels": kern, "other_ops": others, "loss": float(last),
                "same_config_pair": small}
    graph_used = pb.graphed is not None
    extra = None
    if args.others == "auto" and w["name"] == "ecog" and w["rows"] == T * D and args.S == 0 and not args.no_e2e:
        del pb
        torch.cuda.empty_cache()
        extra = other_configs(args, rank, world, dev, timed, peak)
    if rank == 0:
        line["other_configs"] = extra
        if args.cpu_baseline == "auto" and world == 1:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            m = reference_measure(w, 3 if w["rows"] * S > 4096 else 5, 1, 90.0, fit=False)
            line["cpu_baseline"] = {"value": m["value"], "unit": UNIT, "cores": torch.get_num_threads(),
                                    "kind": m["kind"], "sample": m["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        if graph_used:
            # A captured CUDA graph that contains NCCL kernels keeps the communicator busy at teardown (observed: the
            # process group destructor never returns); the line is out and every rank is past the barrier, so leave
            # without running the destructors.
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "sweep":
        import bench_sweep
        sys.modules.setdefault("bench", sys.modules[__name__])
        return bench_sweep.run(args)
    w = resolve(args)
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
