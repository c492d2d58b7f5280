#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-kernel JSON summaries kept under profiles/.

    python profiles/scripts/ncu_summarize.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rN/ncu_..._summary.json

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU), averages every numeric metric of interest
over the captured launches of each kernel (the launches of one kernel in one capture run on the same shapes) and keeps
the launch count.  `bench.py` reads `dram__bytes_read.sum + dram__bytes_write.sum` from these files for `roofline.traffic`.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def summarise(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = rows[0]
    units = rows[1]
    name_col = head.index("Kernel Name")
    per = {}
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        k = r[name_col]
        ent = per.setdefault(k, {"launches": 0, "sums": {}, "units": {}})
        ent["launches"] += 1
        for c, (h, u) in enumerate(zip(head, units)):
            if h in KEEP or (h.startswith(STALL_PREFIX) and h.endswith(STALL_SUFFIX)):
                v = num(r[c])
                if v is None:
                    continue
                if h.startswith("dram__bytes") and u in TO_BYTES:
                    v, u = v * TO_BYTES[u], "byte"
                if h == "gpu__time_duration.sum" and u in TO_MS:
                    v, u = v * TO_MS[u], "ms"
                ent["sums"][h] = ent["sums"].get(h, 0.0) + v
                ent["units"][h] = u
    res = {}
    for k, ent in per.items():
        n = ent["launches"]
        m = {h: s / n for h, s in ent["sums"].items()}
        stalls = sorted(((h[len(STALL_PREFIX):-len(STALL_SUFFIX)], v) for h, v in m.items() if h.startswith(STALL_PREFIX)),
                        key=lambda t: -t[1])
        d = {h: v for h, v in m.items() if not h.startswith(STALL_PREFIX)}
        d["units"] = {h: u for h, u in ent["units"].items() if not h.startswith(STALL_PREFIX)}
        d["launches_captured"] = n
        d["top_stalls"] = [[a, round(b, 3)] for a, b in stalls[:6]]
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            d["dram_bytes_per_launch"] = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        res[k] = d
    return res


if __name__ == "__main__":
    total = {}
    for rep in sys.argv[1:]:
        for k, v in summarise(rep).items():
            v["report"] = rep.split("/")[-1]
            total[k] = v
    json.dump(total, sys.stdout, indent=1, sort_keys=True)
    print()
