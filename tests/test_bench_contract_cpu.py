"""The bench line contract: the committed driver-format lines under profiles/r2 carry every key the driver
reads, with the meanings the contract gives them.  (No GPU: this guards the schema, not the numbers.)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline")


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", "r2", name)).read().strip().splitlines()[-1])


def test_default_line_has_the_contract_keys():
    for name in ("bench_final_1gpu.json", "bench_final_8gpu.json"):
        d = _line(name)
        for k in REQUIRED:
            assert k in d, (name, k)
        assert d["unit"] == "iters/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["vs_baseline"] is None                       # BASELINE.md holds no published number for this metric
        assert abs(d["value"] - 1e3 / d["ms_per_step"]) <= 1e-9 * d["value"]
        e = d["e2e"]
        assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= d["value"] * 1.001
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, k
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert 0.0 < r["executed_frac"] < 1.0
        assert d["gpu_launches"] > 0
        c = d["clocks"]
        assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {
            "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        oc = d["other_configs"]
        assert {"hcp", "sweep"} <= set(oc)
        if d["n_gpus"] == 1:
            assert {"pm25", "sim"} <= set(oc)
            cb = d["cpu_baseline"]
            assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb


def test_scaling_lines_are_consistent():
    v = {n: _line("bench_final_%dgpu.json" % n) for n in (1, 2, 4, 8)}
    for n, d in v.items():
        assert d["n_gpus"] == n and d["metric"] == v[1]["metric"] and d["config"] == v[1]["config"]
    assert v[8]["value"] / v[1]["value"] >= 6.5               # north_star scaling target
    # the same synthetic problem and counter-based noise on every rank count: the losses of the last timed step agree
    for n in (2, 4, 8):
        assert abs(v[n]["loss"] - v[2]["loss"]) <= 1e-9 * abs(v[2]["loss"])
