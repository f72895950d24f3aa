"""The committed bench records of the final build (profiles/bench_r02_v*.json) are real `bench.py` lines: they parse, carry every key of the
bench contract, and their derived fields are consistent with each other (value = units / time, roofline.frac = achieved / peak, e2e <= value,
parity inside the tolerances `bench.py` enforces).  Guards against hand-edited or truncated records; no GPU needed."""
import glob
import json
import os

import pytest

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
RECORDS = sorted(glob.glob(os.path.join(P, "bench_r02_v[34]_*.json")))
KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "e2e", "roofline", "clocks", "gpu_launches"]


def line(path):
    rows = [l for l in open(path) if l.startswith("{")]
    assert rows, path
    return json.loads(rows[-1])


def test_there_are_final_records():
    names = [os.path.basename(r) for r in RECORDS]
    for want in ("bench_r02_v4_int8_full.json", "bench_r02_v3_int8_8gpu.json", "bench_r02_v3_kron_2p26.json", "bench_r02_v3_kron_2p26_8gpu.json"):
        assert want in names


@pytest.mark.parametrize("path", RECORDS, ids=[os.path.basename(r) for r in RECORDS])
def test_record_is_a_consistent_bench_line(path):
    d = line(path)
    for k in KEYS:
        assert k in d, k
    assert d["metric"] == "elbo_grad_points_per_sec" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 3 and d["gpu_launches"] > 0
    N = d["config"]["N"]
    assert abs(d["value"] - N / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]                 # whole-job points per step / step time
    e = d["e2e"]
    assert 0.8 * d["value"] <= e["value"] <= 1.001 * d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9 and 0.3 < r["frac"] < 1.0
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} and c["sm_mhz"] > 1500
    par = d.get("parity")
    if par:                                                                                      # bench.py exits non-zero outside these
        assert par["elbo_rel"] <= 1e-10 and par["grad_rel_max"] <= 1e-8 and par["info"] == 0
    if d["n_gpus"] > 1:
        assert d["scaling"] == "strong" and par and par["ranks"] == d["n_gpus"]
