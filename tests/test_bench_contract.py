"""CPU tier: the committed bench lines (profiles/r2_bench_line*.json, written by bench.py on the GPU box) carry every key the bench
contract names -- a guard against dropping one when bench.py is edited.  No GPU, no oracle."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "gpu_launches"]


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(name + " not committed")
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r2_bench_line.json", 1), ("r2_bench_line_2gpu.json", 2), ("r2_bench_line_4gpu.json", 4), ("r2_bench_line_8gpu.json", 8)])
def test_own_arm_line(name, n):
    d = _line(name)
    for k in BASE_KEYS + ["clocks", "roofline"]:
        assert k in d, k
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["metric"] == "chamfer_nn_point_pairs_per_s" and d["unit"] == "Gpairs/s" and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["config"]["pairs_per_step"] / (d["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * d["value"]   # pairs_per_step is the whole job's
    assert d["config"]["pairs_per_step"] == 2.0 * 32 * 2048 * 16384 * n                                                # weak scaling: B=32 per GPU
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]


def test_reference_arm_line():
    d = _line("r2_bench_line_reference_arm.json")
    for k in BASE_KEYS + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "chamfer_nn_point_pairs_per_s" and d["unit"] == "Gpairs/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "reference"
