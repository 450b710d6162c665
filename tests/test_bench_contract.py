"""The bench line contract, checked on the lines committed under profiles/ (written by bench.py on
B200 boxes): every key the driver reads is present and self-consistent.  CPU only."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_cfg[1-4].json")))


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_single_gpu_bench_line_has_the_contract_keys(path):
    j = json.load(open(path))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e",
                "gpu_launches", "clocks"):
        assert key in j, key
    assert j["metric"] == "patch_queries_per_sec" and j["unit"] == "patch-queries/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["warmup"] >= 3 and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert "workload" in j["config"] and "model" not in j["config"]
    r = j["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and "traffic" in r
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    # achieved = algorithmic flop per launch / measured kernel time
    assert r["achieved"] == pytest.approx(r["flop_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e12, rel=1e-6)
    c = j["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = j["e2e"]
    assert e["unit"] == j["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != j["value"]  # measured separately, through host buffers
    assert j["gpu_launches"] == j["gpu_launches_per_step"] * j["steps"] > 0
    # value = queries per step / step time
    q = j["config"]["queries_per_step_per_gpu"]
    assert j["value"] == pytest.approx(q / (j["ms_per_step"] * 1e-3), rel=1e-6)
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    p = j["parity"]
    assert p["recall_at_30"] >= 0.999 and p["score_max_rel_err"] <= 1e-3 and abs(p["miou_delta_points"]) <= 0.05


def test_reference_arm_line():
    j = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_reference_arm.json")))
    assert j["impl"] == "reference" and j["metric"] == "patch_queries_per_sec" and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["value"] == j["value"] and j["e2e"]["value"] == j["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0


def test_multi_gpu_lines_report_whole_job_throughput():
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_cfg*_n[248].json"))):
        j = json.load(open(path))
        n = j["n_gpus"]
        q = j["config"]["queries_per_step_per_gpu"]
        assert j["value"] == pytest.approx(n * q / (j["ms_per_step"] * 1e-3), rel=1e-6), path
        assert j["scaling"] == "weak" and j["sharded"]["scaling"] == "strong"
        assert j["sharded"]["value"] == pytest.approx(q / (j["sharded"]["ms_per_step"] * 1e-3), rel=1e-6)
