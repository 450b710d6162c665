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


# ---------------------------------------------------------------------------------------- round 2 lines
R02 = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*.json")))
R02_B200 = [p for p in R02 if "reference" not in os.path.basename(p)]


@pytest.mark.parametrize("path", R02_B200, ids=[os.path.basename(p) for p in R02_B200])
def test_round2_bench_line_contract(path):
    j = json.load(open(path))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "roofline_hbm", "cpu_baseline", "parity", "e2e",
                "gpu_launches", "clocks", "by_workload"):
        assert key in j, key
    n = j["n_gpus"]
    assert j["metric"] == "patch_queries_per_sec" and j["unit"] == "patch-queries/s" and j["higher_is_better"] is True
    assert j["warmup"] >= 3 and j["vs_baseline"] is None and j["data"] == "synthetic" and j["dtype"] == "bf16"
    assert "workload" in j["config"] and "model" not in j["config"]
    # whole-job throughput: every rank works on the same queries of a step when the bank is row-sharded
    assert j["value"] == pytest.approx(j["config"]["queries_per_step"] / (j["ms_per_step"] * 1e-3), rel=1e-6)
    assert j["scaling"] == ("weak" if n == 1 else "strong")
    r = j["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and "traffic" in r
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["achieved"] == pytest.approx(r["flop_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e12, rel=1e-6)
    assert r["flop_per_launch"] == pytest.approx(2.0 * j["config"]["bank_rows_per_gpu"] * j["config"]["d"] * j["config"]["queries_per_step"])
    for name, h in j["roofline_hbm"].items():
        assert h["gbs"] == pytest.approx(h["bytes"] / h["ms"] / 1e6, rel=1e-6) and 0 < h["frac"] < 1.2, name
    e = j["e2e"]
    assert e["unit"] == j["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != j["value"]
    assert j["gpu_launches"] == j["gpu_launches_per_step"] * j["steps"] > 0
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    p = j["parity"]
    assert p["ok"] and p["recall_at_30"] >= 0.999 and p["score_max_rel_err"] <= 1e-3 and p["miou_delta_points"] <= 0.05
    assert p["confusion_bit_exact_where_labels_match"] and p["gpu_exact_standin_vs_cpu_oracle_recall"] >= 0.999
    if n == 1:
        c = j["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    else:
        assert j["cpu_baseline"] is None and j["sharded"]["parity"]["ok"] and "exchange" in j["config"]["collective"]
    for name, b in j["by_workload"].items():
        assert b["value"] > 0 and b["ms_per_step"] > 0 and b["unit"] == j["unit"], name


def test_round2_reference_arm_lines():
    for path in [p for p in R02 if "reference" in os.path.basename(p)]:
        j = json.load(open(path))
        assert j["impl"] == "reference" and j["metric"] == "patch_queries_per_sec" and j["gpu_launches"] == 0
        assert j["cpu_baseline"]["value"] == j["value"] and j["e2e"]["value"] == j["value"]
        assert j["cpu_baseline"]["cores"] >= 2, "the reference arm must use the host's cores also under torchrun"


# ---------------------------------------------------------------------------------------- synthetic inputs
def test_synthetic_bank_is_a_pure_function_of_the_row_and_matches_the_oracle_labels():
    """bench_synth: the reference arm builds its label memory with soft_labels() (no one-hot tensor);
    it must equal the oracle's restatement of hbird_eval.py:309-320.  Shard boundaries inside an image
    reproduce exactly the rows of the unsharded bank."""
    import numpy as np
    import torch

    import bench
    import bench_synth as syn
    from oracle import hbird_oracle as O

    w = dict(bench.WORKLOADS["cfg1"], N=5 * 196 + 77)
    cpu = torch.device("cpu")
    whole = [(f.view(-1, w["d"]) if s is None else f.view(-1, w["d"])[s.long()], m, s) for f, m, s, _ in bench.bank_slabs(w, 0, w["N"], cpu)]
    rows = torch.cat([f for f, _, _ in whole])
    assert rows.shape == (w["N"], w["d"])
    a, b = 300, 900  # a shard that starts and ends inside an image
    part = torch.cat([f.view(-1, w["d"]) if s is None else f.view(-1, w["d"])[s.long()] for f, m, s, _ in bench.bank_slabs(w, a, b, cpu, slab_rows=400)])
    assert torch.equal(part, rows[a:b])
    feats, maps = syn.images(w, 0, 3, cpu)
    again, maps2 = syn.images(w, 1, 1, cpu)
    assert torch.equal(again[0], feats[1]) and torch.equal(maps2[0], maps[1])
    bank_maps = torch.where(maps == 255, torch.zeros_like(maps), maps)
    y = (maps.float() / 255.0).unsqueeze(1).numpy()
    fm, lm = O.build_memory([(feats.numpy(), y)], w["C"], w["S"])
    np.testing.assert_allclose(syn.soft_labels(w, bank_maps).numpy(), lm, rtol=0, atol=1e-7)
    np.testing.assert_allclose(O.normalise_rows(feats.view(-1, w["d"]).numpy()), fm, rtol=0, atol=1e-7)
    assert 0.01 < float((maps == 255).float().mean()) < 0.03
    assert bench.cpu_sample_queries(bench.WORKLOADS["cfg3"], 4.0, 21904) == 63
