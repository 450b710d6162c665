#!/bin/bash
# round 2, call 2: GPU test-suite, default bench line (cfg3) with wall time, reference arm at cfg3, K2 probe
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
SECONDS=0
timeout 1200 python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
echo "bench exit $? wall ${SECONDS}s" | tee -a gpurun_out/bench_cfg3.err
tail -5 gpurun_out/bench_cfg3.err
SECONDS=0
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_cfg3.json 2> gpurun_out/bench_ref_cfg3.err
echo "ref exit $? wall ${SECONDS}s" | tee -a gpurun_out/bench_ref_cfg3.err
PROBE_FAST=1 PROBE_TAG=_new2 PROBE_SHAPES=cfg1,cfg2 timeout 600 python tools/gpu_probe.py perf > gpurun_out/probe_perf_new2.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/probe_perf_new2.json"))
    for k, v in d.items():
        if k.startswith("perf_"):
            print(k, round(v["ms"], 3), round(v["frac_sustained"], 3))
except Exception as e:
    print("probe missing", e)
for f in ("bench_cfg3", "bench_ref_cfg3"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
        print(f, j["value"], j.get("ms_per_step"), (j.get("roofline") or {}).get("frac"), (j.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
