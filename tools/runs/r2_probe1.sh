#!/bin/bash
# round 2, call 1: GPU test-suite, K2 A/B against the round-1 library, epilogue cycle stats
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
PROBE_FAST=1 PROBE_TAG=_new timeout 600 python tools/gpu_probe.py perf > gpurun_out/probe_perf_new.log 2>&1
HBIRD_B200_AB_LIB=$PWD/tools/ab/libhbird_b200_r1.so PROBE_FAST=1 PROBE_TAG=_r1 timeout 600 python tools/gpu_probe.py perf > gpurun_out/probe_perf_r1.log 2>&1
timeout 300 python tools/epi_stats.py > gpurun_out/epi_stats_new.log 2>&1
HBIRD_B200_AB_LIB=$PWD/tools/ab/libhbird_b200_r1.so timeout 300 python tools/epi_stats.py > gpurun_out/epi_stats_r1.log 2>&1
tail -4 gpurun_out/epi_stats_new.log gpurun_out/epi_stats_r1.log
python - <<'PY'
import json
for tag in ("_new", "_r1"):
    try:
        d = json.load(open(f"gpurun_out/probe_perf{tag}.json"))
    except Exception as e:
        print(tag, "missing", e); continue
    for k, v in d.items():
        if k.startswith("perf_"):
            print(tag, k, round(v["ms"], 3), round(v["frac_sustained"], 3))
PY
