#!/bin/bash
# Runs the staged GPU probe; each stage in its own process with a timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for st in "$@"; do
  echo "=== stage $st ==="
  timeout 600 python tools/gpu_probe.py "$st" > gpurun_out/probe_${st}.log 2>&1
  echo "exit $?" >> gpurun_out/probe_${st}.log
  tail -c 3000 gpurun_out/probe_${st}.log
done
