#!/bin/bash
# ncu evidence for the search kernel (run under gpurun, 1 GPU): launch list of a short bench run and
# one full-set capture of the tcgen05 kernel.  Outputs land in gpurun_out/.
cd "$(dirname "$0")/.."
WL=${1:-cfg2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${WL}.csv python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${WL}.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_topk -s 3 -c 2 \
    -o gpurun_out/prof_search_${WL} -f python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/prof_search_${WL}.stdout 2>&1
ls -la gpurun_out | tail -8
