#!/bin/bash
# round 2, call 9 (1 GPU): GPU test-suite after the exchange / pipeline changes, ncu of the final K2b+K4a
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 2 -c 1 -f -o gpurun_out/prof_rerank_cfg3 \
    python tools/profile_target.py cfg3 3 > gpurun_out/prof_rerank_cfg3.stdout 2>&1
echo "ncu rerank exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_topk -s 2 -c 1 -f -o gpurun_out/prof_search_cfg3shard \
    python tools/profile_target.py cfg3 3 1280000 > gpurun_out/prof_search_cfg3shard.stdout 2>&1
echo "ncu search shard exit $?"
