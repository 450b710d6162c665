#!/bin/bash
# 2 GPUs, exactly what the driver runs at N = 2 (all by_workload extras), plus the GPU test suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n2.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_n2.log
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/final_bench_cfg3_n2.json 2> gpurun_out/final_bench_cfg3_n2.err
echo "bench n2 exit $? wall ${SECONDS}s"; tail -2 gpurun_out/final_bench_cfg3_n2.err | cut -c1-300
python - <<'PY'
import json
j = json.load(open("gpurun_out/final_bench_cfg3_n2.json")); r = j["roofline"]
print(round(j["value"]), "ms", round(j["ms_per_step"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "parity", j["sharded"]["parity"]["ok"], j["sharded"]["exchange"], j["clocks"]["sm_mhz"])
for k, v in j["by_workload"].items():
    print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), "unpipelined", round(v["ms_per_step_unpipelined"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3), v["layout"][:30])
for k, v in j["roofline_hbm"].items():
    print("  ", k, round(v["ms"], 3), round(v["frac"], 3))
PY
