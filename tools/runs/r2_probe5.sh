#!/bin/bash
# round 2, call 5 (8 GPUs): the driver's scaling command at N=8, reduced extras
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n8.json 2> gpurun_out/bench_cfg3_n8.err
echo "bench n8 exit $? wall ${SECONDS}s" | tee -a gpurun_out/bench_cfg3_n8.err
tail -3 gpurun_out/bench_cfg3_n8.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_cfg3_n8.json"))
print(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["kernel_ms"], j["e2e"]["value"], j["sharded"]["nccl_path_ms_per_step"], j["parity"]["ok"])
for k, v in j["by_workload"].items():
    print(" ", k, round(v["value"]), round(v["ms_per_step"], 3), v["search_kernel_ms"], v["search_kernel_frac_of_sustained_bf16"], v["layout"])
PY
