#!/bin/bash
# 8 GPUs: N=8 headline with the one-warp exchange wait, pipelined vs one-call in the same run
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 --extras cfg2,cfg5 > gpurun_out/bench_cfg3_n8.json 2> gpurun_out/bench_cfg3_n8.err
echo "bench n8 exit $? wall ${SECONDS}s"; tail -2 gpurun_out/bench_cfg3_n8.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_cfg3_n8.json"))
r = j["roofline"]
print(8, round(j["value"]), "ms", round(j["ms_per_step"], 3), "unpipelined", round(r["ms_per_step_unpipelined"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "nccl", j["sharded"]["nccl_path_ms_per_step"], "parity", j["parity"]["ok"], j["clocks"])
for k, v in j["by_workload"].items():
    print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), "unpipelined", round(v["ms_per_step_unpipelined"], 3), "K2", round(v["search_kernel_ms"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3), v["pipelined"])
PY
