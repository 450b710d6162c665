#!/bin/bash
# 8 GPUs: equal vs speed-balanced shards, full-length runs in the same process
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/final_bench_cfg3_n8.json 2> gpurun_out/final_bench_cfg3_n8.err
echo "bench n8 exit $? wall ${SECONDS}s"
python - <<'PY'
import json
j = json.load(open("gpurun_out/final_bench_cfg3_n8.json")); r = j["roofline"]
print(round(j["value"]), "ms", round(j["ms_per_step"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "parity", j["parity"]["ok"], j["clocks"]["sm_mhz"])
print(j["sharded"]["balance"])
for k, v in j["by_workload"].items():
    print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3), v["bank_rows_per_gpu"])
PY
