#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_multi.log
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 3 --extras cfg2 > gpurun_out/bench_cfg3_n2_bal.json 2> gpurun_out/bench_cfg3_n2_bal.err
echo "bench n2 exit $? wall ${SECONDS}s"; tail -3 gpurun_out/bench_cfg3_n2_bal.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_cfg3_n2_bal.json")); r = j["roofline"]
print(round(j["value"]), "ms", round(j["ms_per_step"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "parity", j["parity"]["ok"])
print(j["sharded"]["balance"])
for k, v in j["by_workload"].items():
    print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), v["bank_rows_per_gpu"])
PY
