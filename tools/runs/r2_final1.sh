#!/bin/bash
# final lines, 1 GPU: default bench (as the driver runs it) + reference arm (driver's K/W)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python bench.py > gpurun_out/final_bench_cfg3_n1.json 2> gpurun_out/final_bench_cfg3_n1.err
echo "bench exit $? wall ${SECONDS}s"
SECONDS=0
timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/final_bench_reference_cfg3.json 2> gpurun_out/final_bench_reference_cfg3.err
echo "ref exit $? wall ${SECONDS}s"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
j = json.load(open("gpurun_out/final_bench_cfg3_n1.json")); r = j["roofline"]
print(round(j["value"]), round(j["ms_per_step"], 2), "K2", round(r["kernel_ms"], 2), round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "cpu", j["cpu_baseline"]["value"], j["parity"]["ok"], j["clocks"])
for k, v in j["by_workload"].items():
    print("  ", k, round(v["value"]), round(v["ms_per_step"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3), v["pipelined"], v.get("value_cuda_graph"))
j = json.load(open("gpurun_out/final_bench_reference_cfg3.json")); print("ref", j["value"], j["cpu_baseline"])
PY
