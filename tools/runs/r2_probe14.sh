#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload cfg2 --extras cfg1,cfg5 --no-cpu-baseline > gpurun_out/bench_cfg2_p.json 2> gpurun_out/bench_cfg2_p.err
echo "bench exit $?"; tail -2 gpurun_out/bench_cfg2_p.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_cfg2_p.json")); r = j["roofline"]
print(round(j["value"]), "ms", round(j["ms_per_step"], 3), "unpipelined", round(r["ms_per_step_unpipelined"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), j["parity"]["ok"])
for k, v in j["by_workload"].items():
    print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), "unpipelined", round(v["ms_per_step_unpipelined"], 3), "K2", round(v["search_kernel_ms"], 3), v["pipelined"])
PY
