#!/bin/bash
# 1 GPU: the whole GPU test suite, the default bench line (cfg3 + by_workload), smoke, ncu of the exchange kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
SECONDS=0
timeout 1200 python bench.py > gpurun_out/final_bench_cfg3_n1.json 2> gpurun_out/final_bench_cfg3_n1.err
echo "bench exit $? wall ${SECONDS}s"; tail -2 gpurun_out/final_bench_cfg3_n1.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/final_bench_cfg3_n1.json")); r = j["roofline"]
print(round(j["value"]), "ms", round(j["ms_per_step"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), j["parity"]["ok"], j["clocks"]["sm_mhz"])
for k, v in j["by_workload"].items():
    print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3))
for k, v in j["roofline_hbm"].items():
    print("  ", k, round(v["ms"], 3), round(v["frac"], 3))
PY
for K in shortlist_kernel "rerank_kernel.*Lb1EE"; do
  T=$(echo $K | cut -c1-9)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -f -o gpurun_out/prof_xchg_$T \
      python tools/xchg_probe.py 4 1280000 cfg3 > gpurun_out/prof_xchg_$T.stdout 2>&1
  echo "prof $K exit $?"
done
ls -la gpurun_out/prof_xchg*
