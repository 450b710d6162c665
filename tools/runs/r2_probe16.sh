#!/bin/bash
# 2 GPUs: engine check in every layout / exchange mode, then the bench with the threshold exchange and,
# for comparison, with the full per-shard re-rank (headline only)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "matches_reference" > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?"; tail -3 gpurun_out/pytest_multi.log
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 3 --extras cfg2_sharded,cfg5_1e6_sharded > gpurun_out/bench_cfg3_n2_thr.json 2> gpurun_out/bench_cfg3_n2_thr.err
echo "bench threshold exit $? wall ${SECONDS}s"; tail -2 gpurun_out/bench_cfg3_n2_thr.err
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 3 --exchange full --extras cfg2_sharded,cfg5_1e6_sharded > gpurun_out/bench_cfg3_n2_full.json 2> gpurun_out/bench_cfg3_n2_full.err
echo "bench full exit $? wall ${SECONDS}s"; tail -2 gpurun_out/bench_cfg3_n2_full.err
python - <<'PY'
import json
for f in ("thr", "full"):
    j = json.load(open(f"gpurun_out/bench_cfg3_n2_{f}.json")); r = j["roofline"]
    print(f, round(j["value"]), "ms", round(j["ms_per_step"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]),
          "parity", j["sharded"]["parity"]["ok"], j["sharded"]["exchange"], j["gpu_launches_per_step"], j["clocks"]["sm_mhz"])
    for k, v in j["by_workload"].items():
        print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), "unpipelined", round(v["ms_per_step_unpipelined"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3))
PY
