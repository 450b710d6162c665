#!/bin/bash
# N GPUs (default 8): headline only, threshold exchange vs full per-shard re-rank on the same box
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for mode in threshold full; do
  SECONDS=0
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$((RANDOM % 10)) bench.py --gpus $N --steps 20 --warmup 3 --exchange $mode --no-extras > gpurun_out/bench_cfg3_n${N}_$mode.json 2> gpurun_out/bench_cfg3_n${N}_$mode.err
  echo "bench $mode exit $? wall ${SECONDS}s"; tail -2 gpurun_out/bench_cfg3_n${N}_$mode.err
done
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in ("threshold", "full"):
    j = json.load(open(f"gpurun_out/bench_cfg3_n{N}_{f}.json")); r = j["roofline"]
    print(f, round(j["value"]), "ms", round(j["ms_per_step"], 3), "unpipelined", round(r["ms_per_step_unpipelined"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]),
          "parity", j["sharded"]["parity"]["ok"], j["sharded"]["exchange"], "nccl", j["sharded"]["nccl_path_ms_per_step"], j["clocks"]["sm_mhz"])
    print("   ", j["sharded"]["balance"]["ms_per_step_equal_shards"], j["sharded"]["balance"]["rows_per_gpu"])
PY
