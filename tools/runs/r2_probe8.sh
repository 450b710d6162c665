#!/bin/bash
# round 2, call 8 (8 GPUs): the driver's scaling commands at N=8 and N=4 (4 of the 8 GPUs)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for N in 8 4; do
  SECONDS=0
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err
  echo "bench n$N exit $? wall ${SECONDS}s"
done
python - <<'PY'
import json
for n in (8, 4):
    try:
        j = json.load(open(f"gpurun_out/bench_cfg3_n{n}.json"))
        r = j["roofline"]
        print(n, round(j["value"]), "ms", round(j["ms_per_step"], 3), "unpipelined", round(r["ms_per_step_unpipelined"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "nccl", j["sharded"]["nccl_path_ms_per_step"], "parity", j["parity"]["ok"])
        for k, v in j["by_workload"].items():
            print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), "unpipelined", round(v["ms_per_step_unpipelined"], 3), "K2", round(v["search_kernel_ms"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3), v["pipelined"])
    except Exception as e:
        print(n, "unreadable", e)
PY
