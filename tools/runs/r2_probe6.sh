#!/bin/bash
# round 2, call 6 (2 GPUs): pipelined path — GPU tests on one GPU, engine check + bench on two
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dist_check.py > gpurun_out/dist_check_n2.json 2> gpurun_out/dist_check_n2.err
echo "dist_check exit $?"; tail -c 600 gpurun_out/dist_check_n2.json; echo
SECONDS=0
timeout 900 python bench.py --extras cfg1,cfg2 > gpurun_out/bench_cfg3_p.json 2> gpurun_out/bench_cfg3_p.err
echo "bench n1 exit $? wall ${SECONDS}s"
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --extras cfg2 > gpurun_out/bench_cfg3_n2_p.json 2> gpurun_out/bench_cfg3_n2_p.err
echo "bench n2 exit $? wall ${SECONDS}s"; tail -3 gpurun_out/bench_cfg3_n2_p.err
python - <<'PY'
import json
for f in ("bench_cfg3_p", "bench_cfg3_n2_p"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
        r = j["roofline"]
        print(f, round(j["value"]), "ms", round(j["ms_per_step"], 3), "unpipelined", round(r["ms_per_step_unpipelined"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "parity", j["parity"]["ok"])
        print("  hbm", {k: (round(v["ms"], 3), round(v["frac"], 3)) for k, v in j["roofline_hbm"].items()})
        for k, v in j["by_workload"].items():
            print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), "unpipelined", round(v["ms_per_step_unpipelined"], 3), "K2", round(v["search_kernel_ms"], 3), v["search_kernel_frac_of_sustained_bf16"], v.get("ms_per_step_cuda_graph"))
    except Exception as e:
        print(f, "unreadable", e)
PY
