#!/bin/bash
# round 2, call 3 (2 GPUs): multi-GPU engine check (both layouts), bench at N=2, reference arm under torchrun
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/smi2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -30 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py > gpurun_out/dist_check_n2.json 2> gpurun_out/dist_check_n2.err
echo "dist_check exit $?"; tail -c 1500 gpurun_out/dist_check_n2.json
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n2.json 2> gpurun_out/bench_cfg3_n2.err
echo "bench n2 exit $? wall ${SECONDS}s" | tee -a gpurun_out/bench_cfg3_n2.err
tail -5 gpurun_out/bench_cfg3_n2.err
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref n2 exit $? wall ${SECONDS}s"
python - <<'PY'
import json
for f in ("bench_cfg3_n2", "bench_ref_n2"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
        print(f, j["value"], j.get("ms_per_step"), (j.get("roofline") or {}).get("frac"), (j.get("e2e") or {}).get("value"), (j.get("cpu_baseline") or {}).get("cores"))
        if "sharded" in j:
            print(" parity", j["sharded"]["parity"])
            for k, v in j["by_workload"].items():
                print(" ", k, round(v["value"]), round(v["ms_per_step"], 3), v["search_kernel_frac_of_sustained_bf16"], v["layout"])
    except Exception as e:
        print(f, "unreadable", e)
PY
