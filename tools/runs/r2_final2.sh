#!/bin/bash
# final lines, N GPUs (N = number of visible GPUs; with 8 also N=4): the driver's torchrun commands
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in $NG $( [ "$NG" = "8" ] && echo 4 ); do
  SECONDS=0
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/final_bench_cfg3_n$N.json 2> gpurun_out/final_bench_cfg3_n$N.err
  echo "bench n$N exit $? wall ${SECONDS}s"
done
if [ "$NG" = "2" ]; then
  SECONDS=0
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29590 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/final_bench_reference_cfg3_n2.json 2> gpurun_out/final_bench_reference_cfg3_n2.err
  echo "ref n2 exit $? wall ${SECONDS}s"
fi
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/final_bench_*_n[248].json")):
    j = json.load(open(f))
    if j.get("impl") == "reference":
        print(f, j["value"], j["cpu_baseline"]["cores"]); continue
    r = j["roofline"]
    print(f, round(j["value"]), "ms", round(j["ms_per_step"], 3), "unpipelined", round(r["ms_per_step_unpipelined"], 3), "K2", round(r["kernel_ms"], 3), "frac", round(r["frac"], 4), "e2e", round(j["e2e"]["value"]), "nccl", j["sharded"]["nccl_path_ms_per_step"], "parity", j["parity"]["ok"], j["clocks"]["sm_mhz"])
    for k, v in j["by_workload"].items():
        print("  ", k, round(v["value"]), "ms", round(v["ms_per_step"], 3), round(v["search_kernel_frac_of_sustained_bf16"], 3), v["pipelined"])
PY
