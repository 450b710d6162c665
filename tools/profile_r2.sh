#!/bin/bash
# round 2 ncu evidence (1 GPU, under gpurun): launch lists + one full-set capture per kernel of the step.
#   launches_<wl>.csv     every launch of a short run with its device time (shares, not absolutes)
#   prof_<kernel>_<wl>    ncu --set full of the 3rd launch of that kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for WL in cfg3 cfg2 cfg1; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${WL}.csv \
      python tools/profile_target.py $WL 3 > gpurun_out/launches_${WL}.stdout 2>&1
done
prof() {  # kernel regex, tag, workload, [rows]
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -f -o gpurun_out/prof_$2_$3 \
      python tools/profile_target.py $3 3 $4 > gpurun_out/prof_$2_$3.stdout 2>&1
  echo "prof $2 $3 exit $?"
}
prof search_topk search cfg3
prof search_topk search cfg2
prof search_topk search cfg1
prof rerank_kernel rerank cfg3
prof predict_score tail cfg3
prof predict_score tail cfg1
prof pack_rows pack cfg3
ls -la gpurun_out/*.ncu-rep
