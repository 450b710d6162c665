#!/bin/bash
# ncu full capture of the search kernel on the probe shapes (cfg2-like, cg1 and cg2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/prof_one.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.path.join(os.getcwd(), "open-hummingbird-eval_b200"))
from hbird_b200 import ops
cg, d, Q, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
DEV = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(5)
feats = torch.randn((N, d), generator=g, device=DEV)
bank = ops.MemoryBank(d, 1, 1, N, 0, True)
bank.append_soft(feats, torch.ones((N, 1), device=DEV), normalise=True); bank.finalize()
q = torch.randn((Q, d), generator=g, device=DEV) * 3
bank.configure_search(cta_group=cg); bank.tune_search(4 if cg == 2 else 0, 0)
for _ in range(3): bank.search(q, 30, 64)
torch.cuda.synchronize()
PY
for cfg in "1 384 12544 1024000" "2 768 21904 1024000"; do
  set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:search_topk -s 1 -c 1 \
      -o gpurun_out/prof_cg$1_d$2 -f python /tmp/prof_one.py $cfg > gpurun_out/prof_cg$1_d$2.stdout 2>&1
done
ls -la gpurun_out/*.ncu-rep
