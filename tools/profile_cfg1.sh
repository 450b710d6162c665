#!/bin/bash
# ncu full capture of the search kernel on the cfg1 shape (small bank: threshold start-up transient)
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
bank.configure_search(cta_group=cg)
for _ in range(3): bank.search(q, 30, 64)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:search_topk -s 1 -c 1 \
    -o gpurun_out/prof_cfg1 -f python /tmp/prof_one.py 2 384 12544 102400 > gpurun_out/prof_cfg1.stdout 2>&1
ls -la gpurun_out/prof_cfg1.ncu-rep
