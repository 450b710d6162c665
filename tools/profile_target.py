"""ncu target: a few validation steps of one workload through the 4-launch path (hb_eval_step), plus
one bank-pack call, with nothing else on the GPU — bench.py's cpu/parity/extras legs would only add
replay time under the profiler.  Usage: profile_target.py <workload> [steps]   (run under ncu)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 3:  # optional row count: a shard of the workload's bank (e.g. cfg3 / 8 GPUs)
    w["N"] = int(sys.argv[3])
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
bank = bench.build_bank(w, 0, w["N"], dev)
ring = bench.make_query_ring(w, dev, 0, 2)
conf = torch.zeros((w["C"], w["C"]), dtype=torch.int64, device=dev)
for i in range(steps):
    q, y = ring[i % 2]
    bank.eval_step(q, y, w["S"], conf, w["ignore"], bench.K_NEIGH, bench.K_PRIME, bench.BETA)
torch.cuda.synchronize()
print("conf sum", int(conf.sum()))
