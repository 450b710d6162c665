"""Interleaved A/B timing of step variants on one GPU.  Clocks under the power cap drift by several
per cent within seconds, so variants are run in short alternating blocks (ABCABC...) for a long
total time and compared by their mean block time.  Bench synthetic data.

    python tools/ab_pipeline.py [rows] [cfg] [seconds]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
import bench  # noqa: E402
from hbird_b200 import pipeline as hpipe  # noqa: E402
from hbird_b200.pipeline import EvalPipeline  # noqa: E402

DEV = torch.device("cuda", 0)
torch.cuda.set_device(DEV)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_280_000
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 12.0
W = dict(bench.WORKLOADS[cfg], N=rows)
ring = bench.make_query_ring(W, DEV)
bank = bench.build_bank(W, 0, rows, DEV)
table = bank.label_table()
conf = torch.zeros((W["C"], W["C"]), dtype=torch.int64, device=DEV)
K, KP, BETA = bench.K_NEIGH, bench.K_PRIME, bench.BETA
streams = hpipe.make_streams(DEV)
counter = [0]


def one_call(n):
    for _ in range(n):
        q, y = ring[counter[0] % len(ring)]
        counter[0] += 1
        bank.eval_step(q, y, W["S"], conf, W["ignore"], K, KP, BETA)


def pipelined(n):
    pipe = EvalPipeline(bank, table, W["S"], conf, W["ignore"], K, KP, BETA, streams=streams)
    for _ in range(n):
        q, y = ring[counter[0] % len(ring)]
        counter[0] += 1
        pipe.submit(q, y, W["B"])
    pipe.flush()


def k2_only(n):
    for _ in range(n):
        bank.search_begin(ring[counter[0] % len(ring)][0], KP, 0)
        bank.search_abort()
        counter[0] += 1


VARIANTS = {
    "one_call": (one_call, (False, 4, -1)),
    "one_call_lean": (one_call, (True, 4, -1)),
    "pipelined": (pipelined, (False, 4, -1)),
    "pipelined_lean": (pipelined, (True, 4, -1)),
    "k2_only": (k2_only, (False, 4, -1)),
    "k2_only_lean": (k2_only, (True, 4, -1)),
}
names = [a for a in (sys.argv[4].split(",") if len(sys.argv) > 4 else VARIANTS)]
# block length: ~0.25 s of work
one_call(3)
torch.cuda.synchronize()
t0 = time.perf_counter()
one_call(4)
torch.cuda.synchronize()
per = (time.perf_counter() - t0) / 4
block = max(4, int(0.25 / per))
times = {n: [] for n in names}
t_end = time.perf_counter() + seconds
rnd = 0
while time.perf_counter() < t_end:
    order = names[rnd % len(names):] + names[:rnd % len(names)]  # rotate so no variant always follows the same one
    rnd += 1
    for n in order:
        fn, cfgv = VARIANTS[n]
        bank.configure_coresidency(*cfgv)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn(block)
        e1.record()
        torch.cuda.synchronize()
        times[n].append(e0.elapsed_time(e1) / block)
res = {"rows": rows, "cfg": cfg, "block_steps": block}
for n in names:
    v = sorted(times[n][1:] or times[n])
    res[n] = {"mean_ms": sum(v) / len(v), "median_ms": v[len(v) // 2], "min_ms": v[0], "blocks": len(v)}
    print(f"{cfg} rows={rows} {n:16s} mean {res[n]['mean_ms']:8.3f} median {res[n]['median_ms']:8.3f} min {res[n]['min_ms']:8.3f} ms/step over {len(v)} blocks of {block}", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"ab_pipeline_{cfg}_{rows}.json"), "w"), indent=1)
