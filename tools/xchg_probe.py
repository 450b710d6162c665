"""Cost of the shard-side post-processing in the two exchange modes, on ONE GPU with G simulated
ranks (hb_exchange_connect_local) holding cfg3 row shards: mode 0 = every shard re-ranks its whole
bf16 top-k'; mode 1 = threshold exchange (shortlist + statistics, then re-rank of the survivors).
Times the three loops (scatter for all ranks, phase 2 for all ranks, merge for all ranks) with CUDA
events, interleaving the modes.  Writes gpurun_out/xchg_probe.json.

    python tools/xchg_probe.py [G] [rows_per_shard] [cfg]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
import bench  # noqa: E402
from hbird_b200 import distributed as hdist  # noqa: E402
from hbird_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
torch.cuda.set_device(DEV)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_280_000
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
W = dict(bench.WORKLOADS[cfg], N=rows * G)
K, KP, BETA = bench.K_NEIGH, bench.K_PRIME, bench.BETA
ring = bench.make_query_ring(W, DEV, n=4)
per = W["S"] ** 2
qsplit = hdist.query_split(W["B"], per, G)
banks = [bench.build_bank(W, r * rows, (r + 1) * rows, DEV) for r in range(G)]
table = torch.zeros((rows * G, W["C"]), dtype=torch.int16, device=DEV)
cap = max(qsplit[r + 1] - qsplit[r] for r in range(G))


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(mode, q):
    xs = run.xs[mode]
    e = [ev() for _ in range(4)]
    e[0].record()
    qns = []
    for r in range(G):
        qns.append(xs[r].search_scatter(banks[r], q, qsplit, K, KP, idx_offset=r * rows))
    e[1].record()
    if mode == 1:
        for r in range(G):
            xs[r].rerank()
    e[2].record()
    outs = []
    for r in range(G):
        a, b = qsplit[r], qsplit[r + 1]
        outs.append(xs[r].merge_transfer(table, W["ps"] ** 2, qns[r][a:b], BETA, return_neighbours=True))
    e[3].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(3)], outs


run.xs = {}
for mode in (0, 1):
    xs = [ops.ShardExchange(r, G, cap, K, 0) for r in range(G)]
    ops.ShardExchange.connect_local(xs)
    for x in xs:
        x.configure(bool(mode))
    run.xs[mode] = xs
acc = {0: [], 1: []}
agree = []
for it in range(6):
    q = ring[it % len(ring)][0]
    res = {}
    for mode in ((0, 1) if it % 2 == 0 else (1, 0)):
        t, outs = run(mode, q)
        res[mode] = outs
        if it > 0:
            acc[mode].append(t)
    same = [float((res[0][r][2] == res[1][r][2]).float().mean()) for r in range(G) if res[0][r][2].numel()]
    agree.append(min(same))
out = {"G": G, "rows_per_shard": rows, "cfg": cfg, "index_agreement_min": min(agree)}
for mode in (0, 1):
    m = [sum(t[i] for t in acc[mode]) / len(acc[mode]) for i in range(3)]
    out[f"mode{mode}_ms_all_ranks"] = {"scatter(K2+shard post)": m[0], "phase2": m[1], "merge": m[2], "total": sum(m)}
    print("mode", mode, "ms for all", G, "ranks:", [round(v, 3) for v in m], "total", round(sum(m), 3), flush=True)
d = (out["mode0_ms_all_ranks"]["total"] - out["mode1_ms_all_ranks"]["total"]) / G
out["saved_ms_per_rank_and_step"] = d
print("saved per rank and step:", round(d, 3), "ms; neighbour ids identical for >=", min(agree), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"xchg_probe_{cfg}_G{G}_{rows}.json"), "w"), indent=1)
