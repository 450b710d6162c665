"""Timeline of the two-stream batch pipeline on one GPU (CUDA events on both streams): when does
each batch's search kernel start/end, when do its re-rank (K2b+K4a) and tail start/end?  Shows
whether the post-processing really executes UNDER the next search.  Bench synthetic data, a cfg3
row shard by default.

    python tools/pipe_timeline.py [rows] [cfg]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
import bench  # noqa: E402
from hbird_b200 import ops  # noqa: E402
from hbird_b200 import pipeline as hpipe  # noqa: E402

DEV = torch.device("cuda", 0)
torch.cuda.set_device(DEV)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_280_000
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
W = dict(bench.WORKLOADS[cfg], N=rows)
ring = bench.make_query_ring(W, DEV)
bank = bench.build_bank(W, 0, rows, DEV)
table = bank.label_table()
conf = torch.zeros((W["C"], W["C"]), dtype=torch.int64, device=DEV)
S, B, H = W["S"], W["B"], W["S"] * W["ps"]
K, KP, BETA = bench.K_NEIGH, bench.K_PRIME, bench.BETA


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(n_steps, mode):
    """mode: 'lib' = post released at the prepared event (pipeline.py today); 'early' = post issued
    before the next search_begin; 'serial' = one stream."""
    mma, post = hpipe.make_streams(DEV)
    marks = []
    pending = None
    base = ev()
    base.record()
    mma.wait_stream(torch.cuda.current_stream())
    post.wait_stream(torch.cuda.current_stream())

    def do_post(item):
        slot, q, y, qn, searched, m = item
        post.wait_event(searched)
        with torch.cuda.stream(post):
            m["k2b0"] = ev(); m["k2b0"].record(post)
            lh, _, _ = bank.search_finish(slot, q, K, 0, BETA, None)
            m["k2b1"] = ev(); m["k2b1"].record(post)
            ops.predict_score(lh, B, S, H, H, conf, y=y, ignore_index=W["ignore"])
            m["tail1"] = ev(); m["tail1"].record(post)

    for i in range(n_steps):
        q, y = ring[i % len(ring)]
        slot = i & 1
        m = {}
        marks.append(m)
        if mode == "early" and pending is not None:
            do_post(pending)
            pending = None
        prepared = torch.cuda.Event()
        with torch.cuda.stream(mma):
            m["k2_0"] = ev(); m["k2_0"].record(mma)
            qn = bank.search_begin(q, KP, slot, prepared)
            m["k2_1"] = ev(); m["k2_1"].record(mma)
            searched = torch.cuda.Event(); searched.record(mma)
        if pending is not None:
            post.wait_event(prepared)
            do_post(pending)
        pending = (slot, q, y, qn, searched, m)
        if mode == "serial":
            do_post(pending)
            pending = None
            mma.wait_stream(post)
    if pending is not None:
        do_post(pending)
    torch.cuda.synchronize()
    out = []
    for m in marks:
        out.append({k: round(base.elapsed_time(v), 3) for k, v in m.items()})
    return out


res = {}
COMBOS = [("lib", 0, 4, -1), ("lib", 1, 4, -1), ("lib", 1, 4, 100), ("lib", 0, 4, 100), ("lib", 1, 2, 100), ("lib", 1, 1, 100),
          ("serial", 0, 4, -1), ("serial", 1, 4, -1), ("serial", 1, 4, 100)]
for mode, lean, wpb, carve in COMBOS:
    bank.configure_coresidency(bool(lean), wpb, carve)
    run(6, mode)  # warm
    tl = run(10, mode)
    res[f"{mode}_lean{lean}_wpb{wpb}_carve{carve}"] = tl
    print("mode", mode, "lean search", lean, "re-rank warps per CTA", wpb, "carve-out", carve)
    for i, m in enumerate(tl):
        if i not in (1, 2, 3, 8):
            continue
        prev = tl[i - 1]
        print(f"  step {i}: K2 {m['k2_0']:9.3f} -> {m['k2_1']:9.3f} ({m['k2_1'] - m['k2_0']:7.3f} ms) | "
              f"K2b {m['k2b0']:9.3f} -> {m['k2b1']:9.3f} ({m['k2b1'] - m['k2b0']:7.3f}) | tail end {m['tail1']:9.3f} (+{m['tail1'] - m['k2b1']:6.3f})")
    span = tl[-1]["k2_1"] - tl[1]["k2_1"]
    print(f"  per step over the last {len(tl) - 2} K2 ends: {span / (len(tl) - 2):.3f} ms", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"pipe_timeline_{cfg}_{rows}.json"), "w"), indent=1)
