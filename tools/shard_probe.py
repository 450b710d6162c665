"""Where does a short cfg3 row shard (10.24 M / 8 = 1.28 M rows x 768) lose against a long one?
One GPU, the bench's synthetic data.  For each bank size: the search kernel alone (prep + K2, no
finish) in its normal, GEMM-only (ablate 1) and scan-only (ablate 2) builds, each run back to back
for a few seconds (sustained clocks, SM clock sampled), then the whole step one-call and pipelined,
and the planner's chunk count varied.  Writes gpurun_out/shard_probe.json.

    python tools/shard_probe.py [rows ...]
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
from hbird_b200 import ops  # noqa: E402
from hbird_b200.pipeline import EvalPipeline  # noqa: E402

DEV = torch.device("cuda", 0)
torch.cuda.set_device(DEV)
W = dict(bench.WORKLOADS["cfg3"])
Q = W["B"] * W["S"] ** 2
SECONDS = float(os.environ.get("PROBE_SECONDS", "2.5"))
res = {}


def clocked(fn, seconds=SECONDS):
    """Run fn() back to back for `seconds` (after 0.7 s of the same load), return (median SM MHz, count)."""
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.7:
        fn()
        torch.cuda.synchronize()
    cs = bench.ClockSampler(0)
    cs.start()
    n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn()
        n += 1
        if n % 4 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return cs.stop()["sm_mhz"], n


def k2_only(bank, ring, label):
    i = [0]

    def fn():
        bank.search_begin(ring[i[0] % len(ring)][0], bench.K_PRIME, 0)
        bank.search_abort()  # host-side flag only: the slot may be begun again
        i[0] += 1

    bank.enable_kernel_timing(True)
    mhz, n = clocked(fn)
    ms, cnt = bank.kernel_time_ms()
    bank.enable_kernel_timing(False)
    bank.search_abort()
    tf = 2.0 * bank.rows * W["d"] * Q / (ms * 1e-3) / 1e12
    out = dict(k2_ms=ms, tflops=tf, mhz=mhz, ns_per_krow=ms * 1e6 / (bank.rows / 1e3), launches=cnt,
               flop_per_clk_sm=tf * 1e12 / (mhz * 1e6) / 148 if mhz else None)
    print(label, json.dumps(out), flush=True)
    return out


def steps(bank, table, ring, label, pipelined):
    conf = torch.zeros((W["C"], W["C"]), dtype=torch.int64, device=DEV)
    pipe = EvalPipeline(bank, table, W["S"], conf, W["ignore"], bench.K_NEIGH, bench.K_PRIME, bench.BETA)
    i = [0]

    def fn():
        q, y = ring[i[0] % len(ring)]
        i[0] += 1
        if pipelined:
            pipe.submit(q, y, W["B"])
        else:
            bank.eval_step(q, y, W["S"], conf, W["ignore"], bench.K_NEIGH, bench.K_PRIME, bench.BETA)

    for _ in range(3):
        fn()
    pipe.flush()
    torch.cuda.synchronize()
    bank.enable_kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mhz, n = clocked(fn)
    pipe.flush()
    e1.record()
    torch.cuda.synchronize()
    ms, _ = bank.kernel_time_ms()
    rr, _ = bank.rerank_time_ms()
    bank.enable_kernel_timing(False)
    out = dict(k2_ms=ms, rerank_ms=rr, mhz=mhz, steps=n)
    print(label, json.dumps(out), flush=True)
    return out


def step_ms(bank, table, ring, pipelined, n=60):
    """Plain event-timed mean step (n steps after a 1 s warm-up of the same)."""
    conf = torch.zeros((W["C"], W["C"]), dtype=torch.int64, device=DEV)
    pipe = EvalPipeline(bank, table, W["S"], conf, W["ignore"], bench.K_NEIGH, bench.K_PRIME, bench.BETA)

    def fn(i):
        q, y = ring[i % len(ring)]
        if pipelined:
            pipe.submit(q, y, W["B"])
        else:
            bank.eval_step(q, y, W["S"], conf, W["ignore"], bench.K_NEIGH, bench.K_PRIME, bench.BETA)

    t0 = time.perf_counter()
    i = 0
    while time.perf_counter() - t0 < 1.0:
        fn(i)
        i += 1
        torch.cuda.synchronize()
    pipe.flush()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(n):
        fn(i + j)
    pipe.flush()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ring = bench.make_query_ring(W, DEV)
sizes = [int(a) for a in sys.argv[1:]] or [1_280_000, 5_120_000]
for rows in sizes:
    w = dict(W, N=rows)
    bank = bench.build_bank(w, 0, rows, DEV)
    table = bank.label_table()
    r = res[str(rows)] = {}
    r["normal"] = k2_only(bank, ring, f"{rows} K2 alone")
    for ab in (1, 2):
        bank.tune_search(-1, ab)
        r[f"ablate{ab}"] = k2_only(bank, ring, f"{rows} K2 ablate={ab}")
    bank.tune_search(-1, 0)
    if rows <= 2_000_000:
        for mc in (3, 4, 5, 12, 13, 20):
            bank.configure_search(0, mc)
            r[f"max_chunks{mc}"] = k2_only(bank, ring, f"{rows} K2 max_chunks={mc}")
        bank.configure_search(0, 0)
        r["step_one_call_ms"] = step_ms(bank, table, ring, False)
        r["step_pipelined_ms"] = step_ms(bank, table, ring, True)
        print(rows, "step one-call", r["step_one_call_ms"], "pipelined", r["step_pipelined_ms"], flush=True)
        r["pipelined_detail"] = steps(bank, table, ring, f"{rows} pipelined detail", True)
        r["one_call_detail"] = steps(bank, table, ring, f"{rows} one-call detail", False)
    del table
    bank.close()
    torch.cuda.empty_cache()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "shard_probe.json"), "w"), indent=1)
