"""cfg5 on N GPUs: search-only patch-queries/s vs TOTAL bank size with the bank row-sharded over the
ranks (one process per GPU, torchrun): per-shard tcgen05 search -> NCCL all-gather of (score, idx)[Q,k]
-> k-way merge kernel.  Banks are generated on the device per shard (1e8 x 768 never exists on a host)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
from hbird_b200 import distributed as hdist  # noqa: E402
from hbird_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, nargs="+", default=[100_000, 1_000_000, 10_000_000, 100_000_000])
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--Q", type=int, default=65536)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
DEV = torch.device("cuda", local)
torch.cuda.set_device(DEV)
if world > 1:
    dist.init_process_group("nccl", device_id=DEV)
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
except Exception:
    PEAK = 1400.0
out = {}
gq = torch.Generator(device=DEV).manual_seed(2)  # every rank sees the same queries
q = torch.randn((args.Q, args.d), generator=gq, device=DEV) * 3
g = torch.Generator(device=DEV).manual_seed(100 + rank)
for N in args.rows:
    a, b = hdist.shard_bounds(N, world, rank)
    n = b - a
    bank = ops.MemoryBank(args.d, 1, 1, max(n, 1), local, True)
    one = torch.ones((1 << 20, 1), device=DEV)
    for s0 in range(0, n, 1 << 20):
        m = min(1 << 20, n - s0)
        bank.append_soft(torch.randn((m, args.d), generator=g, device=DEV), one[:m], normalise=True)
    bank.finalize()

    def step():
        s, i, _ = bank.search(q, 30, 64, idx_offset=a)
        if world > 1:
            gs, gi = hdist.all_gather_topk(s, i)
            s, i = ops.merge_topk(gs, gi)
        return s, i

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    iters = 3 if N >= 50_000_000 else 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        s, i = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=DEV)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    ok = bool((s[:, :-1] >= s[:, 1:]).all()) and bool(((i >= 0) & (i < N)).all())
    tf = 2.0 * N * args.d * args.Q / (ms * 1e-3) / 1e12
    out[str(N)] = dict(total_rows=N, rows_per_gpu=n, gpus=world, d=args.d, Q=args.Q, ms=ms, qps=args.Q / ms * 1e3,
                       aggregate_tflops=tf, frac_of_sustained_peak_per_gpu=tf / world / PEAK, sorted_and_in_range=ok)
    if rank == 0:
        print(N, json.dumps(out[str(N)]), flush=True)
    bank.close()
    torch.cuda.empty_cache()
if rank == 0:
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_sharded_d{args.d}_n{world}.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
