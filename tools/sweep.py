"""cfg5: search-only throughput vs bank size (d=768, k=30) on one GPU shard; banks are generated on the
device.  `--rows` are per-GPU shard sizes (1e8 over 8 GPUs = 12.5M rows per shard)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
from hbird_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, nargs="+", default=[100_000, 1_000_000, 10_000_000, 12_500_000])
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--Q", type=int, default=65536)
args = ap.parse_args()
DEV = torch.device("cuda", 0)
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
except Exception:
    PEAK = 1400.0
out = {}
g = torch.Generator(device=DEV).manual_seed(2)
q = torch.randn((args.Q, args.d), generator=g, device=DEV) * 3
for N in args.rows:
    bank = ops.MemoryBank(args.d, 1, 1, N, 0, True)
    one = torch.ones((1 << 20, 1), device=DEV)
    for a in range(0, N, 1 << 20):
        n = min(1 << 20, N - a)
        bank.append_soft(torch.randn((n, args.d), generator=g, device=DEV), one[:n], normalise=True)
    bank.finalize()
    bank.enable_kernel_timing(True)
    iters = 3 if N >= 5_000_000 else 10
    for _ in range(2):
        bank.search(q, 30, 64)
    torch.cuda.synchronize()
    bank.enable_kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        bank.search(q, 30, 64)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    kms, _ = bank.kernel_time_ms()
    tf = 2.0 * N * args.d * args.Q / (kms * 1e-3) / 1e12
    out[str(N)] = dict(rows=N, d=args.d, Q=args.Q, ms=ms, qps=args.Q / ms * 1e3, kernel_ms=kms, kernel_tflops=tf, frac_sustained=tf / PEAK)
    print(N, json.dumps(out[str(N)]))
    bank.close()
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_d{args.d}.json"), "w"), indent=1)
