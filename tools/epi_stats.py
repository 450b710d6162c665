"""Per-warp cycle breakdown of the search epilogue (instrumented build)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
from hbird_b200 import ops
DEV = torch.device("cuda", 0)
for (name, Q, N, d) in [("cfg2", 12544, 1024000, 384), ("d768_1M", 21904, 1024000, 768)]:
    g = torch.Generator(device=DEV).manual_seed(5)
    bank = ops.MemoryBank(d, 1, 1, N, 0, True)
    bank.append_soft(torch.randn((N, d), generator=g, device=DEV), torch.ones((N, 1), device=DEV), normalise=True)
    bank.finalize()
    q = torch.randn((Q, d), generator=g, device=DEV) * 3
    for cg in (1, 2):
        bank.configure_search(cta_group=cg)
        buf = torch.zeros(148 * 8 * 8, dtype=torch.int64, device=DEV)
        bank.set_stats_buffer(buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bank.search(q, 30, 64); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        bank.set_stats_buffer(None)
        s = buf.view(148, 8, 8).double()
        tiles = s[:, :, 6].clamp_min(1)
        per_tile = {k: float((s[:, :, i] / tiles).mean()) for i, k in enumerate(["wait", "load", "slow", "fold_post"])}
        mx_tile = {k: float((s[:, :, i] / tiles).max()) for i, k in enumerate(["wait", "load", "slow", "fold_post"])}
        print(name, "cg", cg, "ms", round(ms, 3), "tiles/warp", float(tiles.mean()), "cycles per tile per warp (mean):", {k: round(v) for k, v in per_tile.items()},
              "max-warp:", {k: round(v) for k, v in mx_tile.items()}, "slow entries/tile", round(float((s[:, :, 4] / tiles).mean()), 2),
              "folds/tile", round(float((s[:, :, 5] / tiles).mean()), 3))
    bank.close()
