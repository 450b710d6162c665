"""Small-bank sweep: search-kernel time vs the number of bank chunks per query block (hb_search_config).
Note: variants run back to back under the power cap, so later ones see lower clocks (the same plan
measured 0.958 ms early and 1.040 ms late in one run); compare neighbours, not first with last."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
from hbird_b200 import ops
DEV = torch.device("cuda", 0)
PEAK = 1425.8
res = {}
for name, Q, N, d in (("cfg1", 12544, 102400, 384), ("shard128k", 12544, 128000, 384), ("shard256k", 12544, 256000, 384)):
    g = torch.Generator(device="cuda").manual_seed(5)
    feats = torch.randn((N, d), generator=g, device=DEV)
    bank = ops.MemoryBank(d, 1, 1, N, 0, True)
    bank.append_soft(feats, torch.ones((N, 1), device=DEV), normalise=True); bank.finalize()
    q = torch.randn((Q, d), generator=g, device=DEV) * 3
    for cg in (1, 2):
        for mc in (1, 2, 3, 4, 6, 0):
            bank.configure_search(cta_group=cg, max_chunks=mc)
            for _ in range(3): bank.search(q, 30, 64)
            bank.enable_kernel_timing(True)
            for _ in range(20): bank.search(q, 30, 64)
            torch.cuda.synchronize()
            ms, n = bank.kernel_time_ms()
            bank.enable_kernel_timing(False)
            frac = 2.0 * Q * N * d / (ms * 1e-3) / 1e12 / PEAK
            res[f"{name}_cg{cg}_mc{mc}"] = dict(ms=ms, frac=frac)
            print(f"{name} cg={cg} max_chunks={mc}: {ms:.3f} ms frac={frac:.3f}", flush=True)
    bank.close()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "chunk_sweep.json"), "w"), indent=1)
