"""HBM-bound kernels of the path (K1 pack, K2b re-rank, K4a label transfer, K4b upsample+argmax, K5
confusion, mask decode) timed alone on the BASELINE shapes: algorithmic GB/s vs the measured HBM
copy bandwidth (MEASURED_PEAKS.json hbm_gbs).  Run under gpurun; writes gpurun_out/kernels.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
sys.path.insert(0, ROOT)
from hbird_b200 import ops  # noqa: E402
from bench import WORKLOADS, synth_images  # noqa: E402

DEV = torch.device("cuda", 0)
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    HBM = 6650.0


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # L2 flush between timed iterations
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


res = {"hbm_gbs_peak": HBM}
for name in ("cfg2", "cfg3", "cfg4"):
    w = WORKLOADS[name]
    S, ps, C, d, B = w["S"], w["ps"], w["C"], w["d"], w["B"]
    H = S * ps
    Q = B * S * S
    gen = torch.Generator(device=DEV).manual_seed(1)
    n_img = max(B, (1 << 19) // (S * S))
    feats, maps = synth_images(w, n_img, gen, DEV)
    rows = n_img * S * S
    # K1: pack
    bank = ops.MemoryBank(d, C, ps * ps, rows * 25, 0, True)
    def k1():
        if bank.rows + rows > bank.capacity:
            return
        bank.append(feats, maps, S, ps)
    ms = timeit(k1, 20)
    bytes_k1 = rows * (4 * d + ps * ps + 2 * bank.d + 0 + 4 * d + 2 * C)
    dpad = (d + 63) // 64 * 64
    bytes_k1 = rows * (4 * d + ps * ps + 2 * dpad + 4 * d + 2 * C)
    res[f"{name}_K1_pack"] = dict(ms=ms, rows=rows, gbs=bytes_k1 / ms / 1e6, frac=bytes_k1 / ms / 1e6 / HBM)
    bank.finalize()
    table = bank.label_table()
    N = bank.rows
    # decode
    y = (maps[:B].float() / 255).unsqueeze(1).contiguous()
    ms = timeit(lambda: ops.decode_mask(y, False))
    res[f"{name}_decode"] = dict(ms=ms, gbs=y.numel() * 5 / ms / 1e6, frac=y.numel() * 5 / ms / 1e6 / HBM)
    # K4a: label transfer on random neighbour ids
    idx = torch.randint(0, N, (Q, 30), generator=gen, device=DEV)
    sc = torch.rand((Q, 30), generator=gen, device=DEV) * 3
    qn = torch.rand((Q,), generator=gen, device=DEV) + 3
    ms = timeit(lambda: ops.label_transfer(table, ps * ps, sc, idx, qn, 0.02))
    by = Q * (30 * 12 + 30 * 2 * C + 4 * C)
    res[f"{name}_K4a_label_transfer"] = dict(ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / HBM)
    lh = ops.label_transfer(table, ps * ps, sc, idx, qn, 0.02)
    # K4b
    ms = timeit(lambda: ops.upsample_argmax(lh, B, S, H, H))
    by = B * (4 * S * S * C + H * H)
    res[f"{name}_K4b_upsample_argmax"] = dict(ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / HBM, pixels_per_s=B * H * H / ms * 1e3)
    pred = ops.upsample_argmax(lh, B, S, H, H)
    # K5 on a val-set sized pixel stream
    gt = maps[:B].contiguous()
    reps = max(1, (256 << 20) // gt.numel())
    gtl, prl = gt.flatten().repeat(reps), pred.flatten().repeat(reps)
    conf = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    ms = timeit(lambda: ops.confusion_accumulate(conf, gtl, prl, w["ignore"]))
    res[f"{name}_K5_confusion_noise_pred"] = dict(ms=ms, pixels=gtl.numel(), gbs=2 * gtl.numel() / ms / 1e6, frac=2 * gtl.numel() / ms / 1e6 / HBM)
    # realistic predictions are piecewise constant like the ground truth (upsampled patch labels):
    # the gt map shifted by 3 pixels disagrees with gt only along region borders
    prl = torch.roll(gt, 3, dims=-1).flatten().repeat(reps).clamp_(max=C - 1)
    ms = timeit(lambda: ops.confusion_accumulate(conf, gtl, prl, w["ignore"]))
    res[f"{name}_K5_confusion"] = dict(ms=ms, pixels=gtl.numel(), gbs=2 * gtl.numel() / ms / 1e6, frac=2 * gtl.numel() / ms / 1e6 / HBM)
    bank.close()
    del feats, maps, bank, table
    torch.cuda.empty_cache()
for k, v in res.items():
    print(k, json.dumps(v))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kernels.json"), "w"), indent=1)
