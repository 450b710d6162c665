"""ncu target: K5 alone on a val-set sized smooth pixel stream (run under tools' ncu command)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200")); sys.path.insert(0, ROOT)
from hbird_b200 import ops
from bench import WORKLOADS, synth_images
DEV = torch.device("cuda", 0)
w = WORKLOADS["cfg2"]
gen = torch.Generator(device=DEV).manual_seed(1)
_, maps = synth_images(w, w["B"], gen, DEV)
gt = maps.contiguous()
reps = (256 << 20) // gt.numel()
gtl = gt.flatten().repeat(reps)
prl = torch.roll(gt, 3, dims=-1).flatten().repeat(reps).clamp_(max=w["C"] - 1)
conf = torch.zeros((w["C"], w["C"]), dtype=torch.int64, device=DEV)
for _ in range(3):
    ops.confusion_accumulate(conf, gtl, prl, w["ignore"])
torch.cuda.synchronize()
