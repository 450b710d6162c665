import json
import os

import numpy as np

from hbird_b200.data import SyntheticSegmentationData

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    cfg = json.loads(str(z["cfg"]))
    return cfg, {k: z[k] for k in z.files if k != "cfg"}


def batches_np(data: SyntheticSegmentationData, loader):
    """(features (B, S*S, d) fp32, y (B,1,H,W) fp32) numpy batches, features through the same
    table-lookup extractor the CUDA path uses."""
    out = []
    for x, y in loader:
        f, _ = data.ftr_extr_fn(data.model, x)
        out.append((f.numpy(), y.numpy()))
    return out


def recall(idx, ref_idx):
    hit = (idx[:, :, None] == ref_idx[:, None, :]).any(axis=2)
    return float(hit.mean())
