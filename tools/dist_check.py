"""Run under torchrun (one rank per GPU): the row-sharded engine must reproduce the reference's
golden result (tests/golden) — per-shard search, shard exchange (fused peer-memory exchange and
the NCCL all-gather + K3 merge path, which must agree bit for bit), replicated label table,
all-reduced confusion matrix."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-hummingbird-eval_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import load_golden  # noqa: E402
from hbird_b200 import HbirdEvaluation  # noqa: E402
from hbird_b200.data import SyntheticSegmentationData  # noqa: E402
from hbird_b200.models import FeatureExtractorSimple  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
report = {}
for name in ("voc_tiny", "ade_tiny"):
    cfg, g = load_golden(name)
    data = SyntheticSegmentationData(**cfg)
    confs = {}
    for mode in ("p2p", "nccl"):
        fe = FeatureExtractorSimple(data.model, data.ftr_extr_fn, data.S, data.d)
        ev = HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=30,
                             device=f"cuda:{local}", nn_method="b200", nn_params={"exchange": mode},
                             dataset_size=data.get_train_dataset_size())
        miou, det = ev.evaluate(data.val_dataloader(), data.S, return_knn_details=True, ignore_index=data.ignore_index)
        conf = ev.last_confusion
        confs[mode] = conf
        # details cover the whole val set on every rank, as the reference returns them; global row ids
        # are ordered rank-major here, so compare the gathered neighbour FEATURES and labels, not ids
        ref_knn_f = g["feature_memory"][g["knn_idx"]].reshape(det["knns"].shape)
        ref_knn_l = g["label_memory"][g["knn_idx"]].reshape(det["knns_labels"].shape)
        same_f = (np.abs(det["knns"].numpy() - ref_knn_f).max(-1) <= 1e-6).mean()
        same_l = (np.abs(det["knns_labels"].numpy() - ref_knn_l).max(-1) <= 1e-6).mean()
        details_ok = same_f >= 0.999 and same_l >= 0.999 and \
            np.abs(det["knns_ca_labels"].numpy() - g["label_hat"]).max() <= 2e-5
        rows = ev.shard_counts
        fused = getattr(ev, "_xchg", None) is not None
        good = abs(miou - float(g["miou"])) <= 5e-4 and conf.sum() == g["conf"].sum() and \
            np.abs(conf - g["conf"]).sum() <= 2e-4 * conf.sum() and sum(rows) == g["feature_memory"].shape[0] and \
            len(rows) == world and fused == (mode == "p2p") and bool(details_ok)
        report[f"{name}_{mode}"] = {"miou": miou, "ref": float(g["miou"]), "shard_rows": rows,
                                    "fused_exchange": fused, "details_ok": bool(details_ok), "ok": bool(good)}
        ok = ok and good
        ev.close()
    # memory files with a sharded bank: rank 0 writes ONE (N, d) / (N, C) pair (hbird_eval.py:371-378);
    # reloading re-shards it by contiguous row ranges and must reproduce the evaluation
    tmp = f"/tmp/hb_dist_check_{os.environ.get('MASTER_PORT', '0')}_{name}"
    if rank == 0:
        os.makedirs(tmp, exist_ok=True)
    dist.barrier()
    fe = FeatureExtractorSimple(data.model, data.ftr_extr_fn, data.S, data.d)
    ev = HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=30, device=f"cuda:{local}",
                         nn_method="b200", dataset_size=data.get_train_dataset_size(),
                         f_mem_p=os.path.join(tmp, "f.pt"), l_mem_p=os.path.join(tmp, "l.pt"))
    saved_f, saved_l = torch.load(os.path.join(tmp, "f.pt")).numpy(), torch.load(os.path.join(tmp, "l.pt")).numpy()
    order_s = np.lexsort(np.concatenate([saved_f, saved_l], 1).T[::-1])
    order_g = np.lexsort(np.concatenate([g["feature_memory"], g["label_memory"]], 1).T[::-1])
    files_ok = saved_f.shape == g["feature_memory"].shape and \
        np.abs(saved_f[order_s] - g["feature_memory"][order_g]).max() <= 2e-7 and \
        np.array_equal(saved_l[order_s], g["label_memory"][order_g])
    reloaded = ev.load_memory()
    miou2 = ev.evaluate(data.val_dataloader(), data.S, ignore_index=data.ignore_index)
    reload_ok = bool(reloaded) and sum(ev.shard_counts) == g["feature_memory"].shape[0] and \
        abs(miou2 - float(g["miou"])) <= 5e-4
    report[f"{name}_save_load"] = {"files_ok": bool(files_ok), "reload_ok": bool(reload_ok), "miou": miou2,
                                   "shard_rows": ev.shard_counts}
    ok = ok and bool(files_ok) and reload_ok
    ev.close()
    same = bool((confs["p2p"] == confs["nccl"]).all())
    report[f"{name}_paths_identical"] = same
    ok = ok and same
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "ok": bool(flag.item()), **report}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
