"""Run under torchrun (one rank per GPU): the multi-GPU engine must reproduce the reference's golden
result (tests/golden) in both of the reference's faiss layouts (search_faiss.py:53-74):
row shards (nn_params idx_shard=True: per-shard search, shard exchange — the peer-memory threshold
exchange, the peer-memory exchange with a full per-shard re-rank and the NCCL all-gather + K3 merge path,
which must agree bit for bit — replicated label table,
features extracted once across the ranks and all-gathered) and replicas (idx_shard=False, the
default: full bank on every rank, validation batches dealt round-robin); all-reduced confusion
matrix in both.  Also: augmentation epochs with a loader length that does not divide by the world
size (per-rank capacity), memory save / load with a sharded bank."""
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
# HB_DIST_ONE_GPU=1: every rank uses cuda:0 and the collectives go through gloo (NCCL refuses two ranks on
# one device).  The kernels, the CUDA-IPC exchange windows and the engine logic are the same as on N GPUs —
# the processes merely time-share the device — so the multi-process path can be checked on a 1-GPU box.
ONE_GPU = os.environ.get("HB_DIST_ONE_GPU") == "1"
if ONE_GPU:
    local = 0
torch.cuda.set_device(local)
if ONE_GPU:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
report = {}
for name in ("voc_tiny", "ade_tiny"):
    cfg, g = load_golden(name)
    data = SyntheticSegmentationData(**cfg)
    confs = {}
    for mode in ("p2p", "p2p_full", "nccl", "replicas"):
        fe = FeatureExtractorSimple(data.model, data.ftr_extr_fn, data.S, data.d)
        nn_params = {"idx_shard": False} if mode == "replicas" else {"idx_shard": True, "exchange": mode}
        ev = HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=30,
                             device=f"cuda:{local}", nn_method="b200", nn_params=nn_params,
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
            len(rows) == (1 if mode == "replicas" else world) and fused == (mode in ("p2p", "p2p_full")) and bool(details_ok) and \
            ev.bank.rows == (sum(rows) if mode == "replicas" else rows[rank])
        report[f"{name}_{mode}"] = {"miou": miou, "ref": float(g["miou"]), "shard_rows": rows,
                                    "fused_exchange": fused, "details_ok": bool(details_ok), "ok": bool(good)}
        ok = ok and good
        if mode == "p2p":
            # speed-balanced shards move rows between ranks (here: by hand, 37 rows from rank 0 to the last
            # rank); global row ids stay, so the evaluation must not change by a single pixel
            moved = list(rows)
            moved[0] -= 37
            moved[-1] += 37
            ev.rebalance(moved)
            miou_b = ev.evaluate(data.val_dataloader(), data.S, ignore_index=data.ignore_index)
            same = ev.shard_counts == moved and ev.bank.rows == moved[rank] and bool((ev.last_confusion == conf).all()) \
                and miou_b == miou
            report[f"{name}_rebalanced"] = {"shard_rows": ev.shard_counts, "identical": bool(same)}
            ok = ok and same
        ev.close()
    # memory files with a sharded bank: rank 0 writes ONE (N, d) / (N, C) pair (hbird_eval.py:371-378);
    # reloading re-shards it by contiguous row ranges and must reproduce the evaluation
    tmp = f"/tmp/hb_dist_check_{os.environ.get('MASTER_PORT', '0')}_{name}"
    if rank == 0:
        os.makedirs(tmp, exist_ok=True)
    dist.barrier()
    fe = FeatureExtractorSimple(data.model, data.ftr_extr_fn, data.S, data.d)
    ev = HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=30, device=f"cuda:{local}",
                         nn_method="b200", nn_params={"idx_shard": True}, dataset_size=data.get_train_dataset_size(),
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
    same = bool((confs["p2p"] == confs["nccl"]).all()) and bool((confs["p2p"] == confs["replicas"]).all()) and \
        bool((confs["p2p"] == confs["p2p_full"]).all())
    report[f"{name}_paths_identical"] = same
    ok = ok and same
# augmentation epochs x a loader whose length does not divide by the world size: the batch counter runs
# on across epochs, so ranks take different numbers of batches; capacity must cover each of them and the
# doubled bank must still reproduce the single-epoch result (duplicate rows, ties by row id)
cfg, g = load_golden("ade_tiny")  # 3 training batches: with 2 ranks and 2 epochs rank 1 takes 3 of the 6
data = SyntheticSegmentationData(**cfg)
aug_conf = {}
for shard in (True, False):
    fe = FeatureExtractorSimple(data.model, data.ftr_extr_fn, data.S, data.d)
    ev = HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=30, augmentation_epoch=2,
                         device=f"cuda:{local}", nn_method="b200", nn_params={"idx_shard": shard},
                         dataset_size=data.get_train_dataset_size())
    miou = ev.evaluate(data.val_dataloader(), data.S, ignore_index=data.ignore_index)
    aug_conf[shard] = ev.last_confusion
    good = ev.total_rows == 2 * g["feature_memory"].shape[0] and 0.0 < miou <= 1.0 and \
        ev.last_confusion.sum() == g["conf"].sum()
    report[f"aug2_{'shards' if shard else 'replicas'}"] = {"loader_len": len(data.train_dataloader()), "rows": ev.shard_counts,
                                                          "miou": miou, "ok": bool(good)}
    ok = ok and good
    ev.close()
# duplicate rows tie on the score and resolve by row id, which both layouts number alike (rank-major)
report["aug2_layouts_identical"] = bool((aug_conf[True] == aug_conf[False]).all())
ok = ok and report["aug2_layouts_identical"]
flag = torch.tensor([1 if ok else 0], device="cpu" if ONE_GPU else "cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "one_gpu": ONE_GPU, "ok": bool(flag.item()), **report}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
