"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE ONLY — dev-container script (the GPU box has no /root/reference); the small
fixtures it writes are committed, so tests never need the reference at run time.

The reference imports two packages that are not installed here: `pytorch_lightning` (only so that
hbird.data imports; the data layer is never exercised because HbirdEvaluation takes loaders
directly) and `faiss`.  Both are shimmed in sys.modules before import; the faiss shim's
GpuIndexFlatIP.search is fp32 `q @ X.T` + topk(k, sorted) — the definition of an exact
inner-product flat index.  Everything else (_create_memory, _sample_features,
_find_nearest_key_to_query, _cross_attention, interpolate/argmax, PredsmIoU) is the reference's own
code, executed verbatim on the CPU.

    python oracle/make_golden.py                # all fixtures under tests/golden/:
                                                #   ref_{voc,ade}_tiny[_bounded].npz  whole-path runs of HbirdEvaluation
                                                #   ref_plugin_metrics.npz            NearestNeighborSearchFaiss, IP and L2
                                                #   ref_kats.json, ref_kats_matrix.json  PredsmIoU known answers
    python oracle/make_golden.py --plugin-only  # only ref_plugin_metrics.npz
    python oracle/make_golden.py --kats-only    # only ref_kats_matrix.json
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HBIRD_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def install_shims():
    pl = types.ModuleType("pytorch_lightning")

    class LightningDataModule:  # noqa: D401 - import-only stub
        def __init__(self, *a, **k):
            pass

    pl.LightningDataModule = LightningDataModule
    sys.modules["pytorch_lightning"] = pl

    faiss = types.ModuleType("faiss")
    faiss.SEARCH_LOG = []  # (D, I) of every search, for the fixtures

    class StandardGpuResources:
        pass

    class GpuIndexFlatConfig:
        useFloat16 = False
        device = 0

    class _Flat:
        def __init__(self, res, d, cfg):
            self.d, self.x = d, None

        def add(self, x):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
            self.x = x if self.x is None else torch.cat([self.x, x])

    class GpuIndexFlatIP(_Flat):
        def search(self, q, k):
            s = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)) @ self.x.T
            v, i = torch.topk(s, k, dim=1, largest=True, sorted=True)
            return v.numpy(), i.numpy().astype(np.int64)

    class GpuIndexFlatL2(_Flat):
        def search(self, q, k):
            qq = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
            dist = torch.cdist(qq, self.x) ** 2
            v, i = torch.topk(dist, k, dim=1, largest=False, sorted=True)
            return v.numpy(), i.numpy().astype(np.int64)

    class IndexReplicas:
        def __init__(self):
            self.subs = []

        def addIndex(self, s):
            self.subs.append(s)

        def add(self, x):
            for s in self.subs:
                s.add(x)

        def search(self, q, k):
            D, I = self.subs[0].search(q, k)
            faiss.SEARCH_LOG.append((D, I))
            return D, I

    class IndexShards:
        def __init__(self, d):
            self.subs, self.threaded, self.offsets = [], False, []

        def add_shard(self, s):
            self.subs.append(s)

        def add(self, x):
            n, g = x.shape[0], len(self.subs)
            self.offsets = [(n * r) // g for r in range(g + 1)]
            for r, s in enumerate(self.subs):
                s.add(x[self.offsets[r]:self.offsets[r + 1]])

        def search(self, q, k):
            Ds, Is = [], []
            for r, s in enumerate(self.subs):
                D, I = s.search(q, min(k, self.offsets[r + 1] - self.offsets[r]))
                Ds.append(D)
                Is.append(I + self.offsets[r])
            D, I = np.concatenate(Ds, 1), np.concatenate(Is, 1)
            order = np.argsort(-D, axis=1, kind="stable")[:, :k]
            D, I = np.take_along_axis(D, order, 1), np.take_along_axis(I, order, 1)
            faiss.SEARCH_LOG.append((D, I))
            return D, I

    faiss.get_num_gpus = lambda: 1
    faiss.StandardGpuResources = StandardGpuResources
    faiss.GpuIndexFlatConfig = GpuIndexFlatConfig
    faiss.GpuIndexFlatIP = GpuIndexFlatIP
    faiss.GpuIndexFlatL2 = GpuIndexFlatL2
    faiss.IndexReplicas = IndexReplicas
    faiss.IndexShards = IndexShards
    sys.modules["faiss"] = faiss
    return faiss


def run_reference(cfg: dict, faiss, memory_size=None, seed=None):
    sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
    from hbird_b200.data import SyntheticSegmentationData  # synthetic inputs only (pure torch)

    from hbird.hbird_eval import HbirdEvaluation
    from hbird.models import FeatureExtractorSimple
    from hbird.utils import eval_metrics as ref_metrics

    data = SyntheticSegmentationData(**cfg)
    fe = FeatureExtractorSimple(data.model, ftr_extr_fn=data.ftr_extr_fn, eval_spatial_resolution=data.S,
                                d_model=data.d)
    if seed is not None:
        torch.manual_seed(seed)
    faiss.SEARCH_LOG.clear()
    captured = {}
    orig_update = ref_metrics.PredsmIoU.update

    def spy_update(self, gt, pred):
        captured["gt"], captured["pred"] = gt.clone(), pred.clone()
        orig_update(self, gt, pred)
        captured["conf"] = self._conf_mat.clone()

    ref_metrics.PredsmIoU.update = spy_update
    try:
        ev = HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=cfg.get("k", 30),
                             augmentation_epoch=1, device="cpu", nn_method="faiss", nn_params={},
                             memory_size=memory_size, dataset_size=data.get_train_dataset_size())
        miou, det = ev.evaluate(data.val_dataloader(), eval_spatial_resolution=data.S, return_knn_details=True,
                                ignore_index=data.ignore_index)
    finally:
        ref_metrics.PredsmIoU.update = orig_update
    D = np.concatenate([d for d, _ in faiss.SEARCH_LOG])
    I = np.concatenate([i for _, i in faiss.SEARCH_LOG])
    return {
        "feature_memory": ev.feature_memory.numpy(), "label_memory": ev.label_memory.numpy(),
        "knn_idx": I, "knn_dist": D, "label_hat": det["knns_ca_labels"].numpy(),
        "pred": captured["pred"].numpy().astype(np.uint8), "gt": captured["gt"].numpy().astype(np.int16),
        "conf": captured["conf"].numpy(), "miou": np.float64(miou),
    }


CONFIGS = {
    # VOC-shaped miniature: S=8, ps=8 (64 px), d=64, C=6, ignore 255
    "voc_tiny": dict(num_train=6, num_val=3, input_size=64, patch_size=8, d_model=64, num_classes=6,
                     batch_size=4, ignore_index=255, cells=4, seed=0),
    # ADE-shaped miniature: ignore_index 0, odd patch size (ps=7 -> mean = hist/49), ragged last batch
    "ade_tiny": dict(num_train=5, num_val=3, input_size=56, patch_size=7, d_model=40, num_classes=9,
                     batch_size=2, ignore_index=0, cells=3, seed=5),
}


def kats():
    """Known-answer vectors for PredsmIoU, from the reference class itself."""
    from hbird.utils.eval_metrics import PredsmIoU

    cases = [
        dict(C=3, ignore=255, gt=[0, 0, 1, 1, 2, 2, 255, 1], pred=[0, 1, 1, 1, 2, 0, 2, 1]),
        dict(C=3, ignore=255, gt=[0, 0, 1, 1, 2, 2], pred=[1, 1, 2, 2, 0, 0]),
        dict(C=3, ignore=0, gt=[0, 0, 1, 1, 2, 2], pred=[1, 0, 1, 1, 2, 2]),
        dict(C=3, ignore=255, gt=[0, 0, 1, 1], pred=[0, 0, 1, 1]),
        dict(C=3, ignore=255, gt=[0, 1, 1], pred=[0, 7, 1]),
    ]
    out = []
    for c in cases:
        for mode in ("hungarian", "linear_probe", "many_to_one"):
            m = PredsmIoU(c["C"], c["C"], device=torch.device("cpu"), ignore_index=c["ignore"])
            m.update(torch.tensor(c["gt"]), torch.tensor(c["pred"]))
            miou, tp, fp, fn, reordered, bg = m.compute(True, many_to_one=mode == "many_to_one",
                                                        linear_probe=mode == "linear_probe")
            out.append(dict(c, mode=mode, conf=m._conf_mat.tolist(), miou=miou, tp=tp, fp=fp, fn=fn,
                            reordered=reordered, bg=bg))
    return out


def matrix_kats():
    """PredsmIoU of the reference on random pixel streams with rectangular class counts and every
    matching mode (incl. precision_based), stored as confusion matrix + results."""
    from hbird.utils.eval_metrics import PredsmIoU

    rng = np.random.default_rng(7)
    out = []
    for (P, G, n, ignore) in ((8, 5, 5000, 255), (5, 8, 5000, 255), (21, 21, 20000, 255), (12, 12, 3000, 0)):
        gt = rng.integers(0, G, size=n)
        # predictions correlated with gt through a random relabelling, plus noise and out-of-range ids
        relabel = rng.integers(0, P, size=G)
        pred = np.where(rng.random(n) < 0.7, relabel[gt], rng.integers(0, P + 1, size=n))
        if ignore == 255:
            gt = np.where(rng.random(n) < 0.03, 255, gt)
        for mode in ("hungarian", "many_to_one", "many_to_one_precision", "linear_probe"):
            m = PredsmIoU(P, G, device=torch.device("cpu"), ignore_index=ignore)
            m.update(torch.from_numpy(gt), torch.from_numpy(pred))
            miou, tp, fp, fn, _, bg = m.compute(True, many_to_one=mode.startswith("many_to_one"),
                                                precision_based=mode.endswith("precision"),
                                                linear_probe=mode == "linear_probe", return_reordered=False)
            out.append(dict(P=P, G=G, ignore=ignore, mode=mode, conf=m._conf_mat.tolist(), miou=miou,
                            tp=tp, fp=fp, fn=fn, bg=bg))
    return out


def plugin_fixture():
    """The reference plugin class itself (hbird/nn/search_faiss.py:6-90) on an UN-normalised bank,
    for both distance measures: pins the (indices, distances) order and the L2 convention
    (squared distances, ascending)."""
    from hbird.nn.search_faiss import NearestNeighborSearchFaiss

    g = torch.Generator().manual_seed(11)
    bank = torch.randn(2000, 48, generator=g) * (0.5 + torch.rand(2000, 1, generator=g))
    q = torch.randn(257, 48, generator=g) * 1.7
    out = {"bank": bank.numpy(), "q": q.numpy()}
    for name in ("dot_product", "l2"):
        nn = NearestNeighborSearchFaiss(bank, n_neighbors=30, distance_measure=name)
        idx, dist = nn.find_nearest_neighbors(q)
        out[f"idx_{name}"], out[f"dist_{name}"] = idx, dist
    return out


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference tree not found at {REF}: this script only runs in the dev container")
    faiss = install_shims()
    sys.path.insert(0, REF)
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order in the fixtures
    if "--kats-only" in sys.argv:
        with open(os.path.join(GOLD, "ref_kats_matrix.json"), "w") as f:
            json.dump(matrix_kats(), f)
        sys.exit(0)
    np.savez_compressed(os.path.join(GOLD, "ref_plugin_metrics.npz"), **plugin_fixture())
    if "--plugin-only" in sys.argv:
        sys.exit(0)
    with open(os.path.join(GOLD, "ref_kats_matrix.json"), "w") as f:
        json.dump(matrix_kats(), f)
    for name, cfg in CONFIGS.items():
        res = run_reference(cfg, faiss)
        np.savez_compressed(os.path.join(GOLD, f"ref_{name}.npz"), cfg=json.dumps(cfg), **res)
        print(name, "unbounded mIoU", float(res["miou"]), "bank", res["feature_memory"].shape)
        ms = cfg["num_train"] * 20
        resb = run_reference(cfg, faiss, memory_size=ms, seed=123)
        np.savez_compressed(os.path.join(GOLD, f"ref_{name}_bounded.npz"), cfg=json.dumps(cfg),
                            memory_size=ms, seed=123, **resb)
        print(name, "bounded mIoU", float(resb["miou"]), "bank", resb["feature_memory"].shape)
    with open(os.path.join(GOLD, "ref_kats.json"), "w") as f:
        json.dump(kats(), f, indent=1)
    print("wrote", sorted(os.listdir(GOLD)))
