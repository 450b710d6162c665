"""Pins oracle/hbird_oracle.py (the numpy restatement) against outputs of the UNMODIFIED reference
(tests/golden/ref_*.npz, written by oracle/make_golden.py) and against the PredsmIoU known-answer
vectors.  CPU only."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, batches_np, load_golden, recall
from hbird_b200.data import SyntheticSegmentationData
from oracle import hbird_oracle as O

CASES = ["voc_tiny", "ade_tiny"]


@pytest.fixture(scope="module", params=CASES)
def case(request):
    cfg, g = load_golden(request.param)
    data = SyntheticSegmentationData(**cfg)
    return cfg, g, data


def test_decode_mask_exact_for_all_byte_values():
    ids = np.arange(256, dtype=np.float32)
    y = (ids / np.float32(255)).astype(np.float32)
    assert np.array_equal(O.decode_mask(y, False), np.arange(256))
    r = O.decode_mask(y, True)
    assert r[255] == 0 and np.array_equal(r[:255], np.arange(255))


def test_bank_matches_reference(case):
    cfg, g, data = case
    fm, lm = O.build_memory(batches_np(data, data.train_dataloader()), data.C, data.S)
    assert fm.shape == g["feature_memory"].shape
    np.testing.assert_allclose(fm, g["feature_memory"], rtol=0, atol=2e-7)
    np.testing.assert_array_equal(lm, g["label_memory"])  # counts / ps^2 is exact in fp32


def test_search_matches_reference(case):
    cfg, g, data = case
    q = np.concatenate([f.reshape(-1, f.shape[-1]) for f, _ in batches_np(data, data.val_dataloader())])
    idx, dist = O.search_exact_ip(q, g["feature_memory"], 30)
    np.testing.assert_allclose(dist, g["knn_dist"], rtol=2e-6, atol=1e-6)
    assert recall(idx, g["knn_idx"]) >= 0.9995
    assert (np.diff(dist, axis=1) <= 0).all()


def test_label_transfer_upsample_argmax_match_reference(case):
    cfg, g, data = case
    vb = batches_np(data, data.val_dataloader())
    label_hats, preds, off = [], [], 0
    for f, y in vb:
        B, N, _ = f.shape
        idx = g["knn_idx"][off:off + B * N]
        off += B * N
        lh = O.transfer_labels(f, g["feature_memory"], g["label_memory"], idx)
        label_hats.append(lh)
        preds.append(O.predict_map(lh, data.S, y.shape[-2], y.shape[-1]))
    lh = np.concatenate(label_hats)
    np.testing.assert_allclose(lh, g["label_hat"], rtol=0, atol=5e-6)
    pred = np.concatenate(preds)
    gt = np.concatenate([O.decode_mask(y, False) for _, y in vb])
    np.testing.assert_array_equal(gt, g["gt"])            # loader-contract decode, eval side
    assert pred.shape == g["pred"].shape                  # (n_img, 1, H, W), as metric.update gets it
    assert (pred == g["pred"]).mean() >= 0.9999


def test_confusion_and_miou_match_reference(case):
    cfg, g, data = case
    conf = O.confusion_matrix(g["gt"], g["pred"], data.C, data.C, data.ignore_index)
    np.testing.assert_array_equal(conf, g["conf"])
    miou, tp, fp, fn, bg = O.miou_from_confusion(conf)
    assert miou == pytest.approx(float(g["miou"]), abs=1e-12)


def test_end_to_end_oracle_matches_reference(case):
    cfg, g, data = case
    fm, lm = O.build_memory(batches_np(data, data.train_dataloader()), data.C, data.S)
    miou, conf = O.evaluate(fm, lm, batches_np(data, data.val_dataloader()), data.C, data.S, 30,
                            data.ignore_index)
    assert abs(miou - float(g["miou"])) <= 5e-4  # 0.05 points
    assert np.abs(conf - g["conf"]).sum() <= 2e-4 * g["conf"].sum()


@pytest.mark.parametrize("name", CASES)
def test_bounded_sampler_matches_reference(name):
    """The reference's bounded bank (memory_size, torch.manual_seed(123)) is reproduced by the
    oracle sampler fed the same CPU uniform stream (hbird_eval.py:497-508)."""
    import torch

    cfg, g = load_golden(name + "_bounded")
    data = SyntheticSegmentationData(**cfg)
    K = max(1, int(np.load(os.path.join(GOLDEN, f"ref_{name}_bounded.npz"))["memory_size"]) // data.get_train_dataset_size())
    torch.manual_seed(123)
    rows_f, rows_l = [], []
    for f, y in batches_np(data, data.train_dataloader()):
        ids = O.decode_mask(y, True)
        pg = O.patchify_gt(ids, data.ps)
        B, S0, S1, _ = pg.shape
        u = torch.rand(B * S0 * S1).numpy()  # every patch is non-empty once 255 -> 0
        sel = O.sample_patches(pg, data.C, K, u)
        lab = O.soft_labels(pg, data.C).reshape(B, S0 * S1, -1)
        nf = O.normalise_rows(np.take_along_axis(f, sel[:, :, None], axis=1))
        rows_f.append(nf.reshape(-1, f.shape[-1]))
        rows_l.append(np.take_along_axis(lab, sel[:, :, None], axis=1).reshape(-1, data.C))
    # same set of sampled rows per image (topk order among equal scores may differ)
    fm, lm = np.concatenate(rows_f), np.concatenate(rows_l)
    assert fm.shape == g["feature_memory"].shape
    for b in range(data.get_train_dataset_size()):
        mine, ref = fm[b * K:(b + 1) * K], g["feature_memory"][b * K:(b + 1) * K]
        dist = np.abs(mine[:, None, :] - ref[None, :, :]).max(axis=2)  # (K, K) row distances
        assert (dist.min(axis=1) <= 1e-6).all() and (dist.min(axis=0) <= 1e-6).all()


def test_predsmiou_known_answers():
    kats = json.load(open(os.path.join(GOLDEN, "ref_kats.json")))
    assert len(kats) == 15
    for k in kats:
        conf = O.confusion_matrix(np.array(k["gt"]), np.array(k["pred"]), k["C"], k["C"], k["ignore"])
        assert conf.tolist() == k["conf"]
        miou, tp, fp, fn, bg = O.miou_from_confusion(conf, linear_probe=k["mode"] == "linear_probe",
                                                     many_to_one=k["mode"] == "many_to_one")
        assert miou == pytest.approx(k["miou"], abs=1e-12)
        assert (tp, fp, fn) == (k["tp"], k["fp"], k["fn"])
        assert bg == pytest.approx(k["bg"])
    # the survey's headline KAT: mIoU 0.52777..
    assert kats[0]["miou"] == pytest.approx(0.5277777777777778)


@pytest.mark.parametrize("measure", ["dot_product", "l2"])
def test_search_oracle_matches_reference_plugin_fixture(measure):
    """ref_plugin_metrics.npz = the reference's NearestNeighborSearchFaiss (search_faiss.py:6-90)
    run on an un-normalised bank by oracle/make_golden.py."""
    z = np.load(os.path.join(GOLDEN, "ref_plugin_metrics.npz"))
    fn = O.search_exact_ip if measure == "dot_product" else O.search_exact_l2
    idx, dist = fn(z["q"], z["bank"], 30)
    ri, rd = z[f"idx_{measure}"], z[f"dist_{measure}"]
    assert (idx == ri).mean() >= 0.999  # fp32 near-ties may swap neighbours
    np.testing.assert_allclose(dist, rd, rtol=2e-5, atol=2e-5)


def test_merge_shards_equals_unsharded():
    rng = np.random.default_rng(0)
    bank = O.normalise_rows(rng.standard_normal((500, 32)).astype(np.float32))
    q = rng.standard_normal((40, 32)).astype(np.float32) * 3
    idx, dist = O.search_exact_ip(q, bank, 30)
    parts_i, parts_d = [], []
    for r in range(4):
        a, b = 500 * r // 4, 500 * (r + 1) // 4
        i, d = O.search_exact_ip(q, bank[a:b], 30)
        parts_i.append(i + a)
        parts_d.append(d)
    mi, md = O.merge_shards(np.stack(parts_i), np.stack(parts_d), 30)
    np.testing.assert_array_equal(mi, idx)
    np.testing.assert_array_equal(md, dist)


def test_search_pads_when_bank_smaller_than_k():
    bank = O.normalise_rows(np.eye(4, 8, dtype=np.float32))
    idx, dist = O.search_exact_ip(np.ones((2, 8), np.float32), bank, 6)
    assert (idx[:, 4:] == -1).all() and np.isneginf(dist[:, 4:]).all()
