"""The plain-C oracle (oracle/hbird_oracle.c, built with gcc) against the golden fixtures of the
unmodified reference and against the numpy oracle.  Integer / byte work must agree bit for bit;
the scalar fp32 search and label transfer within summation-order tolerance.  CPU only."""
import numpy as np
import pytest

from helpers import batches_np, load_golden, recall
from hbird_b200.data import SyntheticSegmentationData
from oracle import c_oracle as C
from oracle import hbird_oracle as O

CASES = ["voc_tiny", "ade_tiny"]


@pytest.fixture(params=CASES)
def case(request):
    cfg, g = load_golden(request.param)
    return cfg, g, SyntheticSegmentationData(**cfg)


def test_decode_mask_every_byte_value():
    y = (np.arange(256, dtype=np.float32) / np.float32(255)).astype(np.float32)
    for remap in (False, True):
        np.testing.assert_array_equal(C.decode_mask(y, remap), O.decode_mask(y, remap).astype(np.uint8))


def test_soft_labels_and_eval_masks_match_reference(case):
    cfg, g, data = case
    hists = []
    for _, y in batches_np(data, data.train_dataloader()):
        mask = C.decode_mask(y, True).reshape(y.shape[0], y.shape[-2], y.shape[-1])
        hists.append(C.patch_histogram(mask, data.S, data.ps, data.C))
    hist = np.concatenate(hists)
    # the reference's label_memory is one_hot(...).mean(3) = counts / ps^2, exact in fp32
    np.testing.assert_array_equal(hist.astype(np.float32) / np.float32(data.ps * data.ps), g["label_memory"])
    gt = np.concatenate([C.decode_mask(y, False) for _, y in batches_np(data, data.val_dataloader())])
    np.testing.assert_array_equal(gt, g["gt"])


def test_confusion_matches_reference_and_numpy_oracle(case):
    cfg, g, data = case
    conf = C.confusion(g["gt"], g["pred"], data.C, data.C, data.ignore_index)
    np.testing.assert_array_equal(conf, g["conf"])
    rng = np.random.default_rng(3)
    gt = rng.integers(0, 256, size=100_003).astype(np.uint8)
    pred = rng.integers(0, 12, size=100_003).astype(np.uint8)
    for G, P, ign in ((9, 7, 255), (7, 9, 0), (12, 12, None)):
        np.testing.assert_array_equal(C.confusion(gt, pred, G, P, ign), O.confusion_matrix(gt, pred, G, P, ign))


def test_upsample_argmax_matches_reference_bit_for_bit(case):
    """Fed the reference's own label_hat, the C bilinear + argmax reproduces the reference's
    prediction maps exactly (same fp32 operation order as ATen's upsample_bilinear2d)."""
    cfg, g, data = case
    H = data.S * data.ps
    n_img = g["pred"].shape[0]
    pred = C.upsample_argmax(g["label_hat"].reshape(n_img, data.S * data.S, data.C), n_img, data.S, H, H)
    np.testing.assert_array_equal(pred, g["pred"].reshape(n_img, H, H))
    np.testing.assert_array_equal(pred.reshape(n_img, 1, H, H), O.predict_map(g["label_hat"].reshape(n_img, data.S * data.S, data.C), data.S, H, H))


def test_search_and_label_transfer_match_reference(case):
    cfg, g, data = case
    q = np.concatenate([f.reshape(-1, f.shape[-1]) for f, _ in batches_np(data, data.val_dataloader())])
    idx, dist = C.search_ip(q, g["feature_memory"], 30)
    np.testing.assert_allclose(dist, g["knn_dist"], rtol=5e-6, atol=2e-6)
    assert recall(idx, g["knn_idx"]) >= 0.9995 and (np.diff(dist, axis=1) <= 0).all()
    lh = C.label_transfer(q, g["feature_memory"], g["label_memory"], g["knn_idx"])
    np.testing.assert_allclose(lh, g["label_hat"].reshape(lh.shape), rtol=0, atol=2e-5)


def test_search_pads_and_orders_ties_like_faiss():
    bank = np.eye(4, 8, dtype=np.float32)
    bank = np.concatenate([bank, bank[:1]])  # row 4 duplicates row 0: a tie
    idx, dist = C.search_ip(np.ones((1, 8), np.float32) * np.array([[1, 0, 0, 0, 0, 0, 0, 0]], np.float32), bank, 7)
    assert idx[0, 0] == 0 and idx[0, 1] == 4  # equal scores: the smaller index first
    assert (idx[0, 5:] == -1).all() and np.isneginf(dist[0, 5:]).all()
    oi, od = O.search_exact_ip(np.array([[1, 0, 0, 0, 0, 0, 0, 0]], np.float32), bank, 7)
    np.testing.assert_array_equal(idx, oi)
    np.testing.assert_array_equal(dist, od)


@pytest.mark.parametrize("S,ps,n_cls", [(14, 16, 21), (37, 14, 21), (5, 7, 151), (3, 10, 2), (1, 8, 4)])
def test_upsample_argmax_c_equals_numpy_on_random_label_maps(S, ps, n_cls):
    """The two restatements of ATen's align_corners=False bilinear + argmax agree pixel for pixel at the
    BASELINE geometries (S=14/ps=16, S=37/ps=14) and at degenerate ones (S=1)."""
    rng = np.random.default_rng(S * 100 + ps)
    B, H = 2, S * ps
    lh = rng.random((B, S * S, n_cls)).astype(np.float32)
    lh[0, :, 0] = lh[0, :, 1]  # exact ties between two classes: the first maximum must win in both
    got = C.upsample_argmax(lh, B, S, H, H)
    ref = O.predict_map(lh, S, H, H).reshape(B, H, H)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("S,ps,n_cls", [(14, 16, 21), (37, 14, 21), (8, 7, 9)])
def test_upsample_argmax_oracles_equal_aten_interpolate(S, ps, n_cls):
    """Ground truth for A9 is ATen itself (hbird_eval.py:235-243): permute -> F.interpolate(bilinear,
    align_corners=False) -> argmax, run here on the CPU.  Both oracles must reproduce it exactly."""
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(7 * S + ps)
    B, H = 2, S * ps
    lh = rng.random((B, S * S, n_cls)).astype(np.float32)
    x = torch.from_numpy(lh).view(B, S, S, n_cls).permute(0, 3, 1, 2)
    aten = F.interpolate(x, size=(H, H), mode="bilinear").argmax(dim=1).numpy().astype(np.uint8)
    np.testing.assert_array_equal(C.upsample_argmax(lh, B, S, H, H), aten)
    np.testing.assert_array_equal(O.predict_map(lh, S, H, H).reshape(B, H, H), aten)


def test_decode_and_soft_labels_equal_the_torch_ops_the_reference_runs():
    """(y*255).long() with y = id/255 as the loader delivers it (image_transformations.py:39-49) and
    one_hot(...).float().mean(3) over patch pixels (hbird_eval.py:319-320), computed with torch on the
    CPU, against both oracles."""
    import torch
    import torch.nn.functional as F

    ids = torch.arange(256, dtype=torch.float32)
    y = ids / 255
    want = (y * 255).long().numpy()
    np.testing.assert_array_equal(C.decode_mask(y.numpy(), False), want.astype(np.uint8))
    np.testing.assert_array_equal(O.decode_mask(y.numpy(), False), want)
    rng = np.random.default_rng(5)
    B, S, ps, n_cls = 2, 5, 7, 11
    mask = rng.integers(0, n_cls, size=(B, S * ps, S * ps)).astype(np.uint8)
    pg = torch.from_numpy(mask.astype(np.int64)).view(B, S, ps, S, ps).permute(0, 1, 3, 2, 4).reshape(B, S, S, ps * ps)
    soft = F.one_hot(pg, n_cls).float().mean(3).reshape(B * S * S, n_cls).numpy()
    hist = C.patch_histogram(mask, S, ps, n_cls)
    np.testing.assert_array_equal(hist.astype(np.float32) / np.float32(ps * ps), soft)
    np.testing.assert_array_equal(O.soft_labels(pg.numpy(), n_cls).reshape(B * S * S, n_cls), soft)
