"""Host-side logic that needs no GPU: metric math vs the reference's known answers, the registry,
the bounded sampler's RNG parity with the reference, the search work decomposition, the synthetic
data contract."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_golden
from hbird_b200 import _capi
from hbird_b200.data import SyntheticSegmentationData
from hbird_b200.registry import (NN_BACKENDS, create_nn_backend, nn_method_choices, parse_nn_params,
                                register_nn_backend)
from hbird_b200.utils.eval_metrics import miou_from_confusion
from oracle import hbird_oracle as O


def test_metric_math_reproduces_reference_known_answers():
    for k in json.load(open(os.path.join(GOLDEN, "ref_kats.json"))):
        conf = np.array(k["conf"], dtype=np.int64)
        miou, tp, fp, fn, mapping, bg = miou_from_confusion(
            conf, many_to_one=k["mode"] == "many_to_one", linear_probe=k["mode"] == "linear_probe")
        assert miou == pytest.approx(k["miou"], abs=1e-12), k
        assert (tp, fp, fn) == (k["tp"], k["fp"], k["fn"]), k
        assert bg == pytest.approx(k["bg"])


def test_metric_math_matches_oracle_on_random_matrices():
    rng = np.random.default_rng(1)
    for C in (3, 21, 151):
        conf = rng.integers(0, 1000, size=(C, C)).astype(np.int64)
        conf[rng.integers(0, C)] = 0  # an absent class still counts in the mean
        a = miou_from_confusion(conf)
        b = O.miou_from_confusion(conf)
        assert a[0] == pytest.approx(b[0], abs=1e-12) and a[1:4] == b[1:4]


def test_registry_dispatch_and_errors():
    assert {"b200", "faiss", "scann"} <= set(NN_BACKENDS)
    with pytest.raises(ValueError, match="Unsupported NN method"):
        create_nn_backend("annoy", torch.zeros(2, 8))
    with pytest.raises(ValueError, match="reference package"):
        create_nn_backend("faiss", torch.zeros(2, 8))  # legacy names defer to the reference classes
    register_nn_backend("dummy", lambda fm, n_neighbors=30, **kw: ("dummy", n_neighbors, kw))
    assert create_nn_backend("dummy", None, n_neighbors=7, a=1) == ("dummy", 7, {"a": 1})
    NN_BACKENDS.pop("dummy")


def _plan(rows, Q, cg, sms=148, max_chunks=0):
    out = (ctypes.c_int * 4)()
    assert _capi.lib.hb_plan_search(rows, Q, cg, sms, max_chunks, out) == 0
    return list(out)


@pytest.mark.parametrize("rows,Q", [(1, 1), (255, 128), (257, 129), (102400, 12544), (1024000, 12544),
                                    (10240000, 21904), (12500000, 65536), (1000003, 7)])
@pytest.mark.parametrize("cg", [1, 2])
def test_search_plan_covers_every_tile_once(rows, Q, cg):
    n_tiles, n_qblocks, n_chunks, n_units = _plan(rows, Q, cg)
    assert n_tiles == (rows + 255) // 256 and n_qblocks == -(-Q // (128 * cg)) and n_units == 148 // cg
    assert 1 <= n_chunks <= min(64, n_tiles)
    bounds = [n_tiles * c // n_chunks for c in range(n_chunks + 1)]
    assert bounds[0] == 0 and bounds[-1] == n_tiles
    sizes = np.diff(bounds)
    assert (sizes >= 1).all() and sizes.max() - sizes.min() <= 1  # balanced, no empty chunk


def test_search_plan_fills_waves_for_headline_configs():
    for rows, Q in [(1024000, 12544), (10240000, 21904), (1280000, 21904)]:
        n_tiles, n_qblocks, n_chunks, n_units = _plan(rows, Q, 2)
        items = n_qblocks * n_chunks
        waves = -(-items // n_units)
        assert items / (waves * n_units) >= 0.95
    assert _plan(1024000, 12544, 2, max_chunks=1)[2] == 1


def test_synthetic_data_contract():
    d = SyntheticSegmentationData(num_train=3, num_val=2, input_size=32, patch_size=8, d_model=16, num_classes=4,
                                  batch_size=2, ignore_index=255, cells=2)
    (x, y), = d.val_dataloader()
    assert x.shape == (2, 3, 32, 32) and y.shape == (2, 1, 32, 32) and y.dtype == torch.float32
    ids = (y * 255).long()
    assert set(ids.unique().tolist()) <= {0, 1, 2, 3, 255}
    f, aux = d.ftr_extr_fn(d.model, x)
    assert f.shape == (2, 16, 16) and aux is None
    assert (f.norm(dim=-1) > 1.5).all()  # queries are NOT unit norm (hbird_eval.py:222-224)


def test_parse_nn_params_follows_reference_cli_coercion():
    """Expected dict = output of the reference's _parse_nn_params (eval.py:444-462) on these items."""
    items = ["k_prime=64", "keep_f32=False", "idx_shard=TRUE", "beta=0.02", "distance_measure=l2",
             " leaves = 200 ", "x=1e3", "name=a=b", "neg=-5"]
    assert parse_nn_params(items) == {
        "k_prime": 64, "keep_f32": False, "idx_shard": True, "beta": 0.02, "distance_measure": "l2",
        "leaves": 200, "x": 1000.0, "name": "a=b", "neg": -5}
    assert parse_nn_params(None) == {} and parse_nn_params([]) == {}
    with pytest.raises(ValueError, match="KEY=VALUE"):
        parse_nn_params(["novalue"])
    assert {"b200", "faiss", "scann"} <= set(nn_method_choices())


def test_metric_math_rectangular_and_precision_modes_match_reference():
    """tests/golden/ref_kats_matrix.json: the reference's PredsmIoU on random pixel streams with
    num_pred != num_gt, ignore_index 255 / 0 and all four matching modes (eval_metrics.py:112-288).
    Both the product's host math and the oracle must reproduce it."""
    kats = json.load(open(os.path.join(GOLDEN, "ref_kats_matrix.json")))
    assert len(kats) == 16
    for k in kats:
        conf = np.array(k["conf"], dtype=np.int64)
        assert conf.shape == (k["G"], k["P"])
        kw = dict(many_to_one=k["mode"].startswith("many_to_one"), precision_based=k["mode"].endswith("precision"),
                  linear_probe=k["mode"] == "linear_probe")
        miou, tp, fp, fn, _, bg = miou_from_confusion(conf, **kw)
        assert miou == pytest.approx(k["miou"], abs=1e-12), (k["P"], k["G"], k["mode"])
        assert (tp, fp, fn) == (k["tp"], k["fp"], k["fn"]) and bg == pytest.approx(k["bg"])
        omiou, otp, ofp, ofn, obg = O.miou_from_confusion(conf, **kw)
        assert omiou == pytest.approx(k["miou"], abs=1e-12) and (otp, ofp, ofn) == (k["tp"], k["fp"], k["fn"])
        assert obg == pytest.approx(k["bg"])


@pytest.mark.parametrize("L,W,aug,bounded", [(3, 2, 2, False), (3, 2, 2, True), (5, 4, 3, False), (2, 2, 1, False), (7, 8, 1, True), (1, 1, 4, False)])
def test_per_rank_bank_capacity_covers_the_batches_a_rank_takes(L, W, aug, bounded):
    """_create_memory deals batch number `step` to rank step % W and `step` runs on across augmentation
    epochs (hbird_b200/hbird_eval.py); every rank's capacity must cover what it is dealt — and, in
    bounded mode, no rank reserves the whole memory_size."""
    from types import SimpleNamespace

    from hbird_b200.hbird_eval import HbirdEvaluation

    B, S, K = 4, 7, 5
    per_image = K if bounded else S * S
    total = aug * L * B * per_image
    for rank in range(W):
        me = SimpleNamespace(memory_size=(total if bounded else None), num_sampled_features=K, augmentation_epoch=aug,
                             rank=rank, world=W)
        cap = HbirdEvaluation._capacity_rows(me, L, B, S)
        dealt = sum(1 for step in range(aug * L) if step % W == rank)
        assert cap >= max(1, dealt * B * per_image)
        assert cap <= max(1, (dealt * B * per_image)) or W == 1
        if bounded and W > 1 and dealt:
            assert cap < total


def test_balanced_shard_counts_equalise_the_search_time():
    """distributed.balanced_counts: shard sizes proportional to measured speed, total preserved, shifts
    bounded; degenerate inputs leave the shards alone."""
    from hbird_b200.distributed import balanced_counts

    counts = [1_280_000] * 8
    times = [30.0, 30.5, 31.0, 29.5, 30.2, 30.1, 31.4, 29.9]
    new = balanced_counts(counts, times)
    assert sum(new) == sum(counts) and all(abs(n - c) <= 0.1 * c + 1 for n, c in zip(new, counts))
    predicted = [t * n / c for t, n, c in zip(times, new, counts)]  # search time is linear in the rows
    assert max(predicted) - min(predicted) < 0.02 * max(predicted)
    assert max(predicted) < max(times)
    assert balanced_counts([100, 100], [1.0, 5.0]) == [110, 90]          # clamped to +-10 %
    assert balanced_counts([7], [3.0]) == [7] and balanced_counts([5, 5], [0.0, 1.0]) == [5, 5]
    assert balanced_counts([3, 1_000_000], [1.0, 1.0]) == [3, 1_000_000]  # equal speed per row... stays within bounds


@pytest.mark.parametrize("G,kp", [(2, 64), (4, 64), (8, 64), (8, 32), (3, 128), (16, 64)])
def test_threshold_exchange_bound_never_prunes_the_global_top_kprime(G, kp):
    """The rule behind hb_exchange_config mode 1, on the oracle's restatement of it: the bound is never
    above the kp-th best score of the union of the shards' candidate lists, so re-ranking only the
    candidates at or above it keeps a superset of the global top-kp; and it is tight enough to matter
    (far fewer than G*kp survivors) whether the neighbours are spread evenly or sit in one shard."""
    rng = np.random.default_rng(100 * G + kp)
    Q, per = 200, 3000
    scores = rng.standard_normal((G, Q, per)).astype(np.float32)
    scores[0, :50, :2 * kp] += 6.0    # 50 queries whose best 2*kp rows all live in shard 0
    scores[1, 50:60, kp // 2:] = -np.inf   # a shard with fewer than kp candidates for some queries
    top = -np.sort(-scores, axis=2)[:, :, :kp]
    bound = O.shortlist_bound(top, kp)
    union = -np.sort(-top.transpose(1, 0, 2).reshape(Q, -1), axis=1)
    kth = union[:, kp - 1]
    assert (bound <= kth).all()
    kept = (top >= bound[None, :, None]).sum(axis=(0, 2))
    assert (kept >= kp).all()
    assert kept[:50].mean() <= 1.5 * kp            # concentrated: about one shard's list
    if G >= 4:
        assert kept[60:].mean() <= 0.45 * G * kp   # spread out: a fraction of the G lists
