"""Parity of the fused kernels added on top of the per-step ones: K2b+K4a (label transfer inside the
re-rank / merge warp), the fused tail (mask decode + upsample + argmax + confusion in one pass),
the one-call validation step, and the candidate bound of the tcgen05 pass (threshold board, hashed
column sets).  Everything goes through the C-ABI and is checked against the CPU oracle, the
reference's golden fixtures, and the unfused kernels (which are themselves pinned to both)."""
import numpy as np
import pytest
import torch

from helpers import batches_np, load_golden, recall
from hbird_b200 import ops
from hbird_b200.data import SyntheticSegmentationData
from oracle import hbird_oracle as O
from test_gpu_parity import bank_from_rows, build_bank_from_loader, cuda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["voc_tiny", "ade_tiny"]


@pytest.fixture(scope="module", params=CASES)
def case(request):
    cfg, g = load_golden(request.param)
    return cfg, g, SyntheticSegmentationData(**cfg)


# ------------------------------------------------------------------ K2b + K4a
def test_search_transfer_equals_search_then_label_transfer(case):
    cfg, g, data = case
    bank = build_bank_from_loader(data)
    q = cuda(np.concatenate([f.reshape(-1, f.shape[-1]) for f, _ in batches_np(data, data.val_dataloader())]))
    s, i, qn = bank.search(q, 30, 64)
    lh_ref = ops.label_transfer(bank.label_table(), data.ps * data.ps, s, i, qn, 0.02)
    lh, qn2, s2, i2 = bank.search_transfer(q, 30, 64, return_neighbours=True)
    assert torch.equal(s2, s) and torch.equal(i2, i) and torch.equal(qn2, qn)
    assert torch.equal(lh, lh_ref)  # same arithmetic in the same order
    lh3, _, s3, i3 = bank.search_transfer(q, 30, 64, label_table=bank.label_table())  # explicit table, no neighbour output
    assert s3 is None and i3 is None and torch.equal(lh3, lh_ref)
    ref = g["label_hat"].reshape(-1, data.C)
    np.testing.assert_allclose(lh.cpu().numpy(), ref, rtol=0, atol=1e-5)
    bank.close()


def test_search_transfer_rejects_l2_banks_and_bad_arguments():
    rows = torch.randn((3000, 64), device=DEV)
    bank = ops.MemoryBank(64, 1, 1, 3000, 0, True, metric="l2")
    bank.append_soft(rows, torch.ones((3000, 1), device=DEV), normalise=False)
    bank.finalize()
    with pytest.raises(ValueError, match="inner-product"):
        bank.search_transfer(torch.zeros((4, 64), device=DEV))
    bank.close()
    bank = bank_from_rows(rows)
    with pytest.raises(ValueError):
        bank.search_transfer(torch.zeros((4, 32), device=DEV))  # wrong d
    with pytest.raises(ValueError):
        bank.search_transfer(torch.zeros((4, 64), device=DEV), 30, 64, beta=0.0)
    lh, qn, _, _ = bank.search_transfer(torch.zeros((0, 64), device=DEV))
    assert lh.shape == (0, 1)
    bank.close()


def test_merge_transfer_paths_equal_merge_then_label_transfer():
    """K3 and K3x with the fused label transfer == merge, then the stand-alone K4a; 4 simulated
    ranks on one GPU (hb_exchange_connect_local), ragged slices, one of them empty."""
    g = torch.Generator().manual_seed(19)
    N, d, G, k, C, pp = 21013, 64, 4, 30, 7, 16
    rows = torch.randn((N, d), generator=g)
    soft = torch.zeros((N, C))
    soft[torch.arange(N), torch.randint(0, C, (N,), generator=g)] = 0.75
    soft[torch.arange(N), torch.randint(0, C, (N,), generator=g)] += 0.25
    bounds = [N * r // G for r in range(G + 1)]
    shards = []
    for r in range(G):
        b = ops.MemoryBank(d, C, pp, bounds[r + 1] - bounds[r], 0, True)
        b.append_soft(rows[bounds[r]:bounds[r + 1]].to(DEV), soft[bounds[r]:bounds[r + 1]].to(DEV), normalise=True)
        b.finalize()
        shards.append(b)
    table = torch.cat([b.label_table() for b in shards]).contiguous()
    xs = [ops.ShardExchange(r, G, 200, k, 0) for r in range(G)]
    ops.ShardExchange.connect_local(xs)
    qsplit = [0, 90, 90, 231, 300]
    q = (torch.randn((300, d), generator=g) * 2).to(DEV)
    ss, si = [], []
    for r in range(G):
        s, i, qn = shards[r].search(q, k, 64, idx_offset=bounds[r])
        ss.append(s), si.append(i)
    ms, mi = ops.merge_topk(torch.stack(ss), torch.stack(si))
    lh_ref = ops.label_transfer(table, pp, ms, mi, qn, 0.02)
    lh, fs, fi = ops.merge_topk_transfer(torch.stack(ss), torch.stack(si), table, pp, qn, 0.02, return_neighbours=True)
    assert torch.equal(lh, lh_ref) and torch.equal(fs, ms) and torch.equal(fi, mi)
    lh2, n1, n2 = ops.merge_topk_transfer(torch.stack(ss), torch.stack(si), table, pp, qn, 0.02)
    assert n1 is None and n2 is None and torch.equal(lh2, lh_ref)
    for r in range(G):
        xs[r].search_scatter(shards[r], q, qsplit, k, 64, idx_offset=bounds[r])
    for r in range(G):
        a, b = qsplit[r], qsplit[r + 1]
        lhr, fs, fi = xs[r].merge_transfer(table, pp, qn[a:b].contiguous(), 0.02, return_neighbours=True)
        assert lhr.shape == (b - a, C)
        assert torch.equal(lhr, lh_ref[a:b]) and torch.equal(fs, ms[a:b]) and torch.equal(fi, mi[a:b])
    # oracle: soft labels through the reference's cross-attention on the merged neighbours
    fb = torch.cat([b.export()[0] for b in shards]).cpu().numpy()
    lm = torch.cat([b.export()[1] for b in shards]).cpu().numpy()
    ref = O.transfer_labels(q.cpu().numpy()[None], fb, lm, mi.cpu().numpy(), 0.02)[0]
    np.testing.assert_allclose(lh.cpu().numpy(), ref, rtol=0, atol=2e-5)
    with pytest.raises(ValueError, match="row-sharded"):
        l2 = ops.MemoryBank(d, 1, 1, 100, 0, True, metric="l2")
        l2.append_soft(rows[:100].to(DEV), torch.ones((100, 1), device=DEV))
        l2.finalize()
        xs[0].search_scatter(l2, q, qsplit, k, 64)
    for o in xs + shards:
        o.close()


# ------------------------------------------------------------------ fused tail
@pytest.mark.parametrize("B,S,C,H,ignore", [(2, 14, 21, 224, 255), (1, 37, 151, 518, 0), (3, 5, 2, 33, None),
                                            (5, 16, 24, 256, 255), (2, 37, 21, 518, 255), (1, 9, 33, 100, 255)])
def test_predict_score_equals_unfused_kernels_and_oracle(B, S, C, H, ignore):
    rng = np.random.default_rng(S * C)
    lh = rng.random((B, S * S, C)).astype(np.float32)
    lh[rng.random(lh.shape) < 0.6] = 0.0  # soft labels are sparse; exact ties at 0 exercise first-max
    ids = np.repeat(np.repeat(rng.integers(0, min(C + 2, 256), size=(B, H // 4 + 1, H // 4 + 1)), 4, 1), 4, 2)[:, :H, :H]
    ids = np.where(rng.random(ids.shape) < 0.03, 255, ids).astype(np.uint8)
    y = (ids.astype(np.float32) / np.float32(255)).astype(np.float32)[:, None]
    t_lh, t_y = cuda(lh.reshape(-1, C)), cuda(y)
    pred_ref = ops.upsample_argmax(t_lh, B, S, H, H)
    gt = ops.decode_mask(t_y, False).view(B, H, H)
    conf_ref = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    ops.confusion_accumulate(conf_ref, gt, pred_ref, ignore)
    # y in, fused decode
    conf = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    pred = ops.predict_score(t_lh, B, S, H, H, conf, y=t_y, ignore_index=ignore, return_pred=True)
    assert torch.equal(pred, pred_ref) and torch.equal(conf, conf_ref)
    # decoded gt in, no prediction map out, accumulates on top
    assert ops.predict_score(t_lh, B, S, H, H, conf, gt_u8=gt, ignore_index=ignore) is None
    assert torch.equal(conf, 2 * conf_ref)
    # prediction only
    assert torch.equal(ops.predict_score(t_lh, B, S, H, H), pred_ref)
    # oracle: torch-exact upsample + argmax (ATen-pinned), bincount on identical predictions
    ref_pred = O.predict_map(lh, S, H, H)[:, 0]
    assert (pred.cpu().numpy() == ref_pred).mean() >= 0.9999
    np.testing.assert_array_equal(conf_ref.cpu().numpy(),
                                  O.confusion_matrix(ids.reshape(-1), pred.cpu().numpy().reshape(-1), C, C, ignore))


def test_predict_score_matches_reference_golden(case):
    cfg, g, data = case
    n_img = g["pred"].shape[0]
    lh = cuda(g["label_hat"].reshape(-1, data.C))
    ys = cuda(np.concatenate([y for _, y in batches_np(data, data.val_dataloader())]))
    conf = torch.zeros((data.C, data.C), dtype=torch.int64, device=DEV)
    pred = ops.predict_score(lh, n_img, data.S, data.H, data.H, conf, y=ys, ignore_index=data.ignore_index, return_pred=True)
    agree = (pred.cpu().numpy() == g["pred"][:, 0])
    assert agree.mean() >= 0.9999
    if agree.all():
        np.testing.assert_array_equal(conf.cpu().numpy(), g["conf"])
    else:  # a handful of argmax ties may differ; the matrix then differs by exactly those pixels
        assert np.abs(conf.cpu().numpy() - g["conf"]).sum() <= 2 * (~agree).sum()


def test_eval_step_is_the_unfused_path_in_one_call(case):
    """hb_eval_step (4 launches) == decode, search, label transfer, upsample+argmax, confusion called
    one by one; the whole validation set reproduces the reference's confusion matrix and mIoU."""
    from hbird_b200.utils.eval_metrics import miou_from_confusion

    cfg, g, data = case
    bank = build_bank_from_loader(data)
    C, S, H = data.C, data.S, data.H
    conf = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    conf_ref = torch.zeros_like(conf)
    for f, y in batches_np(data, data.val_dataloader()):
        q, ty = cuda(f.reshape(-1, f.shape[-1])), cuda(y)
        B = y.shape[0]
        s, i, qn = bank.search(q, 30, 64)
        lh_ref = ops.label_transfer(bank.label_table(), data.ps ** 2, s, i, qn, 0.02)
        pred_ref = ops.upsample_argmax(lh_ref, B, S, H, H)
        ops.confusion_accumulate(conf_ref, ops.decode_mask(ty, False).view(B, H, H), pred_ref, data.ignore_index)
        pred = torch.empty((B, H, H), dtype=torch.uint8, device=DEV)
        sc = torch.empty((B * S * S, 30), dtype=torch.float32, device=DEV)
        ix = torch.empty((B * S * S, 30), dtype=torch.int64, device=DEV)
        lh = bank.eval_step(q, ty, S, conf, data.ignore_index, pred=pred, scores=sc, idx=ix)
        assert torch.equal(lh, lh_ref) and torch.equal(pred, pred_ref) and torch.equal(sc, s) and torch.equal(ix, i)
        assert bank.last_search_launches() == 4
    assert torch.equal(conf, conf_ref)
    miou = miou_from_confusion(conf.cpu().numpy())[0]
    assert abs(miou - float(g["miou"])) <= 5e-4
    assert np.abs(conf.cpu().numpy() - g["conf"]).sum() <= 5e-4 * g["conf"].sum()
    bank.close()


def test_eval_step_replays_from_a_cuda_graph():
    """The step is allocation-free once the scratch is sized: capture it, replay it on new inputs."""
    data = SyntheticSegmentationData(num_train=12, num_val=4, input_size=112, patch_size=16, d_model=128,
                                     num_classes=7, batch_size=4, ignore_index=255, cells=4, seed=3)
    bank = build_bank_from_loader(data)
    (f, y), = batches_np(data, data.val_dataloader())
    C, S, H, B = data.C, data.S, data.H, y.shape[0]
    q, ty = cuda(f.reshape(-1, f.shape[-1])), cuda(y)
    conf_ref = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    bank.eval_step(q, ty, S, conf_ref, 255)  # also sizes the scratch
    sq, sy = torch.zeros_like(q), torch.zeros_like(ty)
    conf = torch.zeros_like(conf_ref)
    lh = torch.empty((B * S * S, C), dtype=torch.float32, device=DEV)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        bank.eval_step(sq, sy, S, conf, 255, label_hat=lh)
    torch.cuda.current_stream().wait_stream(side)
    conf.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        bank.eval_step(sq, sy, S, conf, 255, label_hat=lh)
    conf.zero_()
    sq.copy_(q), sy.copy_(ty)
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(conf, 2 * conf_ref)
    bank.close()


# ------------------------------------------------------------------ candidate bound of the tcgen05 pass
def _list_of(tile, col):
    """Which of a chunk's two lists scans column `col` of bank tile `tile` (search.cu epilogue)."""
    flip = ((tile * 0x9E3779B1) & 0xFFFFFFFF) >> 31
    return ((col // 32) & 1) ^ flip


def _planted(N, d, Q, positions, seed):
    """Bank with, per query, a set of near-duplicate rows at the given positions (well above the
    background), distinct enough that bf16 and fp32 rank them alike."""
    g = torch.Generator().manual_seed(seed)
    rows = torch.randn((N, d), generator=g)
    protos = torch.randn((Q, d), generator=g)
    protos = protos / protos.norm(dim=1, keepdim=True)
    for qi in range(Q):
        pos = positions(qi)
        sigma = torch.linspace(0.10, 0.45, len(pos)).unsqueeze(1)
        rows[pos] = protos[qi] * 6.0 + sigma * torch.randn((len(pos), d), generator=g)
    q = protos * 3.3
    return rows, q


def test_regular_stride_neighbours_are_spread_over_the_lists():
    """Same patch position in consecutive images = bank rows with a constant stride.  With stride 256
    every planted row sits in the same column of its tile; the per-tile flip of the two column sets
    must spread them over both lists, so all 40 are returned although a list holds only 32."""
    N, d, Q, n_dup = 65536, 128, 64, 40
    rows, q = _planted(N, d, Q, lambda qi: torch.arange(n_dup) * 256 + (qi * 7) % 256, seed=31)
    bank = bank_from_rows(rows.to(DEV))
    bank.configure_search(max_chunks=2)  # 128 tiles per chunk: the 40 rows share one chunk
    s, i, _ = bank.search(q.to(DEV), n_dup, 64)
    want = np.stack([(np.arange(n_dup) * 256 + (qi * 7) % 256) for qi in range(Q)])
    found = (i.cpu().numpy()[:, :, None] == want[:, None, :]).any(axis=1).mean()
    assert found >= 0.999, found
    bank.close()


def test_adversarial_placement_bound_and_strict_mode():
    """All 40 near-duplicates of a query inside ONE list of ONE chunk (placed with knowledge of the
    column-set hash, a single chunk forced): a 32-entry list keeps exactly the best 32 by bf16 score
    — the documented bound (include/hbird_b200.h, hb_search) — and k_prime = 128 keeps all 40."""
    N, d, Q, n_dup = 32768, 128, 32, 40
    slots = [(t, c) for t in range(N // 256) for c in range(0, 256, 32) if _list_of(t, c) == 0]

    def positions(qi):
        return torch.tensor([t * 256 + c + (qi % 32) for t, c in slots[qi:qi + 3 * n_dup:3]])

    rows, q = _planted(N, d, Q, positions, seed=33)
    bank = bank_from_rows(rows.to(DEV))
    fb = bank.export()[0]
    exact = (q.to(DEV) @ fb.T).topk(n_dup, dim=1)
    bank.configure_search(max_chunks=1)
    s64, i64, _ = bank.search(q.to(DEV), n_dup, 64)
    hit = (i64.unsqueeze(2) == exact.indices.unsqueeze(1)).any(1)  # (Q, rank): exact neighbour found?
    assert bool(hit[:, :30].float().mean() >= 0.99)      # the best ranks survive ...
    assert float(hit.float().sum(1).min()) >= 32          # ... exactly a list's worth of them, at least
    s128, i128, _ = bank.search(q.to(DEV), n_dup, 128)
    assert recall(i128.cpu().numpy(), exact.indices.cpu().numpy()) >= 0.999
    # the library default (>= 2 chunks) on the same bank finds them all with k_prime = 64 as well
    bank.configure_search(max_chunks=0)
    s, i, _ = bank.search(q.to(DEV), n_dup, 64)
    assert recall(i.cpu().numpy(), exact.indices.cpu().numpy()) >= 0.97
    bank.close()


def test_small_banks_run_with_wide_lists():
    """Banks of <= 16384 rows always use 64-entry lists: the bf16 top-64 is strict there, so k up to
    64 comes back complete even with k_prime = 64."""
    g = torch.Generator().manual_seed(35)
    rows = torch.randn((9000, 96), generator=g)
    q = torch.randn((200, 96), generator=g)
    bank = bank_from_rows(rows.to(DEV))
    s, i, _ = bank.search(q.to(DEV), 60, 64)
    fb = bank.export()[0]
    ref = (q.to(DEV) @ fb.T).topk(60, dim=1)
    assert recall(i.cpu().numpy(), ref.indices.cpu().numpy()) >= 0.999
    np.testing.assert_allclose(s.cpu().numpy(), ref.values.cpu().numpy(), rtol=1e-3, atol=1e-5)
    bank.close()


# ------------------------------------------------------------------ engine: parameters, legacy plugins, prebuilt banks
def _engine(data, nn_method="b200", **kw):
    from hbird_b200 import HbirdEvaluation
    from hbird_b200.models import FeatureExtractorSimple

    fe = FeatureExtractorSimple(data.model, data.ftr_extr_fn, data.S, data.d)
    return HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=kw.pop("n_neighbours", 30),
                           device=DEV, nn_method=nn_method, dataset_size=data.get_train_dataset_size(), **kw)


def test_engine_rejects_unknown_nn_params_and_accepts_the_faiss_knobs():
    cfg, g = load_golden("voc_tiny")
    data = SyntheticSegmentationData(**cfg)
    with pytest.raises(TypeError, match="k_prim"):
        _engine(data, nn_params={"k_prim": 64})  # a typo must not be swallowed (search_faiss.py:7 swallows it)
    # the faiss backend's own knobs are understood: idx_shard is a layout choice (single process: none),
    # use_fp16 drops the fp32 copy of the rows (the re-rank then reads the bf16 rows)
    ev = _engine(data, nn_params={"idx_shard": True, "use_fp16": True, "gpu_ids": [0]})
    assert ev.keep_f32 is False and ev.idx_shard is False
    miou = ev.evaluate(data.val_dataloader(), data.S, ignore_index=data.ignore_index)
    assert abs(miou - float(g["miou"])) <= 5e-3  # bf16 rows: the reference's use_fp16 trade
    ev.close()
    # k' follows n_neighbours: 64 up to 32 neighbours, 128 beyond; more than 128 is out of range
    ev = _engine(data, n_neighbours=40)
    assert ev.k_prime == 128 and ev.NN_algorithm.k_prime == 128
    ev.close()
    with pytest.raises(ValueError, match="n_neighbours"):
        _engine(data, n_neighbours=129)
    from hbird_b200 import NearestNeighborSearchB200

    with pytest.raises(TypeError, match="unexpected keyword"):
        NearestNeighborSearchB200(torch.randn(300, 64), n_neighbors=5, nprobe=3)


def test_legacy_plugin_path_ignores_the_backend_distances():
    """A third-party backend behind the reference's ABC (here: exact ids, but L2 distances — the wrong
    sign and scale for a softmax over cosines).  The engine must use the ids only and recompute the
    inner products from the bank rows, as the reference's _cross_attention does (hbird_eval.py:594-609)."""
    from hbird_b200 import register_nn_backend
    from hbird_b200.nn.search_base import NearestNeighborSearchBase

    class ExactIdsL2Distances(NearestNeighborSearchBase):
        def __init__(self, feature_memory, n_neighbors=30, distance_measure="dot_product", **kwargs):
            self.fm, self.k = feature_memory.clone(), n_neighbors

        def _initialize_index(self):
            return None

        def _add_features_to_index(self):
            return None

        def find_nearest_neighbors(self, q, k=None):
            ip = q.float() @ self.fm.T
            idx = ip.topk(self.k, dim=1).indices
            dist = torch.cdist(q.float(), self.fm).gather(1, idx) ** 2
            return idx.numpy(), dist.numpy()

    register_nn_backend("exact-ids-l2", ExactIdsL2Distances)
    cfg, g = load_golden("voc_tiny")
    data = SyntheticSegmentationData(**cfg)
    ev = _engine(data, nn_method="exact-ids-l2")
    miou, det = ev.evaluate(data.val_dataloader(), data.S, return_knn_details=True, ignore_index=data.ignore_index)
    assert abs(miou - float(g["miou"])) <= 5e-4
    np.testing.assert_allclose(det["knns_ca_labels"].numpy().reshape(g["label_hat"].shape), g["label_hat"], rtol=0, atol=2e-5)
    ev.close()


def test_evaluator_from_a_prebuilt_bank_takes_host_features():
    """HbirdEvaluation.from_bank + a loader of pinned host (features, masks): the e2e call bench.py
    times.  Same result as the engine that built the bank itself from images."""
    from hbird_b200 import HbirdEvaluation
    from hbird_b200.models import FeatureExtractorSimple

    cfg, g = load_golden("ade_tiny")
    data = SyntheticSegmentationData(**cfg)
    bank = build_bank_from_loader(data)
    fe = FeatureExtractorSimple(torch.nn.Identity(), lambda m, x: (x, None), data.S, data.d)
    ev = HbirdEvaluation.from_bank(fe, bank, data.C, 30, DEV, {"k_prime": 64})
    loader = [(torch.from_numpy(f).pin_memory(), torch.from_numpy(y).pin_memory()) for f, y in batches_np(data, data.val_dataloader())]
    miou = ev.evaluate(loader, data.S, ignore_index=data.ignore_index)
    assert abs(miou - float(g["miou"])) <= 5e-4
    assert np.abs(ev.last_confusion - g["conf"]).sum() <= 5e-4 * g["conf"].sum()
    ev.close()


def test_bench_synthetic_inputs_are_bit_identical_on_cpu_and_gpu():
    """bench.py's two arms must look at the same bank: bench_synth is integer hashing plus single fp32
    operations, so the CPU (reference arm) and the GPU (B200 arm) produce the same bits."""
    import bench
    import bench_synth as syn

    for name in ("cfg1", "cfg4"):
        w = bench.WORKLOADS[name]
        fc, mc = syn.images(w, 37, 3, torch.device("cpu"), stream=1)
        fg, mg = syn.images(w, 37, 3, torch.device(DEV), stream=1)
        assert torch.equal(fg.cpu(), fc) and torch.equal(mg.cpu(), mc)
        assert torch.equal(syn.prototypes(w, torch.device(DEV)).cpu(), syn.prototypes(w, torch.device("cpu")))


# ------------------------------------------------------------------ split search + two-stream pipeline
def test_split_search_two_slots_equal_the_one_call_search():
    g = torch.Generator().manual_seed(41)
    rows = torch.randn((60000, 128), generator=g).to(DEV)
    bank = bank_from_rows(rows)
    qa, qb = (torch.randn((700, 128), generator=g) * 2).to(DEV), (torch.randn((333, 128), generator=g) * 2).to(DEV)
    sa, ia, na = bank.search(qa, 30, 64)
    sb, ib, nb = bank.search(qb, 30, 64)
    # both slots in flight at once, finished out of order, on another stream
    other = torch.cuda.Stream()
    n0 = bank.search_begin(qa, 64, 0)
    n1 = bank.search_begin(qb, 64, 1)
    done = torch.cuda.Event()
    done.record()
    other.wait_event(done)
    with torch.cuda.stream(other):
        _, s1, i1 = bank.search_finish(1, qb, 30, want_label_hat=False, want_neighbours=True)
        lh0, s0, i0 = bank.search_finish(0, qa, 30, want_neighbours=True)
    other.synchronize()
    assert torch.equal(s0, sa) and torch.equal(i0, ia) and torch.equal(n0, na)
    assert torch.equal(s1, sb) and torch.equal(i1, ib) and torch.equal(n1, nb)
    assert torch.equal(lh0, ops.label_transfer(bank.label_table(), 1, sa, ia, na, 0.02))
    with pytest.raises(RuntimeError, match="no begun search"):
        bank.search_finish(0, qa, 30)
    bank.search_begin(qa, 64, 0)
    with pytest.raises(RuntimeError, match="already holds"):
        bank.search_begin(qb, 64, 0)
    with pytest.raises(ValueError, match="slot"):
        bank.search_begin(qb, 64, 2)
    bank.search_finish(0, qa, 30)
    torch.cuda.synchronize()
    bank.close()


def test_pipeline_accumulates_the_same_confusion_matrix_as_one_call_steps(case):
    """EvalPipeline (K2 of batch i+1 on one stream over the post-processing of batch i on another) is a
    scheduling change only: bit-identical confusion matrix, over several passes of the validation set."""
    from hbird_b200.pipeline import EvalPipeline

    cfg, g, data = case
    bank = build_bank_from_loader(data)
    C, S = data.C, data.S
    batches = [(cuda(f.reshape(-1, f.shape[-1])), cuda(y)) for f, y in batches_np(data, data.val_dataloader())] * 4
    ref = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    for q, y in batches:
        bank.eval_step(q, y, S, ref, data.ignore_index)
    conf = torch.zeros_like(ref)
    pipe = EvalPipeline(bank, bank.label_table(), S, conf, data.ignore_index, 30, 64, 0.02)
    for q, y in batches:
        pipe.submit(q, y, y.shape[0])
    pipe.flush()
    torch.cuda.synchronize()
    assert torch.equal(conf, ref)
    assert np.abs(conf.cpu().numpy() - 4 * g["conf"]).sum() <= 5e-4 * 4 * g["conf"].sum()
    bank.close()


def test_exchange_gives_up_on_a_missing_peer_without_killing_the_context():
    """Two simulated ranks, only rank 0 scatters: rank 0's merge waits for rank 1 at most the configured
    time, then returns without writing; check_status names the missing rank; the CUDA context, the bank
    and the exchange stay usable (round 1 trapped the context after 10 minutes)."""
    g = torch.Generator().manual_seed(47)
    rows = torch.randn((9000, 64), generator=g).to(DEV)
    bank = bank_from_rows(rows)
    xs = [ops.ShardExchange(r, 2, 64, 30, 0) for r in range(2)]
    ops.ShardExchange.connect_local(xs)
    q = (torch.randn((100, 64), generator=g) * 2).to(DEV)
    xs[0].set_timeout(30)
    xs[0].search_scatter(bank, q, [0, 50, 100], 30, 64)
    s, i = xs[0].merge()
    with pytest.raises(RuntimeError, match="rank 1 did not publish"):
        xs[0].check_status()
    xs[0].check_status()  # reported once
    # the late peer arrives: the same step can still be merged
    xs[1].search_scatter(bank, q, [0, 50, 100], 30, 64, idx_offset=9000)
    s0, i0, _ = bank.search(q, 30, 64)  # context and bank are alive
    assert torch.isfinite(s0).all()
    with pytest.raises(ValueError, match="positive"):
        xs[0].set_timeout(0)
    for o in xs + [bank]:
        o.close()


def test_failed_evaluation_leaves_the_bank_usable():
    """An exception in the middle of a pipelined evaluation (here: a batch of the wrong feature width)
    must not leave a pipeline slot 'begun': the next evaluation on the same bank works."""
    from hbird_b200 import HbirdEvaluation
    from hbird_b200.models import FeatureExtractorSimple

    cfg, g = load_golden("voc_tiny")
    data = SyntheticSegmentationData(**cfg)
    bank = build_bank_from_loader(data)
    fe = FeatureExtractorSimple(torch.nn.Identity(), lambda m, x: (x, None), data.S, data.d)
    ev = HbirdEvaluation.from_bank(fe, bank, data.C, 30, DEV, {"k_prime": 64})
    good = [(torch.from_numpy(f), torch.from_numpy(y)) for f, y in batches_np(data, data.val_dataloader())]
    bad = [good[0], (torch.zeros((2, data.S * data.S, data.d + 8)), good[0][1][:2])]
    with pytest.raises(ValueError):
        ev.evaluate(bad, data.S, ignore_index=data.ignore_index)
    miou = ev.evaluate(good, data.S, ignore_index=data.ignore_index)
    assert abs(miou - float(g["miou"])) <= 5e-4
    ev.close()
