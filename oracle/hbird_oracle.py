"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain numpy restatement of the reference's dense nearest-neighbour evaluation path
(vpariza/open-hummingbird-eval), used as the checker for the CUDA kernels.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it;
the product package (open-hummingbird-eval_b200/) never does.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference ITSELF, executed unmodified in the dev container by
`oracle/make_golden.py` (faiss replaced by an exact inner-product shim, the definition of
IndexFlatIP) and committed under `tests/golden/`; `tests/test_oracle_golden.py` checks every
function below against those fixtures and against the PredsmIoU known-answer vectors.

Each function cites the reference lines it restates (paths relative to the reference root).
Arithmetic is float32 wherever the reference's is.  `hbird_oracle.c` (loaded by `c_oracle.py`) is an
independent plain-C twin of the byte/integer steps, held to the same fixtures.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

F32 = np.float32


# ------------------------------------------------------------------ loader contract (A0)
def decode_mask(y: np.ndarray, remap_255_to_0: bool) -> np.ndarray:
    """hbird/hbird_eval.py:219 and :309-310 — `(y * 255).long()`; bank side maps 255 -> 0."""
    ids = (y.astype(F32) * F32(255)).astype(np.int64)  # fp32 multiply, truncation toward zero
    if remap_255_to_0:
        ids = ids.copy()
        ids[ids == 255] = 0
    return ids


# ------------------------------------------------------------------ bank construction (A1-A3)
def patchify_gt(gt: np.ndarray, patch_size: int) -> np.ndarray:
    """hbird/hbird_eval.py:554-573 — (bs, c, h, w) -> (bs, h/ps, w/ps, c*ps*ps)."""
    bs, c, h, w = gt.shape
    g = gt.reshape(bs, c, h // patch_size, patch_size, w // patch_size, patch_size)
    g = g.transpose(0, 2, 4, 1, 3, 5)
    return g.reshape(bs, h // patch_size, w // patch_size, c * patch_size * patch_size)


def soft_labels(patchified: np.ndarray, num_classes: int) -> np.ndarray:
    """hbird/hbird_eval.py:319-320 — one_hot(...).float().mean(dim=3): class histogram / ps^2."""
    pp = patchified.shape[-1]
    flat = patchified.reshape(-1, pp)
    hist = np.zeros((flat.shape[0], num_classes), dtype=np.int64)
    rows = np.repeat(np.arange(flat.shape[0]), pp)
    np.add.at(hist, (rows, flat.reshape(-1)), 1)
    lab = hist.astype(F32) / F32(pp)
    return lab.reshape(patchified.shape[:-1] + (num_classes,))


def normalise_rows(f: np.ndarray) -> np.ndarray:
    """hbird/hbird_eval.py:324 — features / ||features||_2, no epsilon."""
    f = f.astype(F32)
    nrm = np.sqrt(np.sum(f * f, axis=-1, keepdims=True, dtype=F32)).astype(F32)
    return (f / nrm).astype(F32)


def build_memory(batches, num_classes: int, S: int) -> Tuple[np.ndarray, np.ndarray]:
    """hbird/hbird_eval.py:303-329,357-366 — unbounded memory: returns
    (feature_memory (N, d) fp32 unit rows, label_memory (N, C) fp32).
    `batches` yields (features (B, S*S, d) fp32, y (B, 1, H, W) fp32 = id/255)."""
    fm, lm = [], []
    for feats, y in batches:
        ids = decode_mask(y, True)
        ps = y.shape[-1] // S
        lab = soft_labels(patchify_gt(ids, ps), num_classes)
        fm.append(normalise_rows(feats).reshape(-1, feats.shape[-1]))
        lm.append(lab.reshape(-1, num_classes))
    return np.concatenate(fm), np.concatenate(lm)


def sample_patches(patchified: np.ndarray, num_classes: int, K: int, uniform: np.ndarray) -> np.ndarray:
    """hbird/hbird_eval.py:447-517 — bounded-memory sampler: per image the K patches with the
    smallest score*U, score = sum over classes present in the patch of (#patches containing the
    class); empty patches get 1e6.  `uniform` is the CPU RNG stream the reference draws with
    torch.rand (one value per non-empty patch, image order, :497-508).  Returns (B, K) indices."""
    B, S0, S1, P = patchified.shape
    SS = S0 * S1
    flat = patchified.reshape(B, SS, P)
    presence = np.zeros((B, SS, num_classes), dtype=bool)
    bi = np.repeat(np.arange(B), SS * P)
    pi = np.tile(np.repeat(np.arange(SS), P), B)
    presence[bi, pi, flat.reshape(-1)] = True
    class_freq = presence.sum(axis=1).astype(F32)
    scores = np.einsum("bpc,bc->bp", presence.astype(F32), class_freq).astype(F32)
    nonzero = presence.any(axis=2)
    scores[~nonzero] = F32(1e6)
    rand_map = np.ones_like(scores)
    start = 0
    for b in range(B):
        cnt = int(nonzero[b].sum())
        rand_map[b, nonzero[b]] = uniform[start:start + cnt]
        start += cnt
    scores = (scores * rand_map).astype(F32)
    # torch.topk(largest=False): ascending by value; ties broken by lower index here
    return np.argsort(scores, axis=1, kind="stable")[:, :K]


# ------------------------------------------------------------------ search (A6)
def _topk_rows(s: np.ndarray, kk: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row exact top-kk of a score block: descending score, ties by ascending index."""
    N = s.shape[1]
    if kk < N:
        part = np.argpartition(-s, kk - 1, axis=1)[:, :kk]
    else:
        part = np.tile(np.arange(N), (s.shape[0], 1))
    ps = np.take_along_axis(s, part, axis=1)
    order = np.lexsort((part, -ps), axis=1)
    return np.take_along_axis(part, order, axis=1), np.take_along_axis(ps, order, axis=1)


def search_exact_ip(q: np.ndarray, bank: np.ndarray, k: int, block: Optional[int] = None,
                    threads: Optional[int] = None, scratch_bytes: float = 6e9) -> Tuple[np.ndarray, np.ndarray]:
    """hbird/nn/search_faiss.py:39-41,83-90 — GpuIndexFlatIP.search: exact top-k by inner product
    of raw queries with the unit-norm bank, descending; faiss pads with (-inf, -1) when N < k.
    Returns (indices int64 (Q, k), distances fp32 (Q, k)) — indices first, as the plugin does.
    The GEMM uses the BLAS thread pool; the per-row selection is spread over `threads` host threads
    (default: all cores) so that the CPU baseline uses the whole host, as faiss-cpu would."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    q = np.ascontiguousarray(q, dtype=F32)
    bank = np.ascontiguousarray(bank, dtype=F32)
    Q, N = q.shape[0], bank.shape[0]
    kk = min(k, N)
    idx = np.full((Q, k), -1, dtype=np.int64)
    dist = np.full((Q, k), -np.inf, dtype=F32)
    threads = threads or os.cpu_count() or 1
    if block is None:
        # a block of B query rows costs ~16 B per (row, bank row): fp32 scores, their negation and the
        # int64 argpartition output; keep that within `scratch_bytes` of host memory (48 rows at
        # N = 10.24 M, 366 at N = 1.024 M) so that the baseline cannot exhaust the host
        block = int(max(8, min(4096, scratch_bytes / (16.0 * max(N, 1)))))
    for a in range(0, Q, block):
        s = q[a:a + block] @ bank.T
        rows = s.shape[0]
        nsplit = max(1, min(threads, rows // 2))
        if nsplit == 1:
            i, d = _topk_rows(s, kk)
            idx[a:a + rows, :kk], dist[a:a + rows, :kk] = i, d
        else:
            bounds = [rows * t // nsplit for t in range(nsplit + 1)]
            with ThreadPoolExecutor(nsplit) as pool:
                parts = list(pool.map(lambda t: _topk_rows(s[bounds[t]:bounds[t + 1]], kk), range(nsplit)))
            for t, (i, d) in enumerate(parts):
                idx[a + bounds[t]:a + bounds[t + 1], :kk] = i
                dist[a + bounds[t]:a + bounds[t + 1], :kk] = d
    return idx, dist


def search_exact_l2(q: np.ndarray, bank: np.ndarray, k: int, block: int = 4096) -> Tuple[np.ndarray, np.ndarray]:
    """hbird/nn/search_faiss.py:45-46,83-90 — GpuIndexFlatL2.search: exact top-k by SQUARED
    Euclidean distance, ascending (faiss reports squared L2).  Distances are accumulated as
    sum((q-x)^2) in float64 and rounded to fp32, so the oracle carries no cancellation error."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    bank = np.ascontiguousarray(bank, dtype=np.float64)
    Q, N = q.shape[0], bank.shape[0]
    kk = min(k, N)
    idx = np.full((Q, k), -1, dtype=np.int64)
    dist = np.full((Q, k), np.inf, dtype=F32)
    bn = (bank * bank).sum(1)
    for a in range(0, Q, block):
        qq = q[a:a + block]
        d2 = (qq * qq).sum(1)[:, None] + bn[None, :] - 2.0 * (qq @ bank.T)
        i, negd = _topk_rows(-d2, kk)
        # exact distances for the winners, summed the direct way
        exact = ((qq[:, None, :] - bank[i]) ** 2).sum(-1)
        order = np.lexsort((i, exact), axis=1)
        idx[a:a + qq.shape[0], :kk] = np.take_along_axis(i, order, 1)
        dist[a:a + qq.shape[0], :kk] = np.take_along_axis(exact, order, 1).astype(F32)
    return idx, dist


def merge_shards(shard_idx: np.ndarray, shard_dist: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """hbird/nn/search_faiss.py:53-63 — faiss.IndexShards merges per-shard results on the host:
    top-k of the union, descending, ties by smaller (global) index.  Inputs are (G, Q, k)."""
    G, Q, _ = shard_idx.shape
    ai = shard_idx.transpose(1, 0, 2).reshape(Q, -1)
    ad = shard_dist.transpose(1, 0, 2).reshape(Q, -1).astype(F32)
    ad = np.where(ai < 0, -np.inf, ad)
    order = np.lexsort((ai, -ad), axis=1)[:, :k]
    return np.take_along_axis(ai, order, axis=1), np.take_along_axis(ad, order, axis=1)


def shortlist_bound(shard_scores_desc: np.ndarray, kp: int) -> np.ndarray:
    """The bound the threshold exchange derives on every shard (rerank.cu, rerank_query<SUBSET>): from the
    scores at ranks kp, kp/2, kp/4, kp/8 of each shard's sorted candidate list — if j shards hold kp/j
    candidates >= x each, the union holds kp candidates >= x — take the largest such x over j = 1, 2, 4, 8.
    shard_scores_desc: (G, Q, kp) per-shard candidate scores, descending, -inf where a shard has fewer.
    Returns (Q,) lower bounds of the kp-th best score of the union (-inf = no bound).  What
    faiss.IndexShards (search_faiss.py:53-63) needs from a shard is only its candidates at or above it."""
    G = shard_scores_desc.shape[0]
    bound = np.full(shard_scores_desc.shape[1], -np.inf, dtype=F32)
    for j in (1, 2, 4, 8):
        if G < j or kp // j < 1:
            continue
        stat = shard_scores_desc[:, :, kp // j - 1]           # (G, Q): every shard's (kp/j)-th best
        jth = -np.sort(-stat, axis=0)[j - 1]                  # j-th largest over the shards
        bound = np.maximum(bound, jth.astype(F32))
    return bound


# ------------------------------------------------------------------ label transfer (A7-A8)
def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """torch.nn.functional.normalize(dim=-1): x / max(||x||, eps)."""
    n = np.sqrt(np.sum(x.astype(F32) ** 2, axis=-1, keepdims=True, dtype=F32))
    return (x / np.maximum(n, F32(eps))).astype(F32)


def cross_attention(q: np.ndarray, key_feats: np.ndarray, key_labels: np.ndarray, beta: float = 0.02) -> np.ndarray:
    """hbird/hbird_eval.py:575-609 — q (B,N,D), k (B,N,K,D), v (B,N,K,C) -> (B,N,C)."""
    qn = l2_normalize(q)
    kn = l2_normalize(key_feats)
    attn = np.einsum("bnd,bnkd->bnk", qn, kn).astype(F32) / F32(beta)
    attn = attn - attn.max(axis=-1, keepdims=True)
    e = np.exp(attn).astype(F32)
    attn = e / e.sum(axis=-1, keepdims=True, dtype=F32)
    return np.einsum("bnk,bnkc->bnc", attn, key_labels.astype(F32)).astype(F32)


def transfer_labels(q: np.ndarray, feature_memory: np.ndarray, label_memory: np.ndarray, idx: np.ndarray,
                    beta: float = 0.02) -> np.ndarray:
    """hbird/hbird_eval.py:611-637 then :227 — gather neighbour features/labels, cross-attend.
    q (B, N, D); idx (B*N, k).  Returns label_hat (B, N, C)."""
    B, N, D = q.shape
    k = idx.shape[1]
    kf = feature_memory[idx.reshape(-1)].reshape(B, N, k, D)
    kl = label_memory[idx.reshape(-1)].reshape(B, N, k, -1)
    return cross_attention(q, kf, kl, beta)


# ------------------------------------------------------------------ upsample + argmax (A9)
def _linear_index_weights(out_size: int, in_size: int):
    """ATen area_pixel_compute_source_index (align_corners=False) + guard_index_and_lambda."""
    scale = F32(in_size) / F32(out_size)
    dst = np.arange(out_size, dtype=F32)
    src = scale * (dst + F32(0.5)) - F32(0.5)
    src = np.maximum(src, F32(0)).astype(F32)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = np.clip(src - i0.astype(F32), 0, 1).astype(F32)
    l0 = (F32(1) - l1).astype(F32)
    return i0, i1, l0, l1


def upsample_bilinear(label_hat: np.ndarray, S: int, H: int, W: int) -> np.ndarray:
    """hbird/hbird_eval.py:235-240 — (B, S*S, C) -> reshape (B,S,S,C) -> permute (B,C,S,S) ->
    F.interpolate(size=(H,W), mode='bilinear', align_corners=False) -> (B, C, H, W) fp32."""
    B, _, C = label_hat.shape
    t = label_hat.reshape(B, S, S, C).transpose(0, 3, 1, 2).astype(F32)
    y0, y1, ly0, ly1 = _linear_index_weights(H, S)
    x0, x1, lx0, lx1 = _linear_index_weights(W, S)
    top = t[:, :, y0][:, :, :, x0] * lx0 + t[:, :, y0][:, :, :, x1] * lx1
    bot = t[:, :, y1][:, :, :, x0] * lx0 + t[:, :, y1][:, :, :, x1] * lx1
    return (top * ly0[None, None, :, None] + bot * ly1[None, None, :, None]).astype(F32)


def predict_map(label_hat: np.ndarray, S: int, H: int, W: int) -> np.ndarray:
    """hbird/hbird_eval.py:243 — argmax over classes (first maximum), (B, 1, H, W) int64."""
    return upsample_bilinear(label_hat, S, H, W).argmax(axis=1)[:, None]


# ------------------------------------------------------------------ scoring (A11-A12)
def confusion_matrix(gt: np.ndarray, pred: np.ndarray, num_gt: int, num_pred: int,
                     ignore_index: Optional[int]) -> np.ndarray:
    """hbird/utils/eval_metrics.py:73-104 — mask gt != ignore, drop out-of-range, bincount."""
    gt = gt.reshape(-1).astype(np.int64)
    pred = pred.reshape(-1).astype(np.int64)
    if ignore_index is not None:
        m = gt != ignore_index
        gt, pred = gt[m], pred[m]
    valid = (gt >= 0) & (gt < num_gt) & (pred >= 0) & (pred < num_pred)
    gt, pred = gt[valid], pred[valid]
    return np.bincount(gt * num_pred + pred, minlength=num_gt * num_pred).reshape(num_gt, num_pred).astype(np.int64)


def iou_matrix(conf: np.ndarray) -> np.ndarray:
    """hbird/utils/eval_metrics.py:112-131 — TP / clamp(row + col - TP, 1e-8), float64."""
    c = conf.astype(np.float64)
    denom = c.sum(1, keepdims=True) + c.sum(0, keepdims=True) - c
    return c / np.maximum(denom, 1e-8)


def hungarian_mapping(conf: np.ndarray) -> np.ndarray:
    """hbird/utils/eval_metrics.py:143-159 — linear_sum_assignment(1 - IoU); unmatched -> 0."""
    from scipy.optimize import linear_sum_assignment

    r, c = linear_sum_assignment(1.0 - iou_matrix(conf))
    mapping = np.zeros(conf.shape[1], dtype=np.int64)
    mapping[c] = r
    return mapping


def miou_from_confusion(conf: np.ndarray, linear_probe: bool = False, many_to_one: bool = False,
                        precision_based: bool = False):
    """hbird/utils/eval_metrics.py:162-288 — (mIoU, tp, fp, fn, matched_bg_fraction) with the
    default Hungarian matching (what evaluate() uses, hbird_eval.py:253)."""
    G, P = conf.shape
    row_sum = conf.sum(1)
    if linear_probe:
        col_sum = conf.sum(0)
        tp = np.array([conf[i, i] if i < P else 0 for i in range(G)], dtype=np.int64)
        fp = np.array([col_sum[i] - conf[i, i] if i < P else 0 for i in range(G)], dtype=np.int64)
        fn = row_sum - tp
        bg = 0.0
    else:
        if many_to_one:
            c = conf.astype(np.float64)
            score = c / np.maximum(c.sum(0, keepdims=True), 1e-8) if precision_based else iou_matrix(conf)
            mapping = score.argmax(axis=0)
            bg = float((mapping == 0).sum() / max(P, 1))
        else:
            mapping = hungarian_mapping(conf)
            bg = 1.0 / max(G, 1)
        mapped = np.zeros((G, G), dtype=np.int64)
        np.add.at(mapped, (slice(None), mapping), conf)  # index_add_ over columns (:197-198)
        tp = np.diag(mapped).copy()
        fp = mapped.sum(0) - tp
        fn = row_sum - tp
    denom = (tp + fp + fn).astype(np.float64)
    iou = tp.astype(np.float64) / np.maximum(denom, 1e-8)
    return float(iou.mean()), tp.tolist(), fp.tolist(), fn.tolist(), bg


# ------------------------------------------------------------------ end to end (hbird_eval.py:184-265)
def evaluate(feature_memory: np.ndarray, label_memory: np.ndarray, val_batches, num_classes: int, S: int,
             k: int = 30, ignore_index: int = 255, beta: float = 0.02, return_details: bool = False):
    """Reference evaluate(): search -> gather -> cross-attention -> upsample -> argmax -> mIoU.
    `val_batches` yields (features (B, S*S, d) fp32 raw, y (B, 1, H, W) fp32 = id/255)."""
    conf = np.zeros((num_classes, num_classes), dtype=np.int64)
    det = {"idx": [], "dist": [], "label_hat": [], "pred": []}
    for feats, y in val_batches:
        B, N, D = feats.shape
        H, W = y.shape[-2:]
        gt = decode_mask(y, False)
        idx, dist = search_exact_ip(feats.reshape(B * N, D), feature_memory, k)
        label_hat = transfer_labels(feats, feature_memory, label_memory, idx, beta)
        pred = predict_map(label_hat, S, H, W)
        conf += confusion_matrix(gt, pred, num_classes, num_classes, ignore_index)
        if return_details:
            det["idx"].append(idx)
            det["dist"].append(dist)
            det["label_hat"].append(label_hat)
            det["pred"].append(pred)
    miou, tp, fp, fn, bg = miou_from_confusion(conf)
    if return_details:
        return miou, conf, det
    return miou, conf
