"""One-process-per-GPU plumbing for a row-sharded memory bank.

The bank is split row-wise: rank r owns a contiguous block of global rows.  Every rank searches
every query against its shard (K2/K2b).  The exchange step faiss.IndexShards performs on the host
(search_faiss.py:53-63) is fused into the kernels around it: K2b stores each query's shard top-k
into the window of the rank that post-processes it (NVLink peer stores, ops.ShardExchange) and
that rank's merge kernel waits for all sources — no collective call.  Where peer mapping is not
available the per-shard lists are all-gathered over NCCL and merged by K3 (ops.merge_topk).  The label table is replicated by
an all-gather at build time; the (C, C) confusion matrix is all-reduced once at the end
(eval_metrics.py:251-252).  All functions below are backend-agnostic torch.distributed calls, so
the host logic is covered on CPU with gloo (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def dist_info(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _host_transport(group=None) -> bool:
    """NCCL moves device tensors itself; any other backend (gloo: CPU tests, or several ranks
    time-sharing one GPU) is handed host copies and the result is moved back."""
    return dist.get_backend(group) != "nccl"


def _all_gather_into(out: torch.Tensor, inp: torch.Tensor, group=None) -> None:
    if inp.is_cuda and _host_transport(group):
        tmp = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(tmp, inp.cpu(), group=group)
        out.copy_(tmp)
    else:
        dist.all_gather_into_tensor(out, inp, group=group)


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce (the (C, C) confusion matrix, eval_metrics.py:251-252)."""
    if t.is_cuda and _host_transport(group):
        tmp = t.cpu()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        t.copy_(tmp)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def broadcast(t: torch.Tensor, src: int, group=None) -> torch.Tensor:
    if t.is_cuda and _host_transport(group):
        tmp = t.cpu()
        dist.broadcast(tmp, src=src, group=group)
        t.copy_(tmp)
    else:
        dist.broadcast(t, src=src, group=group)
    return t


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of n_rows global rows: rank r owns [begin, end)."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def balanced_counts(counts: List[int], times_ms: List[float], max_shift: float = 0.10, gain: float = 1.0) -> List[int]:
    """Shard sizes proportional to each rank's measured search speed (rows per ms on its current
    shard), so that all ranks finish a batch together: with a row-sharded bank every step ends with
    everybody's results, i.e. the job runs at the pace of the slowest GPU, and GPUs under the same
    power cap differ by a few per cent.  Each shard stays within +-max_shift of its current size;
    the total is preserved exactly.  gain < 1 applies only that fraction of the correction.  Pure
    integer/float host logic."""
    total = sum(counts)
    if len(counts) < 2 or total == 0 or any(t <= 0 for t in times_ms) or any(c <= 0 for c in counts):
        return list(counts)
    if gain != 1.0:  # damped correction: move only part of the way towards equal times
        mean_t = sum(times_ms) / len(times_ms)
        times_ms = [mean_t + gain * (t - mean_t) for t in times_ms]
    speed = [c / t for c, t in zip(counts, times_ms)]
    share = [v / sum(speed) for v in speed]
    want = [min(max(total * sh, c * (1.0 - max_shift)), c * (1.0 + max_shift)) for sh, c in zip(share, counts)]
    scale = total / sum(want)
    new = [max(1, int(round(x * scale))) for x in want]
    new[new.index(max(new))] += total - sum(new)  # rounding remainder goes to the largest shard
    return new


def gather_floats(value: float, device, group=None) -> List[float]:
    _, world = dist_info(group)
    if world == 1:
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = torch.empty((world,), dtype=torch.float64, device=device)
    _all_gather_into(out, t, group)
    return [float(v) for v in out.tolist()]


def gather_counts(n_local: int, device, group=None) -> List[int]:
    """Row count of every rank's shard (shards built from a strided loader split can differ)."""
    _, world = dist_info(group)
    if world == 1:
        return [int(n_local)]
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    out = torch.empty((world,), dtype=torch.int64, device=device)
    _all_gather_into(out, t, group)
    return [int(v) for v in out.tolist()]


def offsets_from_counts(counts: List[int]) -> List[int]:
    off, acc = [], 0
    for c in counts:
        off.append(acc)
        acc += c
    return off


def all_gather_rows(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """Concatenate per-rank (n_r, ...) row blocks in rank order (ragged n_r allowed).  Used to
    replicate the int16 label table; it is moved as raw bytes because NCCL has no int16."""
    rank, world = dist_info(group)
    if world == 1:
        return local
    row_shape = tuple(local.shape[1:])
    row_bytes = local.element_size()
    for s in row_shape:
        row_bytes *= s
    nmax = max(counts)
    buf = torch.zeros((nmax, row_bytes), dtype=torch.uint8, device=local.device)
    if counts[rank]:
        buf[:counts[rank]] = local.contiguous().view(torch.uint8).reshape(counts[rank], row_bytes)
    # (world * nmax, row_bytes): the concatenated output form both NCCL and gloo accept
    gathered = torch.empty((world * nmax, row_bytes), dtype=torch.uint8, device=local.device)
    _all_gather_into(gathered, buf, group)
    gathered = gathered.view(world, nmax, row_bytes)
    parts = [gathered[r, :counts[r]] for r in range(world)]
    flat = torch.cat(parts, dim=0).contiguous()
    return flat.view(local.dtype).reshape((sum(counts),) + row_shape)


def all_gather_topk(scores: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(Q, k) per-shard results -> (G, Q, k) stacked in rank order: the layout hb_merge_topk takes."""
    _, world = dist_info(group)
    if world == 1:
        return scores.unsqueeze(0), idx.unsqueeze(0)
    Q, k = scores.shape
    gs = torch.empty((world * Q, k), dtype=scores.dtype, device=scores.device)
    gi = torch.empty((world * Q, k), dtype=idx.dtype, device=idx.device)
    _all_gather_into(gs, scores.contiguous(), group)
    _all_gather_into(gi, idx.contiguous(), group)
    return gs.view(world, Q, k), gi.view(world, Q, k)


def all_gather_bytes(blob: bytes, device, group=None) -> List[bytes]:
    """Fixed-size byte strings of every rank, in rank order (IPC handles of the shard exchange)."""
    _, world = dist_info(group)
    if world == 1:
        return [blob]
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    out = torch.empty((world * len(blob),), dtype=torch.uint8, device=device)
    _all_gather_into(out, t, group)
    raw = out.cpu().numpy().tobytes()
    return [raw[r * len(blob):(r + 1) * len(blob)] for r in range(world)]


def all_ranks_ok(ok: bool, device, group=None) -> bool:
    """True only if `ok` holds on every rank (so that all ranks pick the same exchange path)."""
    _, world = dist_info(group)
    if world == 1:
        return bool(ok)
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cpu" if _host_transport(group) else device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def query_split(n_images: int, queries_per_image: int, world: int) -> List[int]:
    """qsplit for the fused exchange: rank p post-processes images split_range(n_images, world, p),
    i.e. queries [qsplit[p], qsplit[p+1])."""
    return [shard_bounds(n_images, world, p)[0] * queries_per_image for p in range(world)] + \
        [n_images * queries_per_image]


def connect_shard_exchange(slice_capacity: int, max_k: int, device: torch.device, group=None,
                           threshold_exchange: bool = True):
    """Create this rank's ShardExchange window and map every peer's (CUDA IPC handles travel by
    all-gather).  Returns None — on ALL ranks — if any rank cannot map its peers, in which case the
    caller uses the NCCL all-gather + merge path.  threshold_exchange (same on all ranks): shards swap
    order statistics of their bf16 shortlists and re-rank only the candidates at or above the common
    bound (ops.ShardExchange.configure), instead of their whole top-k' each."""
    from . import ops

    rank, world = dist_info(group)
    xchg, ok = None, True
    try:
        xchg = ops.ShardExchange(rank, world, slice_capacity, max_k, device.index or 0)
        handle = xchg.handle()
    except RuntimeError:
        ok, handle = False, bytes(64)
    handles = all_gather_bytes(handle, device, group)
    if ok:
        try:
            xchg.connect(handles)
        except RuntimeError:
            ok = False
    if not all_ranks_ok(ok, device, group):
        if xchg is not None:
            xchg.close()
        return None
    xchg.configure(bool(threshold_exchange))
    return xchg


def close_shard_exchange(xchg, group=None) -> None:
    """Orderly teardown: unmap peers everywhere, barrier, then free the own window."""
    if xchg is None:
        return
    xchg.disconnect()
    _, world = dist_info(group)
    if world > 1:
        dist.barrier(group=group)
    xchg.close()


def split_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """The slice of a batch (images or queries) a rank post-processes after the merge."""
    return shard_bounds(n, world, rank)
