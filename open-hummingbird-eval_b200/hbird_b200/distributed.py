"""One-process-per-GPU plumbing for a row-sharded memory bank.

The bank is split row-wise: rank r owns a contiguous block of global rows.  Every rank searches
every query against its shard (K2/K2b); the per-shard (score, global index) top-k lists are
all-gathered over NCCL/NVLink and merged by the K3 kernel (ops.merge_topk) — the exchange step
faiss.IndexShards performs on the host (search_faiss.py:53-63).  The label table is replicated by
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


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of n_rows global rows: rank r owns [begin, end)."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def gather_counts(n_local: int, device, group=None) -> List[int]:
    """Row count of every rank's shard (shards built from a strided loader split can differ)."""
    _, world = dist_info(group)
    if world == 1:
        return [int(n_local)]
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    out = torch.empty((world,), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
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
    dist.all_gather_into_tensor(gathered, buf, group=group)
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
    dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, idx.contiguous(), group=group)
    return gs.view(world, Q, k), gi.view(world, Q, k)


def split_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """The slice of a batch (images or queries) a rank post-processes after the merge."""
    return shard_bounds(n, world, rank)
