"""world_size-2 gloo run of the one-process-per-GPU host logic (hbird_b200.distributed): shard
bounds, ragged label-table replication, the all-gather layout hb_merge_topk consumes (checked with
the oracle's shard merge), and the confusion-matrix all-reduce.  CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hbird_b200 import distributed as hdist
from oracle import hbird_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert hdist.dist_info() == (rank, world)
        rng = np.random.default_rng(0)  # same data on every rank
        N, d, Q, k, C = 301, 16, 23, 30, 5
        bank = O.normalise_rows(rng.standard_normal((N, d)).astype(np.float32))
        q = rng.standard_normal((Q, d)).astype(np.float32) * 3
        labels = rng.integers(0, 50, size=(N, C)).astype(np.int16)
        a, b = hdist.shard_bounds(N, world, rank)
        # ragged shards: counts gathered, offsets derived
        counts = hdist.gather_counts(b - a, torch.device("cpu"))
        assert counts == [hdist.shard_bounds(N, world, r)[1] - hdist.shard_bounds(N, world, r)[0] for r in range(world)]
        off = hdist.offsets_from_counts(counts)
        assert off[rank] == a
        # label table replication (int16 rows moved as bytes)
        table = hdist.all_gather_rows(torch.from_numpy(labels[a:b]), counts)
        assert table.dtype == torch.int16 and np.array_equal(table.numpy(), labels)
        # per-shard search (oracle stands in for the kernel), gather, merge
        li, ld = O.search_exact_ip(q, bank[a:b], k)
        gs, gi = hdist.all_gather_topk(torch.from_numpy(ld), torch.from_numpy(li + a))
        assert gs.shape == (world, Q, k) and gi.shape == (world, Q, k)
        mi, md = O.merge_shards(gi.numpy(), gs.numpy(), k)
        ri, rd = O.search_exact_ip(q, bank, k)
        assert np.array_equal(mi, ri) and np.array_equal(md, rd)
        # confusion matrices add up (eval_metrics.py:251-252)
        conf = torch.full((C, C), rank + 1, dtype=torch.int64)
        dist.all_reduce(conf, op=dist.ReduceOp.SUM)
        assert int(conf[0, 0]) == sum(range(1, world + 1))
        # plumbing of the fused shard exchange: handles travel as fixed-size byte strings, every rank
        # takes the same p2p-or-NCCL decision, and the query split matches the image split
        blobs = hdist.all_gather_bytes(bytes([rank]) * 64, torch.device("cpu"))
        assert blobs == [bytes([r]) * 64 for r in range(world)]
        assert hdist.all_ranks_ok(True, torch.device("cpu")) is True
        assert hdist.all_ranks_ok(rank != 1, torch.device("cpu")) is False
        qs = hdist.query_split(7, 196, world)
        assert qs[0] == 0 and qs[-1] == 7 * 196 and len(qs) == world + 1
        assert (qs[rank], qs[rank + 1]) == tuple(196 * v for v in hdist.split_range(7, world, rank))
        # query-parallel feature extraction with a row-sharded bank: every rank extracts the features of
        # its image slice and the ragged (n_r * S*S, d) blocks are all-gathered in image order; with
        # fewer images than ranks a slice is empty
        for n_img in (5, 1):
            per_img, dd = 9, 8
            allq = np.arange(n_img * per_img * dd, dtype=np.float32).reshape(n_img * per_img, dd)
            spans = [hdist.split_range(n_img, world, r) for r in range(world)]
            i0, i1 = spans[rank]
            mine = torch.from_numpy(allq[i0 * per_img:i1 * per_img].copy())
            got = hdist.all_gather_rows(mine, [(e - a) * per_img for a, e in spans])
            assert got.dtype == torch.float32 and np.array_equal(got.numpy(), allq)
        # replicas: validation batches are dealt round-robin, every batch to exactly one rank
        dealt = [step for step in range(7) if step % world == rank]
        every = [None] * world
        dist.all_gather_object(every, dealt)
        assert sorted(v for part in every for v in part) == list(range(7))
        # no CUDA device here: window creation fails on every rank and all ranks fall back together
        assert hdist.connect_shard_exchange(100, 30, torch.device("cpu")) is None
        b0, b1 = hdist.split_range(7, world, rank)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"{b0},{b1}")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    spans = [tuple(map(int, open(tmp_path / f"ok{r}").read().split(","))) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == 7 and spans[0][1] == spans[1][0]


def test_single_process_helpers():
    assert hdist.dist_info() == (0, 1)
    assert hdist.shard_bounds(10, 3, 0) == (0, 3) and hdist.shard_bounds(10, 3, 2) == (6, 10)
    s, i = torch.zeros(4, 3), torch.zeros(4, 3, dtype=torch.int64)
    gs, gi = hdist.all_gather_topk(s, i)
    assert gs.shape == (1, 4, 3) and gi.shape == (1, 4, 3)
    assert hdist.query_split(16, 1369, 8) == [2 * 1369 * r for r in range(9)]
    assert hdist.query_split(3, 10, 4) == [0, 0, 10, 20, 30]  # fewer images than ranks: empty slices
    assert hdist.all_gather_bytes(b"x" * 64, torch.device("cpu")) == [b"x" * 64]
    assert hdist.all_ranks_ok(False, torch.device("cpu")) is False
