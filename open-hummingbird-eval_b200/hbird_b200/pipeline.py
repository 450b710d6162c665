"""Two-stream software pipeline over validation batches.

A batch has a tensor-bound half (query prep + K2, the tcgen05 pass over the whole bank shard) and an
HBM/latency-bound half (K2b exact re-rank with the label transfer fused in — or, with a row-sharded
bank, shortlist / K2b scatter -> exchange wait -> merge + label transfer — and the fused tail).  The
second half of batch i does not feed the first half of batch i+1, so the two are issued on different
streams: K2 of batch i+1 on a high-priority stream, the post-processing of batch i on a second one.

What that buys, measured (tools/pipe_timeline.py, tools/ab_pipeline.py; profiles/r02_pipeline_timeline_*,
r02_ab_pipeline.json): the post-processing kernels are NOT co-resident with a search CTA — its ten
168-register warps fill the register file of two of an SM's four sub-partitions and a CTA is placed only
if all its warps find room — so K2b of batch i makes little progress while K2 of batch i+1 runs and
completes as that kernel's CTAs finish; the gain is the ramp-down of one search kernel and the launch
gap before the next one being filled with useful work (interleaved A/B: cfg1 1.334 -> 1.299 ms per step,
cfg2 8.00 -> 7.91, a 1.28 M x 768 shard 31.81 -> 31.67; with the threshold exchange on 8 GPUs 31.95 ->
31.33).  A 128-register build of the search kernel (MemoryBank.configure_coresidency) does run K2b beside
it (5.5 ms instead of 30 ms under a 31 ms search) and gains nothing: the step is power-bound, and the
4.3 GB row gather costs the same energy wherever it is scheduled.  With a sharded bank a rank that is
ahead of its peers starts the next search instead of idling in the exchange.  (Same arithmetic as the
one-call step: tests/test_gpu_fused.py.)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import distributed as hdist
from . import ops


# When is the second stream worth it?  The post-processing of a batch costs ~k'*4d bytes of HBM gather
# per query, the search 2*rows*d flop per query: their ratio is ~28 600 / rows on a B200, independent of
# d, so the possible gain shrinks with the bank.  Interleaved A/B on B200 (tools/ab_pipeline.py; one-call
# -> pipelined, ms per step): 102 k rows 1.334 -> 1.299, 1.02 M 8.00 -> 7.91, 1.28 M x 768 31.81 -> 31.67,
# 2.56 M 61.40 -> 61.31, 5.12 M 121.04 -> 120.64.  Never a loss where measured that way (an earlier
# sequential comparison at 10.24 M rows, 239.5 -> 240.4 ms, was within the clock drift); the bound below
# covers the measured range.
MAX_ROWS_PER_GPU = 1 << 23


def make_streams(device):
    """(search stream, post-processing stream): the search gets the higher priority, so that wherever
    CTAs of both are pending the tensor-core pass is placed first and the small kernels fill in."""
    return torch.cuda.Stream(device, priority=-1), torch.cuda.Stream(device, priority=0)


def worthwhile(bank: ops.MemoryBank) -> bool:
    return bank.rows <= MAX_ROWS_PER_GPU


class EvalPipeline:
    def __init__(self, bank: ops.MemoryBank, label_table: torch.Tensor, S: int, conf: torch.Tensor,
                 ignore_index: Optional[int], k: int = 30, k_prime: int = 64, beta: float = 0.02, idx_offset: int = 0,
                 world: int = 1, rank: int = 0, exchange: Optional[ops.ShardExchange] = None, streams=None):
        """conf: int64 (C, C) device matrix accumulated in place.  world > 1: `bank` is this rank's row
        shard, `label_table` the replicated table, `exchange` the connected ShardExchange (None = NCCL
        all-gather + merge).  streams: (search stream, post-processing stream) to reuse — creating streams
        per evaluation would make the caching allocator cudaMalloc fresh blocks for them every time."""
        self.bank, self.table, self.S, self.conf, self.ignore = bank, label_table, int(S), conf, ignore_index
        self.k, self.kp, self.beta, self.offset = int(k), int(k_prime), float(beta), int(idx_offset)
        self.world, self.rank, self.xchg = int(world), int(rank), exchange
        dev = conf.device
        if streams is None:
            streams = make_streams(dev)
        self.mma_stream, self.post_stream = streams
        self.pending = None
        self.count = 0

    def submit(self, q: torch.Tensor, y: torch.Tensor, n_images: int) -> None:
        """q fp32 (n_images*S*S, d): ALL queries of the batch; y fp32 id/255 masks of the images this rank
        post-processes (all of them, or its split_range slice when the bank is sharded), (.., H, W)."""
        cur = torch.cuda.current_stream(self.conf.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        slot = self.count & 1
        self.count += 1
        self.mma_stream.wait_event(ready)
        prepared = torch.cuda.Event()
        with torch.cuda.stream(self.mma_stream):
            qn = self.bank.search_begin(q, self.kp, slot, prepared)
            searched = torch.cuda.Event()
            searched.record(self.mma_stream)
        if self.pending is not None:
            # released together with this batch's search kernel (not in the gap before it)
            self.post_stream.wait_event(prepared)
            self._post(self.pending)
        self.pending = (slot, q, y, qn, n_images, searched, ready)
        for t in (q, y):
            t.record_stream(self.mma_stream)
            t.record_stream(self.post_stream)

    def _post(self, item) -> None:
        slot, q, y, qn, B, searched, ready = item
        S, per = self.S, self.S * self.S
        H, W = int(y.shape[-2]), int(y.shape[-1])
        self.post_stream.wait_event(searched)
        self.post_stream.wait_event(ready)
        with torch.cuda.stream(self.post_stream):
            qn.record_stream(self.post_stream)
            pp = self.bank.patch_pixels
            if self.world == 1:
                lh, _, _ = self.bank.search_finish(slot, q, self.k, 0, self.beta, None)
                ops.predict_score(lh, B, S, H, W, self.conf, y=y, ignore_index=self.ignore)
                return
            b0, b1 = hdist.split_range(B, self.world, self.rank)
            if self.xchg is not None:
                self.xchg.finish_scatter(self.bank, slot, q, hdist.query_split(B, per, self.world), self.k, self.offset)
                lh, _, _ = self.xchg.merge_transfer(self.table, pp, qn[b0 * per:b1 * per], self.beta)
            else:
                _, s, i = self.bank.search_finish(slot, q, self.k, self.offset, want_label_hat=False, want_neighbours=True)
                gs, gi = hdist.all_gather_topk(s, i)
                sl = slice(b0 * per, b1 * per)
                lh, _, _ = ops.merge_topk_transfer(gs[:, sl].contiguous(), gi[:, sl].contiguous(), self.table, pp,
                                                   qn[sl].contiguous(), self.beta)
            if b1 > b0:
                ops.predict_score(lh, b1 - b0, S, H, W, self.conf, y=y, ignore_index=self.ignore)

    def abort(self) -> None:
        """After an exception: drop the batch in flight and free the bank's pipeline slots, so that the
        bank can be searched again."""
        self.pending = None
        torch.cuda.synchronize(self.conf.device)
        self.bank.search_abort()

    def flush(self) -> None:
        """Post-process the last batch and make the caller's stream wait for everything."""
        if self.pending is not None:
            self._post(self.pending)
            self.pending = None
        cur = torch.cuda.current_stream(self.conf.device)
        cur.wait_stream(self.post_stream)
        cur.wait_stream(self.mma_stream)
