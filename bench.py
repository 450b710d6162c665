"""bench.py — patch-queries/sec of the dense nearest-neighbour evaluation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one validation batch of synthetic VOC-shaped input:
query prep -> kNN search against the HBM-resident memory bank (tcgen05 GEMM + fused top-k', exact
fp32 re-rank with the label transfer fused in) -> fused tail (mask decode + bilinear upsample +
argmax + confusion-matrix update); 4 kernel launches.

Default workload = BASELINE.json configs[2], the north-star configuration ("cfg3": DINOv2 ViT-B/14
518 px, 10,240,000-patch bank, d=768, k=30, 16 images = 21,904 patch-queries per step).

N = 1: the bank lives on one GPU.  N > 1 (one process per GPU): the headline `value` is the
ROW-SHARDED bank (faiss.IndexShards, search_faiss.py:53-63; N/G rows per GPU, every rank searches
every query, shard results exchanged by the fused NVLink exchange, each rank post-processes its
image slice) — strong scaling.  Replicas (faiss.IndexReplicas, :65-74; weak scaling) and the other
BASELINE configs are reported under `by_workload` in the same line ("q/s vs bank size").

`e2e` = the same metric through the engine's public call with HOST inputs: one
`HbirdEvaluation.evaluate([batch])` per step (pinned host features and masks in, Python float mIoU
out; H2D copies, confusion-matrix read-back and Hungarian matching inside the timed region).

`parity` = the timed inputs through the CUDA path against the CPU oracle (recall@30, score error on
a query sample; mIoU / confusion matrix / pixel agreement on whole images).

`--impl reference` times the CPU oracle port of the reference path (numpy/BLAS on all host cores)
on a bounded sample of the same workload, with the same synthetic bank generated on the host; it
does not import the product library.  Under torchrun only rank 0 runs it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "open-hummingbird-eval_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: bank rows, d, S, patch px, classes, ignore, images per step
    "cfg1": dict(N=102_400, d=384, S=14, ps=16, C=21, ignore=255, B=64,
                 desc="DINO ViT-S/16 224px, 102,400-patch bank, d=384, k=30 (BASELINE configs[0])"),
    "cfg2": dict(N=1_024_000, d=384, S=14, ps=16, C=21, ignore=255, B=64,
                 desc="DINO ViT-S/16 224px, 1,024,000-patch bank, d=384, k=30 (BASELINE configs[1])"),
    "cfg3": dict(N=10_240_000, d=768, S=37, ps=14, C=21, ignore=255, B=16,
                 desc="DINOv2 ViT-B/14 518px, 10,240,000-patch bank, d=768, k=30 (BASELINE configs[2])"),
    "cfg4": dict(N=10_240_000, d=1024, S=37, ps=14, C=151, ignore=0, B=16,
                 desc="DINOv2 ViT-L/14 518px, ADE20K-shaped (150 classes), 10,240,000-patch bank, d=1024, k=30 (BASELINE configs[3])"),
}
K_NEIGH, K_PRIME, BETA = 30, 64, 0.02
RING = 8  # distinct query batches cycled through (with the bank: inputs far larger than the 126 MB L2)
PATH = ("query prep -> tcgen05 bf16 GEMM + fused top-k' -> fp32 exact re-rank + label transfer -> "
        "fused tail (decode + upsample + argmax + confusion)")


def load_peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"tf": float(pk["bf16_tflops_sustained"]), "tf_burst": float(pk["bf16_tflops"]), "hbm": float(pk["hbm_gbs"]),
                "src": "measured (MEASURED_PEAKS.json: bf16_tflops_sustained, hbm_gbs)"}
    except Exception:
        return {"tf": 1400.0, "tf_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- synthetic inputs
def bank_slabs(w, row0, row1, device, slab_rows=1 << 19):
    """Yield (features (n, S*S, d), bank-side maps (n, H, H) uint8 with 255 -> 0, sel) covering the
    global bank rows [row0, row1): row r is patch r % S*S of training image r // S*S.  sel = int32
    slab-local row picks when the slab is not taken whole (shard boundaries inside an image)."""
    import torch

    import bench_synth as syn

    per_img = w["S"] ** 2
    protos = syn.prototypes(w, device)
    step_img = max(1, slab_rows // per_img)
    img = row0 // per_img
    while img * per_img < row1:
        n_img = min(step_img, -(-row1 // per_img) - img)
        feats, maps = syn.images(w, img, n_img, device, protos, stream=0)
        maps = torch.where(maps == 255, torch.zeros_like(maps), maps)  # hbird_eval.py:310
        lo, hi = max(row0, img * per_img), min(row1, (img + n_img) * per_img)
        sel = None
        if lo != img * per_img or hi != (img + n_img) * per_img:
            sel = torch.arange(lo - img * per_img, hi - img * per_img, device=device, dtype=torch.int32)
        yield feats, maps, sel, hi - lo
        img += n_img


def build_bank(w, row0, row1, device, keep_f32=True):
    from hbird_b200 import ops

    bank = ops.MemoryBank(w["d"], w["C"], w["ps"] ** 2, row1 - row0, device.index, keep_f32=keep_f32)
    for feats, maps, sel, _ in bank_slabs(w, row0, row1, device):
        bank.append(feats, maps, w["S"], w["ps"], sel)
    bank.finalize()
    return bank


def make_query_ring(w, device, first_batch=0, n=RING):
    """Validation batches first_batch .. first_batch+n-1: (features (B*S*S, d), y (B, 1, H, H) = id/255)."""
    import bench_synth as syn

    protos = syn.prototypes(w, device)
    ring = []
    for b in range(first_batch, first_batch + n):
        feats, maps = syn.images(w, b * w["B"], w["B"], device, protos, stream=1)
        y = (maps.float() / 255.0).unsqueeze(1).contiguous()  # loader contract: id/255
        ring.append((feats.view(-1, w["d"]).contiguous(), y))
    return ring


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def pause(self):
        """Stop sampling (the timed region is over); stop() then only reports."""
        self._stop.set()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- the hot path
def run_step(ops, bank, table, w, q, y, conf, shard=None):
    """One pass over one validation batch.  shard = None: the bank is whole on this GPU (4 launches).
    shard given: row-sharded bank — per-shard search, exchange, merge with the label transfer fused
    in, fused tail on this rank's image slice (5 launches; `table` is the replicated label table)."""
    B, S, H = w["B"], w["S"], w["S"] * w["ps"]
    if shard is None:
        bank.eval_step(q, y, S, conf, w["ignore"], K_NEIGH, K_PRIME, BETA)
        return
    from hbird_b200 import distributed as hdist

    b0, b1 = hdist.split_range(B, shard["world"], shard["rank"])
    n = S * S
    if shard.get("xchg") is not None:  # fused exchange: K2b scatters over NVLink, the merge waits
        qn = shard["xchg"].search_scatter(bank, q, shard["qsplit"], K_NEIGH, K_PRIME, shard["offset"])
        lh, _, _ = shard["xchg"].merge_transfer(table, w["ps"] ** 2, qn[b0 * n:b1 * n], BETA)
    else:
        scores, idx, qn = bank.search(q, K_NEIGH, K_PRIME, shard["offset"])
        gs, gi = hdist.all_gather_topk(scores, idx)
        lh, _, _ = ops.merge_topk_transfer(gs[:, b0 * n:b1 * n].contiguous(), gi[:, b0 * n:b1 * n].contiguous(), table,
                                           w["ps"] ** 2, qn[b0 * n:b1 * n].contiguous(), BETA)
    if b1 > b0:
        ops.predict_score(lh, b1 - b0, S, H, H, conf, y=y[b0:b1], ignore_index=w["ignore"])


def timed_loop(torch, dist, world, fn, steps, warmup, drain=None):
    """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks.
    drain: called after the last step inside the timed region (a pipelined step function still owes
    the post-processing of its last batch)."""
    for i in range(warmup):
        fn(i)
    if drain is not None:
        drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    if drain is not None:
        drain()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def measure_workload(torch, dist, ops, w, bank, table, ring, steps, warmup, world, peaks, shard=None, graph=False,
                     sampler=None):
    """Device-resident throughput of one workload + the search kernel's share and roofline fraction.
    The timed loop is the two-stream pipeline (hbird_b200.pipeline.EvalPipeline: K2 of batch i+1 on one
    stream, the HBM-bound post-processing of batch i on another, under it); the one-call-per-batch
    form (run_step) is timed next to it for comparison."""
    from hbird_b200 import distributed as hdist
    from hbird_b200 import pipeline as hpipe
    from hbird_b200.pipeline import EvalPipeline

    Q = w["B"] * w["S"] ** 2
    conf = torch.zeros((w["C"], w["C"]), dtype=torch.int64, device=ring[0][0].device)
    rank = 0 if shard is None else shard["rank"]
    b0, b1 = hdist.split_range(w["B"], world, rank) if shard is not None else (0, w["B"])
    pipe = EvalPipeline(bank, table, w["S"], conf, w["ignore"], K_NEIGH, K_PRIME, BETA,
                        0 if shard is None else shard["offset"], world if shard is not None else 1, rank,
                        None if shard is None else shard.get("xchg"))
    slices = [(q, y[b0:b1].contiguous()) for q, y in ring]
    # the engine's own policy: pipeline when the bank (shard) is small enough for the overlap to pay
    pipelined = hpipe.worthwhile(bank)

    def step(i):
        if pipelined:
            q, y = slices[i % len(slices)]
            pipe.submit(q, y, w["B"])
        else:
            run_step(ops, bank, table, w, *ring[i % len(ring)], conf, shard)

    for i in range(warmup):
        step(i)
    pipe.flush()
    torch.cuda.synchronize()
    bank.enable_kernel_timing(True)
    if sampler is not None:
        sampler.start()  # clocks are sampled during the timed region only
    ms = timed_loop(torch, dist, world, step, steps, 0, drain=pipe.flush) / steps
    if sampler is not None:
        sampler.pause()
    k2_ms, k2_n = bank.kernel_time_ms()
    k2b_under_ms, _ = bank.rerank_time_ms()
    # the same batches, one blocking call sequence per batch (no overlap between batches)
    bank.enable_kernel_timing(True)
    plain_steps = max(3, steps // 4)
    ms_plain = timed_loop(torch, dist, world, lambda i: run_step(ops, bank, table, w, *ring[i % len(ring)], conf, shard),
                          plain_steps, 2) / plain_steps
    k2b_ms, _ = bank.rerank_time_ms()
    bank.enable_kernel_timing(False)
    rows = bank.rows
    flop = 2.0 * rows * w["d"] * Q  # per GPU: this shard's rows x all queries
    tf = flop / (k2_ms * 1e-3) / 1e12 if k2_ms > 0 else None
    out = {"ms_per_step": ms, "pipelined": pipelined, "ms_per_step_unpipelined": ms_plain, "search_kernel_ms": k2_ms, "rerank_kernel_ms": k2b_ms,
           "rerank_kernel_ms_under_next_search": k2b_under_ms, "kernel_launches_timed": k2_n,
           "search_tflops": tf, "frac": (tf / peaks["tf"]) if tf else None,
           "kernel_share_of_step": (k2_ms / ms) if ms else None, "flop_per_launch": flop}
    if graph and shard is None:
        # the one-call step (4 launches) replayed from CUDA graphs, one per ring slot
        lh = torch.empty((Q, w["C"]), dtype=torch.float32, device=conf.device)
        graphs = []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for q, y in ring:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    bank.eval_step(q, y, w["S"], conf, w["ignore"], K_NEIGH, K_PRIME, BETA, label_hat=lh)
                graphs.append(g)
        torch.cuda.current_stream().wait_stream(side)
        out["ms_per_step_cuda_graph"] = timed_loop(torch, dist, world, lambda i: graphs[i % len(graphs)].replay(), steps, warmup) / steps
    return out, conf


# ----------------------------------------------------------------------------- CPU oracle legs
def host_threads():
    """All host cores for BLAS / OpenMP / torch, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1).  Call before numpy / torch are imported."""
    n = os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[k] = str(n)
    return n


def pin_host_threads(n):
    import torch

    torch.set_num_threads(n)
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=n)
    except Exception:
        pass
    used = n
    try:
        from threadpoolctl import threadpool_info

        blas = [t["num_threads"] for t in threadpool_info() if t.get("user_api") == "blas"]
        used = max(blas) if blas else n
    except Exception:
        pass
    return used


def cpu_sample_queries(w, seconds, cap):
    """How many queries a ~`seconds` CPU sample holds, assuming ~0.4 TFLOP/s of host GEMM with ~60 % on
    top for the exact top-k pass (measured on this pool's hosts in round 1); at least 32."""
    per_query = 2.0 * w["N"] * w["d"] * 1.6
    return int(max(32, min(cap, seconds * 4e11 / per_query)))


def cpu_sample(O, w, fm, lm, q, y_one, repeats=1):
    """The oracle port of the reference path on the host cores: exact fp32 IP search + neighbour
    gather + cross-attention for the n_q sampled queries against the FULL bank (the per-query
    stages, where all the time goes), plus the pixel stages (upsample, argmax, bincount) of one
    whole image.  Returns (seconds, (idx, dist, label_hat))."""
    import numpy as np

    S, C, H = w["S"], w["C"], w["S"] * w["ps"]
    t0 = time.perf_counter()
    for _ in range(repeats):
        idx, dist = O.search_exact_ip(q, fm, K_NEIGH)
        cpu_sample.search_seconds = (time.perf_counter() - t0) if repeats == 1 else None
        lh = O.transfer_labels(q[None], fm, lm, idx, BETA)[0]
        full = np.zeros((1, S * S, C), dtype=np.float32)
        full[0, :min(len(lh), S * S)] = lh[:S * S]
        pred = O.predict_map(full, S, H, H)
        O.confusion_matrix(O.decode_mask(y_one, False).reshape(-1), pred.reshape(-1), C, C, w["ignore"])
    return (time.perf_counter() - t0) / repeats, (idx, dist, lh)


def sample_queries(w, ring_host, n_q):
    """n_q queries spread evenly over the first validation batch (numpy) + the mask of its first image."""
    import numpy as np

    q_all, y_all = ring_host
    pick = np.linspace(0, q_all.shape[0] - 1, num=min(n_q, q_all.shape[0])).astype(np.int64)
    return pick, np.ascontiguousarray(q_all[pick]), y_all[:1]


def parity_block(torch, dist, ops, O, w, bank, table, ring, world, rank, offset, n_q, n_img, cpu_time=False, shard=None):
    """The timed inputs through the CUDA path against the oracle.
    (1) n_q queries sampled from the first validation batch: the CPU oracle's exact search runs on
        every rank's host against that rank's rows of the bank (exported from HBM) and the per-shard
        lists are merged with the oracle's IndexShards restatement (search_faiss.py:53-63) ->
        recall@30 and score error of the CUDA result.
    (2) n_img whole images: mIoU, confusion matrix and pixel agreement against the oracle's label
        transfer / upsample / argmax / bincount (numpy).  The oracle's exact search would take ~35 s
        per 518-px image on the host, so for these images the exact fp32 neighbours are computed with
        torch.matmul (TF32 off) on the GPUs — and that stand-in is itself checked against the CPU
        oracle on the queries of (1).
    Returns (dict or None on ranks > 0, seconds of CPU oracle time of (1) on this rank)."""
    import numpy as np

    from hbird_b200 import distributed as hdist
    from hbird_b200.utils.eval_metrics import miou_from_confusion

    dev = ring[0][0].device
    S, d, C, H, k = w["S"], w["d"], w["C"], w["S"] * w["ps"], K_NEIGH
    per = S * S
    q0, y0 = ring[0]
    q_host, y_host = q0.cpu().numpy(), y0.cpu().numpy()
    pick, qs, y_one = sample_queries(w, (q_host, y_host), n_q)
    # this rank's bank rows on the host, in slabs
    rows = bank.rows
    fm = np.empty((rows, d), dtype=np.float32)
    for a in range(0, rows, 1 << 20):
        m = min(1 << 20, rows - a)
        fm[a:a + m] = bank.export(a, m, labels=False)[0].cpu().numpy()
    t0 = time.perf_counter()
    li, ld = O.search_exact_ip(qs, fm, min(k, rows))
    cpu_s = time.perf_counter() - t0
    li = li + offset

    def exact_gpu(q_dev):
        """exact fp32 top-k of q_dev against this rank's rows (torch.matmul, TF32 off), global ids"""
        bs, bi = None, None
        for a in range(0, rows, 1 << 18):
            m = min(1 << 18, rows - a)
            f = bank.export(a, m, labels=False)[0]
            sc = q_dev @ f.T
            v, i = sc.topk(min(k, m), dim=1)
            i = i + (a + offset)
            bs, bi = (v, i) if bs is None else (torch.cat([bs, v], 1), torch.cat([bi, i], 1))
            if bs.shape[1] > k:
                v, j = bs.topk(k, dim=1)
                bs, bi = v, bi.gather(1, j)
        order = torch.argsort(bs, dim=1, descending=True, stable=True)
        return bs.gather(1, order), bi.gather(1, order)

    n_img = min(n_img, w["B"])
    qi_dev = q0[:n_img * per].contiguous()
    gs, gi = exact_gpu(torch.cat([q0[torch.from_numpy(pick).to(dev)], qi_dev]))
    if world > 1:
        kk = gs.shape[1]
        pad = k - kk
        if pad:
            gs = torch.cat([gs, torch.full((gs.shape[0], pad), float("-inf"), device=dev)], 1)
            gi = torch.cat([gi, torch.full((gi.shape[0], pad), -1, device=dev, dtype=torch.int64)], 1)
            li = np.concatenate([li, np.full((li.shape[0], pad), -1, np.int64)], 1)
            ld = np.concatenate([ld, np.full((ld.shape[0], pad), -np.inf, np.float32)], 1)
        ags, agi = hdist.all_gather_topk(gs, gi)
        als, ali = hdist.all_gather_topk(torch.from_numpy(ld).to(dev), torch.from_numpy(li).to(dev))
        mi, md = O.merge_shards(agi.cpu().numpy(), ags.cpu().numpy(), k)
        oi, od = O.merge_shards(ali.cpu().numpy(), als.cpu().numpy(), k)
    else:
        mi, md = gi.cpu().numpy(), gs.cpu().numpy()
        oi, od = li, ld
    # the product path on the same queries
    conf = torch.zeros((C, C), dtype=torch.int64, device=dev)
    if world == 1:
        lh, qn, s_all, i_all = bank.search_transfer(q0, k, K_PRIME, 0, BETA, None, True)
        got_s, got_i = s_all[torch.from_numpy(pick).to(dev)].cpu().numpy(), i_all[torch.from_numpy(pick).to(dev)].cpu().numpy()
        pred = ops.predict_score(lh[:n_img * per].contiguous(), n_img, S, H, H, conf, y=y0[:n_img], ignore_index=w["ignore"],
                                 return_pred=True)
        lh_img = lh[:n_img * per]
    elif shard is not None and shard.get("xchg") is not None:
        # row-sharded, THE TIMED PATH: per-shard K2, shard exchange over NVLink (threshold exchange or full
        # re-rank, as configured), merge with the label transfer fused in on the rank that owns the query;
        # the per-rank slices are then collected for the comparison
        xc, qsp = shard["xchg"], shard["qsplit"]
        qn = xc.search_scatter(bank, q0, qsp, k, K_PRIME, offset)
        a, b = qsp[rank], qsp[rank + 1]
        lh_s, s_s, i_s = xc.merge_transfer(table, w["ps"] ** 2, qn[a:b].contiguous(), BETA, True)
        xc.check_status()
        cnt = [qsp[r + 1] - qsp[r] for r in range(world)]
        lh, s_all, i_all = (hdist.all_gather_rows(t, cnt) for t in (lh_s, s_s, i_s))
        got_s, got_i = s_all[torch.from_numpy(pick).to(dev)].cpu().numpy(), i_all[torch.from_numpy(pick).to(dev)].cpu().numpy()
        pred = ops.predict_score(lh[:n_img * per].contiguous(), n_img, S, H, H, conf, y=y0[:n_img], ignore_index=w["ignore"],
                                 return_pred=True)
        lh_img = lh[:n_img * per]
    else:
        # row-sharded without peer memory: per-shard K2/K2b, NCCL all-gather, merge with the label transfer fused in
        s, i, qn = bank.search(q0, k, K_PRIME, offset)
        ags2, agi2 = hdist.all_gather_topk(s, i)
        lh, s_all, i_all = ops.merge_topk_transfer(ags2, agi2, table, w["ps"] ** 2, qn, BETA, True)
        got_s, got_i = s_all[torch.from_numpy(pick).to(dev)].cpu().numpy(), i_all[torch.from_numpy(pick).to(dev)].cpu().numpy()
        pred = ops.predict_score(lh[:n_img * per].contiguous(), n_img, S, H, H, conf, y=y0[:n_img], ignore_index=w["ignore"],
                                 return_pred=True)
        lh_img = lh[:n_img * per]
    # neighbour rows for the oracle's cross-attention (it re-normalises the gathered keys, :594-609)
    nb = torch.from_numpy(mi[len(pick):]).to(dev).reshape(-1)
    local = nb - offset
    mine = (local >= 0) & (local < rows)
    kf = torch.zeros((nb.numel(), d), dtype=torch.float32, device=dev)
    sel = local[mine]
    for a in range(0, rows, 1 << 20):
        m = min(1 << 20, rows - a)
        inside = (sel >= a) & (sel < a + m)
        if bool(inside.any()):
            f = bank.export(a, m, labels=False)[0]
            where = mine.nonzero().squeeze(1)[inside]
            kf[where] = f[sel[inside] - a]
    if world > 1:
        dist.all_reduce(kf)
    if rank != 0:
        return None, cpu_s
    tbl = torch.as_tensor(table)
    kl = (tbl[nb.clamp_min(0)].to(torch.float32) / float(w["ps"] ** 2)).cpu().numpy()
    ref_lh = O.cross_attention(qi_dev.cpu().numpy()[None], kf.cpu().numpy().reshape(1, -1, k, d), kl.reshape(1, -1, k, C), BETA)[0]
    ref_pred = O.predict_map(ref_lh.reshape(n_img, per, C), S, H, H)[:, 0]
    gt = O.decode_mask(y_host[:n_img], False).reshape(n_img, H, H)
    ref_conf = O.confusion_matrix(gt.reshape(-1), ref_pred.reshape(-1), C, C, w["ignore"])
    ref_miou = O.miou_from_confusion(ref_conf)[0]
    got_conf = conf.cpu().numpy()
    miou = miou_from_confusion(got_conf)[0]
    got_pred = pred.cpu().numpy()
    same = got_pred == ref_pred
    # confusion matrix restricted to the pixels whose predicted label matches: must be bit-exact
    conf_same_ref = O.confusion_matrix(gt[same], ref_pred[same], C, C, w["ignore"])
    conf_same_got = O.confusion_matrix(gt[same], got_pred[same], C, C, w["ignore"])

    def rec(a, b):
        return float((a[:, :, None] == b[:, None, :]).any(axis=2).mean())

    out = {
        "queries_vs_cpu_oracle": int(len(pick)), "recall_at_30": rec(got_i, oi),
        "score_max_rel_err": float((np.abs(got_s - od) / np.maximum(np.abs(od), 1e-6)).max()),
        "gpu_exact_standin_vs_cpu_oracle_recall": rec(mi[:len(pick)], oi),
        "gpu_exact_standin_score_max_rel_err": float((np.abs(md[:len(pick)] - od) / np.maximum(np.abs(od), 1e-6)).max()),
        "images": int(n_img), "miou_b200": miou, "miou_oracle": ref_miou, "miou_delta_points": abs(miou - ref_miou) * 100.0,
        "label_hat_max_abs_err": float(np.abs(lh_img.cpu().numpy() - ref_lh).max()),
        "pred_pixel_agreement": float(same.mean()), "confusion_abs_diff": int(np.abs(got_conf - ref_conf).sum()),
        "confusion_bit_exact_where_labels_match": bool(np.array_equal(conf_same_ref, conf_same_got)),
        "recall_at_30_whole_images": rec(i_all[:n_img * per].cpu().numpy(), mi[len(pick):]),
        "oracle": "numpy port of the reference path; whole-image neighbours by fp32 torch.matmul (TF32 off), checked "
                  "against the CPU oracle on the sampled queries",
        "gates": "recall>=0.999, rel<=1e-3, |dmIoU|<=0.05 points, confusion bit-exact where labels match",
    }
    out["ok"] = bool(out["recall_at_30"] >= 0.999 and out["score_max_rel_err"] <= 1e-3 and out["miou_delta_points"] <= 0.05
                     and out["confusion_bit_exact_where_labels_match"])
    return out, cpu_s


# ----------------------------------------------------------------------------- HBM-bound kernels
def hbm_kernels(torch, ops, w, bank, table, ring, peaks, rerank_ms):
    """K1 pack, K2b+K4a (from the timed steps), fused tail: algorithmic GB/s vs the measured HBM copy
    bandwidth.  Timed alone with an L2 flush (256 MB memset) between iterations."""
    dev = ring[0][0].device
    S, ps, C, d, B = w["S"], w["ps"], w["C"], w["d"], w["B"]
    H, Q = S * ps, B * S * S
    dpad = (d + 63) // 64 * 64
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    out = {}
    feats, maps, _, n_rows = next(bank_slabs(w, 0, 1 << 19, dev))
    tmp = ops.MemoryBank(d, C, ps * ps, n_rows * 14, dev.index, keep_f32=True)
    ms = timeit(lambda: tmp.append(feats, maps, S, ps))
    by = n_rows * (4 * d + ps * ps + 2 * dpad + 4 * d + 2 * C)
    out["K1_pack"] = {"ms": ms, "rows": n_rows, "bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peaks["hbm"]}
    tmp.close()
    by = Q * (K_PRIME * 4 * d + K_NEIGH * 2 * C + 4 * C)
    q, y = ring[0]
    if rerank_ms is None:
        # row-sharded run: the timed steps re-rank through the exchange (threshold exchange: a shortlist kernel
        # and a re-rank of the survivors), so the full K2b + K4a on this shard is timed here, in one-call searches
        bank.enable_kernel_timing(True)
        for i in range(6):
            bank.search_transfer(ring[i % len(ring)][0], K_NEIGH, K_PRIME, 0, BETA, table)
        torch.cuda.synchronize()
        rerank_ms, _ = bank.rerank_time_ms()
        bank.enable_kernel_timing(False)
        timed = "one-call searches on this GPU's shard, after the timed steps"
    else:
        timed = "inside the timed steps (one-call form)"
    if rerank_ms:
        out["K2b_rerank_plus_K4a_label_transfer"] = {"ms": rerank_ms, "queries": Q, "bytes": by, "gbs": by / rerank_ms / 1e6,
                                                     "frac": by / rerank_ms / 1e6 / peaks["hbm"], "timed": timed}
    lh, _, _, _ = bank.search_transfer(q, K_NEIGH, K_PRIME, 0, BETA, table)
    conf = torch.zeros((C, C), dtype=torch.int64, device=dev)
    ms = timeit(lambda: ops.predict_score(lh, B, S, H, H, conf, y=y, ignore_index=w["ignore"]))
    by = B * (4 * S * S * C + 4 * H * H)
    out["tail_decode_upsample_argmax_confusion"] = {
        "ms": ms, "pixels": B * H * H, "bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peaks["hbm"],
        "note": "instruction-bound (6 instructions per class per pixel in torch's rounding order), not HBM-bound"}
    return out


def _claim_stdout():
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr and return a
    file object on the real stdout, so that the only thing on stdout is the final JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


# ----------------------------------------------------------------------------- reference arm
def reference_arm(args, w, real_stdout):
    """The CPU oracle port of the reference path on the host cores, on a bounded sample of the
    workload.  The bank is the same synthetic bank as the B200 arm's (bench_synth is a pure function
    of the row number), generated and normalised on the host; no product library is loaded."""
    n_threads = host_threads()
    import numpy as np
    import torch

    import bench_synth as syn
    from oracle import hbird_oracle as O

    threads = pin_host_threads(n_threads)
    cpu = torch.device("cpu")
    N, d, C = w["N"], w["d"], w["C"]
    fm = np.empty((N, d), dtype=np.float32)
    lm = np.empty((N, C), dtype=np.float32)
    r = 0
    for feats, maps, sel, n in bank_slabs(w, 0, N, cpu, slab_rows=1 << 18):
        f = feats.view(-1, d)
        l = syn.soft_labels(w, maps)  # == oracle.build_memory's one_hot().mean() (tests/test_bench_contract.py)
        if sel is not None:
            f, l = f[sel.long()], l[sel.long()]
        fm[r:r + n] = O.normalise_rows(f.numpy())
        lm[r:r + n] = l.numpy()
        r += n
    ring = make_query_ring(w, cpu, 0, 1)
    q_host, y_host = ring[0][0].numpy(), ring[0][1].numpy()
    n_q = cpu_sample_queries(w, seconds=4.0, cap=q_host.shape[0])  # ~4 s of host work per step
    _, qs, y_one = sample_queries(w, (q_host, y_host), n_q)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_sample(O, w, fm, lm, qs, y_one)
        if i >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = len(qs) * len(times) / tot
    sample = (f"{len(qs)} patch-queries per step (spread over the first validation batch) against the full {N:,}-row bank: "
              f"exact fp32 IP search + gather + cross-attention, plus the pixel stages of one image")
    line = {
        "impl": "reference", "metric": "patch_queries_per_sec", "value": value, "unit": "patch-queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, len(times)),
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "bank_rows": N, "d": d, "k": K_NEIGH, "classes": C,
                   "path": "oracle port of the reference CPU path (exact fp32 IP search + gather + cross-attention + "
                           "bilinear upsample + argmax + bincount), numpy/BLAS",
                   "parallelism": f"host CPU, {threads} threads"},
        "cpu_baseline": {"value": value, "unit": "patch-queries/s", "cores": threads, "kind": "port", "sample": sample,
                         "search_only_value": (len(qs) / cpu_sample.search_seconds) if cpu_sample.search_seconds else None},
        "e2e": {"value": value, "unit": "patch-queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=real_stdout, flush=True)
    return 0


# ----------------------------------------------------------------------------- main
def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no by_workload block)")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep equal row shards (no speed balancing)")
    ap.add_argument("--exchange", default="threshold", choices=["threshold", "full"],
                    help="N > 1: shards re-rank only candidates above the cross-shard bound (default), or their whole top-k'")
    ap.add_argument("--extras", default="", help="comma list of by_workload entries to run (default: all)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = dict(WORKLOADS[args.workload])
    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, w, real_stdout)

    n_threads = host_threads()  # the CPU oracle legs use every host core, also under torchrun
    import numpy as np
    import torch
    import torch.distributed as dist

    warmup = max(args.warmup, 3)
    Q = w["B"] * w["S"] ** 2
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the hot path has no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    torch.backends.cuda.matmul.allow_tf32 = False
    from hbird_b200 import HbirdEvaluation, ops
    from hbird_b200 import distributed as hdist
    from hbird_b200.models import FeatureExtractorSimple
    from oracle import hbird_oracle as O  # the checker: parity block and cpu_baseline leg only

    threads = pin_host_threads(n_threads)
    ops.device_check(device.index)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()
    extras = [e for e in args.extras.split(",") if e]

    def want(name):
        return not args.no_extras and (not extras or name in extras)

    # ------------------------------------------------------------------ headline workload
    a, b = hdist.shard_bounds(w["N"], world, rank)
    bank = build_bank(w, a, b, device)
    ring = make_query_ring(w, device)  # every rank sees every validation batch
    shard = None
    table = bank.label_table()
    xchg = None
    collective = None
    balance = None
    shares = None  # fraction of a bank's rows each rank holds (None = equal)
    if world > 1:
        counts = hdist.gather_counts(bank.rows, device)
        table = hdist.all_gather_rows(bank.label_table(), counts)
        shard = {"offset": hdist.offsets_from_counts(counts)[rank], "world": world, "rank": rank}
        per_rank = -(-w["B"] // world) * w["S"] ** 2
        xchg = hdist.connect_shard_exchange(per_rank, K_NEIGH, device, threshold_exchange=args.exchange == "threshold")
        collective = "NCCL all-gather of (score f32, idx i64)[Q,k] + k-way merge kernel with the label transfer fused in"
        if xchg is not None:
            shard["xchg"], shard["qsplit"] = xchg, hdist.query_split(w["B"], w["S"] ** 2, world)
            collective = ("fused exchange: K2b stores each query's shard top-k into the owner rank's window over NVLink "
                          "(CUDA IPC peer memory), a one-warp wait kernel polls the per-rank step flags, the merge + "
                          "label-transfer kernel follows; no NCCL call")
            if args.exchange == "threshold":
                collective = ("threshold exchange over NVLink peer memory (CUDA IPC), no NCCL call: a shortlist kernel merges each "
                              "query's candidate lists and stores four order statistics of its bf16 top-k' into every rank's "
                              "window; after a one-warp wait every shard derives the same bound of the global k'-th best score "
                              "and gathers fp32 rows only for its candidates above it (instead of k' per query and shard); K2b "
                              "stores the shard top-k into the owner rank's window; wait; merge + label-transfer kernel")
        # Shards sized by measured search speed, as the engine does (HbirdEvaluation._balance_shards): every
        # step ends with everybody's shard results, so the job runs at the pace of the slowest GPU.  The
        # speed figure is each rank's mean K2 time over a full-length run of the SAME steps on equal shards
        # (sustained clocks); that run is reported next to the headline (`ms_per_step_equal_shards`).
        eq, _ = measure_workload(torch, dist, ops, w, bank, table, ring, args.steps, warmup, world, peaks, shard)
        times = hdist.gather_floats(eq["search_kernel_ms"], device)
        new_counts = hdist.balanced_counts(counts, times)
        balance = {"search_ms_per_rank_equal_shards": times, "rows_per_gpu_equal": counts, "rows_per_gpu": counts,
                   "ms_per_step_equal_shards": eq["ms_per_step"], "enabled": not args.no_balance}
        if not args.no_balance and max(abs(n - o) / o for n, o in zip(new_counts, counts)) >= 0.01:
            bank.close()
            off = hdist.offsets_from_counts(new_counts)
            a, b = off[rank], off[rank] + new_counts[rank]
            bank = build_bank(w, a, b, device)  # global row r has the same content wherever it lives
            counts = new_counts
            shard["offset"] = off[rank]
            balance["rows_per_gpu"] = new_counts
            shares = [c / w["N"] for c in new_counts]
    torch.cuda.synchronize()

    sampler = ClockSampler(device.index)
    head, conf = measure_workload(torch, dist, ops, w, bank, table, ring, args.steps, warmup, world, peaks, shard,
                                  sampler=sampler)
    clocks = sampler.stop()
    ms_per_step = head["ms_per_step"]
    value = Q / (ms_per_step * 1e-3)  # world == 1: this GPU; world > 1: all ranks work on the same Q queries
    xchg_used = xchg is not None
    # world == 1: prep, K2, K2b+K4a, tail.  Sharded: prep, K2, K2b-scatter, wait, merge+K4a, tail (6); with the
    # threshold exchange the K2b is a shortlist kernel + a wait + the re-rank of the survivors (8); NCCL path: 6
    launches_per_step = 4 if world == 1 else (8 if (xchg_used and args.exchange == "threshold") else 6)

    nccl_ms = None
    if world > 1 and xchg is not None:
        plain = dict(shard)
        plain.pop("xchg")
        m2, conf_nccl = measure_workload(torch, dist, ops, w, bank, table, ring, max(4, args.steps // 4), warmup, world, peaks, plain)
        nccl_ms = m2["ms_per_step"]

    # ------------------------------------------------------------------ e2e: the engine's public call, host inputs
    fe = FeatureExtractorSimple(torch.nn.Identity(), lambda m, x: (x, None), w["S"], w["d"])
    nn_params = {"k_prime": K_PRIME, "idx_shard": world > 1}
    ev = HbirdEvaluation.from_bank(fe, bank, w["C"], K_NEIGH, device, nn_params,
                                   shard_counts=(counts if world > 1 else None))
    if world > 1 and xchg is not None:
        ev._xchg = xchg  # reuse the connected windows
    host_ring = [(q.view(w["B"], w["S"] ** 2, w["d"]).cpu().pin_memory(), y.cpu().pin_memory()) for q, y in ring]
    last = {}

    def step_e2e(i):
        last["miou"] = ev.evaluate([host_ring[i % RING]], w["S"], ignore_index=w["ignore"])

    ms_e2e = timed_loop(torch, dist, world, step_e2e, args.steps, warmup) / args.steps
    h2d = host_ring[0][0].numel() * 4 + host_ring[0][1].numel() * 4
    if world > 1:
        h2d = h2d // world  # each rank copies its image slice only; the queries then travel over NVLink
    d2h = w["C"] * w["C"] * 8
    e2e_value = Q / (ms_e2e * 1e-3)
    ev._xchg = None

    # ---- the reference's literal plugin call (search_faiss.py:83-90): host queries in, host (indices,
    # distances) out -- what third-party code using the ABC gets
    plugin = None
    if world == 1:
        from hbird_b200 import NearestNeighborSearchB200

        nn = NearestNeighborSearchB200(None, n_neighbors=K_NEIGH, bank=bank, k_prime=K_PRIME, gpu_ids=[device.index])
        q_host = [q.cpu() for q, _ in ring[:2]]
        nn.find_nearest_neighbors(q_host[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_calls = max(3, min(10, args.steps))
        for i in range(n_calls):
            idx_np, dist_np = nn.find_nearest_neighbors(q_host[i % 2])
        dt = (time.perf_counter() - t0) / n_calls
        plugin = {"value": Q / dt, "unit": "patch-queries/s", "ms_per_call": dt * 1e3, "calls": n_calls,
                  "call": "NearestNeighborSearchB200.find_nearest_neighbors(q_cpu) -> (indices, distances) ndarrays",
                  "h2d_bytes_per_call": Q * w["d"] * 4, "d2h_bytes_per_call": int(idx_np.nbytes + dist_np.nbytes)}

    # ------------------------------------------------------------------ parity vs the oracle (+ CPU baseline at N = 1)
    par_q = 256 if w["N"] > 2_000_000 else 1024
    parity, _ = parity_block(torch, dist, ops, O, w, bank, table, ring, world, rank, 0 if shard is None else shard["offset"],
                             n_q=par_q, n_img=2, shard=shard)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rows = bank.rows
        fm = np.empty((rows, w["d"]), dtype=np.float32)
        lm = np.empty((rows, w["C"]), dtype=np.float32)
        for r0 in range(0, rows, 1 << 20):
            m = min(1 << 20, rows - r0)
            f, l = bank.export(r0, m)
            fm[r0:r0 + m], lm[r0:r0 + m] = f.cpu().numpy(), l.cpu().numpy()
        q_host_np, y_host_np = ring[0][0].cpu().numpy(), ring[0][1].cpu().numpy()
        n_q = cpu_sample_queries(w, seconds=15.0, cap=Q)
        _, qs, y_one = sample_queries(w, (q_host_np, y_host_np), n_q)
        dt, _ = cpu_sample(O, w, fm, lm, qs, y_one)
        cpu = {"value": len(qs) / dt, "unit": "patch-queries/s", "cores": threads, "kind": "port",
               "search_only_value": (len(qs) / cpu_sample.search_seconds) if cpu_sample.search_seconds else None,
               "sample": f"{len(qs)} patch-queries (spread over the first validation batch) against the full {rows:,}-row bank: "
                         f"exact fp32 IP search + gather + cross-attention, plus the pixel stages of one image; "
                         f"{dt:.1f} s of oracle (numpy/BLAS) time"}
        del fm, lm

    hbm = hbm_kernels(torch, ops, w, bank, table, ring, peaks, head["rerank_kernel_ms"] if world == 1 else None) if rank == 0 else None
    if world > 1:
        dist.barrier()

    # ------------------------------------------------------------------ the other configurations, same line
    by_workload = {}
    if xchg is not None:
        torch.cuda.synchronize()
        dist.barrier()
        hdist.close_shard_exchange(xchg)
        xchg = None
    ev.bank = None  # the evaluator does not own the bench's bank
    bank.close()
    del bank, table, ring, host_ring
    torch.cuda.empty_cache()
    few = max(5, args.steps // 2)

    def side_workload(name, ww, sharded, keep_f32=True, graph=False, search_only=False):
        """value / kernel fraction of another configuration (single GPU, or row-sharded over the ranks)"""
        if sharded and world > 1 and shares is not None:  # same relative shard sizes as the headline's calibration
            cum = [0.0]
            for sh in shares:
                cum.append(cum[-1] + sh)
            a2, b2 = int(round(ww["N"] * cum[rank])), (ww["N"] if rank == world - 1 else int(round(ww["N"] * cum[rank + 1])))
        else:
            a2, b2 = hdist.shard_bounds(ww["N"], world, rank) if sharded else (0, ww["N"])
        bk = build_bank(ww, a2, b2, device, keep_f32)
        rg = make_query_ring(ww, device, 0, 4)
        tb, sh, xc = bk.label_table(), None, None
        if sharded and world > 1:
            cts = hdist.gather_counts(bk.rows, device)
            tb = hdist.all_gather_rows(bk.label_table(), cts)
            sh = {"offset": hdist.offsets_from_counts(cts)[rank], "world": world, "rank": rank}
            xc = hdist.connect_shard_exchange(-(-ww["B"] // world) * ww["S"] ** 2, K_NEIGH, device,
                                              threshold_exchange=args.exchange == "threshold")
            if xc is not None:
                sh["xchg"], sh["qsplit"] = xc, hdist.query_split(ww["B"], ww["S"] ** 2, world)
        m, _ = measure_workload(torch, dist, ops, ww, bk, tb, rg, few, 3, world, peaks, sh, graph=graph)
        q_step = ww["B"] * ww["S"] ** 2
        per_step = q_step * (1 if sharded or world == 1 else world)  # replicas: every rank its own batch
        res = {"workload": ww["desc"], "bank_rows": ww["N"], "bank_rows_per_gpu": bk.rows, "d": ww["d"],
               "layout": ("row-sharded, fused exchange" if (sharded and world > 1 and xc is not None) else
                          "row-sharded, NCCL all-gather + merge" if (sharded and world > 1) else
                          "single GPU" if world == 1 else "replicas (full bank per GPU, batches split)"),
               "value": per_step / (m["ms_per_step"] * 1e-3), "unit": "patch-queries/s", "ms_per_step": m["ms_per_step"],
               "search_kernel_ms": m["search_kernel_ms"], "search_kernel_frac_of_sustained_bf16": m["frac"],
               "search_kernel_tflops": m["search_tflops"], "scaling": "strong" if sharded and world > 1 else "weak",
               "pipelined": m["pipelined"], "ms_per_step_unpipelined": m["ms_per_step_unpipelined"]}
        if "ms_per_step_cuda_graph" in m:
            res["ms_per_step_cuda_graph"] = m["ms_per_step_cuda_graph"]
            res["value_cuda_graph"] = per_step / (m["ms_per_step_cuda_graph"] * 1e-3)
        if xc is not None:
            torch.cuda.synchronize()
            dist.barrier()
            hdist.close_shard_exchange(xc)
        bk.close()
        torch.cuda.empty_cache()
        by_workload[name] = res

    if world == 1:
        if want("cfg1"):
            side_workload("cfg1", dict(WORKLOADS["cfg1"]), False, graph=True)
        if want("cfg2") and args.workload != "cfg2":
            side_workload("cfg2", dict(WORKLOADS["cfg2"]), False, graph=True)
        if want("cfg4") and args.workload != "cfg4":
            side_workload("cfg4", dict(WORKLOADS["cfg4"]), False)
        for n_rows in (100_000, 1_000_000):  # cfg5 sweep, d = 768; the 1e7 point is the headline (cfg3)
            if want("cfg5"):
                ww = dict(WORKLOADS["cfg3"], N=n_rows, desc=f"bank-size sweep (BASELINE configs[4]): {n_rows:,} rows, d=768, k=30")
                side_workload(f"cfg5_{n_rows:.0e}".replace("+0", ""), ww, False)
    else:
        if want("cfg3_replicas") and args.workload == "cfg3":
            side_workload("cfg3_replicas", dict(WORKLOADS["cfg3"]), False)
        if want("cfg2"):
            side_workload("cfg2_sharded", dict(WORKLOADS["cfg2"]), True)
        if want("cfg4") and args.workload != "cfg4":
            side_workload("cfg4_sharded", dict(WORKLOADS["cfg4"]), True)
        if want("cfg5"):
            for n_rows in (1_000_000, 100_000_000):
                if n_rows // world > 30_000_000:
                    continue  # 1e8 rows need >= 4 GPUs with the fp32 re-rank copy resident
                ww = dict(WORKLOADS["cfg3"], N=n_rows, desc=f"bank-size sweep (BASELINE configs[4]): {n_rows:,} rows, d=768, k=30")
                side_workload(f"cfg5_{n_rows:.0e}".replace("+0", "") + "_sharded", ww, True)

    if rank == 0:
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json"))).get(args.workload)
        except Exception:
            pass
        per_gpu_rows = b - a
        line = {
            "metric": "patch_queries_per_sec", "value": value, "unit": "patch-queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: {w['desc']}", "bank_rows": w["N"], "bank_rows_per_gpu": per_gpu_rows,
                "d": w["d"], "k": K_NEIGH, "k_prime": K_PRIME, "queries_per_step": Q, "classes": w["C"],
                "parallelism": "single GPU" if world == 1 else
                               f"bank row-sharded over {world} GPUs ({per_gpu_rows:,} rows on rank 0; shards sized by measured "
                               f"search speed, +-10 % of equal); every rank searches all {Q} queries of a step, exchange, each "
                               f"rank post-processes its image slice",
                "collective": collective,
                "path": PATH,
                "pipelining": ("two streams: K2 of batch i+1 (high priority) issued beside the re-rank / exchange / merge / tail "
                               "of batch i, which fill the search kernel's ramp-down and the launch gaps; every batch is "
                               "complete inside the timed region") if head["pipelined"] else
                              "none: one blocking call sequence per batch (the engine pipelines banks of <= 8.4M rows per GPU only)",
                "l2": f"inputs larger than L2: bf16 bank shard {per_gpu_rows * w['d'] * 2 / 1e6:.0f} MB streamed every step, "
                      f"{RING} distinct query batches cycled",
            },
            "roofline": {"bound": "tensor", "achieved": head["search_tflops"], "peak": peaks["tf"], "unit": "TFLOP/s",
                         "frac": head["frac"], "traffic": traffic,
                         "frac_of_burst_peak": (head["search_tflops"] / peaks["tf_burst"]) if head["search_tflops"] else None,
                         "frac_of_nominal_2250": (head["search_tflops"] / 2250.0) if head["search_tflops"] else None,
                         "kernel": "search_topk_kernel (tcgen05 GEMM + fused top-k')", "kernel_ms": head["search_kernel_ms"],
                         "kernel_launches_timed": head["kernel_launches_timed"], "flop_per_launch": head["flop_per_launch"],
                         "algorithmic_flop": "2 * bank rows on this GPU * d per query (SURVEY.md 8d)",
                         "peak_source": peaks["src"], "kernel_share_of_step": head["kernel_share_of_step"],
                         "ms_per_step_unpipelined": head["ms_per_step_unpipelined"]},
            "roofline_hbm": hbm,
            "cpu_baseline": cpu,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": "patch-queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": "HbirdEvaluation.evaluate([one batch of pinned host (features, masks)]) -> mIoU float, once per step",
                    "miou_last_step": last.get("miou")},
            "plugin_call": plugin,
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "by_workload": by_workload,
        }
        if world > 1:
            line["sharded"] = {"value": value, "unit": "patch-queries/s", "ms_per_step": ms_per_step, "scaling": "strong",
                               "bank_rows_total": w["N"], "bank_rows_per_gpu": per_gpu_rows, "collective": collective,
                               "exchange": (args.exchange if xchg_used else "nccl"),
                               "nccl_path_ms_per_step": nccl_ms, "parity": parity, "balance": balance}
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
