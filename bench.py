"""bench.py — patch-queries/sec of the dense nearest-neighbour evaluation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one validation batch of synthetic VOC-shaped input:
mask decode -> kNN search against the HBM-resident memory bank (tcgen05 GEMM + fused top-k', exact
fp32 re-rank) -> soft label transfer -> bilinear upsample + argmax -> confusion-matrix update.
Default workload = BASELINE.json configs[1] ("cfg2": DINO ViT-S/16 224 px, 1,024,000-patch bank,
d=384, k=30, 64 images = 12,544 patch-queries per step).

N > 1: one process per GPU.  Primary number = query-parallel replicas (the reference's default
faiss.IndexReplicas layout, search_faiss.py:65-74: full bank on every GPU, each rank evaluates its
own batches, no data-path collective) -> weak scaling.  The row-sharded bank (north star item 3:
per-shard search -> NCCL all-gather -> k-way merge kernel) is timed in the same run and reported
under "sharded".

`--impl reference` times the CPU oracle port of the reference path (numpy/BLAS on all host cores)
on a bounded sample of the same workload; under torchrun only rank 0 runs it.
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
                 desc="DINO ViT-S/16 224px, 102,400-patch bank, d=384, k=30"),
    "cfg2": dict(N=1_024_000, d=384, S=14, ps=16, C=21, ignore=255, B=64,
                 desc="DINO ViT-S/16 224px, 1,024,000-patch bank, d=384, k=30 (BASELINE configs[1])"),
    "cfg3": dict(N=10_240_000, d=768, S=37, ps=14, C=21, ignore=255, B=16,
                 desc="DINOv2 ViT-B/14 518px, 10,240,000-patch bank, d=768, k=30 (BASELINE configs[2])"),
    "cfg4": dict(N=10_240_000, d=1024, S=37, ps=14, C=151, ignore=0, B=16,
                 desc="DINOv2 ViT-L/14 518px, ADE20K-shaped, 10,240,000-patch bank, d=1024, k=30"),
}
K_NEIGH, K_PRIME, BETA = 30, 64, 0.02
RING = 8  # distinct query batches cycled through (with the bank: inputs far larger than the 126 MB L2)


def load_peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md sustained)"


# ----------------------------------------------------------------------------- synthetic inputs
def synth_images(w, n_img, gen, device):
    """(features (n, S*S, d) fp32 raw, masks (n, H, H) uint8) generated on the device: class-prototype
    features + noise, un-normalised; blocky label maps with ~2 % ignore pixels."""
    import torch

    S, ps, C, d = w["S"], w["ps"], w["C"], w["d"]
    H = S * ps
    first = 1 if w["ignore"] == 0 else 0
    cells = 8
    coarse = torch.randint(first, C, (n_img, cells, cells), generator=gen, device=device)
    reps = (H + cells - 1) // cells
    maps = coarse.repeat_interleave(reps, 1).repeat_interleave(reps, 2)[:, :H, :H]
    ign = torch.rand((n_img, H, H), generator=gen, device=device) < 0.02
    maps = torch.where(ign, torch.full_like(maps, w["ignore"]), maps).to(torch.uint8)
    g0 = torch.Generator(device=device).manual_seed(0)
    protos = torch.randn((C, d), generator=g0, device=device)
    centre = maps[:, ps // 2::ps, ps // 2::ps].reshape(n_img, S * S).long().clamp_max(C - 1)
    feats = protos[centre] + 0.8 * torch.randn((n_img, S * S, d), generator=gen, device=device)
    feats = feats * (3.7 * torch.exp(0.25 * torch.randn((n_img, S * S, 1), generator=gen, device=device)))
    return feats.contiguous(), maps.contiguous()


def build_bank(w, rows, device, seed):
    import torch

    from hbird_b200 import ops

    S, ps = w["S"], w["ps"]
    bank = ops.MemoryBank(w["d"], w["C"], ps * ps, rows, device.index, keep_f32=True)
    gen = torch.Generator(device=device).manual_seed(seed)
    per_img = S * S
    slab = max(1, (1 << 19) // per_img)
    left = rows
    while left > 0:
        n_img = min(slab, (left + per_img - 1) // per_img)
        feats, maps = synth_images(w, n_img, gen, device)
        maps = torch.where(maps == 255, torch.zeros_like(maps), maps)  # bank side: 255 -> 0
        take = min(left, n_img * per_img)
        sel = None if take == n_img * per_img else torch.arange(take, device=device, dtype=torch.int32)
        bank.append(feats, maps, S, ps, sel)
        left -= take
    bank.finalize()
    return bank


def make_query_ring(w, device, seed):
    import torch

    gen = torch.Generator(device=device).manual_seed(seed)
    ring = []
    for _ in range(RING):
        feats, maps = synth_images(w, w["B"], gen, device)
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

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- the hot path
def run_step(ops, bank, table, w, q, y, conf, shard=None):
    """One pass: decode -> search -> label transfer -> upsample+argmax -> confusion."""
    B, S, H = w["B"], w["S"], w["S"] * w["ps"]
    gt = ops.decode_mask(y, False)
    if shard is not None and shard.get("xchg") is not None:
        qn = shard["xchg"].search_scatter(bank, q, shard["qsplit"], K_NEIGH, K_PRIME, shard["offset"])
    else:
        scores, idx, qn = bank.search(q, K_NEIGH, K_PRIME, 0 if shard is None else shard["offset"])
    if shard is not None:
        from hbird_b200 import distributed as hdist

        b0, b1 = hdist.split_range(B, shard["world"], shard["rank"])
        n = S * S
        if shard.get("xchg") is not None:  # fused exchange: K2b scatters over NVLink, merge waits
            scores, idx = shard["xchg"].merge()
        else:
            gs, gi = hdist.all_gather_topk(scores, idx)
            scores, idx = ops.merge_topk(gs, gi)
            scores, idx = scores[b0 * n:b1 * n], idx[b0 * n:b1 * n]
        lh = ops.label_transfer(table, w["ps"] ** 2, scores, idx, qn[b0 * n:b1 * n], BETA)
        pred = ops.upsample_argmax(lh, b1 - b0, S, H, H)
        ops.confusion_accumulate(conf, gt.view(B, H, H)[b0:b1], pred, w["ignore"])
    else:
        lh = ops.label_transfer(table, w["ps"] ** 2, scores, idx, qn, BETA)
        pred = ops.upsample_argmax(lh, B, S, H, H)
        ops.confusion_accumulate(conf, gt.view(B, H, H), pred, w["ignore"])


def timed_loop(torch, dist, world, fn, steps, warmup):
    """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def cpu_sample_images(w, seconds):
    """How many images of the workload a ~`seconds` CPU sample holds, assuming ~0.4 TFLOP/s of host
    GEMM + selection (one image at least, the ring's 8 batches at most)."""
    per_img = 2.0 * w["S"] ** 2 * w["N"] * w["d"] * 1.6  # flop, with ~60 % on top for the top-k pass
    return int(max(1, min(w["B"] * RING, seconds * 4e11 / per_img)))


def cpu_reference_sample(w, bank, ring, n_img, repeats=1):
    """The oracle port of the reference path on the host cores, on `n_img` images of the workload's
    first validation batch against the FULL bank.  Returns (queries/s, seconds, threads)."""
    import numpy as np

    from oracle import hbird_oracle as O

    fm_t, lm_t = bank.export()
    fm, lm = fm_t.cpu().numpy(), lm_t.cpu().numpy()
    del fm_t, lm_t
    S, d = w["S"], w["d"]
    n_img = min(n_img, w["B"] * len(ring))
    fl, yl, left = [], [], n_img
    for q, y in ring:  # images are taken from as many of the ring's batches as needed
        take = min(left, w["B"])
        fl.append(q.view(w["B"], S * S, d)[:take].cpu().numpy())
        yl.append(y[:take].cpu().numpy())
        left -= take
        if left == 0:
            break
    feats, yy = np.concatenate(fl), np.concatenate(yl)
    # score blocks of at most ~4 GB of host RAM: (images per block) * S*S * N * 4 B
    per_block = max(1, min(n_img, int(4e9 // (S * S * fm.shape[0] * 4))))
    batches = [(feats[i:i + per_block], yy[i:i + per_block]) for i in range(0, n_img, per_block)]
    threads = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info

        blas = [t["num_threads"] for t in threadpool_info() if t.get("user_api") == "blas"]
        threads = max(blas) if blas else threads
    except Exception:
        pass
    t0 = time.perf_counter()
    for _ in range(repeats):
        ref = O.evaluate(fm, lm, batches, w["C"], S, K_NEIGH, w["ignore"], BETA, return_details=True)
    dt = (time.perf_counter() - t0) / repeats
    return n_img * S * S / dt, dt, threads, ref


def parity_vs_oracle(ops, torch, w, bank, table, ring, n_img, ref):
    """The same n_img images through the CUDA path, compared with what the oracle just computed:
    the "mIoU delta" of the headline metric, measured on this workload's own bank."""
    import numpy as np

    from hbird_b200.utils.eval_metrics import miou_from_confusion

    ref_miou, ref_conf, det = ref
    S, d, H, C = w["S"], w["d"], w["S"] * w["ps"], w["C"]
    n_img = min(n_img, w["B"] * len(ring))
    qs = torch.cat([q for q, _ in ring])[:n_img * S * S].contiguous()
    ys = torch.cat([y for _, y in ring])[:n_img].contiguous()
    scores, idx, qn = bank.search(qs, K_NEIGH, K_PRIME)
    lh = ops.label_transfer(table, w["ps"] ** 2, scores, idx, qn, BETA)
    pred = ops.upsample_argmax(lh, n_img, S, H, H)
    conf = torch.zeros((C, C), dtype=torch.int64, device=qs.device)
    ops.confusion_accumulate(conf, ops.decode_mask(ys, False).view(n_img, H, H), pred, w["ignore"])
    miou = miou_from_confusion(conf.cpu().numpy())[0]
    ref_idx = np.concatenate(det["idx"])
    ref_dist = np.concatenate(det["dist"])
    got_idx, got_s = idx.cpu().numpy(), scores.cpu().numpy()
    recall = float((got_idx[:, :, None] == ref_idx[:, None, :]).any(axis=2).mean())
    rel = float((np.abs(got_s - ref_dist) / np.maximum(np.abs(ref_dist), 1e-6)).max())
    ref_pred = np.concatenate(det["pred"])[:, 0]
    agree = float((pred.cpu().numpy() == ref_pred).mean())
    return {"images": n_img, "queries": n_img * S * S, "miou_b200": miou, "miou_oracle": ref_miou,
            "miou_delta_points": abs(miou - ref_miou) * 100.0, "recall_at_30": recall, "score_max_rel_err": rel,
            "pred_pixel_agreement": agree, "confusion_abs_diff": int(np.abs(conf.cpu().numpy() - ref_conf).sum()),
            "gates": "recall>=0.999, rel<=1e-3, |dmIoU|<=0.05 points"}


def _claim_stdout():
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr and return a
    file object on the real stdout, so that the only thing on stdout is the final JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--cta-group", type=int, default=0)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = dict(WORKLOADS[args.workload])
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    Q = w["B"] * w["S"] * w["S"]

    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the hot path has no CPU fallback (the reference arm also "
                           "builds its synthetic bank with the CUDA pack kernel)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    from hbird_b200 import ops

    ops.device_check(device.index)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        bank = build_bank(w, w["N"], device, seed=1)
        ring = make_query_ring(w, device, seed=2)
        n_img = cpu_sample_images(w, seconds=1.5)  # per step
        qps_list = []
        for i in range(args.warmup + args.steps):
            qps, dt, threads, _ = cpu_reference_sample(w, bank, ring, n_img)
            if i >= args.warmup:
                qps_list.append((qps, dt))
        tot_q = n_img * w["S"] ** 2 * len(qps_list)
        tot_t = sum(dt for _, dt in qps_list)
        value = tot_q / tot_t
        sample = f"{n_img} image(s) = {n_img * w['S'] ** 2} patch-queries per step against the full {w['N']:,}-row bank"
        line = {
            "impl": "reference", "metric": "patch_queries_per_sec", "value": value, "unit": "patch-queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(qps_list)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "k": K_NEIGH, "path": "oracle port of the reference "
                       "CPU path (exact fp32 IP search + gather + cross-attention + bilinear upsample + argmax + bincount)"},
            "cpu_baseline": {"value": value, "unit": "patch-queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "patch-queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), file=real_stdout, flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peak_tf, peak_src = load_peaks()
    bank = build_bank(w, w["N"], device, seed=1)  # replica: the full bank on every GPU
    if args.cta_group:
        bank.configure_search(cta_group=args.cta_group)
    table = bank.label_table()
    ring = make_query_ring(w, device, seed=2 + rank)
    conf = torch.zeros((w["C"], w["C"]), dtype=torch.int64, device=device)
    torch.cuda.synchronize()

    def step(i):
        q, y = ring[i % RING]
        run_step(ops, bank, table, w, q, y, conf)

    sampler = ClockSampler(device.index)
    bank.enable_kernel_timing(True)
    for i in range(warmup):  # warm-up outside the clock-sampling window
        step(i)
    torch.cuda.synchronize()
    bank.enable_kernel_timing(True)  # reset the event ring: only timed steps are averaged
    sampler.start()
    ms_total = timed_loop(torch, dist, world, step, args.steps, 0)
    clocks = sampler.stop()
    kern_ms, kern_n = bank.kernel_time_ms()
    bank.enable_kernel_timing(False)
    launches_per_step = bank.last_search_launches() + 4  # decode, label transfer, upsample+argmax, confusion
    ms_per_step = ms_total / args.steps
    value = world * Q / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, host result out, copies inside the timed region.  Every step copies its
    # own inputs from pinned host memory and reads its result back; the copy of step i+1 is issued on a
    # second stream while step i computes (double-buffered device inputs), as a serving loop would.
    host_ring = [(q.cpu().pin_memory(), y.cpu().pin_memory()) for q, y in ring]
    dev_in = [(torch.empty_like(ring[0][0]), torch.empty_like(ring[0][1])) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    conf_host = torch.zeros((w["C"], w["C"]), dtype=torch.int64).pin_memory()
    issued = {"next": None}

    def issue_copy(i):
        b = i % 2
        qh, yh = host_ring[i % RING]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            dev_in[b][0].copy_(qh, non_blocking=True)
            dev_in[b][1].copy_(yh, non_blocking=True)
            copied[b].record(copy_stream)

    def step_e2e(i):
        if issued["next"] != i:  # first step of a loop: nothing was prefetched
            issue_copy(i)
        issue_copy(i + 1)
        issued["next"] = i + 1
        b = i % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(copied[b])
        run_step(ops, bank, table, w, dev_in[b][0], dev_in[b][1], conf)
        consumed[b].record(cur)
        conf_host.copy_(conf, non_blocking=True)
        cur.synchronize()  # the caller reads the step's result

    for b in range(2):
        consumed[b].record(torch.cuda.current_stream())
    ms_e2e = timed_loop(torch, dist, world, step_e2e, args.steps, warmup) / args.steps
    h2d = host_ring[0][0].numel() * 4 + host_ring[0][1].numel() * 4
    d2h = conf_host.numel() * 8
    e2e_value = world * Q / (ms_e2e * 1e-3)

    # ---- the reference's literal plugin call (search_faiss.py:83-90): pageable host queries in,
    # host (indices, distances) out, every copy synchronous -- what a third party using the ABC gets
    plugin = None
    if world == 1:
        import time

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

    # ---- row-sharded bank: per-shard search -> NCCL all-gather -> merge kernel (strong scaling)
    sharded = None
    if world > 1 and not args.no_sharded:
        from hbird_b200 import distributed as hdist

        a, b = hdist.shard_bounds(w["N"], world, rank)
        shard_bank = build_bank(w, b - a, device, seed=100 + rank)
        counts = hdist.gather_counts(shard_bank.rows, device)
        full_table = hdist.all_gather_rows(shard_bank.label_table(), counts)
        info = {"offset": hdist.offsets_from_counts(counts)[rank], "world": world, "rank": rank}
        shared_ring = make_query_ring(w, device, seed=2)  # every rank sees every query batch
        conf2 = torch.zeros_like(conf)

        def step_sh(i):
            q, y = shared_ring[i % RING]
            run_step(ops, shard_bank, full_table, w, q, y, conf2, shard=info)

        ms_nccl = timed_loop(torch, dist, world, step_sh, args.steps, warmup) / args.steps
        conf_nccl = conf2.clone()
        # fused exchange over NVLink peer memory (K2b peer stores + waiting merge kernel)
        per_rank = -(-w["B"] // world) * w["S"] ** 2
        xchg = hdist.connect_shard_exchange(per_rank, K_NEIGH, device)
        ms_sh, collective = ms_nccl, "NCCL all-gather of (score f32, idx i64)[Q,k] + k-way merge kernel"
        exchange_parity = None
        if xchg is not None:
            info["xchg"], info["qsplit"] = xchg, hdist.query_split(w["B"], w["S"] ** 2, world)
            conf2.zero_()
            ms_sh = timed_loop(torch, dist, world, step_sh, args.steps, warmup) / args.steps
            collective = ("fused: K2b stores each query's shard top-k into the owner rank's window over "
                          "NVLink (CUDA IPC), merge kernel waits on per-rank step flags; no NCCL call")
            exchange_parity = bool(torch.equal(conf2, conf_nccl))  # same steps, same batches
        sharded = {"value": Q / (ms_sh * 1e-3), "unit": "patch-queries/s", "ms_per_step": ms_sh,
                   "bank_rows_total": w["N"], "bank_rows_per_gpu": b - a, "scaling": "strong",
                   "collective": collective, "nccl_path_ms_per_step": ms_nccl,
                   "confusion_equals_nccl_path": exchange_parity}
        if xchg is not None:
            torch.cuda.synchronize()
            dist.barrier()
            hdist.close_shard_exchange(xchg)
        shard_bank.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        n_img = cpu_sample_images(w, seconds=15.0)
        qps, dt, threads, ref = cpu_reference_sample(w, bank, ring, n_img)
        parity = parity_vs_oracle(ops, torch, w, bank, table, ring, n_img, ref)
        cpu = {"value": qps, "unit": "patch-queries/s", "cores": threads, "kind": "port",
               "sample": f"{n_img} images = {n_img * w['S'] ** 2} patch-queries against the full {w['N']:,}-row bank, "
                         f"{dt:.1f} s of oracle (numpy/BLAS) time"}

    if rank == 0:
        flop = 2.0 * w["N"] * w["d"] * Q
        achieved = flop / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json"))).get(args.workload)
        except Exception:
            pass
        line = {
            "metric": "patch_queries_per_sec", "value": value, "unit": "patch-queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: {w['desc']}", "bank_rows": w["N"], "d": w["d"], "k": K_NEIGH,
                "k_prime": K_PRIME, "queries_per_step_per_gpu": Q, "classes": w["C"],
                "parallelism": "single GPU" if world == 1 else f"replicas x{world} (full bank per GPU, queries split; no data-path collective)",
                "path": "decode -> tcgen05 bf16 GEMM + fused top-k' -> fp32 exact re-rank -> label transfer -> upsample+argmax -> confusion",
                "l2": f"inputs larger than L2: bf16 bank {w['N'] * w['d'] * 2 / 1e6:.0f} MB streamed every step, {RING} distinct query batches cycled",
            },
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
                         "frac_of_nominal_2250": (achieved / 2250.0) if achieved else None,
                         "kernel": "search_topk_kernel (tcgen05 GEMM + fused top-k')", "kernel_ms": kern_ms,
                         "kernel_launches_timed": kern_n, "flop_per_launch": flop, "peak_source": peak_src,
                         "kernel_share_of_step": kern_ms / ms_per_step if ms_per_step else None},
            "cpu_baseline": cpu,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": "patch-queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "plugin_call": plugin,
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
        }
        if sharded is not None:
            line["sharded"] = sharded
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
