"""Torch-tensor front end of the C-ABI: device memory and streams come from PyTorch, every
computation is a call into libhbird_b200.so.  No function here has a PyTorch/CPU fallback."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _capi
from ._capi import check, lib, ptr, stream_ptr


def _require_cuda(t: torch.Tensor, name: str, dtype: torch.dtype) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"hbird_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"hbird_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t.contiguous()


def device_check(device: int = 0) -> int:
    """SM count of an sm_100 device; raises RuntimeError otherwise (search_faiss.py:14-16)."""
    import ctypes

    n = ctypes.c_int(0)
    check(lib.hb_device_check(int(device), ctypes.byref(n)))
    return n.value


class _CudaArray:
    """Minimal __cuda_array_interface__ view over library-owned HBM."""

    def __init__(self, address: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (address, False), "version": 3, "strides": None,
        }
        self._owner = owner


class MemoryBank:
    """One HBM-resident shard of the memory bank (K1).  Mirrors feature_memory / label_memory of
    hbird_eval.py:156-161,357-366, packed as bf16 rows (+ fp32 copy) and uint16 class histograms."""

    def __init__(self, d: int, num_classes: int, patch_pixels: int, capacity_rows: int,
                 device: int = 0, keep_f32: bool = True, metric: str = "dot_product"):
        import ctypes

        if metric not in ("dot_product", "l2"):
            raise ValueError(f"Unsupported distance measure: {metric}")  # search_faiss.py:48
        self.metric = metric

        self.d, self.num_classes, self.patch_pixels = int(d), int(num_classes), int(patch_pixels)
        self.device = int(device)
        self._h = ctypes.c_void_p(0)
        flags = (_capi.HB_BANK_KEEP_F32 if keep_f32 else 0) | (_capi.HB_BANK_L2 if metric == "l2" else 0)
        check(lib.hb_bank_create(self.device, self.d, self.num_classes, self.patch_pixels,
                                 int(capacity_rows), flags, ctypes.byref(self._h)))
        self.finalized = False

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.hb_bank_destroy(self._h)
            self._h.value = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rows(self) -> int:
        return int(lib.hb_bank_rows(self._h))

    @property
    def capacity(self) -> int:
        return int(lib.hb_bank_capacity(self._h))

    def append(self, feats: torch.Tensor, mask_u8: torch.Tensor, S: int, ps: int,
               sel: Optional[torch.Tensor] = None) -> None:
        """feats fp32 (B, S*S, d) raw features; mask_u8 uint8 (B, S*ps, S*ps) decoded class ids;
        sel optional int32 (n,) flat source-row picks (bounded sampler)."""
        feats = _require_cuda(feats, "feats", torch.float32)
        mask_u8 = _require_cuda(mask_u8, "mask", torch.uint8)
        B = mask_u8.shape[0]
        if feats.numel() != B * S * S * self.d:
            raise ValueError(f"feats has {feats.numel()} elements, expected B*S*S*d = {B * S * S * self.d}")
        if tuple(mask_u8.shape[-2:]) != (S * ps, S * ps):
            raise ValueError(f"mask spatial size {tuple(mask_u8.shape[-2:])} != (S*ps, S*ps) = {(S * ps, S * ps)}")
        if sel is not None:
            sel = _require_cuda(sel, "sel", torch.int32)
            n = sel.numel()
        else:
            n = B * S * S
        check(lib.hb_bank_append(self._h, ptr(feats), ptr(mask_u8), B, S, ps, ptr(sel), n,
                                 stream_ptr(feats.device)))

    def append_soft(self, feats: torch.Tensor, soft: torch.Tensor, normalise: bool = False) -> None:
        feats = _require_cuda(feats, "feats", torch.float32)
        soft = _require_cuda(soft, "soft", torch.float32)
        n = feats.shape[0]
        if feats.shape != (n, self.d) or soft.shape != (n, self.num_classes):
            raise ValueError("append_soft expects feats (n, d) and soft (n, C)")
        check(lib.hb_bank_append_soft(self._h, ptr(feats), ptr(soft), n, int(normalise), stream_ptr(feats.device)))

    def finalize(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
        check(lib.hb_bank_finalize(self._h))
        self.finalized = True

    def label_table(self) -> torch.Tensor:
        """(rows, C) view of the bank-owned uint16 label histograms (no copy).  Exposed as int16
        (counts are <= patch_pixels < 32768) because torch/NCCL support for uint16 is partial."""
        addr = lib.hb_bank_label_table(self._h)
        arr = _CudaArray(addr, (self.rows, self.num_classes), "<i2", self)
        with torch.cuda.device(self.device):
            return torch.as_tensor(arr, device=f"cuda:{self.device}")

    def export(self, row0: int = 0, n: Optional[int] = None, features: bool = True,
               labels: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """The reference's (feature_memory (n,d) fp32, label_memory (n,C) fp32) for rows [row0,row0+n)."""
        n = self.rows - row0 if n is None else n
        dev = torch.device("cuda", self.device)
        f = torch.empty((n, self.d), dtype=torch.float32, device=dev) if features else None
        l = torch.empty((n, self.num_classes), dtype=torch.float32, device=dev) if labels else None
        check(lib.hb_bank_export(self._h, row0, n, ptr(f), ptr(l), stream_ptr(dev)))
        return f, l

    def configure_search(self, cta_group: int = 0, max_chunks: int = 0) -> None:
        check(lib.hb_search_config(self._h, int(cta_group), int(max_chunks)))

    def search(self, q: torch.Tensor, k: int = 30, k_prime: int = 64, idx_offset: int = 0,
               return_qnorm: bool = True):
        """K2 + K2b.  q fp32 (Q, d) on the bank's device, raw (un-normalised).  Returns
        (scores fp32 (Q,k) descending, idx int64 (Q,k), qnorm fp32 (Q,) or None)."""
        q = _require_cuda(q, "q", torch.float32)
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"queries must be (Q, {self.d}), got {tuple(q.shape)}")
        Q = q.shape[0]
        scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
        idx = torch.empty((Q, k), dtype=torch.int64, device=q.device)
        qn = torch.empty((Q,), dtype=torch.float32, device=q.device) if return_qnorm else None
        check(lib.hb_search(self._h, ptr(q), Q, int(k), int(k_prime), int(idx_offset), ptr(scores), ptr(idx),
                            ptr(qn), stream_ptr(q.device)))
        return scores, idx, qn

    def search_transfer(self, q: torch.Tensor, k: int = 30, k_prime: int = 64, idx_offset: int = 0,
                        beta: float = 0.02, label_table: Optional[torch.Tensor] = None, return_neighbours: bool = False):
        """K2 + K2b with K4a fused into the re-rank warp.  Returns (label_hat fp32 (Q, C), qnorm fp32 (Q,),
        scores, idx) — scores/idx are None unless return_neighbours.  label_table: int16 (rows, C) indexed by
        global row (None = this bank's own table)."""
        q = _require_cuda(q, "q", torch.float32)
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"queries must be (Q, {self.d}), got {tuple(q.shape)}")
        Q = q.shape[0]
        table_rows = 0
        if label_table is not None:
            label_table = _require_cuda(label_table, "label_table", torch.int16)
            if label_table.dim() != 2 or label_table.shape[1] != self.num_classes:
                raise ValueError(f"label_table must be (rows, {self.num_classes})")
            table_rows = label_table.shape[0]
        lh = torch.empty((Q, self.num_classes), dtype=torch.float32, device=q.device)
        qn = torch.empty((Q,), dtype=torch.float32, device=q.device)
        scores = torch.empty((Q, k), dtype=torch.float32, device=q.device) if return_neighbours else None
        idx = torch.empty((Q, k), dtype=torch.int64, device=q.device) if return_neighbours else None
        check(lib.hb_search_transfer(self._h, ptr(label_table), table_rows, ptr(q), Q, int(k), int(k_prime),
                                     int(idx_offset), float(beta), ptr(scores), ptr(idx), ptr(qn), ptr(lh),
                                     stream_ptr(q.device)))
        return lh, qn, scores, idx

    def search_begin(self, q: torch.Tensor, k_prime: int = 64, slot: int = 0,
                     prepared: Optional[torch.cuda.Event] = None) -> torch.Tensor:
        """First half of a search on the CURRENT stream: query prep + K2 (tensor-core pass) into pipeline
        slot 0/1.  Returns the (Q,) query norms.  Pair with search_finish / ShardExchange.finish_scatter,
        which may run on another stream (ordered after this one by an event) while the next batch's
        search_begin already executes — see hbird_b200.pipeline.EvalPipeline.  prepared: an event the
        library records between the query prep and K2 (the previous batch's finish is released there)."""
        q = _require_cuda(q, "q", torch.float32)
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"queries must be (Q, {self.d}), got {tuple(q.shape)}")
        qn = torch.empty((q.shape[0],), dtype=torch.float32, device=q.device)
        ev = None
        if prepared is not None:
            import ctypes

            prepared.record(torch.cuda.current_stream(q.device))  # materialises the lazily created CUDA event
            ev = ctypes.c_void_p(prepared.cuda_event)
        check(lib.hb_search_begin(self._h, ptr(q), q.shape[0], int(k_prime), int(slot), ptr(qn), ev, stream_ptr(q.device)))
        return qn

    def search_finish(self, slot: int, q: torch.Tensor, k: int = 30, idx_offset: int = 0, beta: float = 0.02,
                      label_table: Optional[torch.Tensor] = None, want_label_hat: bool = True,
                      want_neighbours: bool = False):
        """Second half on the CURRENT stream: K2b (+ fused K4a) of the search begun in `slot`.
        Returns (label_hat or None, scores or None, idx or None)."""
        q = _require_cuda(q, "q", torch.float32)
        Q = q.shape[0]
        table_rows = 0
        if label_table is not None:
            label_table = _require_cuda(label_table, "label_table", torch.int16)
            table_rows = label_table.shape[0]
        lh = torch.empty((Q, self.num_classes), dtype=torch.float32, device=q.device) if want_label_hat else None
        scores = torch.empty((Q, k), dtype=torch.float32, device=q.device) if want_neighbours else None
        idx = torch.empty((Q, k), dtype=torch.int64, device=q.device) if want_neighbours else None
        check(lib.hb_search_finish(self._h, int(slot), ptr(q), int(k), int(idx_offset), ptr(label_table), table_rows,
                                   float(beta), ptr(scores), ptr(idx), ptr(lh), stream_ptr(q.device)))
        return lh, scores, idx

    def search_abort(self) -> None:
        """Free both pipeline slots after an error between search_begin and search_finish."""
        check(lib.hb_search_abort(self._h))

    def eval_step(self, q: torch.Tensor, y: torch.Tensor, S: int, conf: torch.Tensor, ignore_index: Optional[int],
                  k: int = 30, k_prime: int = 64, beta: float = 0.02, label_table: Optional[torch.Tensor] = None,
                  idx_offset: int = 0, label_hat: Optional[torch.Tensor] = None, pred: Optional[torch.Tensor] = None,
                  scores: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One validation batch in 4 launches (hb_eval_step): q fp32 (B*S*S, d) raw features, y fp32
        (B, 1, H, W) or (B, H, W) = id/255, conf int64 (C, C) accumulated in place.  Optional
        pre-allocated outputs (label_hat (B*S*S, C) fp32, pred (B, H, W) uint8, scores/idx (B*S*S, k))
        make the call allocation-free, i.e. capturable in a CUDA graph.  Returns label_hat."""
        q = _require_cuda(q, "q", torch.float32)
        y = _require_cuda(y, "y", torch.float32)
        conf = _require_cuda(conf, "conf", torch.int64)
        H, W = int(y.shape[-2]), int(y.shape[-1])
        B = y.numel() // (H * W)
        if q.shape != (B * S * S, self.d):
            raise ValueError(f"queries must be (B*S*S, d) = ({B * S * S}, {self.d}), got {tuple(q.shape)}")
        if tuple(conf.shape) != (self.num_classes, self.num_classes):
            raise ValueError(f"conf must be ({self.num_classes}, {self.num_classes})")
        table_rows = 0
        if label_table is not None:
            label_table = _require_cuda(label_table, "label_table", torch.int16)
            table_rows = label_table.shape[0]
        if label_hat is None:
            label_hat = torch.empty((B * S * S, self.num_classes), dtype=torch.float32, device=q.device)
        if (scores is None) != (idx is None):
            raise ValueError("give both or neither of scores / idx")
        ig = -1 if ignore_index is None else int(ignore_index)
        check(lib.hb_eval_step(self._h, ptr(label_table), table_rows, ptr(q), B, int(S), H, W, ptr(y), int(k),
                               int(k_prime), int(idx_offset), float(beta), ig, ptr(label_hat), ptr(conf), ptr(pred),
                               ptr(scores), ptr(idx), stream_ptr(q.device)))
        return label_hat

    def configure_coresidency(self, lean_search: bool = False, rerank_warps_per_cta: int = 0,
                              rerank_shared_carveout: int = -1) -> None:
        """How the search kernel and the re-rank kernel share an SM when the latter is issued under the
        former (pipeline.py): see hb_coresidency_config."""
        check(lib.hb_coresidency_config(self._h, int(bool(lean_search)), int(rerank_warps_per_cta),
                                        int(rerank_shared_carveout)))

    def tune_search(self, prefetch_tiles: int = -1, ablate: int = 0) -> None:
        """ablate != 0 is for measurement only (wrong results): 1 = GEMM pipeline alone, 2 = scan only."""
        check(lib.hb_search_tune(self._h, int(prefetch_tiles), int(ablate)))

    def set_stats_buffer(self, buf: Optional[torch.Tensor]) -> None:
        """Diagnostics: int64 CUDA tensor of 148*8*8 zeros to collect epilogue cycle counters, or None."""
        self._stats_buf = buf
        check(lib.hb_search_stats(self._h, ptr(buf)))

    def set_pacing(self, enable: bool = True) -> None:
        check(lib.hb_search_pacing(self._h, int(enable)))

    def enable_kernel_timing(self, enable: bool = True) -> None:
        check(lib.hb_search_timing(self._h, int(enable)))

    def kernel_time_ms(self):
        """(mean ms, count) of the tcgen05 search kernel over the searches since timing was enabled."""
        import ctypes

        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        check(lib.hb_search_kernel_time(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def rerank_time_ms(self):
        """(mean ms, count) of K2b (+ fused K4a / scatter) over the same searches."""
        import ctypes

        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        check(lib.hb_search_rerank_time(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def calibrate_search_ms(self, n_queries: int = 8192, k_prime: int = 64, seconds: float = 1.5) -> float:
        """Mean duration (ms) of the tensor-core search kernel over this bank for n_queries random
        queries, measured over ~`seconds` of back-to-back searches: the per-GPU speed figure shard
        balancing is based on (distributed.balanced_counts).  It has to be a SUSTAINED figure — a short
        burst runs above the power cap's steady clocks and ranks the GPUs differently (measured on an
        8-GPU box: burst spread 8 %, sustained 3 %)."""
        import time

        dev = torch.device("cuda", self.device)
        g = torch.Generator(device=dev).manual_seed(1234)
        q = torch.randn((n_queries, self.d), generator=g, device=dev)
        self.search(q, 1, k_prime)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        self.search(q, 1, k_prime)
        torch.cuda.synchronize(dev)
        once = max(time.perf_counter() - t0, 1e-4)
        n = int(max(8, min(2000, seconds / once)))
        for _ in range(max(1, n - 48)):   # settle the clocks; the event ring keeps the last 64 searches
            self.search(q, 1, k_prime)
        self.enable_kernel_timing(True)
        for _ in range(min(n, 48)):
            self.search(q, 1, k_prime)
        ms, _ = self.kernel_time_ms()
        self.enable_kernel_timing(False)
        return ms

    def last_search_launches(self) -> int:
        return int(lib.hb_search_last_launches(self._h))

    def dump_scores(self, q: torch.Tensor, cta_group: int = 0) -> torch.Tensor:
        """Validation only: the full (Q, rows) bf16-input score matrix of the tcgen05 pass."""
        q = _require_cuda(q, "q", torch.float32)
        out = torch.full((q.shape[0], self.rows), float("nan"), dtype=torch.float32, device=q.device)
        check(lib.hb_search_dump_scores(self._h, ptr(q), q.shape[0], ptr(out), int(cta_group), stream_ptr(q.device)))
        return out


def sample_patches(mask_u8: torch.Tensor, S: int, ps: int, num_classes: int, uniform: torch.Tensor, K: int) -> torch.Tensor:
    """N1: bounded-memory sampler (hbird_eval.py:447-517).  mask_u8 uint8 (B, S*ps, S*ps); uniform fp32
    (B, S*S) U(0,1) draws.  Returns int32 (B*K,) flat source rows for MemoryBank.append(sel=...)."""
    mask_u8 = _require_cuda(mask_u8, "mask", torch.uint8)
    uniform = _require_cuda(uniform, "uniform", torch.float32)
    B = mask_u8.shape[0]
    if uniform.numel() != B * S * S:
        raise ValueError(f"uniform must hold B*S*S = {B * S * S} draws, got {uniform.numel()}")
    sel = torch.empty((B * K,), dtype=torch.int32, device=mask_u8.device)
    check(lib.hb_sample_patches(ptr(mask_u8), B, S, ps, int(num_classes), ptr(uniform), int(K), ptr(sel),
                                stream_ptr(mask_u8.device)))
    return sel


class ShardExchange:
    """K3x: one rank's end of the fused shard exchange (include/hbird_b200.h).  K2b stores each
    query's shard top-k into the owner rank's window over NVLink; `merge` waits for all ranks and
    merges the local slice.  Replaces all-gather + merge (faiss.IndexShards, search_faiss.py:53-63)."""

    def __init__(self, rank: int, world: int, slice_capacity: int, max_k: int = 30, device: int = 0):
        import ctypes

        self.rank, self.world, self.device = int(rank), int(world), int(device)
        self.slice_capacity, self.max_k = int(slice_capacity), int(max_k)
        self._h = ctypes.c_void_p(0)
        check(lib.hb_exchange_create(self.device, self.rank, self.world, self.slice_capacity, self.max_k,
                                     ctypes.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.hb_exchange_destroy(self._h)
            self._h.value = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_timeout(self, timeout_ms: int) -> None:
        """Bound of a merge kernel's wait for its peers (default 10 min)."""
        check(lib.hb_exchange_set_timeout(self._h, int(timeout_ms)))

    def configure(self, threshold_exchange: bool) -> None:
        """False (default): every shard re-ranks its whole bf16 top-k' (one hop).  True: the threshold
        exchange — shards first swap order statistics of their shortlists and then re-rank only the
        candidates at or above the common bound (hb_exchange_config mode 1).  Same setting on all ranks."""
        check(lib.hb_exchange_config(self._h, 1 if threshold_exchange else 0))

    def rerank(self) -> None:
        """Threshold exchange only: issue phase 2 of the last scatter on the current stream now (the next
        merge would do it implicitly).  The queries passed to the scatter must still be alive."""
        check(lib.hb_exchange_rerank(self._h, stream_ptr(torch.device("cuda", self.device))))

    def check_status(self) -> None:
        """Synchronise the current stream and raise RuntimeError if a merge since the last check gave up
        waiting for a peer (its outputs were then not written)."""
        check(lib.hb_exchange_status(self._h, stream_ptr(torch.device("cuda", self.device))))

    def disconnect(self) -> None:
        """Unmap the peers' windows (first half of an orderly multi-process teardown)."""
        if self._h.value:
            check(lib.hb_exchange_disconnect(self._h))

    def handle(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's window, to be all-gathered by the host side."""
        import ctypes

        buf = ctypes.create_string_buffer(_capi.HB_EXCHANGE_HANDLE_BYTES)
        check(lib.hb_exchange_handle(self._h, buf, _capi.HB_EXCHANGE_HANDLE_BYTES))
        return buf.raw

    def connect(self, handles) -> None:
        """handles: the `world` handles in rank order (bytes each)."""
        blob = b"".join(handles)
        if len(blob) != self.world * _capi.HB_EXCHANGE_HANDLE_BYTES:
            raise ValueError(f"expected {self.world} handles of {_capi.HB_EXCHANGE_HANDLE_BYTES} bytes")
        check(lib.hb_exchange_connect(self._h, blob, self.world))

    @staticmethod
    def connect_local(exchanges) -> None:
        """Wire exchanges created in this process (one per simulated rank, same device) together."""
        import ctypes

        arr = (ctypes.c_void_p * len(exchanges))(*[x._h.value for x in exchanges])
        for x in exchanges:
            check(lib.hb_exchange_connect_local(x._h, arr, len(exchanges)))

    def search_scatter(self, bank: "MemoryBank", q: torch.Tensor, qsplit, k: int = 30, k_prime: int = 64,
                       idx_offset: int = 0, return_qnorm: bool = True):
        """K2 + K2b with the output rows of queries [qsplit[p], qsplit[p+1]) stored into rank p's
        window.  Returns the (Q,) query norms (or None)."""
        import ctypes

        q = _require_cuda(q, "q", torch.float32)
        if q.dim() != 2 or q.shape[1] != bank.d:
            raise ValueError(f"queries must be (Q, {bank.d}), got {tuple(q.shape)}")
        if len(qsplit) != self.world + 1:
            raise ValueError(f"qsplit needs world+1 = {self.world + 1} entries")
        Q = q.shape[0]
        qn = torch.empty((Q,), dtype=torch.float32, device=q.device) if return_qnorm else None
        arr = (ctypes.c_int64 * (self.world + 1))(*[int(v) for v in qsplit])
        check(lib.hb_search_scatter(bank._h, self._h, ptr(q), Q, int(k), int(k_prime), int(idx_offset), arr,
                                    ptr(qn), stream_ptr(q.device)))
        self._k = int(k)
        return qn

    def finish_scatter(self, bank: "MemoryBank", slot: int, q: torch.Tensor, qsplit, k: int = 30, idx_offset: int = 0) -> None:
        """Second half of search_scatter for a search begun with MemoryBank.search_begin(slot): K2b with the
        scatter into the owner ranks' windows, on the CURRENT stream."""
        import ctypes

        q = _require_cuda(q, "q", torch.float32)
        if len(qsplit) != self.world + 1:
            raise ValueError(f"qsplit needs world+1 = {self.world + 1} entries")
        arr = (ctypes.c_int64 * (self.world + 1))(*[int(v) for v in qsplit])
        check(lib.hb_search_finish_scatter(bank._h, self._h, int(slot), ptr(q), int(k), int(idx_offset), arr,
                                           stream_ptr(q.device)))
        self._k = int(k)

    def merge_transfer(self, label_table: torch.Tensor, patch_pixels: int, qnorm_slice: torch.Tensor,
                       beta: float = 0.02, return_neighbours: bool = False):
        """K3x with K4a fused into the merging warp: (label_hat (rows, C), scores, idx) of this rank's
        slice of the last scatter (scores/idx None unless return_neighbours)."""
        label_table = _require_cuda(label_table, "label_table", torch.int16)
        qnorm_slice = _require_cuda(qnorm_slice, "qnorm_slice", torch.float32)
        rows = int(lib.hb_exchange_slice_rows(self._h))
        if qnorm_slice.numel() != rows:
            raise ValueError(f"qnorm_slice must hold {rows} norms, got {qnorm_slice.numel()}")
        trows, C = label_table.shape
        dev = torch.device("cuda", self.device)
        lh = torch.empty((rows, C), dtype=torch.float32, device=dev)
        out_s = torch.empty((rows, self._k), dtype=torch.float32, device=dev) if return_neighbours else None
        out_i = torch.empty((rows, self._k), dtype=torch.int64, device=dev) if return_neighbours else None
        check(lib.hb_exchange_merge_transfer(self._h, ptr(label_table), trows, C, int(patch_pixels), ptr(qnorm_slice),
                                             float(beta), ptr(out_s), ptr(out_i), ptr(lh), stream_ptr(dev)))
        return lh, out_s, out_i

    def merge(self):
        """(scores fp32 (rows, k), idx int64 (rows, k)) of this rank's slice of the last scatter."""
        rows = int(lib.hb_exchange_slice_rows(self._h))
        dev = torch.device("cuda", self.device)
        out_s = torch.empty((rows, self._k), dtype=torch.float32, device=dev)
        out_i = torch.empty((rows, self._k), dtype=torch.int64, device=dev)
        check(lib.hb_exchange_merge(self._h, ptr(out_s), ptr(out_i), stream_ptr(dev)))
        return out_s, out_i


def predict_score(label_hat: torch.Tensor, B: int, S: int, H: int, W: int, conf: Optional[torch.Tensor] = None,
                  y: Optional[torch.Tensor] = None, gt_u8: Optional[torch.Tensor] = None,
                  ignore_index: Optional[int] = None, return_pred: bool = False) -> Optional[torch.Tensor]:
    """Fused tail (hb_predict_score): mask decode + bilinear upsample + argmax + confusion update in one
    pass (hbird_eval.py:219,235-243; eval_metrics.py:73-109).  label_hat fp32 (B*S*S, C); ground truth as
    y fp32 id/255 (B,1,H,W)/(B,H,W) or gt_u8 uint8 (B,H,W); conf int64 (C, C) accumulated in place.
    Returns the uint8 (B, H, W) prediction map when return_pred, else None."""
    label_hat = _require_cuda(label_hat, "label_hat", torch.float32)
    C = label_hat.shape[-1]
    if label_hat.numel() != B * S * S * C:
        raise ValueError("label_hat must hold B*S*S rows")
    dev = label_hat.device
    if conf is not None:
        conf = _require_cuda(conf, "conf", torch.int64)
        if tuple(conf.shape) != (C, C):
            raise ValueError(f"conf must be ({C}, {C})")
        if y is None and gt_u8 is None:
            raise ValueError("scoring needs y or gt_u8")
    if y is not None:
        y = _require_cuda(y, "y", torch.float32)
        if y.numel() != B * H * W:
            raise ValueError(f"Shapes must match. Got gt={tuple(y.shape)}, pred={(B, H, W)}")
    if gt_u8 is not None:
        gt_u8 = _require_cuda(gt_u8, "gt", torch.uint8)
        if gt_u8.numel() != B * H * W:
            raise ValueError(f"Shapes must match. Got gt={tuple(gt_u8.shape)}, pred={(B, H, W)}")
    pred = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if (return_pred or conf is None) else None
    ig = -1 if ignore_index is None else int(ignore_index)
    check(lib.hb_predict_score(ptr(label_hat), B, S, C, H, W, ptr(y), ptr(gt_u8), ig, ptr(conf), ptr(pred),
                               stream_ptr(dev)))
    return pred


def merge_topk_transfer(shard_scores: torch.Tensor, shard_idx: torch.Tensor, label_table: torch.Tensor,
                        patch_pixels: int, qnorm: torch.Tensor, beta: float = 0.02, return_neighbours: bool = False):
    """K3 with K4a fused: (G, Q, k) gathered per-shard results -> label_hat (Q, C) (+ merged (scores, idx))."""
    shard_scores = _require_cuda(shard_scores, "shard_scores", torch.float32)
    shard_idx = _require_cuda(shard_idx, "shard_idx", torch.int64)
    label_table = _require_cuda(label_table, "label_table", torch.int16)
    qnorm = _require_cuda(qnorm, "qnorm", torch.float32)
    G, Q, k = shard_scores.shape
    rows, C = label_table.shape
    dev = shard_scores.device
    lh = torch.empty((Q, C), dtype=torch.float32, device=dev)
    out_s = torch.empty((Q, k), dtype=torch.float32, device=dev) if return_neighbours else None
    out_i = torch.empty((Q, k), dtype=torch.int64, device=dev) if return_neighbours else None
    check(lib.hb_merge_topk_transfer(ptr(shard_scores), ptr(shard_idx), G, Q, k, ptr(label_table), rows, C,
                                     int(patch_pixels), ptr(qnorm), float(beta), ptr(out_s), ptr(out_i), ptr(lh),
                                     stream_ptr(dev)))
    return lh, out_s, out_i


def merge_topk(shard_scores: torch.Tensor, shard_idx: torch.Tensor):
    """K3: (G, Q, k) gathered per-shard results -> (Q, k) global top-k."""
    shard_scores = _require_cuda(shard_scores, "shard_scores", torch.float32)
    shard_idx = _require_cuda(shard_idx, "shard_idx", torch.int64)
    G, Q, k = shard_scores.shape
    out_s = torch.empty((Q, k), dtype=torch.float32, device=shard_scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=shard_scores.device)
    check(lib.hb_merge_topk(ptr(shard_scores), ptr(shard_idx), G, Q, k, ptr(out_s), ptr(out_i),
                            stream_ptr(shard_scores.device)))
    return out_s, out_i


def label_transfer(label_table: torch.Tensor, patch_pixels: int, scores: torch.Tensor, idx: torch.Tensor,
                   qnorm: torch.Tensor, beta: float = 0.02) -> torch.Tensor:
    """K4a: (Q, C) label_hat from neighbour scores/indices (hbird_eval.py:575-609,611-637)."""
    label_table = _require_cuda(label_table, "label_table", torch.int16)
    scores = _require_cuda(scores, "scores", torch.float32)
    idx = _require_cuda(idx, "idx", torch.int64)
    qnorm = _require_cuda(qnorm, "qnorm", torch.float32)
    Q, k = scores.shape
    rows, C = label_table.shape
    out = torch.empty((Q, C), dtype=torch.float32, device=scores.device)
    check(lib.hb_label_transfer(ptr(label_table), rows, C, int(patch_pixels), ptr(scores), ptr(idx), ptr(qnorm),
                                Q, k, float(beta), ptr(out), stream_ptr(scores.device)))
    return out


def upsample_argmax(label_hat: torch.Tensor, B: int, S: int, H: int, W: int) -> torch.Tensor:
    """K4b: label_hat fp32 (B*S*S, C) -> uint8 (B, H, W) prediction (hbird_eval.py:235-243)."""
    label_hat = _require_cuda(label_hat, "label_hat", torch.float32)
    C = label_hat.shape[-1]
    if label_hat.numel() != B * S * S * C:
        raise ValueError("label_hat must hold B*S*S rows")
    out = torch.empty((B, H, W), dtype=torch.uint8, device=label_hat.device)
    check(lib.hb_upsample_argmax(ptr(label_hat), B, S, C, H, W, ptr(out), stream_ptr(label_hat.device)))
    return out


def decode_mask(y: torch.Tensor, remap_255_to_0: bool) -> torch.Tensor:
    """(y*255).long() as uint8, optional 255->0 (hbird_eval.py:219,309-310)."""
    y = _require_cuda(y, "y", torch.float32)
    out = torch.empty(y.shape, dtype=torch.uint8, device=y.device)
    check(lib.hb_decode_mask(ptr(y), y.numel(), int(remap_255_to_0), ptr(out), stream_ptr(y.device)))
    return out


def confusion_accumulate(conf: torch.Tensor, gt_u8: torch.Tensor, pred_u8: torch.Tensor,
                         ignore_index: Optional[int]) -> None:
    """K5: conf int64 (C_gt, C_pred) += histogram of (gt, pred) pairs (eval_metrics.py:73-109)."""
    conf = _require_cuda(conf, "conf", torch.int64)
    gt_u8 = _require_cuda(gt_u8, "gt", torch.uint8)
    pred_u8 = _require_cuda(pred_u8, "pred", torch.uint8)
    if gt_u8.numel() != pred_u8.numel():
        raise ValueError(f"Shapes must match. Got gt={tuple(gt_u8.shape)}, pred={tuple(pred_u8.shape)}")
    ig = -1 if ignore_index is None else int(ignore_index)
    check(lib.hb_confusion_accumulate(ptr(gt_u8), ptr(pred_u8), gt_u8.numel(), conf.shape[0], conf.shape[1], ig,
                                      ptr(conf), stream_ptr(conf.device)))
