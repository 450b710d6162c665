"""nn_method="b200" — the drop-in backend behind the reference's plugin interface.

Replaces NearestNeighborSearchFaiss (hbird/nn/search_faiss.py:6-90): same constructor shape
(feature_memory, n_neighbors, distance_measure, gpu_ids, **kwargs), same
find_nearest_neighbors(q, k) -> (indices ndarray int64 (Q,k), distances ndarray fp32 (Q,k))
sorted by descending inner product (distance_measure="dot_product", GpuIndexFlatIP) or by
ascending squared L2 distance ("l2"/"euclidean", GpuIndexFlatL2, search_faiss.py:43-48).
The index is an HBM-resident MemoryBank (bf16 rows for the
tcgen05 pass + fp32 rows for the exact re-rank); there is no CPU path.

Extra (device-resident) entry points used by the fused evaluator:
    search_device(q_cuda) -> (scores, idx, qnorm) CUDA tensors, no host round trip.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops
from .search_base import NearestNeighborSearchBase


class NearestNeighborSearchB200(NearestNeighborSearchBase):
    def __init__(self, feature_memory, n_neighbors: int = 30, distance_measure: str = "dot_product",
                 idx_shard: bool = False, use_fp16: bool = False, gpu_ids=None, k_prime: Optional[int] = None,
                 keep_f32: Optional[bool] = None, idx_offset: int = 0,
                 label_memory: Optional[torch.Tensor] = None, patch_pixels: int = 256,
                 bank: Optional[ops.MemoryBank] = None, cta_group: int = 0, max_chunks: int = 0,
                 renormalise: bool = False, **kwargs):
        """feature_memory: fp32 (N, d) unit-norm rows, CPU or CUDA (the reference hands over a CPU
        tensor, hbird_eval.py:178-182) — or None when a pre-built `bank` is adopted.
        idx_shard / use_fp16 / gpu_ids: the faiss backend's knobs (search_faiss.py:7).  This class is
        ONE index on ONE GPU: idx_shard is recorded for the engine, which lays a multi-GPU bank out
        as row shards or replicas with one process per GPU (hbird_b200.hbird_eval); use_fp16=True
        drops the fp32 copy of the rows (the re-rank then reads the bf16 rows).
        k_prime: candidates kept by the bf16 tensor-core pass (32, 64 or 128; default 64 for
        n_neighbors <= 32, else 128 — the best k_prime/2 are strict, include/hbird_b200.h).
        Unknown keyword arguments raise TypeError (the faiss backend ignores them silently)."""
        if kwargs:
            raise TypeError(f"NearestNeighborSearchB200 got unexpected keyword arguments {sorted(kwargs)}")
        self.n_neighbors = int(n_neighbors)
        self.idx_shard, self.use_fp16 = bool(idx_shard), bool(use_fp16)
        self.distance_measure = distance_measure.lower()
        if self.distance_measure not in ("dot_product", "l2", "euclidean"):
            # search_faiss.py:48 / search_scann.py:20
            raise ValueError(f"Unsupported distance measure: {self.distance_measure}")
        self.metric = "dot_product" if self.distance_measure == "dot_product" else "l2"
        if not torch.cuda.is_available():
            raise RuntimeError("No GPUs available for the b200 backend.")  # search_faiss.py:15-16
        n_gpus = torch.cuda.device_count()
        if gpu_ids is None:
            gpu_ids = [torch.cuda.current_device()]
        for g in gpu_ids:
            if g >= n_gpus or g < 0:
                raise ValueError(f"Invalid GPU ID: {g}. Available GPUs: 0-{n_gpus - 1}")  # :25
        if len(gpu_ids) != 1:
            raise ValueError("the b200 backend is one shard per process; launch one rank per GPU "
                             "(torchrun) for a sharded bank")
        self.gpu_id = int(gpu_ids[0])
        ops.device_check(self.gpu_id)
        if not 1 <= self.n_neighbors <= 128:
            raise ValueError(f"n_neighbors={n_neighbors} outside the b200 backend's range [1, 128]")
        if k_prime is None:
            k_prime = 64 if self.n_neighbors <= 32 else 128
        self.k_prime = int(k_prime)
        if self.k_prime not in (32, 64, 128) or self.n_neighbors > self.k_prime:
            raise ValueError(f"k_prime={k_prime} must be 32, 64 or 128 and >= n_neighbors={n_neighbors}")
        self.idx_offset = int(idx_offset)
        self.keep_f32 = (not self.use_fp16) if keep_f32 is None else bool(keep_f32)
        self._label_memory = label_memory
        self._patch_pixels = int(patch_pixels)
        self._renormalise = bool(renormalise)
        self._cfg = (int(cta_group), int(max_chunks))
        self.feature_memory = feature_memory
        self.device = torch.device("cuda", self.gpu_id)
        self.bank = bank
        if bank is None:
            if feature_memory is None:
                raise ValueError("either feature_memory or bank must be given")
            self.embed_d = feature_memory.size(1)
            self.index = self._initialize_index()
            self._add_features_to_index()
        else:
            if bank.metric != self.metric:
                raise ValueError(f"bank was built for metric {bank.metric!r}, not {self.metric!r}")
            self.embed_d = bank.d
            self.index = bank
        if not self.bank.finalized:
            self.bank.finalize()
        self.bank.configure_search(*self._cfg)

    # --- NearestNeighborSearchBase contract ---------------------------------------------
    def _initialize_index(self):
        n, d = self.feature_memory.shape
        c = self._label_memory.shape[1] if self._label_memory is not None else 1
        self.bank = ops.MemoryBank(d, c, self._patch_pixels, max(int(n), 1), self.gpu_id, self.keep_f32,
                                    metric=self.metric)
        return self.bank

    def _add_features_to_index(self):
        fm = self.feature_memory
        n = fm.shape[0]
        step = 1 << 20  # stream the host tensor through HBM in 1M-row slabs
        for a in range(0, n, step):
            f = fm[a:a + step].to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            if self._label_memory is not None:
                l = self._label_memory[a:a + step].to(self.device, dtype=torch.float32).contiguous()
            else:
                l = torch.ones((f.shape[0], 1), dtype=torch.float32, device=self.device)
            self.bank.append_soft(f, l, normalise=self._renormalise)

    def find_nearest_neighbors(self, q, k=None):
        """q: (Q, d) fp32 tensor (CPU or CUDA) or ndarray.  Returns host (indices, distances),
        in that order, as search_faiss.py:83-90."""
        if k is None:
            k = self.n_neighbors
        if isinstance(q, np.ndarray):
            q = torch.from_numpy(q)
        Q = q.shape[0]
        if q.is_cuda:
            q_dev = q.to(self.device, dtype=torch.float32).contiguous()
        else:
            # pinned staging buffers (grown on demand): asynchronous copies both ways, one
            # synchronisation per call instead of three pageable, synchronous transfers
            st = self._staging(Q, int(k))
            st["q"][:Q].copy_(q.to(torch.float32))
            q_dev = st["q_dev"][:Q]
            q_dev.copy_(st["q"][:Q], non_blocking=True)
        scores, idx, _ = self.search_device(q_dev, k)
        if q.is_cuda:
            return idx.cpu().numpy(), scores.cpu().numpy()
        st["idx"][:Q].copy_(idx, non_blocking=True)
        st["scores"][:Q].copy_(scores, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return st["idx"][:Q].numpy().copy(), st["scores"][:Q].numpy().copy()

    def _staging(self, Q: int, k: int):
        st = getattr(self, "_stage", None)
        if st is None or st["q"].shape[0] < Q or st["k"] != k:
            cap = max(Q, 1024)
            st = {"k": k, "q": torch.empty((cap, self.embed_d), dtype=torch.float32).pin_memory(),
                  "q_dev": torch.empty((cap, self.embed_d), dtype=torch.float32, device=self.device),
                  "idx": torch.empty((cap, k), dtype=torch.int64).pin_memory(),
                  "scores": torch.empty((cap, k), dtype=torch.float32).pin_memory()}
            self._stage = st
        return st

    # --- device-resident path -----------------------------------------------------------------
    def search_device(self, q_dev: torch.Tensor, k: Optional[int] = None):
        k = self.n_neighbors if k is None else int(k)
        return self.bank.search(q_dev, k, self.k_prime, self.idx_offset, return_qnorm=True)
