"""HbirdEvaluation / hbird_evaluation — host-side mirror of the reference's evaluation engine
(hbird/hbird_eval.py:54-722) for the dense nearest-neighbour hot path, device resident end to end.

Same constructor, `evaluate(...)` and `hbird_evaluation(...)` signatures and return types as the
reference.  What differs is where the work happens:

  reference (hbird_eval.py)                         here
  ------------------------------------------------  ---------------------------------------------
  :309-329 decode, patchify, one_hot.mean, norm,    K1  hb_decode_mask + hb_bank_append (fused
           D2H append, torch.cat                        normalise/cast/pack + class histogram)
  :182     faiss index build + index.add (H2D)      -   the bank already lives in HBM
  :628     features.cpu() -> faiss search -> host   K2  tcgen05 GEMM + fused top-k', K2b fp32
                                                        re-rank (+ NCCL all-gather, K3 merge)
  :632-636 index_select of neighbour features       -   not needed: cos = score/||q||
           and labels on the CPU                    K4a hb_label_transfer (gathers uint16 hists)
  :594-609 normalise, bmm, softmax, bmm
  :235-243 permute, F.interpolate, argmax           K4b hb_upsample_argmax
  :245-252 cat all pixels, one giant bincount       K5  hb_confusion_accumulate per batch
  :253     Hungarian mIoU on the C x C matrix       host (scipy), unchanged

ViT feature extraction and data loading stay in PyTorch.  There is no CPU fallback.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import distributed as hdist
from . import ops
from .models import FeatureExtractorSimple
from . import pipeline as hpipe
from .pipeline import EvalPipeline
from .registry import NN_BACKENDS, create_nn_backend
from .utils.eval_metrics import PredsmIoU

logger = logging.getLogger(__name__)

BETA = 0.02  # cross-attention temperature, hbird_eval.py:576


class HbirdEvaluation:
    """Build the patch memory bank from `train_loader`, then evaluate `val_loader` by kNN label
    transfer.  Loaders are any iterable of (x fp32 (B,3,H,W), y fp32 = class_id/255 (B,1,H,W)).

    Multi-GPU (torch.distributed initialised, one process per GPU) follows the reference's two faiss
    layouts, selected like there by nn_params["idx_shard"] (search_faiss.py:7,53-74):
      False (default) — replicas (IndexReplicas): every rank ends up with the whole bank and the
                        validation batches are dealt round-robin to the ranks; no data-path
                        collective, one all-reduce of the (C, C) matrix at the end;
      True            — row shards (IndexShards): rank r keeps the rows it built; every rank searches
                        every query against its shard, the shard results are exchanged (fused NVLink
                        exchange or NCCL all-gather + merge) and each rank post-processes its image
                        slice of the batch.  Features are extracted once: rank r runs the ViT on its
                        slice of the batch and the queries are all-gathered."""

    _B200_PARAMS = ("k_prime", "keep_f32", "gpu_ids", "exchange", "idx_shard", "use_fp16", "distance_measure",
                    "cta_group", "max_chunks", "balance_shards")
    _BALANCE_MIN_ROWS = 1 << 17  # below this a shard's search time says nothing about the GPU's speed

    def __init__(self, feature_extractor: torch.nn.Module, train_loader, num_classes: int,
                 n_neighbours: int = 30, augmentation_epoch: int = 1, device: torch.device | str = "cpu",
                 nn_method: str = "scann", nn_params: Optional[Dict[str, Any]] = None,
                 memory_size: Optional[int] = None, dataset_size: Optional[int] = None,
                 f_mem_p: Optional[str] = None, l_mem_p: Optional[str] = None) -> None:
        self.nn_params = dict(nn_params or {})
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("hbird_b200 runs on a CUDA device (sm_100); there is no CPU fallback. "
                               f"Got device={device!r}.")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        ops.device_check(self.device.index)
        if nn_method not in NN_BACKENDS:
            raise ValueError(f"Unsupported NN method. Choose from {set(NN_BACKENDS)}.")
        self.nn_method = nn_method
        self.feature_extractor = feature_extractor.to(self.device)
        self.feature_extractor.eval()
        self.augmentation_epoch = augmentation_epoch
        self.memory_size = memory_size
        self.n_neighbours = int(n_neighbours)
        self.num_classes = num_classes
        self.f_mem_p, self.l_mem_p = f_mem_p, l_mem_p
        self.num_sampled_features: Optional[int] = None
        self.rank, self.world = hdist.dist_info()
        self.idx_shard = bool(self.nn_params.get("idx_shard", False)) and self.world > 1
        if nn_method == "b200":
            unknown = sorted(set(self.nn_params) - set(self._B200_PARAMS))
            if unknown:  # the faiss backend swallows unknown names (search_faiss.py:7); a typo should not be silent
                raise TypeError(f"nn_params not understood by the b200 backend: {unknown} (known: {sorted(self._B200_PARAMS)})")
            if not 1 <= self.n_neighbours <= 128:
                raise ValueError(f"n_neighbours={n_neighbours} outside the b200 backend's range [1, 128]")
        elif self.world > 1:
            raise ValueError(f"nn_method={nn_method!r} is a single-process backend; multi-GPU runs need nn_method='b200'")
        # k' (candidates kept by the bf16 pass): the best k'/2 are strict, see include/hbird_b200.h
        default_kp = 64 if self.n_neighbours <= 32 else 128
        self.k_prime = int(self.nn_params.get("k_prime", default_kp))
        if nn_method == "b200" and self.n_neighbours > self.k_prime // 2:
            logger.warning("n_neighbours=%d > k_prime/2=%d: neighbours beyond rank %d of the bf16 pass are covered "
                           "statistically, not strictly", self.n_neighbours, self.k_prime // 2, self.k_prime // 2)
        self.keep_f32 = bool(self.nn_params.get("keep_f32", not self.nn_params.get("use_fp16", False)))

        S = self.feature_extractor.eval_spatial_resolution
        if self.memory_size is not None:
            if dataset_size is None:
                raise ValueError("dataset_size must be provided when memory_size is set.")
            denom = dataset_size * self.augmentation_epoch
            self.num_sampled_features = max(1, self.memory_size // max(1, denom))
            logger.info("Bounded memory: memory_size=%d => %d sampled patches per image",
                        self.memory_size, self.num_sampled_features)

        self.bank: Optional[ops.MemoryBank] = None
        self._create_memory(train_loader, num_classes, S)
        self._save_memory()
        self._create_nn(self.n_neighbours, nn_method=self.nn_method, **self.nn_params)
        self._balance_shards()

    @classmethod
    def from_bank(cls, feature_extractor: torch.nn.Module, bank: "ops.MemoryBank", num_classes: int,
                  n_neighbours: int = 30, device: torch.device | str = "cuda",
                  nn_params: Optional[Dict[str, Any]] = None, shard_counts=None) -> "HbirdEvaluation":
        """An evaluator around a memory bank that already lives in HBM (built with ops.MemoryBank, or
        kept from an earlier run): `evaluate()` behaves exactly as after a constructor that built the
        bank from a loader.  With a row-sharded bank (torch.distributed initialised, nn_params
        idx_shard=True) `bank` is this rank's shard and the label table is replicated here."""
        self = cls.__new__(cls)
        self.nn_params = dict(nn_params or {})
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("hbird_b200 runs on a CUDA device (sm_100); there is no CPU fallback.")
        if self.device.index is None:
            self.device = torch.device("cuda", bank.device)
        unknown = sorted(set(self.nn_params) - set(cls._B200_PARAMS))
        if unknown:
            raise TypeError(f"nn_params not understood by the b200 backend: {unknown}")
        self.nn_method = "b200"
        self.feature_extractor = feature_extractor.to(self.device)
        self.feature_extractor.eval()
        self.augmentation_epoch, self.memory_size, self.num_sampled_features = 1, None, None
        self.n_neighbours, self.num_classes = int(n_neighbours), num_classes
        self.f_mem_p = self.l_mem_p = None
        self.rank, self.world = hdist.dist_info()
        self.idx_shard = bool(self.nn_params.get("idx_shard", False)) and self.world > 1
        self.k_prime = int(self.nn_params.get("k_prime", 64 if self.n_neighbours <= 32 else 128))
        self.keep_f32 = True
        self.bank = bank
        if not bank.finalized:
            bank.finalize()
        if self.idx_shard:
            counts = list(shard_counts) if shard_counts is not None else hdist.gather_counts(bank.rows, self.device)
            self.idx_offset = hdist.offsets_from_counts(counts)[self.rank]
            self.label_table = hdist.all_gather_rows(bank.label_table(), counts)
        else:
            counts, self.idx_offset = [bank.rows], 0
            self.label_table = bank.label_table()
        self.shard_counts, self.total_rows = counts, sum(counts)
        self._create_nn(self.n_neighbours, nn_method="b200", **self.nn_params)
        return self

    # ------------------------------------------------------------------ bank construction
    def _capacity_rows(self, loader_len: Optional[int], first_batch: int, S: int) -> int:
        """Rows this rank can be asked to hold: its share of the loader's batches over ALL
        augmentation epochs (the batch counter runs on across epochs) times the rows per batch."""
        per_image = S * S if self.memory_size is None else int(self.num_sampled_features)
        if loader_len is None:
            if self.memory_size is None:
                raise ValueError("train_loader must define __len__ when memory_size is None")
            return int(self.memory_size)
        my_batches = (self.augmentation_epoch * loader_len - self.rank + self.world - 1) // self.world
        cap = max(1, my_batches * first_batch * per_image)
        return cap if self.memory_size is None else min(cap, int(self.memory_size))

    @torch.no_grad()
    def _create_memory(self, train_loader, num_classes: int, eval_spatial_resolution: int) -> int:
        """hbird_eval.py:283-369.  With world_size > 1 rank r takes batches r, r+W, ... and builds the
        rows they produce; with idx_shard it keeps them (row-sharded bank), otherwise the shards are
        then replicated to every rank."""
        S = eval_spatial_resolution
        d = self.feature_extractor.d_model
        loader_len = len(train_loader) if hasattr(train_loader, "__len__") else None
        step = 0
        error: Optional[BaseException] = None
        try:
            for _ in range(self.augmentation_epoch):
                for x, y in train_loader:
                    mine = (step % self.world) == self.rank
                    step += 1
                    if not mine:
                        continue
                    x = x.to(self.device)
                    y = y.to(self.device, dtype=torch.float32)
                    B, _, H, W = x.shape
                    ps = x.shape[-1] // S
                    if H != S * ps or W != S * ps:
                        raise ValueError(f"input {H}x{W} is not eval_spatial_resolution*patch = {S}*{ps}")
                    feats, _ = self.feature_extractor.forward_features(x)
                    feats = feats.to(torch.float32).contiguous()
                    mask = ops.decode_mask(y.contiguous(), True).view(B, H, W)
                    if self.bank is None:
                        cap = self._capacity_rows(loader_len, B, S)
                        self.bank = ops.MemoryBank(d, num_classes, ps * ps, cap, self.device.index, self.keep_f32)
                        self._ps = ps
                    if self.memory_size is None:
                        self.bank.append(feats, mask, S, ps)
                    else:
                        sel = self._sample_patches(mask, S, ps, num_classes)
                        room = self.bank.capacity - self.bank.rows
                        if sel.numel() > room:  # the reference's slice assignment would raise here too
                            raise ValueError("memory_size exhausted before the training set was consumed")
                        self.bank.append(feats, mask, S, ps, sel)
            if self.bank is None:
                raise ValueError("train_loader yielded no batches for this rank")
            self.bank.finalize()
        except Exception as e:  # noqa: BLE001 - re-raised below, on every rank
            error = e
        # one rank's failure must not leave the others waiting in the collectives below
        if not hdist.all_ranks_ok(error is None, self.device):
            if error is not None:
                raise error
            raise RuntimeError("memory-bank construction failed on another rank")
        counts = hdist.gather_counts(self.bank.rows, self.device)
        if self.world > 1 and not self.idx_shard:
            self._replicate_bank(counts)
            counts = [self.bank.rows]
            self.idx_offset = 0
        else:
            self.idx_offset = hdist.offsets_from_counts(counts)[self.rank]
        self.shard_counts = counts
        self.total_rows = sum(counts)
        # the label table is indexed by global row: replicated when the bank is sharded
        self.label_table = hdist.all_gather_rows(self.bank.label_table(), counts) if self.idx_shard \
            else self.bank.label_table()
        logger.info("Memory bank: %d rows on this rank, %d total, d=%d (%s)", self.bank.rows, self.total_rows, d,
                    "row shards" if self.idx_shard else ("replicas" if self.world > 1 else "single GPU"))
        return self.bank.rows

    def _replicate_bank(self, counts, slab: int = 1 << 18) -> None:
        """Replicas layout (search_faiss.py:65-74): every rank receives every shard's packed rows
        (slab-wise broadcasts of the fp32 unit rows and soft labels) and re-packs them, so all ranks
        hold the same bank in the same global row order (rank-major)."""
        import torch.distributed as dist

        d, C = self.bank.d, self.num_classes
        full = ops.MemoryBank(d, C, self.bank.patch_pixels, max(1, sum(counts)), self.device.index, self.keep_f32)
        for r, n in enumerate(counts):
            for a in range(0, n, slab):
                m = min(slab, n - a)
                if self.rank == r:
                    f, l = self.bank.export(a, m)
                else:
                    f = torch.empty((m, d), dtype=torch.float32, device=self.device)
                    l = torch.empty((m, C), dtype=torch.float32, device=self.device)
                hdist.broadcast(f, r)
                hdist.broadcast(l, r)
                full.append_soft(f, l, normalise=False)
        full.finalize()
        self.bank.close()
        self.bank = full

    def _balance_shards(self) -> None:
        """Row shards sized by measured search speed (nn_params["balance_shards"], default on).  With a
        sharded bank every batch ends with everybody's shard results, so the evaluation runs at the pace
        of the slowest GPU; GPUs under the same power cap differ by a few per cent.  Each rank times the
        search kernel on its shard, the times are all-gathered, and rows migrate between neighbouring
        ranks until all shards take the same time (distributed.balanced_counts; at most +-10 %)."""
        if not self.idx_shard or not bool(self.nn_params.get("balance_shards", True)):
            return
        if min(self.shard_counts) < self._BALANCE_MIN_ROWS:
            return
        times = hdist.gather_floats(self.bank.calibrate_search_ms(k_prime=self.k_prime), self.device)
        new_counts = hdist.balanced_counts(self.shard_counts, times)
        if max(abs(n - o) / o for n, o in zip(new_counts, self.shard_counts)) < 0.01:
            return  # already even within a per cent
        logger.info("Balancing shards by search speed: %s ms -> rows %s", [round(t, 2) for t in times], new_counts)
        self.rebalance(new_counts)

    def rebalance(self, new_counts) -> None:
        """Re-shard the bank to `new_counts` rows per rank (same total).  Global row order is rank-major and
        contiguous before and after, so global row ids — and the replicated label table — do not change:
        rows only move between ranks.  Every rank rebuilds its shard from the pieces of the old shards
        that fall into its new range (exported as fp32 unit rows + soft labels, re-packed bit-identically)."""
        import torch.distributed as dist

        old = list(self.shard_counts)
        new = [int(c) for c in new_counts]
        if len(new) != self.world or sum(new) != sum(old) or min(new) < 1:
            raise ValueError(f"rebalance: {new} must hold {self.world} positive counts summing to {sum(old)}")
        if new == old:
            return
        old_off, new_off = hdist.offsets_from_counts(old) + [sum(old)], hdist.offsets_from_counts(new) + [sum(new)]
        d, C = self.bank.d, self.num_classes
        via = self.device if dist.get_backend() == "nccl" else torch.device("cpu")
        fresh = ops.MemoryBank(d, C, self.bank.patch_pixels, new[self.rank], self.device.index, self.keep_f32)
        slab = 1 << 18
        for src in range(self.world):      # same order on every rank: matched blocking send / recv pairs,
            for dst in range(self.world):  # and for a given dst the pieces arrive in ascending global rows
                lo, hi = max(old_off[src], new_off[dst]), min(old_off[src + 1], new_off[dst + 1])
                if hi <= lo or self.rank not in (src, dst):
                    continue
                for a in range(lo, hi, slab):
                    m = min(slab, hi - a)
                    if self.rank == src:
                        f, l = self.bank.export(a - old_off[src], m)
                        if dst != src:
                            dist.send(f.to(via), dst=dst)
                            dist.send(l.to(via), dst=dst)
                    if self.rank == dst:
                        if dst != src:
                            f = torch.empty((m, d), dtype=torch.float32, device=via)
                            l = torch.empty((m, C), dtype=torch.float32, device=via)
                            dist.recv(f, src=src)
                            dist.recv(l, src=src)
                        fresh.append_soft(f.to(self.device), l.to(self.device), normalise=False)
        fresh.finalize()
        self.bank.close()
        self.bank = fresh
        self.shard_counts = new
        self.idx_offset = new_off[self.rank]
        self.__dict__.pop("_export_cache", None)
        self._create_nn(self.n_neighbours, nn_method=self.nn_method, **self.nn_params)

    def _sample_patches(self, mask: torch.Tensor, S: int, ps: int, num_classes: int) -> torch.Tensor:
        """Bounded-memory sampler, hbird_eval.py:447-517: per image keep the K patches with the
        smallest score*U(0,1) (hb_sample_patches).  U is drawn with the CPU generator in image order
        exactly as the reference does (:497-508: one value per NON-EMPTY patch, i.e. per patch that
        holds a label in [0, C); a patch of out-of-range labels only consumes no draw and scores 1e6)
        and uploaded, so a seeded run picks the same patches.
        Returns int32 flat source rows (b*S*S + patch) on the device, ascending score per image."""
        B = mask.shape[0]
        K = int(self.num_sampled_features)
        if K > S * S:
            raise ValueError(f"memory_size asks for {K} patches per image but an image has only {S * S}")
        uniform = torch.ones(B * S * S)  # empty patches: 1e6 * 1, as the reference (:494-506)
        if num_classes >= 256:  # uint8 ids are always < C: every patch is non-empty
            uniform = torch.rand(B * S * S)
        else:
            nonempty = (mask.view(B, S, ps, S, ps) < num_classes).any(dim=4).any(dim=2).reshape(-1).cpu()
            n = int(nonempty.sum())
            uniform[nonempty] = torch.rand(n)
        return ops.sample_patches(mask, S, ps, num_classes, uniform.to(mask.device, non_blocking=True), K)

    def _save_memory(self) -> None:
        """hbird_eval.py:371-378 — same on-disk format: fp32 (N,d) and (N,C) tensors.  With a sharded
        bank rank 0 collects the shards (in global row order) and writes the single pair of files."""
        if self.f_mem_p is None and self.l_mem_p is None:
            return
        want_f, want_l = self.f_mem_p is not None, self.l_mem_p is not None
        if not self.idx_shard:
            f = l = None
            if self.rank == 0:
                f, l = self.bank.export(features=want_f, labels=want_l)
                f, l = (f.cpu() if want_f else None), (l.cpu() if want_l else None)
        else:
            f, l = self._collect_memory_on_rank0(want_f, want_l)
        if self.rank == 0:
            if want_f:
                torch.save(f, self.f_mem_p)
            if want_l:
                torch.save(l, self.l_mem_p)
        if self.world > 1:
            torch.distributed.barrier()

    def _collect_memory_on_rank0(self, want_f: bool, want_l: bool, slab: int = 1 << 18):
        """Stream every shard to rank 0 in slabs (point-to-point over NCCL); rank 0 returns the CPU
        tensors (total_rows, d) / (total_rows, C) in global row order, the others (None, None)."""
        import torch.distributed as dist

        d, C = self.bank.d, self.num_classes
        fs, ls = [], []
        # point-to-point transfers of device tensors need NCCL; other backends (gloo) go through the host
        via = self.device if dist.get_backend() == "nccl" else torch.device("cpu")
        for r, n in enumerate(self.shard_counts):
            for a in range(0, n, slab):
                m = min(slab, n - a)
                if self.rank == r:
                    f, l = self.bank.export(a, m, features=want_f, labels=want_l)
                    if r != 0:
                        if want_f:
                            dist.send(f.to(via), dst=0)
                        if want_l:
                            dist.send(l.to(via), dst=0)
                elif self.rank == 0:
                    f = torch.empty((m, d), dtype=torch.float32, device=via) if want_f else None
                    l = torch.empty((m, C), dtype=torch.float32, device=via) if want_l else None
                    if want_f:
                        dist.recv(f, src=r)
                    if want_l:
                        dist.recv(l, src=r)
                if self.rank == 0:
                    if want_f:
                        fs.append(f.cpu())
                    if want_l:
                        ls.append(l.cpu())
        if self.rank != 0:
            return None, None
        return (torch.cat(fs) if want_f else None), (torch.cat(ls) if want_l else None)

    def load_memory(self) -> bool:
        """hbird_eval.py:380-400 — reload the tensors written by _save_memory (fp32 (N, d) unit rows and
        (N, C) soft labels) and rebuild the HBM bank and the search backend from them.  With a sharded
        bank every rank takes its contiguous row range of the files; replicas load all of them."""
        import os

        if not (self.f_mem_p and self.l_mem_p and os.path.isfile(self.f_mem_p) and os.path.isfile(self.l_mem_p)):
            logger.warning("Memory files not found or paths not provided; skipping load.")
            return False
        f = torch.load(self.f_mem_p, mmap=True)
        l = torch.load(self.l_mem_p, mmap=True)
        if f.shape[0] != l.shape[0]:
            raise ValueError(f"feature memory has {f.shape[0]} rows, label memory {l.shape[0]}")
        a, b = hdist.shard_bounds(f.shape[0], self.world, self.rank) if self.idx_shard else (0, f.shape[0])
        pp = self.bank.patch_pixels
        new_bank = ops.MemoryBank(f.shape[1], l.shape[1], pp, max(1, b - a), self.device.index, self.keep_f32)
        step = 1 << 20
        for s0 in range(a, b, step):
            s1 = min(b, s0 + step)
            new_bank.append_soft(f[s0:s1].to(self.device, dtype=torch.float32).contiguous(),
                                 l[s0:s1].to(self.device, dtype=torch.float32).contiguous(), normalise=False)
        new_bank.finalize()
        self.bank.close()
        self.bank = new_bank
        if self.idx_shard:
            counts = hdist.gather_counts(new_bank.rows, self.device)
            self.idx_offset = hdist.offsets_from_counts(counts)[self.rank]
            self.label_table = hdist.all_gather_rows(new_bank.label_table(), counts)
        else:
            counts, self.idx_offset = [new_bank.rows], 0
            self.label_table = new_bank.label_table()
        self.shard_counts, self.total_rows = counts, sum(counts)
        self.__dict__.pop("_export_cache", None)
        self._create_nn(self.n_neighbours, nn_method=self.nn_method, **self.nn_params)
        self._balance_shards()
        return True

    @property
    def feature_memory(self) -> torch.Tensor:
        """The reference's feature_memory (N, d) fp32 CPU tensor, materialised on demand (this
        rank's rows when the bank is sharded)."""
        return self.bank.export(labels=False)[0].cpu()

    @property
    def label_memory(self) -> torch.Tensor:
        return self.bank.export(features=False)[1].cpu()

    def _create_nn(self, n_neighbours: int = 30, nn_method: str = "b200", **kwargs) -> None:
        """hbird_eval.py:267-281, through the registry."""
        if nn_method == "b200":
            kw = {k: v for k, v in kwargs.items() if k in ("cta_group", "max_chunks", "idx_shard", "use_fp16")}  # engine-level names stay here
            measure = str(kwargs.get("distance_measure", "dot_product")).lower()
            if measure not in ("dot_product", "l2", "euclidean"):
                raise ValueError(f"Unsupported distance measure: {measure}")  # search_faiss.py:48
            # Bank rows are unit-norm (hbird_eval.py:324), so ||q-x||^2 = ||q||^2 + 1 - 2 q.x ranks
            # exactly as the inner product does, and the reference drops the distances (:628):
            # "l2" therefore runs the inner-product index here.  The plugin itself implements true
            # squared-L2 for callers that hand it un-normalised rows.
            self.NN_algorithm = create_nn_backend(
                "b200", None, n_neighbors=n_neighbours, bank=self.bank, k_prime=self.k_prime,
                idx_offset=self.idx_offset, gpu_ids=[self.device.index], **kw)
        else:
            # legacy / third-party plugins take the reference's CPU feature tensor
            measure = str(kwargs.get("distance_measure", "dot_product")).lower()
            if measure != "dot_product":
                # the reference ignores the backend's distances and re-derives cosines (:594-609); the
                # ranking of a non-IP backend over unit rows is the same, so only IP is wired here
                raise ValueError("legacy backends are driven with distance_measure='dot_product' only")
            self.NN_algorithm = create_nn_backend(nn_method, self.feature_memory, n_neighbors=n_neighbours, **kwargs)

    # ------------------------------------------------------------------ evaluation
    def _legacy_neighbours(self, q: torch.Tensor):
        """faiss / scann / third-party plugin: host round trip as in the reference (:624-629).  The
        backend's distances are NOT used (they may be approximate or another metric): exact inner
        products with the bank rows are recomputed on the device, as _cross_attention does."""
        idx_np, _ = self.NN_algorithm.find_nearest_neighbors(q.cpu())
        idx = torch.as_tensor(np.asarray(idx_np).astype("int64"), device=self.device)
        if not hasattr(self, "_export_cache"):
            self._export_cache = self.bank.export(labels=False)[0]
        rows = self._export_cache.index_select(0, idx.reshape(-1).clamp_min(0)).view(idx.shape[0], idx.shape[1], -1)
        scores = torch.einsum("qkd,qd->qk", rows, q)
        scores = torch.where(idx >= 0, scores, torch.full_like(scores, float("-inf")))
        return scores, idx, torch.linalg.vector_norm(q, dim=1)

    def _exchange_for(self, n_queries: int, n_images: int):
        """The ShardExchange, (re)built when a batch needs a larger window.  Every rank sees the
        same batches, so every rank takes the same decision here."""
        mode = str(self.nn_params.get("exchange", "p2p")).lower()
        if mode not in ("p2p", "p2p_full", "nccl"):
            raise ValueError(f"nn_params['exchange']={mode!r} must be 'p2p', 'p2p_full' or 'nccl'")
        if mode == "nccl" or getattr(self, "_xchg_failed", False):
            return None
        per_img = n_queries // n_images
        need = -(-n_images // self.world) * per_img
        xchg = getattr(self, "_xchg", None)
        if xchg is None or xchg.slice_capacity < need or xchg.max_k < self.n_neighbours:
            if xchg is not None:
                # peers still map the old window: keep it alive instead of freeing it under them
                self._retired_xchg = getattr(self, "_retired_xchg", []) + [xchg]
            # 'p2p': threshold exchange (shards re-rank only what can reach the global top-k');
            # 'p2p_full': every shard re-ranks its whole top-k' (one hop less, G-fold redundant gathers)
            self._xchg = hdist.connect_shard_exchange(need, self.n_neighbours, self.device,
                                                      threshold_exchange=(mode == "p2p"))
            if self._xchg is None:
                logger.warning("peer-memory exchange unavailable; using NCCL all-gather + merge")
                self._xchg_failed = True
        return self._xchg

    def close(self) -> None:
        """Release the HBM bank and, collectively on all ranks, the shard-exchange windows."""
        for x in getattr(self, "_retired_xchg", []) + [getattr(self, "_xchg", None)]:
            hdist.close_shard_exchange(x)
        self._retired_xchg, self._xchg = [], None
        if self.bank is not None:
            self.bank.close()
            self.bank = None

    def _prefetched(self, loader):
        """(step, x, y) with batch i+1's host->device copies issued on a side stream while batch i is
        being evaluated (asynchronous when the loader yields pinned tensors, as a DataLoader with
        pin_memory=True does).  Only the image slice / batches this rank works on are copied."""
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(self.device)
        replicas = self.world > 1 and not self.idx_shard

        def stage(step, x, y):
            B = x.shape[0]
            b0, b1 = hdist.split_range(B, self.world, self.rank) if self.idx_shard else (0, B)
            main = torch.cuda.current_stream(self.device)
            with torch.cuda.stream(copy_stream):
                xd = x[b0:b1].to(self.device, non_blocking=True)
                yd = y[b0:b1].to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            for t in (xd, yd):
                t.record_stream(main)
            return step, B, b0, b1, xd, yd, ev

        pending = None
        for step, (x, y) in enumerate(loader):
            if replicas and step % self.world != self.rank:
                continue  # replicas: the batches are dealt round-robin, nothing is exchanged
            nxt = stage(step, x, y)
            if pending is not None:
                yield pending
            pending = nxt
        if pending is not None:
            yield pending

    def _features(self, x: torch.Tensor) -> torch.Tensor:
        feats, _ = self.feature_extractor.forward_features(x)
        return feats.to(torch.float32).contiguous()

    def _gathered_queries(self, x_slice: torch.Tensor, B: int, b0: int, b1: int) -> torch.Tensor:
        """Row-sharded bank: features are extracted once across the ranks — each rank runs the extractor
        on its image slice [b0, b1) (x_slice, already on the device) and the query rows of the whole
        batch are all-gathered, because every shard must see every query."""
        img_counts = [hdist.split_range(B, self.world, r) for r in range(self.world)]
        if b1 > b0:
            mine = self._features(x_slice)
            N, d = mine.shape[1], mine.shape[2]
        else:
            N, d = self.feature_extractor.eval_spatial_resolution ** 2, self.feature_extractor.d_model
            mine = torch.empty((0, N, d), dtype=torch.float32, device=self.device)
        return hdist.all_gather_rows(mine.view(-1, d), [(e - a) * N for a, e in img_counts])

    def _sharded_batch(self, x_slice: torch.Tensor, B: int, b0: int, b1: int, want_neighbours: bool):
        """Row-sharded bank: (label_hat, scores, idx, q) for the image slice [b0, b1) of the batch this
        rank post-processes.  Features are extracted once across the ranks: each rank runs the
        extractor on its slice (x_slice, already on the device) and the query rows are all-gathered
        (every shard must see every query).  Then K2/K2b per shard -> exchange -> merge with the
        label transfer fused in."""
        q = self._gathered_queries(x_slice, B, b0, b1)
        N = q.shape[0] // B
        k, kp = self.n_neighbours, self.k_prime
        pp = self.bank.patch_pixels
        if self._exchange_for(B * N, B) is not None:
            qsplit = hdist.query_split(B, N, self.world)
            qn = self._xchg.search_scatter(self.bank, q, qsplit, k, kp, self.idx_offset)
            lh, s, i = self._xchg.merge_transfer(self.label_table, pp, qn[b0 * N:b1 * N].contiguous(), BETA, want_neighbours)
        else:
            s, i, qn = self.bank.search(q, k, kp, self.idx_offset)
            gs, gi = hdist.all_gather_topk(s, i)
            sl = slice(b0 * N, b1 * N)
            lh, s, i = ops.merge_topk_transfer(gs[:, sl].contiguous(), gi[:, sl].contiguous(), self.label_table, pp,
                                               qn[sl].contiguous(), BETA, want_neighbours)
        return lh, s, i, q

    @torch.no_grad()
    def evaluate(self, val_loader, eval_spatial_resolution: int, return_knn_details: bool = False,
                 ignore_index: int = 255):
        """hbird_eval.py:184-265.  Returns the mIoU (Python float in [0, 1]) or (mIoU, details)."""
        S = eval_spatial_resolution
        C = self.num_classes
        metric = PredsmIoU(C, C, device=self.device, ignore_index=ignore_index, store_reordered_preds=False)
        conf = metric.confusion_buffer()
        details = []  # per batch: (batch number, knns, knns_labels, knns_ca_labels) CPU tensors
        replicas = self.world > 1 and not self.idx_shard
        # Banks (or shards) of up to ~8 M rows per GPU go through a two-stream pipeline (pipeline.py):
        # the tensor-core search of batch i+1 is issued beside the HBM-bound post-processing of batch i,
        # which fills the search kernel's ramp-down and the launch gaps.  Larger ones, and
        # return_knn_details (every batch's neighbours go to the host), take the one-call-per-batch path.
        pipe = None
        if self.nn_method == "b200" and not return_knn_details and hpipe.worthwhile(self.bank):
            if getattr(self, "_pipe_streams", None) is None:
                self._pipe_streams = hpipe.make_streams(self.device)
            pipe = EvalPipeline(self.bank, self.label_table, S, conf, ignore_index, self.n_neighbours, self.k_prime, BETA,
                                self.idx_offset, self.world if self.idx_shard else 1, self.rank, None,
                                streams=self._pipe_streams)
        try:
            for step, B, b0, b1, x, ys, copied in self._prefetched(val_loader):
                torch.cuda.current_stream(self.device).wait_event(copied)
                h, w = int(ys.shape[-2]), int(ys.shape[-1])  # the mask has the input's spatial size (:219,:240)
                if pipe is not None:
                    if self.idx_shard:
                        q = self._gathered_queries(x, B, b0, b1)
                        pipe.xchg = self._exchange_for(q.shape[0], B)
                    else:
                        feats = self._features(x)
                        q = feats.view(-1, feats.shape[2])
                    pipe.submit(q, ys, B)
                    continue
                if self.idx_shard:
                    lh, s, i, q = self._sharded_batch(x, B, b0, b1, return_knn_details)
                    if b1 > b0:
                        ops.predict_score(lh, b1 - b0, S, h, w, conf, y=ys, ignore_index=ignore_index)
                    N, d = q.shape[0] // B, q.shape[1]
                else:
                    feats = self._features(x)
                    N, d = feats.shape[1], feats.shape[2]
                    q = feats.view(B * N, d)
                    if self.nn_method != "b200":
                        s, i, qn = self._legacy_neighbours(q)
                        lh = ops.label_transfer(self.label_table, self.bank.patch_pixels, s, i, qn, BETA)
                        ops.predict_score(lh, B, S, h, w, conf, y=ys, ignore_index=ignore_index)
                    elif return_knn_details:
                        lh, _, s, i = self.bank.search_transfer(q, self.n_neighbours, self.k_prime, 0, BETA, None, True)
                        ops.predict_score(lh, B, S, h, w, conf, y=ys, ignore_index=ignore_index)
                    else:  # the whole batch in one call: prep, K2, K2b+K4a, fused tail
                        self.bank.eval_step(q, ys, S, conf, ignore_index, self.n_neighbours, self.k_prime, BETA)
                if return_knn_details:
                    k = self.n_neighbours
                    kf, kl, lhd = self._gather_details(i, lh, B, N)
                    details.append((step, kf.view(-1, N, k, d).cpu(), kl.view(-1, N, k, C).cpu(), lhd.view(-1, N, C).cpu()))
        except BaseException:
            if pipe is not None:
                pipe.abort()  # the bank stays usable after a failed evaluation
            raise
        if pipe is not None:
            pipe.flush()
        if self.idx_shard and getattr(self, "_xchg", None) is not None:
            self._xchg.check_status()  # raises if a merge gave up waiting for a peer (results not written)
        jac, tp, fp, fn, _, _ = metric.compute(is_global_zero=True, sync_distributed=self.world > 1,
                                               return_reordered=False)
        self.last_confusion = metric.confusion_matrix()
        if return_knn_details:
            if replicas:  # every rank returns the whole validation set, in batch order, as the reference does
                gathered = [None] * self.world
                torch.distributed.all_gather_object(gathered, details)
                details = sorted((t for part in gathered for t in part), key=lambda t: t[0])
            return jac, {"knns": torch.cat([t[1] for t in details]), "knns_labels": torch.cat([t[2] for t in details]),
                         "knns_ca_labels": torch.cat([t[3] for t in details])}
        return jac

    def _gather_details(self, idx: torch.Tensor, label_hat: torch.Tensor, n_images: int, per_img: int):
        """return_knn_details support (hbird_eval.py:229-232,255-262): neighbour features, neighbour
        soft labels and label_hat for the WHOLE batch, as the reference returns them.  With a sharded
        bank every rank contributes the feature rows it owns (one all-reduce of the (Q, k, d) tensor:
        exactly one rank holds each row, the others add zeros) and the slices of idx / label_hat are
        all-gathered; the soft labels come from the replicated label table."""
        if self.idx_shard:
            counts = [hdist.split_range(n_images, self.world, r) for r in range(self.world)]
            counts = [(b - a) * per_img for a, b in counts]
            idx = hdist.all_gather_rows(idx, counts)
            label_hat = hdist.all_gather_rows(label_hat, counts)
        if not hasattr(self, "_export_cache"):
            self._export_cache = self.bank.export(labels=False)[0]
        f = self._export_cache
        flat = idx.reshape(-1)
        local = flat - self.idx_offset
        mine = (local >= 0) & (local < self.bank.rows)
        kf = f.index_select(0, local.clamp(0, self.bank.rows - 1)) * mine.unsqueeze(1).to(f.dtype)
        if self.idx_shard:
            hdist.all_reduce_sum(kf)
        table = torch.as_tensor(self.label_table, device=self.device)
        kl = table.index_select(0, flat.clamp_min(0)).to(torch.float32) / float(self.bank.patch_pixels)
        return kf, kl, label_hat


def hbird_evaluation(model, d_model: int, patch_size: int, dataset_name, data_dir: str, batch_size: int = 64,
                     input_size: int = 224, augmentation_epoch: int = 1, device: str | torch.device = "cpu",
                     return_knn_details: bool = False, n_neighbours: int = 30, nn_method: str = "scann",
                     nn_params: Optional[Dict[str, Any]] = None, ftr_extr_fn=None,
                     memory_size: Optional[int] = None, num_workers: int = 8, ignore_index: int = 255,
                     train_fs_path: Optional[str] = None, val_fs_path: Optional[str] = None):
    """Same entry point as hbird_eval.py:640-722.  `dataset_name` is either a registered dataset
    name — resolved by the REFERENCE package's data layer (hbird.data / hbird.utils.transforms),
    which is out of scope here and must be installed for real datasets — or a datamodule-like
    object exposing get_train_dataset_size(), get_num_classes(), train_dataloader(),
    val_dataloader() and optionally `ignore_index` (e.g. hbird_b200.data.SyntheticSegmentationData)."""
    nn_params = dict(nn_params or {})
    S = input_size // patch_size
    if ftr_extr_fn is None:
        try:
            from hbird.models import FeatureExtractor  # reference's auto-detecting wrapper
        except ImportError as e:
            raise ImportError("ftr_extr_fn=None needs the reference package's hbird.models.FeatureExtractor; "
                              "pass ftr_extr_fn=(model, imgs) -> (features, None) instead") from e
        feature_extractor = FeatureExtractor(model, eval_spatial_resolution=S, d_model=d_model)
    else:
        feature_extractor = FeatureExtractorSimple(model, ftr_extr_fn=ftr_extr_fn, eval_spatial_resolution=S,
                                                   d_model=d_model)
    if isinstance(dataset_name, str):
        try:
            from hbird.data import get_dataset
            from hbird.utils.image_transformations import CombTransforms
            from hbird.utils.transforms import get_hbird_train_transforms, get_hbird_val_transforms
        except ImportError as e:
            raise ImportError("named datasets are loaded by the reference package's data layer "
                              f"(hbird.data), which is not importable here: {e}") from e
        tt, vt = get_hbird_train_transforms(input_size), get_hbird_val_transforms(input_size)
        train_tf = CombTransforms(img_transform=tt["img"], tgt_transform=None, img_tgt_transform=tt["shared"])
        val_tf = CombTransforms(img_transform=vt["img"], tgt_transform=None, img_tgt_transform=vt["shared"])
        dataset, ignore_index_local = get_dataset(dataset_name, data_dir, batch_size, num_workers, train_tf,
                                                  val_tf, train_fs_path, val_fs_path)
    else:
        dataset = dataset_name
        ignore_index_local = getattr(dataset, "ignore_index", 255)
    evaluator = HbirdEvaluation(
        feature_extractor, dataset.train_dataloader(), num_classes=dataset.get_num_classes(),
        n_neighbours=n_neighbours, augmentation_epoch=augmentation_epoch, device=device, nn_method=nn_method,
        nn_params=nn_params, memory_size=memory_size, dataset_size=dataset.get_train_dataset_size())
    effective_ignore = ignore_index if ignore_index != 255 else ignore_index_local  # hbird_eval.py:715
    return evaluator.evaluate(dataset.val_dataloader(), eval_spatial_resolution=S,
                              return_knn_details=return_knn_details, ignore_index=effective_ignore)
