"""Synthetic VOC/ADE-shaped inputs for tests, smoke and bench.

The reference's data layer (hbird/data/**) is out of scope; what crosses into the hot path is the
loader tensor contract: batches of (x fp32 (B,3,H,W), y fp32 = class_id/255 (B,1,H,W)) and a
feature extractor returning (B, S*S, d) fp32 patch features.  This module fabricates both,
deterministically, on the CPU:

  * label maps: `cells x cells` random class cells, nearest-upsampled to H x W, with ~2 % ignore
    pixels (255, or class 0 for ADE-shaped data);
  * patch features: f = hist(patch) @ P + sigma * eps, times a positive per-patch scale, with
    class prototypes P ~ N(0,1) — queries are therefore UN-normalised, as in the reference;
  * a table-lookup "ViT": image ids travel in x[:, 0, 0, 0]; `ftr_extr_fn(model, x)` returns the
    pre-computed features of those images on x's device, so the CUDA path and the CPU oracle see
    bit-identical features.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def make_label_maps(n: int, H: int, num_classes: int, ignore_index: int, cells: int, seed: int,
                    ignore_frac: float = 0.02, first_class: int = 0) -> torch.Tensor:
    """(n, H, H) uint8 class maps."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randint(first_class, num_classes, (n, cells, cells), generator=g)
    reps = (H + cells - 1) // cells
    full = coarse.repeat_interleave(reps, dim=1).repeat_interleave(reps, dim=2)[:, :H, :H]
    ign = torch.rand((n, H, H), generator=g) < ignore_frac
    full = torch.where(ign, torch.full_like(full, ignore_index), full)
    return full.to(torch.uint8)


def make_patch_features(maps: torch.Tensor, S: int, ps: int, d: int, num_classes: int, seed: int,
                        sigma: float = 0.8) -> torch.Tensor:
    """(n, S*S, d) fp32 raw patch features from the (n, H, H) class maps."""
    n = maps.shape[0]
    g0 = torch.Generator().manual_seed(0)  # prototypes are shared by bank and queries
    protos = torch.randn((num_classes, d), generator=g0)
    g = torch.Generator().manual_seed(seed)
    ids = maps.long().clamp_max(num_classes - 1)  # ignore pixels count as the last class here
    patches = ids.view(n, S, ps, S, ps).permute(0, 1, 3, 2, 4).reshape(n, S * S, ps * ps)
    hist = torch.zeros((n, S * S, num_classes)).scatter_add_(2, patches, torch.ones_like(patches, dtype=torch.float32))
    hist = hist / float(ps * ps)
    feats = hist @ protos + sigma * torch.randn((n, S * S, d), generator=g)
    scale = 3.7 * torch.exp(0.25 * torch.randn((n, S * S, 1), generator=g))
    return (feats * scale).contiguous()


class TableFeatureModel(torch.nn.Module):
    """Stand-in ViT: looks the features of image id x[b,0,0,0] up in a table."""

    def __init__(self, train_feats: torch.Tensor, val_feats: torch.Tensor):
        super().__init__()
        self.register_buffer("table", torch.cat([train_feats, val_feats]).contiguous())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ids = x[:, 0, 0, 0].round().long()
        return self.table.index_select(0, ids.to(self.table.device))


def table_ftr_extr_fn(model: torch.nn.Module, imgs: torch.Tensor):
    return model(imgs), None


class SyntheticSegmentationData:
    """Datamodule duck type (get_train_dataset_size / get_num_classes / train_dataloader /
    val_dataloader, cf. hbird_eval.py:694-699) over fabricated data."""

    def __init__(self, num_train: int = 16, num_val: int = 8, input_size: int = 224, patch_size: int = 16,
                 d_model: int = 384, num_classes: int = 21, batch_size: int = 8, ignore_index: int = 255,
                 cells: int = 8, seed: int = 0):
        self.S = input_size // patch_size
        self.ps, self.H, self.d, self.C = patch_size, self.S * patch_size, d_model, num_classes
        self.batch_size, self.ignore_index = batch_size, ignore_index
        first = 1 if ignore_index == 0 else 0
        self.train_maps = make_label_maps(num_train, self.H, num_classes, ignore_index, cells, 3 + seed, first_class=first)
        self.val_maps = make_label_maps(num_val, self.H, num_classes, ignore_index, cells, 103 + seed, first_class=first)
        self.train_feats = make_patch_features(self.train_maps, self.S, self.ps, d_model, num_classes, 1 + seed)
        self.val_feats = make_patch_features(self.val_maps, self.S, self.ps, d_model, num_classes, 2 + seed)
        self.model = TableFeatureModel(self.train_feats, self.val_feats)
        self.ftr_extr_fn = table_ftr_extr_fn

    def get_train_dataset_size(self) -> int:
        return self.train_maps.shape[0]

    def get_num_classes(self) -> int:
        return self.C

    def _loader(self, maps: torch.Tensor, id0: int) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        out = []
        for a in range(0, maps.shape[0], self.batch_size):
            m = maps[a:a + self.batch_size]
            b = m.shape[0]
            x = torch.zeros((b, 3, self.H, self.H), dtype=torch.float32)
            x[:, 0, 0, 0] = torch.arange(id0 + a, id0 + a + b, dtype=torch.float32)
            y = (m.float() / 255.0).unsqueeze(1)  # ToTensor on an 8-bit mask: id/255
            out.append((x, y))
        return out

    def train_dataloader(self):
        return self._loader(self.train_maps, 0)

    def val_dataloader(self):
        return self._loader(self.val_maps, self.train_maps.shape[0])
