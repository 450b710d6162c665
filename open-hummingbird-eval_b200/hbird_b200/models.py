"""Feature-extractor duck type.

ViT feature extraction stays in PyTorch (north star).  What the evaluator touches is small:
`.forward_features(x) -> (feats (B, S*S, d), aux)`, `.eval_spatial_resolution`, `.d_model`, `.to()`,
`.eval()` (hbird_eval.py:116-117,133,157,217,312).  `FeatureExtractorSimple` below takes the same
constructor arguments as the reference wrapper of that name (hbird/models.py:70-103) and simply
forwards to the user's `ftr_extr_fn(model, imgs)`.
"""
from typing import Callable

import torch


class FeatureExtractorSimple(torch.nn.Module):
    def __init__(self, vit_model: torch.nn.Module, ftr_extr_fn: Callable, eval_spatial_resolution: int = 14,
                 d_model: int = 768) -> None:
        super().__init__()
        if not callable(ftr_extr_fn):
            raise TypeError("ftr_extr_fn must be callable: (model, imgs) -> features or (features, aux)")
        self.model, self.ftr_extr_fn = vit_model, ftr_extr_fn
        self.eval_spatial_resolution, self.d_model = int(eval_spatial_resolution), int(d_model)

    def forward_features(self, imgs: torch.Tensor):
        result = self.ftr_extr_fn(self.model, imgs)
        if isinstance(result, (tuple, list)):
            return tuple(result)
        return result, None  # a bare feature tensor: no attention map available

    forward = forward_features
