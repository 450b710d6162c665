"""Feature-extractor duck type.  ViT feature extraction stays in PyTorch (north star); the
evaluator only needs `.forward_features(x) -> (feats (B, S*S, d), aux)`, `.eval_spatial_resolution`,
`.d_model`, `.to()`, `.eval()` (hbird_eval.py:116-117,133,157,217,312).  This thin wrapper has the
same constructor as the reference's FeatureExtractorSimple (hbird/models.py:70-103)."""
import torch.nn as nn


class FeatureExtractorSimple(nn.Module):
    def __init__(self, vit_model, ftr_extr_fn, eval_spatial_resolution: int = 14, d_model: int = 768):
        super().__init__()
        self.model = vit_model
        self.ftr_extr_fn = ftr_extr_fn
        self.eval_spatial_resolution = eval_spatial_resolution
        self.d_model = d_model

    def forward_features(self, imgs):
        out = self.ftr_extr_fn(self.model, imgs)
        return out if isinstance(out, tuple) else (out, None)

    def forward(self, imgs):
        return self.forward_features(imgs)
