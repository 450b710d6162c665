"""PredsmIoU — same public contract as the reference's hbird/utils/eval_metrics.py:13-339
(`update(gt, pred)`, `compute(...) -> (miou, tp, fp, fn, reordered_preds, matched_bg_fraction)`,
`reset()`), with the streaming confusion matrix accumulated by the K5 CUDA kernel
(hb_confusion_accumulate) in an int64 (C_gt, C_pred) device tensor.

The C x C post-processing (IoU matrix, Hungarian assignment, TP/FP/FN folding) is microseconds of
host work on a matrix of at most 151 x 151; like the reference it runs on the host (scipy).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import ops

try:
    from scipy.optimize import linear_sum_assignment

    _SCIPY_AVAILABLE = True
except Exception:  # pragma: no cover
    _SCIPY_AVAILABLE = False


# ---- C x C host math (eval_metrics.py:112-218 of the reference) ---------------------------------
def score_matrix(conf: np.ndarray, precision_based: bool = False) -> np.ndarray:
    """IoU = TP / max(row + col - TP, 1e-8) per (gt, pred) cell, or precision = TP / col; float64."""
    c = conf.astype(np.float64)
    if precision_based:
        return c / np.maximum(c.sum(axis=0, keepdims=True), 1e-8)
    union = c.sum(axis=1, keepdims=True) + c.sum(axis=0, keepdims=True) - c
    return c / np.maximum(union, 1e-8)


def hungarian_mapping(conf: np.ndarray) -> np.ndarray:
    """map[pred] -> gt maximising total IoU; unmatched predicted classes -> 0 (background)."""
    if not _SCIPY_AVAILABLE:
        raise RuntimeError("scipy is not available for Hungarian matching. Install scipy or use many_to_one=True.")
    rows, cols = linear_sum_assignment(1.0 - score_matrix(conf))
    mapping = np.zeros(conf.shape[1], dtype=np.int64)
    mapping[cols] = rows
    return mapping


def tp_fp_fn(conf: np.ndarray, mapping: Optional[np.ndarray]):
    G, P = conf.shape
    row_sum = conf.sum(axis=1)
    if mapping is None:  # linear probe: predicted ids are final labels
        diag = np.array([conf[i, i] if i < P else 0 for i in range(G)], dtype=np.int64)
        col = conf.sum(axis=0)
        fp = np.array([col[i] - conf[i, i] if i < P else 0 for i in range(G)], dtype=np.int64)
        return diag, fp, row_sum - diag
    folded = np.zeros((G, G), dtype=np.int64)
    np.add.at(folded, (slice(None), mapping), conf)  # fold predicted columns onto their gt class
    tp = np.diag(folded).copy()
    return tp, folded.sum(axis=0) - tp, row_sum - tp


def miou_from_confusion(conf: np.ndarray, many_to_one: bool = False, precision_based: bool = False,
                        linear_probe: bool = False):
    """(miou, tp, fp, fn, mapping, matched_bg_fraction) from an int64 (C_gt, C_pred) matrix.
    Default = Hungarian matching, the mean runs over ALL gt classes (absent ones count as 0)."""
    G, P = conf.shape
    if linear_probe:
        mapping, bg = None, 0.0
    elif many_to_one:
        mapping = score_matrix(conf, precision_based).argmax(axis=0).astype(np.int64)
        bg = float((mapping == 0).sum() / max(P, 1))
    else:
        mapping = hungarian_mapping(conf)
        bg = 1.0 / max(G, 1)
    tp, fp, fn = tp_fp_fn(conf, mapping)
    denom = (tp + fp + fn).astype(np.float64)
    miou = float((tp.astype(np.float64) / np.maximum(denom, 1e-8)).mean())
    return miou, tp.tolist(), fp.tolist(), fn.tolist(), mapping, bg


class PredsmIoU:
    def __init__(self, num_pred_classes: int, num_gt_classes: int, device: Optional[torch.device] = None,
                 ignore_index: Optional[int] = None, prefer_cuda: bool = True,
                 store_reordered_preds: bool = True):
        self.num_pred_classes = int(num_pred_classes)
        self.num_gt_classes = int(num_gt_classes)
        self.ignore_index = int(ignore_index) if ignore_index is not None else None
        # As the reference, every prediction is kept on the host by default so that compute() can
        # return `reordered_preds` (eval_metrics.py:32,107-109,277-284).  HbirdEvaluation.evaluate()
        # discards that list (hbird_eval.py:253) and passes False.
        self.store_reordered_preds = bool(store_reordered_preds)
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("hbird_b200.PredsmIoU needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self._conf_mat = torch.zeros((self.num_gt_classes, self.num_pred_classes), dtype=torch.int64,
                                     device=self.device)
        self._pred_chunks: List[torch.Tensor] = []

    @torch.no_grad()
    def reset(self) -> None:
        self._conf_mat.zero_()
        self._pred_chunks.clear()

    @torch.no_grad()
    def update(self, gt: torch.Tensor, pred: torch.Tensor) -> None:
        """gt / pred: class-index tensors of identical shape (any integer dtype, CPU or CUDA).
        Values outside uint8 range can never be valid classes (C <= 256) and are dropped, as the
        reference's range mask does (eval_metrics.py:91-95)."""
        if gt.shape != pred.shape:
            raise ValueError(f"Shapes must match. Got gt={gt.shape}, pred={pred.shape}")
        gt8, pr8 = self._to_u8_pair(gt, pred)
        ops.confusion_accumulate(self._conf_mat, gt8, pr8, self.ignore_index)
        if self.store_reordered_preds:
            keep = torch.ones_like(gt8, dtype=torch.bool)
            if self.ignore_index is not None:
                keep &= gt8 != self.ignore_index
            keep &= (gt8 < self.num_gt_classes) & (pr8 < self.num_pred_classes)
            self._pred_chunks.append(pr8[keep].to("cpu", dtype=torch.int32))

    def _to_u8_pair(self, gt: torch.Tensor, pred: torch.Tensor):
        """Flatten to uint8 on the device.  Class ids are < 256 (num classes <= 256), so a pixel
        whose gt or pred does not fit a byte can never be counted: such pixels are removed here,
        everything else is range-checked inside the kernel."""
        gt = gt.to(self.device, non_blocking=True).reshape(-1)
        pred = pred.to(self.device, non_blocking=True).reshape(-1)
        if gt.dtype != torch.uint8 or pred.dtype != torch.uint8:
            ok = (gt >= 0) & (gt <= 255) & (pred >= 0) & (pred <= 255)
            if not bool(ok.all()):
                gt, pred = gt[ok], pred[ok]
            gt, pred = gt.to(torch.uint8), pred.to(torch.uint8)
        return gt.contiguous(), pred.contiguous()

    def confusion_matrix(self) -> np.ndarray:
        return self._conf_mat.cpu().numpy()

    def confusion_buffer(self) -> torch.Tensor:
        """The int64 (C_gt, C_pred) device matrix itself, for kernels that score a batch in place
        (hb_eval_step / hb_predict_score) instead of going through update()."""
        return self._conf_mat

    @torch.no_grad()
    def compute(self, is_global_zero: bool, many_to_one: bool = False, precision_based: bool = False,
                linear_probe: bool = False, sync_distributed: bool = False,
                return_reordered: bool = True) -> Tuple[float, List[int], List[int], List[int], List[int], float]:
        if not is_global_zero:
            return 0.0, [], [], [], [], 0.0
        if sync_distributed and torch.distributed.is_available() and torch.distributed.is_initialized():
            from ..distributed import all_reduce_sum

            all_reduce_sum(self._conf_mat)
        miou, tp, fp, fn, mapping, bg = miou_from_confusion(
            self.confusion_matrix(), many_to_one=many_to_one, precision_based=precision_based,
            linear_probe=linear_probe)
        reordered: List[int] = []
        if return_reordered and not self.store_reordered_preds:
            # eval_metrics.py:272-276
            raise RuntimeError("return_reordered=True requires store_reordered_preds=True at construction time.")
        if return_reordered:
            allp = torch.cat(self._pred_chunks).long().numpy() if self._pred_chunks else np.zeros(0, np.int64)
            reordered = (allp if mapping is None else mapping[allp]).astype(np.int64).tolist()
        return miou, tp, fp, fn, reordered, bg

    @torch.no_grad()
    def compute_miou(self, gt: np.ndarray, pred: np.ndarray, num_pred: int, num_gt: int, many_to_one: bool = False,
                     precision_based: bool = False, linear_probe: bool = False):
        """Single-shot adapter with the reference's (historical) argument order
        (eval_metrics.py:293-339)."""
        self.__init__(num_pred_classes=num_pred, num_gt_classes=num_gt, device=self.device,
                      ignore_index=self.ignore_index, store_reordered_preds=True)
        self.update(torch.from_numpy(gt.astype(np.int64)), torch.from_numpy(pred.astype(np.int64)))
        miou, tp, fp, fn, reordered, bg = self.compute(True, many_to_one, precision_based, linear_probe)
        as64 = lambda xs: [np.int64(x) for x in xs]  # noqa: E731
        return float(miou), as64(tp), as64(fp), as64(fn), as64(reordered), float(bg)
