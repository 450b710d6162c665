"""ctypes loader of the plain-C oracle (oracle/hbird_oracle.c) — TEST INFRASTRUCTURE ONLY.
Builds oracle/_build/libhbird_oracle.so with gcc on first use (oracle/Makefile)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhbird_oracle.so")


def build() -> str:
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB


def _load():
    src = os.path.join(_HERE, "hbird_oracle.c")
    if not os.path.isfile(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        build()
    return ctypes.CDLL(_LIB)


_lib = _load()
_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
for _name in ("hbo_decode_mask", "hbo_patch_histogram", "hbo_confusion", "hbo_upsample_argmax", "hbo_search_ip",
              "hbo_label_transfer"):
    getattr(_lib, _name).restype = None


def decode_mask(y: np.ndarray, remap: bool) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.float32)
    out = np.empty(y.shape, dtype=np.uint8)
    _lib.hbo_decode_mask(_p(y), ctypes.c_int64(y.size), int(remap), _p(out))
    return out


def patch_histogram(mask: np.ndarray, S: int, ps: int, C: int) -> np.ndarray:
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    B = mask.shape[0]
    hist = np.empty((B * S * S, C), dtype=np.uint16)
    _lib.hbo_patch_histogram(_p(mask), B, S, ps, C, _p(hist))
    return hist


def confusion(gt: np.ndarray, pred: np.ndarray, G: int, P: int, ignore_index) -> np.ndarray:
    gt = np.ascontiguousarray(gt.reshape(-1), dtype=np.uint8)
    pred = np.ascontiguousarray(pred.reshape(-1), dtype=np.uint8)
    conf = np.zeros((G, P), dtype=np.int64)
    _lib.hbo_confusion(_p(gt), _p(pred), ctypes.c_int64(gt.size), G, P, -1 if ignore_index is None else int(ignore_index),
                       _p(conf))
    return conf


def upsample_argmax(label_hat: np.ndarray, B: int, S: int, H: int, W: int) -> np.ndarray:
    label_hat = np.ascontiguousarray(label_hat, dtype=np.float32)
    C = label_hat.shape[-1]
    out = np.empty((B, H, W), dtype=np.uint8)
    _lib.hbo_upsample_argmax(_p(label_hat), B, S, C, H, W, _p(out))
    return out


def search_ip(q: np.ndarray, bank: np.ndarray, k: int):
    q = np.ascontiguousarray(q, dtype=np.float32)
    bank = np.ascontiguousarray(bank, dtype=np.float32)
    idx = np.empty((q.shape[0], k), dtype=np.int64)
    dist = np.empty((q.shape[0], k), dtype=np.float32)
    _lib.hbo_search_ip(_p(q), _p(bank), ctypes.c_int64(q.shape[0]), ctypes.c_int64(bank.shape[0]), q.shape[1], k,
                       _p(idx), _p(dist))
    return idx, dist


def label_transfer(q, feature_memory, label_memory, idx, beta: float = 0.02) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float32)
    fm = np.ascontiguousarray(feature_memory, dtype=np.float32)
    lm = np.ascontiguousarray(label_memory, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty((q.shape[0], lm.shape[1]), dtype=np.float32)
    _lib.hbo_label_transfer(_p(q), _p(fm), _p(lm), _p(idx), ctypes.c_int64(q.shape[0]), q.shape[1], lm.shape[1],
                            idx.shape[1], ctypes.c_float(beta), _p(out))
    return out
