"""ctypes binding of the C-ABI library (include/hbird_b200.h).

This is the only place the Python host touches native code.  The library must be built in-tree
(`python __graft_entry__.py` or `make -C open-hummingbird-eval_b200/csrc`); there is no fallback:
if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libhbird_b200.so"))
# Measurement tooling only (tools/gpu_probe.py A/B runs): load another build of the library, e.g. the
# previous round's, whose export list may be shorter.  The product never sets this.
_AB_LIB = os.environ.get("HBIRD_B200_AB_LIB")

HB_OK = 0
HB_ERR_INVALID = -1
HB_ERR_CUDA = -2
HB_ERR_OOM = -3
HB_ERR_UNSUPPORTED = -4
HB_ERR_STATE = -5

HB_BANK_KEEP_F32 = 1
HB_BANK_L2 = 2
HB_EXCHANGE_HANDLE_BYTES = 64

# every symbol include/hbird_b200.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("hb_abi_version", c_int, []),
    ("hb_last_error", c_char_p, []),
    ("hb_device_check", c_int, [c_int, POINTER(c_int)]),
    ("hb_bank_create", c_int, [c_int, c_int, c_int, c_int, c_int64, c_uint, POINTER(c_void_p)]),
    ("hb_bank_destroy", c_int, [c_void_p]),
    ("hb_bank_append", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    ("hb_bank_append_soft", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    ("hb_sample_patches", c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    ("hb_bank_finalize", c_int, [c_void_p]),
    ("hb_bank_rows", c_int64, [c_void_p]),
    ("hb_bank_capacity", c_int64, [c_void_p]),
    ("hb_bank_label_table", c_void_p, [c_void_p]),
    ("hb_bank_export", c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    ("hb_search", c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("hb_search_transfer", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int64, c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("hb_search_begin", c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    ("hb_search_finish", c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_int64, c_float, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    ("hb_search_abort", c_int, [c_void_p]),
    ("hb_coresidency_config", c_int, [c_void_p, c_int, c_int, c_int]),
    ("hb_eval_step", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                             c_int64, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("hb_search_config", c_int, [c_void_p, c_int, c_int]),
    ("hb_search_tune", c_int, [c_void_p, c_int, c_int]),
    ("hb_search_stats", c_int, [c_void_p, c_void_p]),
    ("hb_search_pacing", c_int, [c_void_p, c_int]),
    ("hb_search_last_launches", c_int, [c_void_p]),
    ("hb_search_timing", c_int, [c_void_p, c_int]),
    ("hb_search_kernel_time", c_int, [c_void_p, POINTER(c_float), POINTER(c_int)]),
    ("hb_search_rerank_time", c_int, [c_void_p, POINTER(c_float), POINTER(c_int)]),
    ("hb_plan_search", c_int, [c_int64, c_int64, c_int, c_int, c_int, POINTER(c_int)]),
    ("hb_search_dump_scores", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    ("hb_merge_topk", c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    ("hb_merge_topk_transfer", c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_void_p,
                                       c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("hb_exchange_create", c_int, [c_int, c_int, c_int, c_int64, c_int, POINTER(c_void_p)]),
    ("hb_exchange_destroy", c_int, [c_void_p]),
    ("hb_exchange_handle", c_int, [c_void_p, c_void_p, c_int]),
    ("hb_exchange_connect", c_int, [c_void_p, c_void_p, c_int]),
    ("hb_exchange_connect_local", c_int, [c_void_p, POINTER(c_void_p), c_int]),
    ("hb_exchange_disconnect", c_int, [c_void_p]),
    ("hb_search_scatter", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, POINTER(c_int64), c_void_p, c_void_p]),
    ("hb_search_finish_scatter", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int64, POINTER(c_int64), c_void_p]),
    ("hb_exchange_merge", c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    ("hb_exchange_merge_transfer", c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                           c_void_p, c_void_p]),
    ("hb_exchange_slice_rows", c_int64, [c_void_p]),
    ("hb_exchange_set_timeout", c_int, [c_void_p, c_int64]),
    ("hb_exchange_config", c_int, [c_void_p, c_int]),
    ("hb_exchange_rerank", c_int, [c_void_p, c_void_p]),
    ("hb_exchange_status", c_int, [c_void_p, c_void_p]),
    ("hb_label_transfer", c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p]),
    ("hb_upsample_argmax", c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    ("hb_predict_score", c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                 c_void_p]),
    ("hb_decode_mask", c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    ("hb_confusion_accumulate", c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]


def _load() -> ctypes.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"hbird_b200: native library not found at {LIB_PATH}. Build it with "
            "`python __graft_entry__.py` (or `make -C open-hummingbird-eval_b200/csrc`). "
            "There is no CPU or PyTorch fallback for this path."
        )
    lib = ctypes.CDLL(_AB_LIB or LIB_PATH)
    for name, restype, argtypes in _SIGNATURES:
        if _AB_LIB and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError if the header and the library drifted apart
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def last_error() -> str:
    msg = lib.hb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int) -> None:
    """Map hb_status to the exception types the reference backends raise
    (search_faiss.py:15-16,25,48; hbird_eval.py:281)."""
    if status == HB_OK:
        return
    msg = last_error() or f"hbird_b200 error {status}"
    if status == HB_ERR_INVALID:
        raise ValueError(msg)
    if status == HB_ERR_OOM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(device=None) -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
