"""The C-ABI library loads without a GPU, exports every symbol the header declares, and every
compute entry point fails loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

from hbird_b200 import _capi, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hbird_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_capi.EXPORTED_SYMBOLS) == syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in header_symbols():
        assert getattr(lib, name) is not None, name


def test_library_has_no_cuda_link_dependency():
    # the .so must load on a box without libcuda/libcudart (static cudart, driver entry points
    # resolved at run time), otherwise this import would already have failed on the CPU container
    assert _capi.lib.hb_abi_version() == 2


def test_plan_search_is_host_only():
    out = (ctypes.c_int * 4)()
    assert _capi.lib.hb_plan_search(1024000, 12544, 2, 148, 0, out) == 0
    n_tiles, n_qblocks, n_chunks, n_units = list(out)
    assert n_tiles == 4000 and n_qblocks == 49 and n_units == 74 and 1 <= n_chunks <= 64


def test_error_codes_map_to_reference_exception_types():
    out = (ctypes.c_int * 4)()
    rc = _capi.lib.hb_plan_search(0, 1, 2, 148, 0, out)
    assert rc == _capi.HB_ERR_INVALID
    with pytest.raises(ValueError):
        _capi.check(rc)
    assert "rows" in _capi.last_error()
    with pytest.raises(RuntimeError):
        _capi.check(_capi.HB_ERR_UNSUPPORTED)
    with pytest.raises(MemoryError):
        _capi.check(_capi.HB_ERR_OOM)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(RuntimeError, match="no CPU fallback|No GPUs|CUDA"):
        ops.device_check(0)
    with pytest.raises(RuntimeError):
        ops.MemoryBank(64, 3, 16, 10, 0, True)
    with pytest.raises(RuntimeError):
        ops.decode_mask(torch.zeros(4), False)  # CPU tensors are rejected, not silently computed
    from hbird_b200 import HbirdEvaluation, NearestNeighborSearchB200

    with pytest.raises(RuntimeError):
        NearestNeighborSearchB200(torch.zeros(4, 64))
    with pytest.raises(RuntimeError):
        HbirdEvaluation(torch.nn.Identity(), [], 3, device="cpu")


def test_exchange_argument_checks_are_host_only():
    """hb_exchange_* validate their arguments before touching the device: bad arguments give
    HB_ERR_INVALID (ValueError) even on a CPU box; a valid request without a GPU fails loudly."""
    h = ctypes.c_void_p(0)
    lib = _capi.lib
    assert lib.hb_exchange_create(0, 2, 2, 100, 30, ctypes.byref(h)) == _capi.HB_ERR_INVALID  # rank >= world
    assert "rank" in _capi.last_error()
    assert lib.hb_exchange_create(0, 0, 99, 100, 30, ctypes.byref(h)) == _capi.HB_ERR_INVALID  # world > 16
    assert lib.hb_exchange_create(0, 0, 2, 0, 30, ctypes.byref(h)) == _capi.HB_ERR_INVALID     # no capacity
    assert lib.hb_exchange_create(0, 0, 2, 100, 500, ctypes.byref(h)) == _capi.HB_ERR_INVALID  # k > 128
    assert h.value is None
    assert lib.hb_exchange_destroy(None) == 0
    assert lib.hb_exchange_slice_rows(None) == -1
    assert lib.hb_exchange_merge(None, None, None, None) == _capi.HB_ERR_INVALID
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ops.ShardExchange(0, 2, 100, 30, 0)
