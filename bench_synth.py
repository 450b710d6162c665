"""Synthetic VOC/ADE-shaped inputs for bench.py, as a pure function of (seed, image number).

Counter-based (a 32-bit integer hash of the element's coordinates, no RNG state), written with
torch integer ops only, so the SAME bits come out on the CPU and on a CUDA device: the B200 arm
generates its bank on the GPU, the reference arm (`bench.py --impl reference`) generates the
identical bank on the host without importing the product library, and the parity sample of both
is the same set of images.  (blocky class maps with ~2 % ignore pixels delivered as
y = id/255; features = class prototype + noise, un-normalised, with a per-patch positive scale.)

Nothing here is on the product path: bench.py, tools/ and tests/ import it.
"""
from __future__ import annotations

import torch

_M1, _M2 = -2048144789, -1028477387  # 0x85ebca6b, 0xc2b2ae35 as int32
_GOLD = -1640531535                  # 0x9E3779B1 as int32


def _srl(x: torch.Tensor, n: int) -> torch.Tensor:
    """Logical right shift of an int32 tensor (torch's >> is arithmetic)."""
    return (x >> n) & ((1 << (32 - n)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """murmur3 fmix32 on int32 tensors (multiplication wraps modulo 2^32 on CPU and CUDA alike)."""
    x = x ^ _srl(x, 16)
    x = x * _M1
    x = x ^ _srl(x, 13)
    x = x * _M2
    return x ^ _srl(x, 16)


def _hash(seed: int, *coords: torch.Tensor) -> torch.Tensor:
    """int32 hash of broadcastable int32 coordinate tensors."""
    h = torch.full((), (seed * 0x632BE5AB + 0x1B873593) & 0x7FFFFFFF, dtype=torch.int32, device=coords[0].device)
    for c in coords:
        h = _mix((h ^ c.to(torch.int32)) * _GOLD + 0x3C6EF372)
    return h


def _noise(h: torch.Tensor) -> torch.Tensor:
    """Zero-mean, unit-variance noise from one hash word: the sum of its two 16-bit halves
    (a triangular distribution; its shape is irrelevant here, its reproducibility is not)."""
    lo = (h & 0xFFFF).to(torch.float32)
    hi = _srl(h, 16).to(torch.float32)
    return (lo + hi - 65535.0) * (1.0 / 26754.4)  # sqrt(2 * (65536^2 - 1) / 12)


def _noise_bytes(h: torch.Tensor) -> torch.Tensor:
    """Four zero-mean, unit-variance (uniform, 8-bit) noise values per hash word: (..., n) int32 ->
    (..., 4n) fp32.  A quarter of the hashing of _noise; used for the bulk of the data."""
    parts = [((h >> (8 * j)) & 0xFF).to(torch.float32) for j in range(4)]
    return ((torch.stack(parts, dim=-1) - 127.5) * (1.0 / 73.9)).flatten(-2)


def prototypes(w: dict, device) -> torch.Tensor:
    """(C, d) class prototypes."""
    c = torch.arange(w["C"], dtype=torch.int32, device=device).view(-1, 1)
    k = torch.arange(w["d"], dtype=torch.int32, device=device).view(1, -1)
    return _noise(_hash(0, c, k))


def label_maps(w: dict, img: torch.Tensor) -> torch.Tensor:
    """uint8 (n, H, H) class maps of images `img` (int64 ids): cells x cells random classes
    nearest-upsampled, ~2 % of the pixels set to the ignore id."""
    dev = img.device
    S, ps, C = w["S"], w["ps"], w["C"]
    H, cells = S * ps, 8
    first = 1 if w["ignore"] == 0 else 0
    i = img.to(torch.int32).view(-1, 1, 1)
    cy = torch.arange(cells, dtype=torch.int32, device=dev).view(1, -1, 1)
    cx = torch.arange(cells, dtype=torch.int32, device=dev).view(1, 1, -1)
    coarse = (_srl(_hash(1, i, cy, cx), 8) % (C - first) + first).to(torch.uint8)
    reps = (H + cells - 1) // cells
    maps = coarse.repeat_interleave(reps, 1).repeat_interleave(reps, 2)[:, :H, :H]
    py = torch.arange(H, dtype=torch.int32, device=dev).view(1, -1, 1)
    px = torch.arange(H, dtype=torch.int32, device=dev).view(1, 1, -1)
    ign = _srl(_hash(2, i, py, px), 8) < int(0.02 * (1 << 24))
    return torch.where(ign, torch.full_like(maps, w["ignore"]), maps).contiguous()


def features(w: dict, img: torch.Tensor, maps: torch.Tensor, protos: torch.Tensor, salt: int = 3) -> torch.Tensor:
    """fp32 (n, S*S, d) raw patch features: prototype of the class at the patch centre + 0.8 * noise,
    times a per-patch scale of about 3.7 (queries reach the search un-normalised, hbird_eval.py:625)."""
    dev = img.device
    S, ps, C, d = w["S"], w["ps"], w["C"], w["d"]
    n = img.shape[0]
    centre = maps[:, ps // 2::ps, ps // 2::ps].reshape(n, S * S).long().clamp_max(C - 1)
    i = img.to(torch.int32).view(-1, 1, 1)
    p = torch.arange(S * S, dtype=torch.int32, device=dev).view(1, -1, 1)
    k = torch.arange(d // 4, dtype=torch.int32, device=dev).view(1, 1, -1)  # d % 4 == 0
    f = protos[centre] + 0.8 * _noise_bytes(_hash(salt, i, p, k))
    scale = 3.7 * (1.0 + 0.25 * _noise(_hash(salt + 1, i, p)))  # in [1.4, 6.0]; no transcendental: same bits everywhere
    return (f * scale).contiguous()


def images(w: dict, first: int, n: int, device, protos: torch.Tensor = None, stream: int = 0):
    """(features (n, S*S, d) fp32, maps (n, H, H) uint8) of images first .. first+n-1 of stream
    `stream` (0 = training images that fill the bank, 1 = validation images)."""
    img = torch.arange(first, first + n, dtype=torch.int64, device=device) + stream * (1 << 24)
    if protos is None:
        protos = prototypes(w, device)
    maps = label_maps(w, img)
    return features(w, img, maps, protos), maps


def soft_labels(w: dict, maps: torch.Tensor) -> torch.Tensor:
    """fp32 (n*S*S, C) per-patch class histogram / ps^2 of bank-side maps (255 already 0), i.e.
    one_hot(...).float().mean(3) of hbird_eval.py:319-320 without the one-hot tensor.  Used by the
    reference arm to build label_memory at bank sizes where oracle.build_memory's materialised
    one-hot would take minutes (tests/test_bench_contract.py checks the two agree)."""
    S, ps, C = w["S"], w["ps"], w["C"]
    n = maps.shape[0]
    patch = torch.arange(S, device=maps.device).repeat_interleave(ps)
    pid = (torch.arange(n, device=maps.device).view(-1, 1, 1) * S + patch.view(1, -1, 1)) * S + patch.view(1, 1, -1)
    flat = pid.reshape(-1) * C + maps.reshape(-1).long().clamp_max(C - 1)
    valid = (maps.reshape(-1) < C).to(torch.float32)
    hist = torch.zeros(n * S * S * C, dtype=torch.float32, device=maps.device).index_add_(0, flat, valid)
    return (hist / float(ps * ps)).view(n * S * S, C)
