"""Deterministic on-device synthetic inputs: slabs of
softmax(scale * N(0, 1)) keyed by (seed, image index), and matching references.
Used by bench.py, smoke() and the GPU tests; plays the role of the stochastic
forward passes that feed the hot path."""
from __future__ import annotations

import torch

from . import _lib


def synth_slab(P: int, B: int, C: int, spatial, seed: int = 0, first_image: int = 0, scale: float = 2.0,
               device=None, out: torch.Tensor | None = None) -> torch.Tensor:
    _lib.require_device()
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    spatial = tuple(spatial)
    V = 1
    for s in spatial:
        V *= int(s)
    if out is None:
        out = torch.empty((P, B, C) + spatial, dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(lib.vu_synth_slab(out.data_ptr(), P, B, C, V, seed, first_image, float(scale),
                                     _lib.current_stream_ptr()), "vu_synth_slab")
    return out


def synth_gt(slab: torch.Tensor, R: int, seed: int = 0, first_image: int = 0, flip: float = 0.1,
             ignore_frac: float = 0.0, ignore_value: int = 255) -> torch.Tensor:
    lib = _lib.load()
    P, B, C = slab.shape[:3]
    spatial = tuple(slab.shape[3:])
    V = slab[0, 0, 0].numel()
    out = torch.empty((B, R) + spatial, dtype=torch.uint8, device=slab.device)
    with torch.cuda.device(slab.device):
        _lib.check(lib.vu_synth_gt(out.data_ptr(), slab.data_ptr(), P, B, C, V, R, seed, first_image, float(flip),
                                   float(ignore_frac), int(ignore_value), _lib.current_stream_ptr()), "vu_synth_gt")
    return out
