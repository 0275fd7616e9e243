"""Exact quantiles on the GPU (np.quantile, method "linear") and what the reference builds on them:

    threshold discovery   evaluation/uncertainty_aggregation/find_threshold.py:10-30, 69-112
    (eqACE lives in calibration.py and uses order_statistics from here)

The order statistics come from a 3-pass radix select (libvalunc's vu_radix_hist: 11 + 11 + 10 bits of the
order-preserving float key); the host only walks 2048-entry cumulative counts between the passes and applies NumPy's
interpolation formula to the two neighbouring order statistics.  Any number of maps can be folded into one selection
(find_threshold.py:98-105 concatenates every validation map), and voxels can be weighted by their number of valid
raters (one sample per (rater, pixel) pair, ace.py:378-406).
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


def _key_to_float(key: np.ndarray) -> np.ndarray:
    key = np.asarray(key, np.uint64)
    bits = np.where(key & 0x80000000, key & 0x7FFFFFFF, (~key) & 0xFFFFFFFF).astype(np.uint32)
    return bits.view(np.float32)


def _as_device_f32(t) -> torch.Tensor:
    t = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t))
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    return t.contiguous().reshape(-1)


def _gt_struct(gt: Optional[torch.Tensor], ignore_value, n: int):
    """references (R, n) -> vu_gt for the sample weights (number of valid raters per voxel)."""
    if gt is None:
        return None, None
    g = gt if isinstance(gt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(gt))
    g = g.to(torch.device("cuda", torch.cuda.current_device()))
    g = g.to(torch.uint8 if g.dtype in (torch.uint8, torch.bool) else torch.int64).contiguous().reshape(g.shape[0], -1)
    if g.shape[1] != n:
        raise ValueError("references and values must cover the same voxels")
    s = _lib.Gt()
    s.data, s.dtype, s.R = g.data_ptr(), (_lib.GT_U8 if g.dtype == torch.uint8 else _lib.GT_I64), g.shape[0]
    s.stride_b, s.stride_r, s.stride_v = g.numel(), n, 1
    s.has_ignore, s.ignore_index = (0, 0) if ignore_value is None else (1, int(ignore_value))
    return s, g


_WS = {}


def _workspace(dev):
    """Per-device workspace of the device-side selection: 64 x 2048 histogram counters, the vu_radix_state and its pinned
    host mirror (allocated once: a selection is a handful of tiny launches, allocations would dominate it)."""
    key = (dev.type, dev.index)
    if key not in _WS:
        _WS[key] = (torch.empty((64, 2048), dtype=torch.int64, device=dev),
                    torch.zeros(C.sizeof(_lib.RadixState), dtype=torch.uint8, device=dev),
                    torch.zeros(C.sizeof(_lib.RadixState), dtype=torch.uint8).pin_memory())
    return _WS[key]


class RadixSelect:
    """Exact rank selection over the multiset formed by all ``values`` tensors (float32, any shape).  With ``refs`` (one
    (R, *S) tensor per values tensor) voxel v counts once per reference that is not ``ignore_value``.  NaN sorts last,
    like np.sort.  The level-0 histogram (and with it ``total``) is computed once; ``select`` runs two more passes per
    group of up to 64 ranks."""

    def __init__(self, values: Sequence, refs: Optional[Sequence] = None, ignore_value=None, reduce=None):
        """``reduce``: optional callable applied in place to every device histogram before it is read (``all_reduce_sum``
        below when the maps are sharded over ranks: every rank then selects in the union of all shards -- the ranks it asks
        for must be the same on all of them, which they are when they derive from ``total``)."""
        _lib.require_device()
        self._reduce = reduce
        self._lib = _lib.load()
        self._dev = torch.device("cuda", torch.cuda.current_device())
        self._vals = [_as_device_f32(v) for v in values]
        self._gts = [_gt_struct(None if refs is None else refs[i], ignore_value, self._vals[i].numel())
                     for i in range(len(self._vals))]
        self._h0_cache = None

    @property
    def _h0(self) -> np.ndarray:
        if self._h0_cache is None:
            self._h0_cache = self._run_pass(0, np.zeros(0, np.uint32))
        return self._h0_cache

    @property
    def total(self) -> int:
        return int(self._h0.sum())

    def select_quantile_stats(self, qs, q_is_f32: bool = False, reverse: bool = False):
        """The order statistics np.quantile (method "linear") needs for the fractions ``qs`` (at most 31), selected WITHOUT
        host round trips between the passes: three histogram passes and three one-warp descents (vu_radix_walk) on the
        current stream, then ONE read-back.  Returns (total, lo_stats, hi_stats, last) -- float32 order statistics of rank
        floor((total - 1) q), the next rank, and the largest element (NaN if the data holds a NaN).  ``reverse`` selects the
        mirrored ranks total - 1 - r instead (a confidence that falls with the uncertainty).  Not for sharded data
        (``reduce``): there the histograms have to cross ranks between the passes."""
        if self._reduce is not None:
            raise ValueError("select_quantile_stats works on one device; use select() with reduce")
        qs = np.ascontiguousarray(np.asarray(qs, np.float64).reshape(-1))
        nq = len(qs)
        if nq > 31:
            raise ValueError("at most 31 quantile fractions per selection")
        lib, dev, stream = self._lib, self._dev, _lib.current_stream_ptr()
        hist, state, state_host = _workspace(dev)
        qp = qs.ctypes.data_as(C.POINTER(C.c_double))
        if len(self._vals) == 1:  # one array: the whole selection is one call
            v, (gs, _keep) = self._vals[0], self._gts[0]
            _lib.check(lib.vu_quantile_select(v.data_ptr(), v.numel(), C.byref(gs) if gs is not None else None, qp, nq, int(q_is_f32),
                                              int(reverse), hist.data_ptr(), state.data_ptr(), stream), "vu_quantile_select")
        else:
            for level in range(3):
                hist.zero_()
                for v, (gs, _keep) in zip(self._vals, self._gts):
                    _lib.check(lib.vu_radix_hist_state(v.data_ptr(), v.numel(), C.byref(gs) if gs is not None else None, level,
                                                       state.data_ptr(), hist.data_ptr(), stream), "vu_radix_hist_state")
                _lib.check(lib.vu_radix_walk(hist.data_ptr(), level, qp, nq, int(q_is_f32), int(reverse), state.data_ptr(), stream),
                           "vu_radix_walk")
        state_host.copy_(state, non_blocking=True)  # the one read-back (pinned)
        torch.cuda.current_stream(dev).synchronize()
        st = _lib.RadixState.from_buffer_copy(state_host.numpy().tobytes())
        total = int(st.total)
        if total == 0:
            nan = np.full(nq, np.nan, np.float32)
            return 0, nan, nan.copy(), np.float32(np.nan)
        keys = _key_to_float(np.array(st.key[:2 * nq + 1], np.uint64))
        return total, keys[0:2 * nq:2].copy(), keys[1:2 * nq:2].copy(), keys[2 * nq]

    def _run_pass(self, level: int, prefixes: np.ndarray) -> np.ndarray:
        n_slots = max(1, len(prefixes))
        hist = torch.zeros((n_slots, 2048), dtype=torch.int64, device=self._dev)
        pre = torch.from_numpy(prefixes.astype(np.uint32).view(np.int32)).to(self._dev) if len(prefixes) else None
        stream = _lib.current_stream_ptr()
        for v, (gs, _keep) in zip(self._vals, self._gts):
            _lib.check(self._lib.vu_radix_hist(v.data_ptr(), v.numel(), C.byref(gs) if gs is not None else None, level,
                                               pre.data_ptr() if pre is not None else None, len(prefixes), hist.data_ptr(), stream),
                       "vu_radix_hist")
        if self._reduce is not None:
            self._reduce(hist)
        return hist.cpu().numpy()

    @staticmethod
    def _descend(hist_rows: np.ndarray, slot_of: np.ndarray, residual: np.ndarray):
        """per rank: digit whose cumulative count first exceeds the residual rank, and the new residual"""
        cum = np.cumsum(hist_rows, axis=1)
        digit = np.array([np.searchsorted(cum[s], r, side="right") for s, r in zip(slot_of, residual)])
        below = np.array([cum[s, d - 1] if d > 0 else 0 for s, d in zip(slot_of, digit)])
        return digit.astype(np.uint64), residual - below

    def select(self, ranks: Sequence[int]) -> np.ndarray:
        """float32 elements of rank ``ranks`` (0-based, ascending)."""
        ranks = np.asarray(ranks, np.int64).reshape(-1)
        if self.total == 0:
            return np.full(len(ranks), np.nan, np.float32)
        if len(ranks) == 0:
            return np.zeros(0, np.float32)
        if ranks.min() < 0 or ranks.max() >= self.total:
            raise IndexError("rank out of range")
        d0, res = self._descend(self._h0, np.zeros(len(ranks), np.int64), ranks.copy())
        out_keys = np.zeros(len(ranks), np.uint64)
        for start in range(0, len(ranks), 64):  # at most 64 prefixes per pass
            sl = slice(start, start + 64)
            p1, slot1 = np.unique(d0[sl], return_inverse=True)
            d1, res1 = self._descend(self._run_pass(1, p1), slot1, res[sl])
            pre2 = (d0[sl] << np.uint64(11)) | d1
            p2, slot2 = np.unique(pre2, return_inverse=True)
            d2, _ = self._descend(self._run_pass(2, p2), slot2, res1)
            out_keys[sl] = (pre2 << np.uint64(10)) | d2
        return _key_to_float(out_keys)


def all_reduce_sum(hist: torch.Tensor) -> None:
    """``reduce`` hook for maps sharded over the ranks of the default process group (one int64 all-reduce per histogram
    pass: 16 KB per prefix, latency-bound).  Counts are integers, so the selection is exact and identical on every rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)


def order_statistics(values: Sequence, ranks: Sequence[int], refs: Optional[Sequence] = None, ignore_value=None, reduce=None):
    """(float32 array of the elements of rank ``ranks``, total number of samples); see RadixSelect."""
    sel = RadixSelect(values, refs, ignore_value, reduce)
    return sel.select(ranks), sel.total


def quantile_ranks(total: int, qs: np.ndarray):
    """NumPy's "linear" method: virtual index h = (n - 1) q **in the dtype of q**, its two neighbouring ranks and the
    interpolation weight (numpy/lib/_function_base_impl.py: _compute_virtual_index, _get_indexes, _get_gamma)."""
    h = (total - 1) * qs
    lo = np.minimum(np.floor(h).astype(np.int64), total - 1)  # a float32 h can round above n - 1: NumPy takes the last element
    hi = np.minimum(lo + 1, total - 1)
    return lo, hi, np.asarray(h - lo, dtype=h.dtype)


def lerp(a: np.ndarray, b: np.ndarray, t: np.ndarray) -> np.ndarray:
    """NumPy's _lerp: a + (b - a) t, replaced by b - (b - a)(1 - t) where t >= 0.5."""
    diff = b - a
    return np.where(t >= 0.5, b - diff * (1 - t), a + diff * t)


def quantile(values: Sequence, q, refs: Optional[Sequence] = None, ignore_value=None, dtype=np.float32, reduce=None, index_dtype=None):
    """np.quantile(np.concatenate(values), q) (method "linear") as NumPy 2.x evaluates it for data of ``dtype``: a Python
    scalar ``q`` is cast to ``dtype`` first (so the virtual index of a float32 map is a float32, find_threshold.py:76), an
    array ``q`` keeps its own dtype (float64 for the ``np.linspace`` of ace.py:387).  NaN in the data gives NaN.
    ``index_dtype=np.float64`` evaluates the virtual index (n - 1) q and the interpolation weight in float64 whatever ``q``
    is -- what NumPy 1.24 does, the version the reference pins in requirements.txt: with more than 2^24 pooled samples
    (find_threshold.py:98-105 pools every validation map) a float32 index cannot address every rank and the two NumPy
    generations select different order statistics.  The default follows the NumPy of this environment (2.x); the golden
    vectors under tests/golden were recorded with it.  The rank selection itself is exact either way.
    ``reduce=all_reduce_sum``: the maps of this rank are one shard of the data set (find_threshold.py:98-105 over a sharded
    validation split); every rank gets the quantile of the union."""
    scalar = np.ndim(q) == 0
    qs = np.atleast_1d(np.asarray(q, dtype=dtype) if isinstance(q, (int, float)) else np.asarray(q))
    if index_dtype is not None:
        qs = qs.astype(index_dtype)
    if qs.size and (np.nanmin(qs) < 0 or np.nanmax(qs) > 1 or np.isnan(qs).any()):
        raise ValueError("Quantiles must be in the range [0, 1]")
    sel = RadixSelect(values, refs, ignore_value, reduce)
    if reduce is None and 0 < qs.size <= 31 and qs.dtype in (np.float32, np.float64):
        # one device: three passes and one read-back
        total, s_lo, s_hi, last = sel.select_quantile_stats(qs.astype(np.float64), q_is_f32=qs.dtype == np.float32)
        if total == 0:
            res = np.full(qs.shape, np.nan)
            return float("nan") if scalar else res
        _, _, g = quantile_ranks(total, qs)
        if np.isnan(last):
            res = np.full(qs.shape, np.nan, np.result_type(dtype, g.dtype))
        else:
            res = lerp(s_lo.astype(dtype), s_hi.astype(dtype), g)
        return res[0] if scalar else res
    total = sel.total
    if total == 0:
        res = np.full(qs.shape, np.nan)
        return float("nan") if scalar else res
    lo, hi, g = quantile_ranks(total, qs)
    ranks = np.unique(np.concatenate([lo, hi, [total - 1]]))
    stat = sel.select(ranks).astype(dtype)
    if np.isnan(stat[-1]):
        res = np.full(qs.shape, np.nan, np.result_type(dtype, g.dtype))
    else:
        res = lerp(stat[np.searchsorted(ranks, lo)], stat[np.searchsorted(ranks, hi)], g)
    return res[0] if scalar else res


# ---- threshold discovery (find_threshold.py) --------------------------------------------------------------------------
def calculate_foreground_quantile_image(image) -> float:
    """find_threshold.py:10-12: 1 - (non-zero pixels / pixels) of one predicted segmentation."""
    from .aggregation import _compute_area  # area == number of label values > 0 == non-zero for uint8 labels
    arr = image if isinstance(image, torch.Tensor) else np.asarray(image)
    size = arr.numel() if isinstance(arr, torch.Tensor) else arr.size
    return 1 - (_compute_area(arr != 0) / size)


def calculate_threshold_image(quantile_path, image, method: str, reduce=None) -> float:
    """find_threshold.py:69-77: np.quantile of the (concatenated) uncertainty values at the method's mean foreground
    quantile.  ``image`` may be one array / tensor or a list of them (they are not concatenated)."""
    with open(Path(quantile_path)) as f:
        method_quantile = json.load(f)[method]
    maps = image if isinstance(image, (list, tuple)) else [image]
    return float(quantile(maps, method_quantile, reduce=reduce))
