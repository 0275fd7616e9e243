"""C3 aggregation strategies on the GPU behind the reference's call surface
(evaluation/uncertainty_aggregation/aggregate_uncertainties.py:16-130 and
prediction_shape_stats.py:10-30).

Every function accepts what the reference's Hydra task hands over -- a NumPy
array loaded from disk -- or a CUDA tensor that the fused pass just produced
(then nothing leaves the device).  Unknown keyword arguments are swallowed like
the reference's ``**kwargs``.  Results are plain Python floats / ints.
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import F64, I64

_STATS_CACHE: dict = {}


def _to_device_map(image) -> torch.Tensor:
    _lib.require_device()
    if isinstance(image, torch.Tensor):
        t = image
    else:
        t = torch.from_numpy(np.ascontiguousarray(image))
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    return t.contiguous()


def _map_stats(maps3, labels, flags, thresholds=None, B=1):
    """Run vu_map_stats on up to three (B, V) maps; returns host rows."""
    lib = _lib.load()
    ref = next(m for m in list(maps3) + [labels] if m is not None)
    dev = ref.device
    V = ref.numel() // B
    sf = torch.zeros((B, F64["COLS"]), dtype=torch.float64, device=dev)
    si = torch.zeros((B, I64["COLS"]), dtype=torch.int64, device=dev)
    a = _lib.MapStatsArgs()
    a.struct_size = C.sizeof(_lib.MapStatsArgs)
    a.stat_flags = flags
    a.B, a.V = B, V
    for k, m in enumerate(maps3):
        a.maps[k] = m.data_ptr() if m is not None else None
    a.labels = labels.data_ptr() if labels is not None else None
    if thresholds is not None:
        for k in range(3):
            a.threshold[k] = float(thresholds[k])
    a.stats_f64, a.stats_i64 = sf.data_ptr(), si.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(lib.vu_map_stats(C.byref(a), _lib.current_stream_ptr()), "vu_map_stats")
    return sf.cpu().numpy(), si.cpu().numpy()


def image_level_aggregation(image, mean=True, **kwargs):
    """aggregate_uncertainties.py:37-39."""
    t = _to_device_map(image)
    n = t.numel()
    if n == 0:
        return {"max_score": float("nan") if mean else 0.0}
    f, _ = _map_stats([t, None, None], None, _lib.STAT_IMAGE_SUM)
    total = float(f[0, F64["SUM"]])
    return {"max_score": total / n if mean else total}


def threshold_aggregation(image, threshold=None, threshold_path=None, pred_model=None, unc_type=None, mean=True,
                          **kwargs):
    """aggregate_uncertainties.py:102-130, including the JSON lookup of the
    threshold and the rule that an empty selection returns the (zero) sum."""
    if threshold is None:
        if threshold_path is None:
            raise Exception("A threshold needs to be provided for threshold aggregation!")
        with open(threshold_path) as f:
            threshold_json = json.load(f)
        if pred_model is None or unc_type is None:
            raise Exception("If you want to load the threshold from a json file, you have to provide the "
                            "prediction model and the uncertainty type")
        threshold = threshold_json[pred_model][f"Mean {unc_type.split('_')[0]} threshold"]
    t = _to_device_map(image)
    if t.numel() == 0:
        return {"max_score": 0.0, "threshold": threshold}
    thr32 = _threshold_as_f32(threshold)
    f, i = _map_stats([t, None, None], None, _lib.STAT_THRESHOLD, thresholds=[thr32, 0.0, 0.0])
    total, count = float(f[0, F64["THR_SUM"]]), int(i[0, I64["THR_COUNT"]])
    if mean and count > 0:
        return {"max_score": total / count, "threshold": threshold}
    return {"max_score": total, "threshold": threshold}


def _threshold_as_f32(threshold) -> float:
    """``image >= threshold`` with a float32 image and a Python-float threshold:
    NumPy casts the scalar to float32 (NEP 50), and so does the C ABI -- but only
    if the cast is done to the smallest float32 >= t would a float64 comparison
    agree; NumPy 2 compares in float32 after rounding t, which is what we mirror."""
    return float(np.float32(threshold))


def patch_level_aggregation(image, patch_size, mean=False, **kwargs):
    """aggregate_uncertainties.py:16-34: max box sum and the bounding box of the
    first box that is np.isclose to it (row-major order of the array given)."""
    t = _to_device_map(image)
    ndim = t.dim()
    if type(patch_size) == int:
        patch_size = ndim * [patch_size]
    patch_size = [int(k) for k in patch_size]
    if ndim not in (1, 2, 3) or len(patch_size) != ndim:
        raise NotImplementedError("patch_level_aggregation on the GPU supports 1-, 2- and 3-D maps")
    dims = [1] * (3 - ndim) + list(t.shape)
    ks = [1] * (3 - ndim) + patch_size
    if any(k > d for k, d in zip(ks, dims)):
        # scipy's "valid" mode would swap the operands here; the reference never does this
        raise ValueError("patch_size exceeds the image size")
    out = patch_level_batched(t.reshape(1, *dims), ks, mean=mean)
    lin = int(out["first_index"][0])
    o = [d - k + 1 for d, k in zip(dims, ks)]
    idx3 = [lin // (o[1] * o[2]), (lin // o[2]) % o[1], lin % o[2]][3 - ndim:]
    return {"max_score": float(out["max_score"][0]),
            "bounding_box": [(int(i0), int(i0 + k)) for i0, k in zip(idx3, patch_size)]}


def patch_level_batched(maps: torch.Tensor, box: Sequence[int], mean: bool = False):
    """(B, d0, d1, d2) CUDA fp32 maps -> per-image max box sum and first-isclose
    linear index, without leaving the device until the tiny result copy."""
    _lib.require_device()
    lib = _lib.load()
    if maps.dim() != 4 or not maps.is_cuda or maps.dtype != torch.float32:
        raise ValueError("maps must be a (B, d0, d1, d2) CUDA float32 tensor")
    maps = maps.contiguous()
    B, d0, d1, d2 = maps.shape
    out_max = torch.empty(B, dtype=torch.float64, device=maps.device)
    out_first = torch.empty(B, dtype=torch.int64, device=maps.device)
    k = [int(box[0]), int(box[1]), int(box[2])]
    with torch.cuda.device(maps.device):
        for s in range(0, B, 65535):
            e = min(B, s + 65535)
            ws_bytes = int(lib.vu_patch_workspace_bytes(e - s, d0, d1, d2, *k))  # per-CTA maxima: lets the index pass skip
            ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.int64, device=maps.device)
            _lib.check(lib.vu_patch_max_ws(maps[s:e].data_ptr(), e - s, d0, d1, d2, *k, 1 if mean else 0,
                                           out_max[s:e].data_ptr(), out_first[s:e].data_ptr(),
                                           ws.data_ptr() if ws_bytes else None, ws_bytes, _lib.current_stream_ptr()), "vu_patch_max_ws")
    return {"max_score": out_max.cpu().numpy(), "first_index": out_first.cpu().numpy()}


# ---- prediction shape statistics (prediction_shape_stats.py:10-30) ------------
def prediction_shape_stats(mask):
    """(area, border) of one label map: #(label > 0) and the number of adjacent
    pairs with different labels summed over the axes."""
    _lib.require_device()
    lib = _lib.load()
    t = mask if isinstance(mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(mask))
    if t.numel() == 0:
        return 0.0, 0.0
    if t.dtype != torch.uint8:
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        else:
            if int(t.min()) < 0 or int(t.max()) > 255:
                raise NotImplementedError("label maps must fit uint8 (test_2D.py:818 casts to uint8)")
            t = t.to(torch.uint8)
    if not t.is_cuda:
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    t = t.contiguous()
    if t.dim() > 3:
        raise NotImplementedError("label maps of rank > 3")
    dims = [1] * (3 - t.dim()) + list(t.shape)
    _, i = _map_stats([None, None, None], t.reshape(-1), _lib.STAT_AREA)
    si = torch.zeros((1, I64["COLS"]), dtype=torch.int64, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.vu_border_count(t.data_ptr(), 1, dims[0], dims[1], dims[2], si.data_ptr(),
                                       _lib.current_stream_ptr()), "vu_border_count")
    return float(i[0, I64["AREA"]]), float(si.cpu().numpy()[0, I64["BORDER"]])


def _compute_area(mask) -> float:
    return prediction_shape_stats(mask)[0]


def _compute_border(mask) -> float:
    return prediction_shape_stats(mask)[1]


def _load_prediction_stats(dataset_path, stats_filename):
    if dataset_path is None:
        raise ValueError("Prediction statistics require a dataset-specific path.")
    stats_path = (Path(dataset_path) / stats_filename).resolve()
    cached = _STATS_CACHE.get(stats_path)
    if cached is None:
        if not stats_path.is_file():
            raise FileNotFoundError(f"Missing prediction stats file: {stats_path}. Run the area task first.")
        with open(stats_path) as f:
            cached = json.load(f)
        _STATS_CACHE[stats_path] = cached
    return cached


def _get_stat_value(stats_dict, image_id, stat_key):
    if image_id is None:
        raise ValueError(f"image_id is required to fetch '{stat_key}' statistics")
    entry = stats_dict.get(str(image_id))
    if entry is None or stat_key not in entry:
        raise KeyError(f"Statistic '{stat_key}' missing for image '{image_id}'. Ensure area task completed.")
    return float(entry[stat_key])


def _normalize_uncertainty_sum(image, divisor):
    """aggregate_uncertainties.py:70-74: divisor <= 0 leaves the sum un-normalised."""
    total = image_level_aggregation(image, mean=False)["max_score"]
    return total if divisor <= 0 else total / divisor


def border_normalized_aggregation(image, dataset_path=None, image_id=None, stats_filename="area.json", **kwargs):
    """aggregate_uncertainties.py:77-88."""
    value = _get_stat_value(_load_prediction_stats(dataset_path, stats_filename), image_id, "border")
    return {"max_score": _normalize_uncertainty_sum(image, value), "normalizer": value}


def area_normalized_aggregation(image, dataset_path=None, image_id=None, stats_filename="area.json", **kwargs):
    """aggregate_uncertainties.py:90-100."""
    value = _get_stat_value(_load_prediction_stats(dataset_path, stats_filename), image_id, "area")
    return {"max_score": _normalize_uncertainty_sum(image, value), "normalizer": value}
