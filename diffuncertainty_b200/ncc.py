"""Ambiguity metric (NCC) behind the reference's call surface
(evaluation/metrics/ncc.py:9-28, GT map: evaluation/experiment_dataloader.py:283).

The device accumulates five float64 sums per image (sum g, g^2, u, u^2, g*u); the
closed form below turns them into the reference's value.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import F64, I64


def ncc_from_sums(n, sg, sgg, su, suu, sgu) -> float:
    """ncc.py:17-27 from sums: covariance sum / (n * sigma_g * sigma_u) with
    ddof=1 sigmas; 0.0 when either map is constant (sigma == 0)."""
    n = float(n)
    if n < 2:
        return 0.0
    var_g = (sgg - sg * sg / n)
    var_u = (suu - su * su / n)
    # a constant map gives a centred sum of squares that is zero up to rounding
    if var_g <= 1e-13 * abs(sgg) or var_u <= 1e-13 * abs(suu):
        return 0.0
    cov = sgu - sg * su / n
    sigma_g = np.sqrt(var_g / (n - 1))
    sigma_u = np.sqrt(var_u / (n - 1))
    return float((1 / (n * sigma_g * sigma_u)) * cov)


def compute_ncc(gt_unc_map, pred_unc_map):
    """Drop-in for ncc.py:9-28: a ready GT uncertainty map (e.g. np.var of the
    raters, or GTA's analytic map) against a predicted map."""
    _lib.require_device()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    g = gt_unc_map if isinstance(gt_unc_map, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(gt_unc_map))
    p = pred_unc_map if isinstance(pred_unc_map, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pred_unc_map))
    if g.numel() != p.numel():
        raise ValueError("gt and predicted uncertainty maps must have the same number of elements")
    g = g.to(dev).to(torch.float64).contiguous().reshape(-1)
    p = p.to(dev).to(torch.float32).contiguous().reshape(-1)
    n = p.numel()
    if n == 0:
        return 0.0
    sf = torch.zeros((1, F64["COLS"]), dtype=torch.float64, device=dev)
    si = torch.zeros((1, I64["COLS"]), dtype=torch.int64, device=dev)
    a = _lib.MapStatsArgs()
    a.struct_size = C.sizeof(_lib.MapStatsArgs)
    a.stat_flags = _lib.STAT_NCC
    a.B, a.V = 1, n
    a.maps[0] = p.data_ptr()
    a.ncc_gt_map = g.data_ptr()
    a.stats_f64, a.stats_i64 = sf.data_ptr(), si.data_ptr()
    _lib.check(lib.vu_map_stats(C.byref(a), _lib.current_stream_ptr()), "vu_map_stats")
    f = sf.cpu().numpy()[0]
    return ncc_from_sums(n, f[F64["NCC_G"]], f[F64["NCC_GG"]], f[F64["NCC_U"]], f[F64["NCC_UU"]], f[F64["NCC_GU"]])


def ncc_from_result(result, unc_index: int = 2) -> np.ndarray:
    """Per-image NCC of one uncertainty type (0 TU, 1 AU, 2 EU) from a fused pass
    run with STAT_NCC and the raters as ground truth."""
    s = result.ncc_sums()
    return np.array([ncc_from_sums(result.n_voxels, s["g"][b], s["gg"][b], s["u"][b, unc_index],
                                   s["uu"][b, unc_index], s["gu"][b, unc_index]) for b in range(len(s["g"]))])
