"""Evaluation tasks behind the reference's driver signatures, i.e. what the Hydra ``_target_`` strings of
``evaluation/configs/tasks/*.yaml`` resolve to (``evaluation/eval_experiments.py:336,355``):

    aggregate_uncertainties(exp_dataloader, aggregations)   aggregate_uncertainties.py:133-188 -> aggregated_<unc>.json
    compute_prediction_shape_stats(exp_dataloader, ...)     prediction_shape_stats.py:70-103   -> area.json
    ncc_main(exp_dataloader)                                 ncc.py:46-164                      -> ambiguity_modeling.json
    aurc_main(exp_dataloader)                                aurc.py:130-153                    -> failure_detection.json
    calibration_error(exp_dataloader, ignore_value)          ace.py:463-534                     -> calibration.json
    calibration_main(exp_dataloader, ignore_value)           ace.py:537-545
    get_foreground_quantile / save_foreground_quantiles /    find_threshold.py:15-66, 80-112 -> quantile_analysis.json,
    threshold_images_paths / find_threshold                                                      threshold_analysis.json

They take the reference's ``ExperimentDataloader`` (any object with the same attributes works), load each array once
per image instead of once per aggregation, run the arithmetic on the GPU through libvalunc and write the same JSON
files with the same keys.  Only file discovery, JSON and the image-count sized finalisation run on the host.
``eqace`` (ace.py:378-406, per-image quantile bins) comes from an exact GPU rank selection
(``calibration.eqace_from_maps``); pass ``eqace_fn=None`` to skip it.
"""
from __future__ import annotations

import importlib
import json
import os
from pathlib import Path
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import _lib, aggregation as _agg, aurc as _aurc, calibration as _cal, ncc as _ncc, quantile as _qt

_LOCAL_TARGETS = {
    "patch_level_aggregation": _agg.patch_level_aggregation,
    "image_level_aggregation": _agg.image_level_aggregation,
    "threshold_aggregation": _agg.threshold_aggregation,
    "area_normalized_aggregation": _agg.area_normalized_aggregation,
    "border_normalized_aggregation": _agg.border_normalized_aggregation,
}


def _resolve(aggregation_config) -> tuple:
    """(callable, kwargs) of one aggregation entry: a callable, or a mapping with a Hydra-style ``_target_`` whose last
    component names one of the aggregation functions (the reference's own dotted paths resolve to the GPU versions)."""
    if callable(aggregation_config):
        return aggregation_config, {}
    cfg = dict(aggregation_config)
    target = cfg.pop("_target_", "")
    name = str(target).rsplit(".", 1)[-1]
    if name in _LOCAL_TARGETS:
        return _LOCAL_TARGETS[name], cfg
    module, _, attr = str(target).rpartition(".")
    return getattr(importlib.import_module(module), attr), cfg


def _load_unc_image(exp_dataloader, unc: str, unc_path: Path, image_id) -> np.ndarray:
    """The array the reference's ``medpy.io.load`` would return for this file (aggregate_uncertainties.py:140-142);
    the loader object is asked first (``load_unc_file``), medpy only if it is installed."""
    name = f"{image_id}{exp_dataloader.exp_version.unc_ending}"
    if hasattr(exp_dataloader, "load_unc_file"):
        return np.asarray(exp_dataloader.load_unc_file(unc, image_id))
    from medpy.io import load  # noqa: WPS433 (optional dependency of the reference)
    return load(unc_path / name)[0]


def aggregate_uncertainties(exp_dataloader, aggregations) -> Dict[str, dict]:
    """aggregate_uncertainties.py:133-188."""
    written = {}
    for unc, unc_path in exp_dataloader.unc_path_dict.items():
        all_uncs = {}
        for image_id in exp_dataloader.image_ids:
            key = f"{image_id}{exp_dataloader.exp_version.unc_ending}"
            all_uncs[key] = {}
            image = _load_unc_image(exp_dataloader, unc, Path(unc_path), image_id)
            dev_image = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).cuda()  # one upload for all aggregations
            for aggregation in aggregations:
                fn, kwargs = _resolve(aggregations[aggregation])
                call_kwargs = {"image": dev_image, "pred_model": exp_dataloader.exp_version.pred_model, "unc_type": unc,
                               "image_id": image_id, "dataset_path": exp_dataloader.dataset_path}
                threshold_path_cfg = kwargs.get("threshold_path")
                if fn is _agg.threshold_aggregation and (threshold_path_cfg is None or str(threshold_path_cfg).lower() == "none") \
                        and kwargs.get("threshold") is None:
                    kwargs = {k: v for k, v in kwargs.items() if k != "threshold_path"}
                    call_kwargs["threshold_path"] = exp_dataloader.exp_version.exp_path / "threshold_analysis.json"
                res = fn(**kwargs, **call_kwargs)
                all_uncs[key][aggregation] = {k: (float(v) if isinstance(v, np.floating) else v) for k, v in res.items()}
        save_path = Path(exp_dataloader.dataset_path) / f"aggregated_{unc}.json"
        with open(save_path, "w") as f:
            json.dump(all_uncs, f, indent=4)
        written[unc] = all_uncs
    return written


def compute_prediction_shape_stats(exp_dataloader, mean_pred: bool = True, stats_filename: str = "area.json",
                                   majority_threshold: float = 0.5):
    """prediction_shape_stats.py:70-103."""
    if not 0.0 < majority_threshold <= 1.0:
        raise ValueError("majority_threshold must be within (0, 1].")
    if exp_dataloader.dataset_path is None:
        raise ValueError("Area/border statistics require a single dataset split (no paired splits).")
    stats = {}
    for image_id in exp_dataloader.image_ids:
        pred_segs = [np.asarray(p) for p in exp_dataloader.get_pred_segs(image_id)]
        if not pred_segs:
            raise ValueError("No prediction segmentations available to compute statistics.")
        if mean_pred:
            mask = None
            try:
                mask = exp_dataloader.get_mean_pred_seg(image_id)
            except Exception:
                mask = None
            if mask is None:  # _majority_mask (prediction_shape_stats.py:45-48)
                stack = torch.from_numpy(np.stack(pred_segs, 0)).cuda()
                mask = ((stack > 0).double().mean(0) >= majority_threshold).to(torch.uint8)
            area, border = _agg.prediction_shape_stats(mask)
        else:
            pairs = [_agg.prediction_shape_stats(p) for p in pred_segs]
            area = float(np.mean([p[0] for p in pairs]))
            border = float(np.mean([p[1] for p in pairs]))
        stats[str(image_id)] = {"area": float(area), "border": float(border)}
    stats_path = Path(exp_dataloader.dataset_path) / stats_filename
    stats_path.parent.mkdir(parents=True, exist_ok=True)
    with open(stats_path, "w") as f:
        json.dump(stats, f, indent=2)
    return stats


def ncc_main(exp_dataloader, plot: bool = False):
    """ncc.py:46-164 without the diagnostic plots."""
    ncc_dict = {"mean": {}}
    for unc_type in exp_dataloader.exp_version.unc_types:
        values = []
        for image_id in exp_dataloader.image_ids:
            ncc_dict.setdefault(image_id, {})
            ncc = _ncc.compute_ncc(exp_dataloader.get_gt_unc_map(image_id), exp_dataloader.get_unc_map(image_id, unc_type))
            ncc_dict[image_id][unc_type] = {"metrics": {"ncc": ncc}}
            values.append(ncc)
        ncc_dict["mean"][unc_type] = {"metrics": {"ncc": float(np.mean(np.array(values)))}}
    with open(Path(exp_dataloader.dataset_path) / "ambiguity_modeling.json", "w") as f:
        json.dump(ncc_dict, f, indent=2)
    return ncc_dict


def _metric_entry(metrics: dict, image_id: str) -> dict:
    if image_id in metrics:
        return metrics[image_id]
    keys = [k for k in metrics if k.split("/")[-1].split(".")[0] == image_id]  # aurc.py:73-77
    return metrics[keys[0]]


def aurc_main(exp_dataloader):
    """aurc.py:130-153; metrics.json and aggregated_<unc>.json are read once instead of once per image."""
    dataset_path = Path(exp_dataloader.dataset_path)
    with open(dataset_path / "metrics.json") as f:
        metrics = json.load(f)
    results = {"mean": {}}
    for unc_type in exp_dataloader.exp_version.unc_types:
        with open(dataset_path / f"aggregated_{unc_type}.json") as f:
            agg = json.load(f)
        results["mean"][unc_type] = {}
        for aggregation in exp_dataloader.exp_version.aggregations:
            risks, confids = [], []
            for image in exp_dataloader.image_ids:
                entry = _metric_entry(metrics, image)
                dice = entry["dice"] if "dice" in entry else entry["metrics"]["dice"]
                risks.append(1 - dice)
                confids.append(-agg[f"{image}{exp_dataloader.exp_version.unc_ending}"][aggregation]["max_score"])
            results["mean"][unc_type][aggregation] = {"metrics": {"aurc": _aurc.aurc(np.array(risks), np.array(confids)),
                                                                  "eaurc": _aurc.eaurc(np.array(risks), np.array(confids))}}
    with open(dataset_path / "failure_detection.json", "w") as f:
        json.dump(results, f, indent=2)
    return results


def calibration_error(exp_dataloader, ignore_value=None, eqace_fn: Optional[Callable] = _cal.eqace_from_maps):
    """ace.py:463-534: per image and uncertainty type the 21-slot histogram comes from one vu_map_stats launch on
    (references, prediction, uncertainty map, Platt parameters); ACE / ECE and the dataset accumulator are finalised in
    float64 on the host."""
    lib = _lib.load()
    import ctypes as C
    exp_path = Path(exp_dataloader.exp_version.exp_path)
    platt_file = exp_path / "platt_scale_params.json"
    calib_dict = {"mean": {}}
    for unc_type in exp_dataloader.exp_version.unc_types:
        a_p, b_p = _cal.load_platt_params(platt_file, unc_type)
        edges = _cal.platt_edges(a_p, b_p).as_struct()
        aces, eces, eqaces = [], [], []
        acc = _cal.GlobalCalibAccumulator()
        for image_id in exp_dataloader.image_ids:
            calib_dict.setdefault(image_id, {})
            refs = np.asarray(exp_dataloader.get_reference_segs(image_id))
            pred = np.asarray(exp_dataloader.get_mean_pred_seg(image_id))
            unc = np.asarray(exp_dataloader.get_unc_map(image_id, unc_type))
            if pred.shape != unc.shape:  # 2d unc map is loaded in shape (W, H) (ace.py:479-484)
                unc = np.swapaxes(unc, 0, 1)
            dev = torch.device("cuda", torch.cuda.current_device())
            r = torch.from_numpy(np.ascontiguousarray(refs)).to(dev)
            r = r.to(torch.uint8 if r.dtype in (torch.uint8, torch.bool) else torch.int64).contiguous()
            p = torch.from_numpy(np.ascontiguousarray(pred)).to(dev).to(torch.uint8).contiguous()
            u = torch.from_numpy(np.ascontiguousarray(unc, dtype=np.float32)).to(dev)
            V = p.numel()
            sf = torch.zeros((1, _lib.F64["COLS"]), dtype=torch.float64, device=dev)
            si = torch.zeros((1, _lib.I64["COLS"]), dtype=torch.int64, device=dev)
            a = _lib.MapStatsArgs()
            a.struct_size = C.sizeof(_lib.MapStatsArgs)
            a.stat_flags = _lib.STAT_CALIB
            a.B, a.V = 1, V
            a.maps[0] = u.data_ptr()
            a.labels = p.data_ptr()
            a.gt.data, a.gt.R = r.data_ptr(), r.shape[0]
            a.gt.dtype = _lib.GT_U8 if r.dtype == torch.uint8 else _lib.GT_I64
            a.gt.stride_b, a.gt.stride_r, a.gt.stride_v = r.numel(), V, 1
            a.gt.has_ignore, a.gt.ignore_index = (0, 0) if ignore_value is None else (1, int(ignore_value))
            for k in range(3):
                a.calib[k] = edges
            a.stats_f64, a.stats_i64 = sf.data_ptr(), si.data_ptr()
            _lib.check(lib.vu_map_stats(C.byref(a), _lib.current_stream_ptr()), "vu_map_stats")
            f, i = sf.cpu().numpy()[0], si.cpu().numpy()[0]
            s = f[_lib.F64["BIN_SUMS"]:_lib.F64["BIN_SUMS"] + 21]
            t = i[_lib.I64["BIN_TRUE"]:_lib.I64["BIN_TRUE"] + 21]
            n = i[_lib.I64["BIN_TOTAL"]:_lib.I64["BIN_TOTAL"] + 21]
            ace, ece = _cal.per_image_ace_ece(s, t, n)
            eqace = eqace_fn(r, p, u, a_p, b_p, ignore_value) if eqace_fn is not None else None  # device tensors: no second upload
            acc.accumulate_histogram(s, t, n)
            calib_dict[image_id][unc_type] = {"metrics": {"ace": ace, "ece": ece, "eqace": eqace}}
            aces.append(ace); eces.append(ece)
            if eqace is not None:
                eqaces.append(eqace)
        calib_dict["mean"][unc_type] = {"metrics": {
            "ace": float(np.mean(np.array(aces))), "ece": float(np.mean(np.array(eces))),
            "eqace": float(np.mean(np.array(eqaces))) if eqaces else None,
            "gace": acc.compute_ace(), "gece": acc.compute_ece()}}
    with open(Path(exp_dataloader.dataset_path) / "calibration.json", "w") as f:
        json.dump(calib_dict, f, indent=2)
    return calib_dict


def calibration_main(exp_dataloader, ignore_value=None, val_exp_dataloader=None):
    """ace.py:537-545: fit the Platt parameters on the validation split if the JSON is missing, then evaluate."""
    platt_file = Path(exp_dataloader.exp_version.exp_path) / "platt_scale_params.json"
    if not os.path.isfile(platt_file):
        if val_exp_dataloader is None:
            from evaluation.experiment_dataloader import ExperimentDataloader  # the reference's own loader
            val_exp_dataloader = ExperimentDataloader(exp_dataloader.exp_version, "val")
        _cal.platt_scale_params(val_exp_dataloader, ignore_value=ignore_value)
    return calibration_error(exp_dataloader, ignore_value=ignore_value)


# ---- threshold discovery on the validation split (find_threshold.py) ------------------------------------------------
def get_foreground_quantile(exp_dataloader) -> dict:
    """find_threshold.py:15-30: background share of every member prediction of every image (area counts on the GPU)."""
    version = exp_dataloader.exp_version
    all_quantiles = []
    for image_id in exp_dataloader.image_ids:
        for pred_seg in exp_dataloader.get_pred_segs(image_id):
            all_quantiles.append(_qt.calculate_foreground_quantile_image(pred_seg))
    return {version.pred_model: {version.version_name: {"quantiles": all_quantiles, "exp_path": Path(version.exp_path).as_posix()}}}


def save_foreground_quantiles(results_dict, save_path=None) -> None:
    """find_threshold.py:33-47."""
    for method, versions in results_dict.items():
        for _version_name, version_data in versions.items():
            exp_path = Path(version_data["exp_path"])
            exp_path.mkdir(parents=True, exist_ok=True)
            quantiles = version_data["quantiles"]
            if not quantiles:
                continue
            with open(exp_path / "quantile_analysis.json", "w") as f:
                json.dump({method: float(np.mean(np.array(quantiles)))}, f, indent=2)


def threshold_images_paths(exp_dataloader) -> dict:
    """find_threshold.py:50-66.  The loader rides along under ``"exp_dataloader"`` so that find_threshold can ask it for
    the maps (``load_unc_file``) instead of medpy."""
    version = exp_dataloader.exp_version
    version_dict = {"exp_path": Path(version.exp_path).as_posix(), "unc_paths": {}, "exp_dataloader": exp_dataloader}
    for unc_type in version.unc_types:
        unc_path = Path(exp_dataloader.unc_path_dict[unc_type])
        version_dict["unc_paths"][unc_type] = [(unc_path / f"{image_id}{version.unc_ending}").as_posix()
                                               for image_id in exp_dataloader.image_ids]
    return {version.pred_model: {version.version_name: version_dict}}


def find_threshold(results_dict, quantile_path=None, save_path=None) -> None:
    """find_threshold.py:80-112: per uncertainty type the quantile of *all* validation maps together.  The maps are
    uploaded one by one and folded into one radix selection; the concatenation is never built."""
    for pred_model, versions in results_dict.items():
        for version_name, version_data in versions.items():
            exp_path = Path(version_data["exp_path"])
            exp_path.mkdir(parents=True, exist_ok=True)
            quantile_file = exp_path / "quantile_analysis.json"
            if not quantile_file.is_file():
                raise FileNotFoundError(f"Quantile file not found for {pred_model} {version_name}: {quantile_file}")
            loader = version_data.get("exp_dataloader")
            threshold_entries = {}
            for unc, paths in version_data["unc_paths"].items():
                if not paths:
                    continue
                maps = []
                for path in paths:
                    path = Path(path)
                    if loader is not None and hasattr(loader, "load_unc_file"):
                        image = np.asarray(loader.load_unc_file(unc, path.name[: len(path.name) - len(loader.exp_version.unc_ending)]))
                    else:
                        from medpy.io import load  # noqa: WPS433 (optional dependency of the reference)
                        image = load(path)[0]
                    maps.append(torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).cuda())
                threshold_entries[f"Mean {unc.split('_')[0]} threshold"] = _qt.calculate_threshold_image(quantile_file, maps, method=pred_model)
            if not threshold_entries:
                continue
            with open(exp_path / "threshold_analysis.json", "w") as f:
                json.dump({pred_model: threshold_entries}, f, indent=2)
