"""Import the UNMODIFIED reference hot-path functions from /root/reference.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the reference
tree does not travel to the GPU box).  Used by ``oracle/make_golden.py`` to
generate the committed fixtures under ``tests/golden/`` and by the
``-m "not gpu"`` tests (skipped when /root/reference is absent) to pin
``oracle/oracle.py`` against the real thing.

The hot-path modules import hydra / medpy / jsbeautifier / pytorch_lightning /
torchmetrics / albumentations / ... at module top level but the hot functions
never touch them, so those names are stubbed with
MagicMock in sys.modules before the import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("VU_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "hydra", "hydra.utils", "hydra.core", "hydra.core.global_hydra", "hydra.core.hydra_config",
    "omegaconf", "medpy", "medpy.io", "jsbeautifier", "SimpleITK",
    "pytorch_lightning", "pytorch_lightning.callbacks", "pytorch_lightning.loggers",
    "torchmetrics", "torchmetrics.segmentation", "torchmetrics.utilities",
    "torchmetrics.utilities.enums", "torchmetrics.utilities.checks", "torchmetrics.utilities.data",
    "torchmetrics.functional", "torchmetrics.functional.classification",
    "torchmetrics.functional.classification.stat_scores",
    "albumentations", "albumentations.pytorch", "albumentations.core",
    "albumentations.core.transforms_interface", "tifffile", "matplotlib", "matplotlib.pyplot",
    "skimage", "skimage.io", "seaborn", "wandb",
    # uncertainty_modeling.test_2D (only for Tester._compute_likelihood_stats / _compute_expected_nll)
    "pytorch_lightning.utilities", "pytorch_lightning.utilities.rank_zero", "pytorch_lightning.utilities.types",
    "pytorch_lightning.strategies", "pytorch_lightning.callbacks.progress", "loss_modules",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "evaluation", "metrics"))


def _install_stubs() -> None:
    for name in _STUBS:
        try:
            if name not in sys.modules:
                importlib.import_module(name)
        except Exception:
            m = MagicMock(name=name)
            m.__path__ = []  # behave like a package
            m.__spec__ = None
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load() -> types.SimpleNamespace:
    """Return a namespace with the reference's own hot-path callables."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    tu = importlib.import_module("uncertainty_modeling.unc_mod_utils.test_utils")
    agg = importlib.import_module("evaluation.uncertainty_aggregation.aggregate_uncertainties")
    shp = importlib.import_module("evaluation.uncertainty_aggregation.prediction_shape_stats")
    ace = importlib.import_module("evaluation.metrics.ace")
    ncc = importlib.import_module("evaluation.metrics.ncc")
    aurc = importlib.import_module("evaluation.metrics.aurc")
    thr = importlib.import_module("evaluation.uncertainty_aggregation.find_threshold")
    ged = importlib.import_module("evaluation.metrics.ged_fast")
    try:
        tester = importlib.import_module("uncertainty_modeling.test_2D").Tester
    except Exception:  # pragma: no cover - depends on which optional packages this container has
        tester = None
    return types.SimpleNamespace(
        calculate_uncertainty=tu.calculate_uncertainty,
        calculate_one_minus_msr=tu.calculate_one_minus_msr,
        patch_level_aggregation=agg.patch_level_aggregation,
        image_level_aggregation=agg.image_level_aggregation,
        threshold_aggregation=agg.threshold_aggregation,
        normalize_uncertainty_sum=agg._normalize_uncertainty_sum,
        compute_area=shp._compute_area,
        compute_border=shp._compute_border,
        platt_scale_params=ace.platt_scale_params,
        ace_module=ace,
        agg_module=agg,
        shp_module=shp,
        ncc_module=ncc,
        aurc_module=aurc,
        thr_module=thr,
        ged_binary_fast=ged.ged_binary_fast,
        Tester=tester,
        calculate_foreground_quantile_image=thr.calculate_foreground_quantile_image,
        calculate_threshold_image=thr.calculate_threshold_image,
        platt_scale_confid=ace.platt_scale_confid,
        calib_stats=ace.calib_stats,
        calc_ace=ace.calc_ace,
        calc_ece=ace.calc_ece,
        calc_eqace=ace.calc_eqace,
        GlobalCalibAccumulator=ace.GlobalCalibAccumulator,
        compute_ncc=ncc.compute_ncc,
        rc_curve_stats=aurc.rc_curve_stats,
        aurc=aurc.aurc,
        eaurc=aurc.eaurc,
    )
