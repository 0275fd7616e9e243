"""Failure-detection finalisation on the host (evaluation/metrics/aurc.py:14-67,
uncertainty_modeling/test_2D.py:884-899).  The per-image risks (1 - Dice from the
integer counts the fused pass emits) and confidences (-score) are image-count
sized, so the O(n log n) curve stays on the CPU in float64 by design (SURVEY
section 8a, row a16).
"""
from __future__ import annotations

import numpy as np


def binary_dice_from_counts(tp, pred_sum, gt_sum) -> np.ndarray:
    """test_2D.py:884-899 for (B, R) count arrays -> (B,) mean Dice over raters,
    in float32 like the reference: both empty -> 1, exactly one empty -> 0."""
    tp = np.asarray(tp, np.float32)
    ps = np.asarray(pred_sum, np.float32)
    gs = np.asarray(gt_sum, np.float32)
    denom = 2 * tp + (ps - tp) + (gs - tp)
    dice = np.zeros_like(denom)
    dice[(ps == 0) & (gs == 0)] = 1.0
    regular = (ps != 0) & (gs != 0) & (denom > 0)
    dice[regular] = (2 * tp[regular]) / denom[regular]
    return dice.mean(axis=-1)


def multiclass_dice_from_counts(tp, pred_sum, gt_sum, include_background: bool = False) -> np.ndarray:
    """The general path of ``Tester.calculate_test_metrics`` (test_2D.py:901-918) on the integer counts of
    ``STAT_CLASS_COUNTS``: for every rater ``dice(pred, rater, ignore_index, include_background=False, average="macro")``
    (evaluation/metrics/dice_wrapped.py:17-104), then the mean over raters.  (..., R, C) count arrays -> (...,) float64.

    dice_wrapped moves the ignored pixels into channel 0 of prediction AND target and drops that channel (together with
    class 0 when ``include_background`` is False): what is left per class c is tp = #(pred == c & gt == c), #(pred == c) and
    #(gt == c) over the pixels the rater does not ignore -- the counts the kernel emits.  The macro reduction is
    torchmetrics' ``DiceScore(average="macro", aggregation_level="global", include_background=False)``: 2 tp / (pred + gt) per
    class, classes that occur in neither prediction nor target (0 / 0) skipped by a nan-mean.  torchmetrics is absent from
    the reference tree and from this image, so THIS reduction is restated from the published algorithm (torchmetrics >= 1.6,
    ``segmentation/dice.py::_dice_score_compute``) and not pinned against it; the counts it consumes are pinned bit-exactly.
    dice_wrapped's early returns: every pixel ignored -> 1; prediction and target all background -> 1."""
    tp = np.asarray(tp, np.float64)
    ps = np.asarray(pred_sum, np.float64)
    gs = np.asarray(gt_sum, np.float64)
    first = 0 if include_background else 1
    valid_pixels = ps.sum(axis=-1)                       # pixels the rater does not ignore (every pixel has one label)
    num, den = 2.0 * tp[..., first:], ps[..., first:] + gs[..., first:]
    present = den > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        per_class = np.where(present, num / np.where(present, den, 1.0), np.nan)
        n_present = present.sum(axis=-1)
        macro = np.where(n_present > 0, np.nansum(per_class, axis=-1) / np.maximum(n_present, 1), np.nan)
    macro = np.where(valid_pixels == 0, 1.0, macro)     # dice_wrapped.py:93-94
    if not include_background:
        macro = np.where((valid_pixels > 0) & (n_present == 0), 1.0, macro)  # :95-99: only background on both sides
    return macro.mean(axis=-1)


def rc_curve_stats(risks, confids):
    """aurc.py:14-51 vectorised: images sorted by confidence; a curve point is
    emitted after dropping image i only if i == 0 or its confidence differs from
    the previous one's; weights are the numbers of images dropped in between."""
    risks = np.asarray(risks, np.float64)
    confids = np.asarray(confids, np.float64)
    assert risks.ndim == 1 and confids.ndim == 1 and len(risks) == len(confids)
    n = len(risks)
    order = np.argsort(confids)
    r, c = risks[order], confids[order]
    total = float(sum(r))
    coverages = [1.0]
    sel = [total / n]
    weights = []
    if n > 1:
        i = np.arange(n - 1)
        # error sum after removing images 0..i, in the reference's left-to-right order
        err = total - np.cumsum(r[:-1])
        emit = np.ones(n - 1, bool)
        emit[1:] = c[1:n - 1] != c[0:n - 2]
        idx = i[emit]
        coverages += list((n - 1 - idx) / n)
        sel += list(err[idx] / (n - 1 - idx))
        prev = np.concatenate([[-1], idx[:-1]])
        weights += list((idx - prev) / n)
        pending = (n - 2) - idx[-1]
        if pending > 0:
            coverages.append(0)
            sel.append(sel[-1])
            weights.append(pending / n)
    return coverages, sel, weights


def aurc(risks, confids) -> float:
    """aurc.py:54-58."""
    _, r, w = rc_curve_stats(risks, confids)
    r = np.asarray(r)
    w = np.asarray(w)
    return float(np.sum((r[:-1] + r[1:]) * 0.5 * w)) if len(w) else 0.0


def eaurc(risks, confids) -> float:
    """aurc.py:61-67."""
    risks = np.asarray(risks, np.float64)
    n = len(risks)
    best = np.sort(risks).cumsum() / np.arange(1, n + 1)
    return aurc(risks, confids) - float(best.sum() / n)
