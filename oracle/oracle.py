"""CPU oracle for the ValUES per-pixel uncertainty hot path.

TEST INFRASTRUCTURE ONLY.  This module restates, on the CPU with torch/NumPy,
the algorithm of the reference (JakobLC/DiffUncertainty) for the one path this
repository accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it;
the product package ``diffuncertainty_b200`` never does (it fails loudly when
the CUDA library is missing instead of falling back to anything here).

Parity pinning: every function below is checked in ``tests/test_oracle.py``
against (a) the committed fixtures in ``tests/golden/`` that were produced by
importing the *unmodified* reference functions (``oracle/make_golden.py`` via
``oracle/ref_shim.py``) and (b), when ``/root/reference`` is mounted, the
reference functions themselves on fresh seeded inputs.

All ``file:line`` citations are relative to the reference repository root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

N_CALIB_BINS = 20  # evaluation/metrics/ace.py:334,417


# --------------------------------------------------------------------------
# mean over the member axis (uncertainty_modeling/test_2D.py:971,
# unc_mod_utils/test_utils.py:835,854)
# --------------------------------------------------------------------------
def cascade_sum_f32(x: np.ndarray) -> np.ndarray:
    """fp32 sum over axis 0 in the exact order torch's CPU kernel uses.

    ``torch.mean(t, dim=0)`` (test_2D.py:971, test_utils.py:835) reduces an
    outer dimension with ATen's cascade sum: members are added sequentially
    into a level-0 accumulator that starts at zero; every ``2**k`` members
    (k = max(4, ceil(log2 P) // 4), i.e. 16 for P <= 2**19) the level-0 value
    is added into level 1 and reset, level 1 is flushed into level 2 every
    ``2**(2k)`` members, and so on for 4 levels; the leftover members go to
    level 0 and the levels are finally added 0 <- 1 <- 2 <- 3.
    Probed bit-exact against torch 2.11 for P in 2..600 (tests/test_oracle.py).

    torch deviates from this order only for the last ``numel % 32`` elements of
    a contiguous reduction row (its SIMD remainder loop adds 4 interleaved
    partial sums, see ``interleaved_tail_sum_f32``); the canonical order below
    is what the CUDA kernel reproduces.
    """
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[0]
    k = max(4, (math.ceil(math.log2(n)) if n > 1 else 0) // 4)
    step = 1 << k
    level = [np.zeros(x.shape[1:], dtype=np.float32) for _ in range(4)]
    i = 0
    while i + step <= n:
        for _ in range(step):
            level[0] = level[0] + x[i]
            i += 1
        for j in range(1, 4):
            level[j] = level[j] + level[j - 1]
            level[j - 1] = np.zeros_like(level[0])
            if i & ((step - 1) << (j * k)):
                break
    while i < n:
        level[0] = level[0] + x[i]
        i += 1
    for j in range(1, 4):
        level[0] = level[0] + level[j]
    return level[0]


def interleaved_tail_sum_f32(x: np.ndarray) -> np.ndarray:
    """The order torch uses for the SIMD remainder columns of an outer sum:
    four partial cascade sums over members 0,4,8.. / 1,5,9.. / 2.. / 3..,
    leftovers added to partial 0, then partials added 0 <- 1 <- 2 <- 3."""
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[0]
    q = n // 4
    if q > 0:
        parts = cascade_sum_f32(x[: 4 * q].reshape(q, 4, *x.shape[1:]))
    else:
        parts = np.zeros((4,) + x.shape[1:], dtype=np.float32)
    acc = parts[0]
    for i in range(4 * q, n):
        acc = acc + x[i]
    for j in range(1, 4):
        acc = acc + parts[j]
    return acc


def mean_members_f32(x: np.ndarray) -> np.ndarray:
    """``torch.mean(x, dim=0)`` in canonical order: cascade sum, then a true
    fp32 division by P (not a multiply by the reciprocal)."""
    x = np.asarray(x, dtype=np.float32)
    return cascade_sum_f32(x) / np.float32(x.shape[0])


def argmax_first_nan_max(m: np.ndarray) -> np.ndarray:
    """``torch.argmax(dim=0)`` semantics (test_2D.py:817,871): first maximal
    index, a NaN compares as the maximum (first NaN wins)."""
    m = np.asarray(m)
    nan = np.isnan(m)
    has_nan = nan.any(axis=0)
    first_nan = np.argmax(nan, axis=0)
    plain = np.argmax(np.where(nan, -np.inf, m), axis=0)
    return np.where(has_nan, first_nan, plain).astype(np.int64)


def mean_and_label(image_preds: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's own ops for one image (test_2D.py:969-971, :871):
    mean over members with torch, argmax over classes, cast as save_prediction
    does (test_2D.py:817-818)."""
    mean_softmax = torch.mean(image_preds, dim=0)
    label = mean_softmax.argmax(dim=0)
    return mean_softmax, label.to(torch.uint8)


# --------------------------------------------------------------------------
# C2 uncertainty measures (unc_mod_utils/test_utils.py:833-864)
# --------------------------------------------------------------------------
def _neg_plogp_sum(stack: torch.Tensor, out_shape, device) -> torch.Tensor:
    """-sum_c p_c log p_c for a (C, *S) stack the way test_utils.py:836-841 /
    :848-852 do it: class by class in fp32, NaN products skipped (p == 0,
    p < 0 and NaN contribute nothing), sign flipped at the end."""
    acc = torch.zeros(*out_shape, device=device)
    for c in range(stack.shape[0]):
        term = stack[c] * torch.log(stack[c])
        keep = ~torch.isnan(term)
        acc[keep] += term[keep]
    acc *= -1
    return acc


def calculate_uncertainty(softmax_preds: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Restates test_utils.py:833-859.  (P, C, *S) -> TU, AU, EU of shape S,
    always fp32 (the accumulators are default-dtype zeros, :836,843-847)."""
    spatial = softmax_preds.shape[2:]
    dev = softmax_preds.device
    mean_softmax = torch.mean(softmax_preds, dim=0)  # :835
    total = _neg_plogp_sum(mean_softmax, spatial, mean_softmax.device)  # :836-841
    per_member = torch.zeros(softmax_preds.shape[0], *spatial, device=dev)  # :843-845
    for p in range(softmax_preds.shape[0]):  # :846-853
        per_member[p] = _neg_plogp_sum(softmax_preds[p], spatial, dev)
    aleatoric = torch.mean(per_member, dim=0)  # :854
    return {"TU": total, "AU": aleatoric, "EU": total - aleatoric}  # :855-858


def calculate_one_minus_msr(softmax_pred: torch.Tensor) -> Dict[str, torch.Tensor]:
    """test_utils.py:862-864: 1 - max_c p_c for a single member, key "pred_entropy"."""
    return {"pred_entropy": 1 - softmax_pred.max(dim=0)[0]}


def process_image(image_preds: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The per-image body of ``Tester.process_output`` restricted to the hot
    path (test_2D.py:969-971, :1004-1007, :817): mean, label, uncertainty."""
    mean_softmax, label = mean_and_label(image_preds)
    if image_preds.shape[0] > 1:
        unc = calculate_uncertainty(image_preds)
    else:
        unc = calculate_one_minus_msr(image_preds.squeeze(0))
    out = dict(unc)
    out["mean"] = mean_softmax
    out["label"] = label
    return out


# --------------------------------------------------------------------------
# C3 aggregation (evaluation/uncertainty_aggregation/aggregate_uncertainties.py)
# --------------------------------------------------------------------------
def image_level_aggregation(image: np.ndarray, mean: bool = True) -> Dict[str, float]:
    """aggregate_uncertainties.py:37-39."""
    total = np.sum(image)
    return {"max_score": float(total / image.size) if mean else float(total)}


def patch_level_aggregation(image: np.ndarray, patch_size, mean: bool = False) -> Dict[str, object]:
    """aggregate_uncertainties.py:16-34.  Box sum by ``scipy.signal.convolve``
    with a float64 ones kernel ("valid"), max, and the box whose corner is the
    first entry -- per axis, as ``np.where`` orders them -- that is
    ``np.isclose`` to the max.  Note the reference takes ``indices[0]`` of each
    axis' index array separately, which is the row-major first hit."""
    from scipy.signal import convolve

    if isinstance(patch_size, int):
        patch_size = image.ndim * [patch_size]
    box = convolve(image, np.ones(patch_size), mode="valid")
    if mean:
        box = box / np.prod(patch_size)
    peak = np.max(box)
    hits = np.where(np.isclose(box, peak))
    corner = [int(axis_hits[0]) for axis_hits in hits]
    return {
        "max_score": float(peak),
        "bounding_box": [(c, c + int(k)) for c, k in zip(corner, patch_size)],
    }


def box_sum_direct_f64(image: np.ndarray, patch_size: Sequence[int]) -> np.ndarray:
    """Direct (non-FFT) float64 "valid" box sum via an integral image; what
    the CUDA kernel computes.  Differs from the FFT path by ~1e-13 relative."""
    acc = np.asarray(image, dtype=np.float64)
    for axis, k in enumerate(patch_size):
        c = np.cumsum(acc, axis=axis)
        pad_shape = list(c.shape)
        pad_shape[axis] = 1
        c = np.concatenate([np.zeros(pad_shape), c], axis=axis)
        hi = [slice(None)] * acc.ndim
        lo = [slice(None)] * acc.ndim
        hi[axis] = slice(k, None)
        lo[axis] = slice(0, c.shape[axis] - k)
        acc = c[tuple(hi)] - c[tuple(lo)]
    return acc


def threshold_aggregation(image: np.ndarray, threshold: float, mean: bool = True) -> Dict[str, object]:
    """aggregate_uncertainties.py:124-130 (threshold given explicitly): sum and
    count of the pixels >= t; the mean only when the count is positive,
    otherwise the (zero) sum is returned."""
    sel = image >= threshold
    total = image[sel].sum()
    count = sel.sum()
    if mean and count > 0:
        return {"max_score": total / count, "threshold": threshold}
    return {"max_score": total, "threshold": threshold}


def normalized_sum(image: np.ndarray, divisor: float) -> float:
    """aggregate_uncertainties.py:70-74: sum / divisor, un-normalised when the
    divisor is <= 0."""
    total = float(np.sum(image))
    return total if divisor <= 0 else total / divisor


def compute_area(mask: np.ndarray) -> float:
    """prediction_shape_stats.py:10-12."""
    return float(np.count_nonzero(np.asarray(mask) > 0))


def compute_border(mask: np.ndarray) -> float:
    """prediction_shape_stats.py:15-30: per axis, count adjacent pairs whose
    labels differ; summed over axes."""
    mask = np.asarray(mask)
    if mask.size == 0:
        return 0.0
    total = 0
    for axis in range(mask.ndim):
        if mask.shape[axis] < 2:
            continue
        a = np.take(mask, range(0, mask.shape[axis] - 1), axis=axis)
        b = np.take(mask, range(1, mask.shape[axis]), axis=axis)
        total += int(np.count_nonzero(a != b))
    return float(total)


# --------------------------------------------------------------------------
# calibration (evaluation/metrics/ace.py)
# --------------------------------------------------------------------------
def platt_scale_confid(uncalib_confid: np.ndarray, a: float, b: float) -> np.ndarray:
    """ace.py:325-329 with the JSON lookup already done: 1 / (1 + exp(x a + b)),
    x = -uncertainty; a, b are Python floats so float32 input stays float32."""
    return 1 / (1 + np.exp(uncalib_confid * a + b))


def calib_bin_edges() -> np.ndarray:
    """ace.py:350: 21 float64 edges over [0, 1 + 1e-8]."""
    return np.linspace(0.0, 1.0 + 1e-8, N_CALIB_BINS + 1)


def _binarize_like_sklearn(correct: np.ndarray) -> np.ndarray:
    """ace.py:343-348: ``label_binarize(y, classes=unique(y))[:, 0]``.
    Two classes -> indicator of the larger one; a single class (all correct or
    all wrong) -> all zeros; more than two -> ValueError."""
    labels = np.unique(correct)
    if len(labels) > 2:
        raise ValueError(f"Only binary classification is supported. Provided labels {labels}.")
    if len(labels) < 2:
        return np.zeros(correct.shape[0], dtype=np.int64)
    return (correct == labels[1]).astype(np.int64)


def calib_histogram(correct: np.ndarray, calib_confids: np.ndarray, binarize: bool = True):
    """The three ``np.bincount`` arrays of ace.py:352-356 (per image, with the
    single-class quirk) or ace.py:431-437 (``binarize=False``, the global
    accumulator).  Returns (bin_sums f64[21], bin_true f64[21], bin_total i64[21])."""
    conf = np.clip(np.ravel(calib_confids), 0, 1)
    y = np.ravel(correct)
    y = _binarize_like_sklearn(y) if binarize else y.astype(np.float64)
    edges = calib_bin_edges()
    ids = np.digitize(conf, edges) - 1
    n = len(edges)
    return (
        np.bincount(ids, weights=conf, minlength=n),
        np.bincount(ids, weights=y, minlength=n),
        np.bincount(ids, minlength=n),
    )


def ace_ece_from_histogram(bin_sums, bin_true, bin_total) -> Tuple[float, float]:
    """ace.py:357-375 (and :439-460): ACE = mean |acc - conf| over non-empty
    bins, ECE = the same weighted by the bin's share of samples.  NaN when
    there are no samples (ace.py:443-444,453-454)."""
    bin_total = np.asarray(bin_total)
    filled = bin_total != 0
    n_filled = int(filled.sum())
    if n_filled == 0:
        return float("nan"), float("nan")
    acc = np.asarray(bin_true, dtype=np.float64)[filled] / bin_total[filled]
    conf = np.asarray(bin_sums, dtype=np.float64)[filled] / bin_total[filled]
    gap = np.abs(acc - conf)
    share = bin_total[filled] / bin_total.sum()
    return float((1 / n_filled) * np.sum(gap)), float(np.sum(gap * share))


def calc_ace(correct, calib_confids) -> float:
    """ace.py:368-370."""
    return ace_ece_from_histogram(*calib_histogram(correct, calib_confids))[0]


def calc_ece(correct, calib_confids) -> float:
    """ace.py:373-375."""
    return ace_ece_from_histogram(*calib_histogram(correct, calib_confids))[1]


class GlobalCalibAccumulator:
    """ace.py:409-460: dataset-level running histogram (no single-class quirk)."""

    def __init__(self) -> None:
        n = N_CALIB_BINS + 1
        self.bin_sums = np.zeros(n, dtype=np.float64)
        self.bin_true = np.zeros(n, dtype=np.float64)
        self.bin_total = np.zeros(n, dtype=np.int64)

    def accumulate(self, correct, calib_confids) -> None:
        s, t, n = calib_histogram(correct, calib_confids, binarize=False)
        self.bin_sums += s
        self.bin_true += t
        self.bin_total += n

    def compute_ace(self) -> float:
        return ace_ece_from_histogram(self.bin_sums, self.bin_true, self.bin_total)[0]

    def compute_ece(self) -> float:
        return ace_ece_from_histogram(self.bin_sums, self.bin_true, self.bin_total)[1]


def calibration_inputs(refs: np.ndarray, pred: np.ndarray, unc: np.ndarray, a: float, b: float,
                       ignore_value=None) -> Tuple[np.ndarray, np.ndarray]:
    """The per-image preparation in ``calibration_error`` (ace.py:484-515):
    rater-wise ``correct = (ref == pred)``, optional ignore mask on the
    references, Platt-scaled confidence of ``-unc``."""
    n_gt = refs.shape[0]
    pred_rep = np.repeat(pred[np.newaxis], n_gt, 0)
    unc_rep = np.repeat(unc[np.newaxis], n_gt, 0)
    correct = (refs == pred_rep).astype(int)
    if ignore_value is not None:
        keep = refs != ignore_value
        return correct[keep], platt_scale_confid(-unc_rep[keep], a, b)
    return correct.flatten(), platt_scale_confid(-unc_rep.flatten(), a, b)


# --------------------------------------------------------------------------
# quantiles: eqACE (evaluation/metrics/ace.py:378-406) and threshold discovery
# (evaluation/uncertainty_aggregation/find_threshold.py:10-30, 69-112)
# --------------------------------------------------------------------------
def quantile_linear(x: np.ndarray, q) -> np.ndarray:
    """``np.quantile(x, q)`` with the default method "linear" (NumPy is a third-party dependency of the reference; call
    sites ace.py:388 and find_threshold.py:76), restated from its published algorithm as NumPy 2.3 (installed here, the
    version the golden vectors were recorded with) runs it: a Python-scalar ``q`` is first cast to the data's float
    dtype, an array ``q`` keeps its own; virtual index ``h = (n - 1) q`` in that dtype; neighbours ``floor(h)`` and
    ``floor(h) + 1`` of the sorted data (NaN sorts last and poisons every quantile); weight ``g = h - floor(h)``; the
    lerp ``a + (b - a) g`` is replaced by ``b - (b - a)(1 - g)`` where ``g >= 0.5``.
    Version drift: numpy 1.24.3 (the reference's requirements.txt) keeps ``q`` and ``h`` in float64 for float32 data, so
    its threshold differs from this one by about one float32 rounding."""
    x = np.sort(np.ravel(np.asarray(x)))
    scalar = np.ndim(q) == 0
    if isinstance(q, (int, float)) and x.dtype.kind == "f":
        qs = np.atleast_1d(np.asarray(q, dtype=x.dtype))
    else:
        qs = np.atleast_1d(np.asarray(q))
    n = x.size
    h = (n - 1) * qs
    if n and np.isnan(x[-1]):
        out = np.full(qs.shape, np.nan, np.result_type(x.dtype, h.dtype))
        return out[0] if scalar else out
    lo = np.minimum(np.floor(h).astype(np.intp), n - 1)  # _get_indexes: indexes above bounds take the last element
    hi = np.minimum(lo + 1, n - 1)
    g = np.asarray(h - lo, dtype=h.dtype)
    a, b = x[lo], x[hi]
    diff = b - a
    out = np.where(g >= 0.5, b - diff * (1 - g), a + diff * g)
    return out[0] if scalar else out


def calc_eqace(correct, calib_confids, n_bins: int = N_CALIB_BINS) -> float:
    """ace.py:378-406: adaptive calibration error on the image's own quantile
    bins.  Confidences are clipped to [0, 1] and ranked in float64; the outer
    edges are replaced by 0 and 1 + 1e-8 and the edges made non-decreasing."""
    conf = np.clip(np.ravel(calib_confids), 0.0, 1.0).astype(np.float64)
    y = np.ravel(correct).astype(np.float64)
    if conf.size == 0:
        return float("nan")
    edges = quantile_linear(conf, np.linspace(0.0, 1.0, n_bins + 1))
    edges[0] = 0.0
    edges[-1] = 1.0 + 1e-8
    edges = np.maximum.accumulate(edges)
    ids = np.clip(np.digitize(conf, edges) - 1, 0, n_bins - 1)
    bin_sums = np.bincount(ids, weights=conf, minlength=n_bins)
    bin_true = np.bincount(ids, weights=y, minlength=n_bins)
    bin_total = np.bincount(ids, minlength=n_bins)
    filled = bin_total > 0
    if not filled.any():
        return float("nan")
    gap = np.abs(bin_true[filled] / bin_total[filled] - bin_sums[filled] / bin_total[filled])
    return float((1.0 / int(filled.sum())) * np.sum(gap))


def foreground_quantile(image: np.ndarray) -> float:
    """find_threshold.py:10-12: share of background pixels of one predicted segmentation."""
    image = np.asarray(image)
    return 1 - (np.count_nonzero(image) / image.size)


def mean_foreground_quantile(pred_segs: Sequence[np.ndarray]) -> float:
    """find_threshold.py:15-47: mean over every member prediction of every image."""
    return float(np.mean(np.array([foreground_quantile(p) for p in pred_segs])))


def uncertainty_threshold(unc_maps: Sequence[np.ndarray], method_quantile: float) -> float:
    """find_threshold.py:69-77, 96-105: ``np.quantile`` of the concatenation of all
    maps of one uncertainty type at the method's mean foreground quantile."""
    return float(quantile_linear(np.concatenate([np.ravel(m) for m in unc_maps]), method_quantile))


# --------------------------------------------------------------------------
# Platt-scaling fit on the validation split (evaluation/metrics/ace.py:14-285)
# --------------------------------------------------------------------------
N_PLATT_BINS = 256  # ace.py:17


def platt_fit_bin_edges(n_bins: int = N_PLATT_BINS) -> np.ndarray:
    """ace.py:31: 257 float64 edges, log-spaced over [1e-12, 1e2]."""
    return np.logspace(-12, 2, num=n_bins + 1, dtype=np.float64)


def platt_fit_histogram(refs: np.ndarray, pred: np.ndarray, unc: np.ndarray, ignore_value=None, n_bins: int = N_PLATT_BINS):
    """One image's contribution to the compressed fit data (ace.py:80-137): every valid (rater, pixel)
    pair is binned by the magnitude of its uncertainty (``np.digitize`` on the log-spaced edges, values
    outside clamped to the end bins); per bin the sample count, the correct / wrong counts and the sum of
    uncertainties.  Returns (total i64, pos i64, neg i64, sum_unc f64), each of length n_bins."""
    refs = np.asarray(refs)
    pred = np.asarray(pred)
    correct = refs == pred[None, ...]
    valid = (refs != ignore_value) if ignore_value is not None else np.ones(refs.shape, dtype=bool)
    u = np.broadcast_to(unc[None, ...], refs.shape)[valid].ravel()
    c = correct[valid].ravel().astype(np.int8)
    total = np.zeros(n_bins, np.int64)
    pos = np.zeros(n_bins, np.int64)
    neg = np.zeros(n_bins, np.int64)
    sums = np.zeros(n_bins, np.float64)
    if u.size == 0:
        return total, pos, neg, sums
    idx = np.digitize(u, platt_fit_bin_edges(n_bins)) - 1
    idx[idx < 0] = 0
    idx[idx >= n_bins] = n_bins - 1
    sums += np.bincount(idx, weights=u, minlength=n_bins)
    total += np.bincount(idx, minlength=n_bins)
    pos += np.bincount(idx[c == 1], minlength=n_bins)
    neg += np.bincount(idx[c == 0], minlength=n_bins)
    return total, pos, neg, sums


def platt_fit_samples(total, pos, neg, sum_unc):
    """ace.py:150-169: at most two weighted samples per non-empty bin, at F = -(mean uncertainty of the bin)."""
    total = np.asarray(total)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean_unc = np.divide(sum_unc, total, out=np.zeros_like(np.asarray(sum_unc, np.float64)), where=total > 0)
    F: List[float] = []
    y: List[int] = []
    w: List[int] = []
    for b in range(len(total)):
        if total[b] == 0:
            continue
        if pos[b] > 0:
            F.append(-mean_unc[b]); y.append(1); w.append(int(pos[b]))
        if neg[b] > 0:
            F.append(-mean_unc[b]); y.append(0); w.append(int(neg[b]))
    return np.asarray(F, np.float64), np.asarray(y, np.float64), np.asarray(w, np.float64)


def platt_fit(total, pos, neg, sum_unc) -> Tuple[float, float]:
    """ace.py:171-178: sklearn's Platt fit on the compressed data; (0, 0) when there are no samples."""
    from sklearn.calibration import _sigmoid_calibration
    F, y, w = platt_fit_samples(total, pos, neg, sum_unc)
    if len(F) == 0:
        return 0.0, 0.0
    a, b = _sigmoid_calibration(F, y, sample_weight=w)
    return float(a), float(b)


# --------------------------------------------------------------------------
# member-level scores on the same slab: GED (evaluation/metrics/ged_fast.py:5-142)
# and the likelihood statistics (uncertainty_modeling/test_2D.py:1043-1120)
# --------------------------------------------------------------------------
def ged_counts(member_labels: np.ndarray, gt: np.ndarray, ignore_index=None, mean_label: np.ndarray | None = None) -> Dict[str, np.ndarray]:
    """Every integer the GED of ged_fast.py is made of, from (P, V) member labels
    (argmax of each member, ged_fast.py:44) and (G, V) references:
    pg_tp / pg_pred (P, G) and g_sum (G) on each reference's valid voxels (:51-62), pp_tp (P, P) and pos (P) (:84-88),
    gg_tp / gg_sum (G, G) with [i, j] masked by reference j's validity (:95-103), and the majority counts (:121-131)."""
    lab = np.asarray(member_labels).reshape(member_labels.shape[0], -1)
    g = np.asarray(gt).reshape(gt.shape[0], -1)
    valid = np.ones(g.shape, bool) if ignore_index is None else (g != ignore_index)
    pred1 = lab == 1
    gt1 = g == 1
    gpos = gt1 & valid
    out = {
        "pg_tp": np.einsum("pv,gv->pg", pred1.astype(np.int64), gpos.astype(np.int64)),
        "pg_pred": np.einsum("pv,gv->pg", pred1.astype(np.int64), valid.astype(np.int64)),
        "g_sum": gpos.sum(1).astype(np.int64),
        "pp_tp": np.einsum("pv,qv->pq", pred1.astype(np.int64), pred1.astype(np.int64)),
        "pos": pred1.sum(1).astype(np.int64),
        "gg_tp": np.einsum("iv,jv->ij", gt1.astype(np.int64), gpos.astype(np.int64)),
        "gg_sum": np.einsum("iv,jv->ij", gt1.astype(np.int64), valid.astype(np.int64)),
    }
    if mean_label is not None:
        m1 = np.asarray(mean_label).reshape(-1) == 1
        maj = gt1.astype(np.float32).mean(0) >= 0.5  # ged_fast.py:121-122
        va = valid.all(0) if ignore_index is not None else np.ones(m1.shape, bool)
        out["major"] = np.array([(m1 & maj & va).sum(), (m1 & va).sum(), (maj & va).sum()], np.int64)
    return out


def ged_from_counts(c: Dict[str, np.ndarray], additional_metrics=("dice",)) -> Dict[str, float]:
    """ged_fast.py:60-140 from the counts, in the reference's float32 arithmetic."""
    f = np.float32
    tp, ps, gs = c["pg_tp"].astype(f), c["pg_pred"].astype(f), np.broadcast_to(c["g_sum"].astype(f), c["pg_tp"].shape)
    denom = f(2) * tp + (ps - tp) + (gs - tp)
    both_empty = (ps == 0) & (gs == 0)
    one_empty = (ps == 0) ^ (gs == 0)
    dice_pg = np.zeros(tp.shape, f)
    dice_pg[both_empty] = 1.0
    idx = ~(both_empty | one_empty) & (denom > 0)
    dice_pg[idx] = (f(2) * tp[idx]) / denom[idx]
    d_gp = float(np.mean(f(1) - dice_pg, dtype=f))
    pos = c["pos"].astype(f)
    den_pp = pos[:, None] + pos[None, :]
    dice_pp = np.ones(den_pp.shape, f)
    m = den_pp > 0
    dice_pp[m] = (f(2) * c["pp_tp"].astype(f)[m]) / den_pp[m]
    d_pp = float(np.mean(f(1) - dice_pp, dtype=f))
    G = c["g_sum"].shape[0]
    per_j = []
    for j in range(G):
        den = c["gg_sum"][:, j].astype(f) + f(c["g_sum"][j])
        dice_g = np.ones(G, f)
        mg = den > 0
        dice_g[mg] = (f(2) * c["gg_tp"][:, j].astype(f)[mg]) / den[mg]
        per_j.append(f(1) - np.mean(dice_g, dtype=f))
    d_gg = float(np.mean(np.array(per_j, f), dtype=f)) if per_j else 0.0
    res = {"ged": float(2 * d_gp - d_pp - d_gg)}
    if "dice" in additional_metrics:
        res["dice"] = float(np.mean(dice_pg, dtype=f))
    if "max_dice_pred" in additional_metrics:
        res["max_dice_pred"] = float(np.mean(dice_pg.max(1), dtype=f))
    if "max_dice_gt" in additional_metrics:
        res["max_dice_gt"] = float(np.mean(dice_pg.max(0), dtype=f))
    if "major_dice" in additional_metrics:
        tpm, psm, gsm = (f(x) for x in c["major"])
        if psm == 0 and gsm == 0:
            res["major_dice"] = 1.0
        elif psm == 0 or gsm == 0:
            res["major_dice"] = 0.0
        else:
            res["major_dice"] = float(f(2) * tpm / (psm + gsm))
    return res


def ged_binary_fast(output_softmax: torch.Tensor, ground_truth, ignore_index=None, additional_metrics=None) -> Dict[str, float]:
    """ged_fast.py:5-142 for (P, 2, H, W) probabilities and (G, H, W) references."""
    if additional_metrics is None:
        additional_metrics = ["dice"]
    if output_softmax.ndim != 4 or output_softmax.shape[1] != 2:
        raise ValueError("ged_binary_fast expects (P, 2, H, W) softmax input for binary segmentation")
    gt = np.asarray(ground_truth)
    if gt.ndim != 3:
        raise ValueError("ged_binary_fast expects ground_truth of shape (G, H, W)")
    x = output_softmax.detach().cpu().numpy()
    labels = np.stack([argmax_first_nan_max(m) for m in x])
    mean_label = argmax_first_nan_max(mean_members_f32(x)) if "major_dice" in additional_metrics else None
    return ged_from_counts(ged_counts(labels, gt, ignore_index, mean_label), additional_metrics)


def likelihood_sums(image_preds: np.ndarray, gt: np.ndarray, ignore_index: int, eps: float = 1e-12):
    """The reductions of test_2D.py:1043-1075: per reference g and member p the sum over valid voxels of
    log(clamp(p[member, gt, voxel], eps)) (float64 here; the reference reduces in float32) and the valid count
    (every voxel when ignore_index < 0, :1055-1060)."""
    x = np.asarray(image_preds, np.float32)
    P, C = x.shape[:2]
    x = x.reshape(P, C, -1)
    g = np.asarray(gt).reshape(gt.shape[0], -1).astype(np.int64)
    logp = np.log(np.maximum(x, np.float32(eps)))  # torch.clamp(min=eps) then log, float32
    logp = np.where(np.isnan(x), np.float32(np.nan), logp)
    sums = np.zeros((g.shape[0], P), np.float64)
    counts = np.zeros(g.shape[0], np.int64)
    for r in range(g.shape[0]):
        valid = g[r] != ignore_index if ignore_index >= 0 else np.ones(g.shape[1], bool)
        counts[r] = int(valid.sum())
        if counts[r] == 0:
            continue
        idx = np.where(valid, g[r], 0)
        picked = np.take_along_axis(logp, np.broadcast_to(idx[None, None, :], (P, 1, idx.size)), axis=1)[:, 0]
        sums[r] = (picked.astype(np.float64) * valid).sum(1)
    return sums, counts


def compute_likelihood_stats(image_preds: np.ndarray, gt: np.ndarray, ignore_index: int, eps: float = 1e-12):
    """test_2D.py:1043-1083: (gt_model_nll [G][P], gt_nll [G], mean_nll)."""
    sums, counts = likelihood_sums(image_preds, gt, ignore_index, eps)
    nll = np.where(counts[:, None] > 0, -(sums / np.maximum(counts, 1)[:, None]), 0.0)
    gt_model_nll = [[float(np.float32(v)) for v in row] for row in nll]
    gt_nll = [float(np.mean(np.array(row, np.float32), dtype=np.float32)) for row in gt_model_nll]
    flat = [v for row in gt_model_nll for v in row]
    return gt_model_nll, gt_nll, (float(np.mean(np.array(flat))) if flat else 0.0)


def compute_expected_nll(pred_samples: np.ndarray, gt: np.ndarray, ignore_index: int, eps: float = 1e-12) -> float:
    """test_2D.py:1085-1120: mean over references and samples of the per-sample NLL."""
    sums, counts = likelihood_sums(pred_samples, gt, ignore_index, eps)
    if sums.size == 0:
        return 0.0
    nll = np.where(counts[:, None] > 0, -(sums / np.maximum(counts, 1)[:, None]), 0.0).astype(np.float32)
    return float(np.mean(nll, dtype=np.float32))


# --------------------------------------------------------------------------
# ambiguity: NCC (evaluation/metrics/ncc.py:9-28, experiment_dataloader.py:283)
# --------------------------------------------------------------------------
def rater_variance_map(refs: np.ndarray) -> np.ndarray:
    """experiment_dataloader.py:283: ``np.var(reference_segs, axis=0)`` (ddof=0)."""
    return np.var(refs, axis=0)


def compute_ncc(gt_unc_map: np.ndarray, pred_unc_map: np.ndarray):
    """ncc.py:17-28: covariance sum over (n * sigma_gt * sigma_pred) with the
    sigmas taken with ddof=1; 0.0 if either sigma is exactly 0."""
    g = gt_unc_map - np.mean(gt_unc_map)
    p = pred_unc_map - np.mean(pred_unc_map)
    s_g = np.std(gt_unc_map, ddof=1)
    s_p = np.std(pred_unc_map, ddof=1)
    if s_g == 0 or s_p == 0:
        return 0.0
    # same operation order (and result dtype) as ncc.py:27
    return (1 / (np.size(gt_unc_map) * s_g * s_p)) * np.sum(np.multiply(g, p))


# --------------------------------------------------------------------------
# failure detection inputs: binary Dice (test_2D.py:873-899) and AURC
# (evaluation/metrics/aurc.py:14-67)
# --------------------------------------------------------------------------
def binary_dice_counts(label: np.ndarray, gt: np.ndarray, ignore_index: int):
    """Integer TP / |pred| / |gt| per rater on valid pixels (test_2D.py:878-886)."""
    valid = gt != ignore_index
    pred_pos = (label[np.newaxis] == 1) & valid
    gt_pos = (gt == 1) & valid
    axes = tuple(range(1, gt.ndim))
    return (pred_pos & gt_pos).sum(axis=axes), pred_pos.sum(axis=axes), gt_pos.sum(axis=axes)


def binary_dice_from_counts(tp, pred_sum, gt_sum) -> float:
    """test_2D.py:884-899 in fp32: both empty -> 1, exactly one empty -> 0,
    else 2TP / (2TP + FP + FN); mean over raters."""
    tp = np.asarray(tp, dtype=np.float32)
    ps = np.asarray(pred_sum, dtype=np.float32)
    gs = np.asarray(gt_sum, dtype=np.float32)
    denom = 2 * tp + (ps - tp) + (gs - tp)
    dice = np.zeros_like(denom)
    dice[(ps == 0) & (gs == 0)] = 1.0
    regular = (ps != 0) & (gs != 0) & (denom > 0)
    dice[regular] = (2 * tp[regular]) / denom[regular]
    return float(dice.mean())


def rc_curve_stats(risks: np.ndarray, confids: np.ndarray):
    """aurc.py:14-51: selective-risk curve; a point is emitted only where the
    sorted confidence changes (and at i == 0)."""
    risks = np.asarray(risks)
    confids = np.asarray(confids)
    assert risks.ndim == 1 and confids.ndim == 1 and len(risks) == len(confids)
    n = len(risks)
    order = np.argsort(confids)
    remaining = n
    err = sum(risks[order])
    coverages: List[float] = [remaining / n]
    sel_risks: List[float] = [err / n]
    weights: List[float] = []
    pending = 0
    for i in range(n - 1):
        remaining -= 1
        err = err - risks[order[i]]
        pending += 1
        if i == 0 or confids[order[i]] != confids[order[i - 1]]:
            coverages.append(remaining / n)
            sel_risks.append(err / (n - 1 - i))
            weights.append(pending / n)
            pending = 0
    if pending > 0:
        coverages.append(0)
        sel_risks.append(sel_risks[-1])
        weights.append(pending / n)
    return coverages, sel_risks, weights


def aurc(risks: np.ndarray, confids: np.ndarray) -> float:
    """aurc.py:54-58: trapezoid over the emitted points."""
    _, r, w = rc_curve_stats(risks, confids)
    return sum((r[i] + r[i + 1]) * 0.5 * w[i] for i in range(len(w)))


def eaurc(risks: np.ndarray, confids: np.ndarray) -> float:
    """aurc.py:61-67: AURC minus the AURC of the risk-sorted oracle."""
    n = len(risks)
    best = np.sort(risks).cumsum() / np.arange(1, n + 1)
    return aurc(risks, confids) - best.sum() / n


# --------------------------------------------------------------------------
# the reference's per-image call sequence, as timed by bench.py's CPU legs
# 
# --------------------------------------------------------------------------
def reference_pipeline_image(image_preds: torch.Tensor, gt: np.ndarray | None = None, *,
                             thresholds: Sequence[float] | None = None, patch_size: int | None = None,
                             platt: Sequence[Tuple[float, float]] | None = None,
                             ignore_value=None, ncc: bool = False) -> Dict[str, object]:
    """mean + argmax + calculate_uncertainty, then the aggregations / metric
    inputs the way the reference evaluates them, map by map on the CPU."""
    out = process_image(image_preds)
    label = out["label"].numpy()
    res: Dict[str, object] = {"label": label}
    names = ("TU", "AU", "EU") if "TU" in out else ("pred_entropy",)
    for k, name in enumerate(names):
        m = out[name].numpy()
        res[name] = m
        res[f"{name}/image"] = image_level_aggregation(m)["max_score"]
        if thresholds is not None:
            res[f"{name}/threshold"] = threshold_aggregation(m, thresholds[k])["max_score"]
        if patch_size is not None:
            res[f"{name}/patch"] = patch_level_aggregation(m, patch_size)
        if gt is not None and platt is not None:
            correct, conf = calibration_inputs(gt, label, m, platt[k][0], platt[k][1], ignore_value)
            res[f"{name}/ace"] = calc_ace(correct, conf)
            res[f"{name}/ece"] = calc_ece(correct, conf)
            res[f"{name}/hist"] = calib_histogram(correct, conf, binarize=False)
        if gt is not None and ncc:
            res[f"{name}/ncc"] = compute_ncc(rater_variance_map(gt), m)
    res["area"] = compute_area(label)
    res["border"] = compute_border(label)
    return res


# ---------------------------------------------------------------------------
# multi-class Dice inputs (uncertainty_modeling/test_2D.py:901-918 -> evaluation/metrics/dice_wrapped.py:17-104)
# ---------------------------------------------------------------------------
def class_counts(label, gt, n_classes, ignore_value=None):
    """Per rater and class the integers dice_wrapped's macro Dice is made of: tp = #(label == c & gt == c), pred =
    #(label == c), gt = #(gt == c), over the pixels the rater does not ignore (dice_wrapped.py:47,74-75 move the ignored
    pixels into a channel that DiceScore(include_background=False) drops).  label: (*S) ints, gt: (R, *S) ints.
    Returns three (R, n_classes) int64 arrays."""
    label = np.asarray(label).astype(np.int64).ravel()
    gt = np.asarray(gt).astype(np.int64).reshape(np.asarray(gt).shape[0], -1)
    R = gt.shape[0]
    tp = np.zeros((R, n_classes), np.int64)
    ps = np.zeros((R, n_classes), np.int64)
    gs = np.zeros((R, n_classes), np.int64)
    for r in range(R):
        valid = np.ones(label.shape, bool) if ignore_value is None else gt[r] != ignore_value
        valid = valid & (gt[r] >= 0) & (gt[r] < n_classes)  # (dice() raises on other values; the kernel skips them)
        lv, gv = label[valid], gt[r][valid]
        ps[r] = np.bincount(lv, minlength=n_classes)[:n_classes]
        gs[r] = np.bincount(gv, minlength=n_classes)[:n_classes]
        tp[r] = np.bincount(lv[lv == gv], minlength=n_classes)[:n_classes]
    return tp, ps, gs


def macro_dice_reference_semantics(label, gt_rater, n_classes, ignore_value=None):
    """ONE rater through dice_wrapped.dice(..., include_background=False, average="macro") with torchmetrics' DiceScore
    restated on one-hot arrays (torchmetrics >= 1.6 _dice_score_update / _dice_score_compute: per class 2 * intersection /
    (pred + target) over the channels 1..C-1, nan-mean over the classes with a non-zero denominator).  torchmetrics is not
    installed here: parity of this reduction is unpinned; it cross-checks the count-based form."""
    pred = np.asarray(label).astype(np.int64).ravel().copy()
    tgt = np.asarray(gt_rater).astype(np.int64).ravel().copy()
    ign = np.zeros(tgt.shape, bool) if ignore_value is None else tgt == ignore_value
    if ign.all():
        return 1.0
    pred[ign] = 0
    tgt[ign] = 0
    if (pred[~ign] == 0).all() and (tgt[~ign] == 0).all():
        return 1.0
    vals = []
    for c in range(1, n_classes):
        p, t = pred == c, tgt == c
        den = p.sum() + t.sum()
        if den > 0:
            vals.append(2.0 * (p & t).sum() / den)
    return float(np.mean(vals)) if vals else float("nan")


# ---------------------------------------------------------------------------
# upstream producers of the slab (uncertainty_modeling/test_2D.py:188-194, :1272-1277)
# ---------------------------------------------------------------------------
def renormalize_probabilities(probs: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """AlbumentationsTTABackend._renormalize_probabilities (test_2D.py:188-194), the reference's own torch expression: (B, C, *S) -> same."""
    normalizer = probs.sum(dim=1, keepdim=True)
    safe_normalizer = torch.clamp(normalizer, min=eps)
    renormalized = probs / safe_normalizer
    return torch.where(normalizer > eps, renormalized, probs)


def build_softmax_pred(groups, discretize: bool = False) -> torch.Tensor:
    """The tail of Tester._build_batch_predictions (test_2D.py:1272-1277): optional one-hot of every draw, then
    torch.stack(groups).mean(dim=1).  groups: list of (n_g, B, C, H, W) tensors -> (G, B, C, H, W)."""
    import torch.nn.functional as F
    if discretize:
        groups = [F.one_hot(torch.argmax(g, dim=2), num_classes=g.shape[2]).permute(0, 1, 4, 2, 3).float() for g in groups]
    return torch.stack(list(groups)).mean(dim=1)


def renormalize_probabilities_canonical(probs: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """The same in NumPy with the summation order spelled out (cascade sum over the class axis, IEEE float32 division):
    what the kernel implements.  probs: (B, C, *S) float32."""
    p = np.asarray(probs, np.float32)
    nrm = cascade_sum_f32(np.moveaxis(p, 1, 0))[:, None]
    safe = np.maximum(nrm, np.float32(eps))
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(nrm > np.float32(eps), (p / safe).astype(np.float32), p)


def softmax_logits(logits: torch.Tensor, dim: int = 1) -> torch.Tensor:
    """The step that turns network outputs into the probabilities of the slab: ``F.softmax(output, dim=1)``
    (uncertainty_modeling/test_2D.py:1181, 1185, 1225, 1241, 1256) -- torch's own CPU kernel in float32, which is what the
    reference runs.  logits: (..., C, *S) with the class axis at ``dim``."""
    import torch.nn.functional as F
    return F.softmax(logits.float(), dim=dim)
