"""Dataset sweep: the hot path over many images, sharded by image across the GPUs
of one box (one process per GPU), with one exchange step at the end.

Replaces, for a whole test set, the two loops of the reference -- the per-image
loop of ``Tester.process_output`` (uncertainty_modeling/test_2D.py:968-1041) and
the per-uncertainty / per-image loops of the evaluation tasks
(evaluation/uncertainty_aggregation/aggregate_uncertainties.py:133-188,
evaluation/metrics/ace.py:463-534, evaluation/metrics/aurc.py:130-153) -- by one
fused launch per batch that accumulates per-image rows on the device.  Images are
independent, so ranks own contiguous blocks of image indices and no data-path
collective is needed; only the per-image rows (histogram partials, scores, Dice
counts: the ECE / ACE and AURC inputs) are combined, with ONE int64 all-reduce
(NCCL on GPUs; the same code runs over gloo on CPU tensors in the tests) that
acts as an exact all-gather (``Partials``), so the combined rows and the
dataset-level histograms are bit-identical at any GPU count.

The finalisation (ACE / ECE, AURC / E-AURC) is image-count sized and stays on
the host in float64, mirroring ace.py:357-375,439-460 and aurc.py:14-67.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, aurc as _aurc, calibration
from ._lib import F64, I64

UNC = ("TU", "AU", "EU")


@dataclass
class SweepConfig:
    P: int
    C: int
    spatial: Tuple[int, ...]
    n_images: int
    batch: int                      # images per launch
    R: int = 1                      # reference segmentations per image
    ignore_index: Optional[int] = 255
    thresholds: Sequence[float] = (0.3, 0.2, 0.02)
    platt: Sequence[Tuple[float, float]] = ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))
    stats: int = (_lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CALIB)
    # synthetic source: images keyed by (seed, image index)
    seed: int = 0
    scale: float = 3.0
    flip: float = 0.2
    ignore_frac: float = 0.02
    keep_maps: bool = False         # keep TU/AU/EU + labels of every local image on the device

    @property
    def V(self) -> int:
        return int(np.prod(self.spatial))


def shard_bounds(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of image indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# ---------------------------------------------------------------------------
# exchange step
# ---------------------------------------------------------------------------
class Partials:
    """Everything that crosses GPUs, in ONE int64 buffer:

        [ n_images x 156 int64 rows | n_images x 80 float64 rows, viewed as int64 ]

    ``vu_fused_pass`` accumulates straight into the rows of the images a rank owns (``local``); every other row stays zero,
    so a SUM all-reduce over the int64 view is an exact all-gather -- each element has one non-zero contributor, and adding
    zeros to the bit pattern of a float64 leaves it untouched.  One latency-bound ``ncclInt64`` all-reduce per exchange, no
    packing kernels, and the combined rows are bit-identical at any GPU count.  The dataset-level histograms are the column
    sums of the gathered rows and are formed on the host (``result``)."""

    def __init__(self, n_images: int, device, class_shape: Optional[Tuple[int, int]] = None):
        """``class_shape`` = (R, C): the buffer also carries the per-rater, per-class tp / pred / gt counts of
        ``STAT_CLASS_COUNTS`` (n_images x R x C x 3 int64) behind the rows."""
        self.n_images = n_images
        n_i, n_f = n_images * I64["COLS"], n_images * F64["COLS"]
        n_c = n_images * class_shape[0] * class_shape[1] * 3 if class_shape else 0
        self.buf = torch.zeros(n_i + n_f + n_c, dtype=torch.int64, device=device)
        self.rows_i = self.buf[:n_i].view(n_images, I64["COLS"])
        self.rows_f = self.buf[n_i:n_i + n_f].view(torch.float64).view(n_images, F64["COLS"])
        self.rows_c = self.buf[n_i + n_f:].view(n_images, class_shape[0], class_shape[1], 3) if class_shape else None

    def local(self, lo: int, hi: int):
        """(stats_f64, stats_i64) of the images lo .. hi - 1: the ``stats_out`` argument of ``fused_pass``."""
        return self.rows_f[lo:hi], self.rows_i[lo:hi]

    def zero_(self) -> None:
        self.buf.zero_()

    def exchange(self, async_op: bool = False):
        """The only collective of the sweep.  Returns the work handle when ``async_op`` (else None)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    def result(self, n_voxels: int, n_raters: int) -> "SweepResult":
        i = self.rows_i.cpu().numpy().copy()
        f = self.rows_f.cpu().numpy().copy()
        return SweepResult(n_images=self.n_images, n_voxels=n_voxels, n_raters=n_raters,
                           bin_total=i[:, I64["BIN_TOTAL"]:I64["BIN_TOTAL"] + 63].sum(0).reshape(3, 21),
                           bin_true=i[:, I64["BIN_TRUE"]:I64["BIN_TRUE"] + 63].sum(0).reshape(3, 21),
                           bin_sums=f[:, F64["BIN_SUMS"]:F64["BIN_SUMS"] + 63].sum(0).reshape(3, 21), rows_f64=f, rows_i64=i,
                           class_counts=None if self.rows_c is None else self.rows_c.cpu().numpy().copy())


def pack_partials(rows_f64: torch.Tensor, rows_i64: torch.Tensor, lo: int, n_images: int) -> Partials:
    """A ``Partials`` buffer holding this rank's rows (local image order) at their global positions, zeros elsewhere.
    (The sweep itself lets the kernel write into ``Partials.local``; this is for rows that already exist.)"""
    p = Partials(n_images, rows_f64.device)
    n_local = rows_f64.shape[0]
    if n_local:
        p.rows_i[lo:lo + n_local] = rows_i64
        p.rows_f[lo:lo + n_local] = rows_f64
    return p


def exchange(partials: Partials) -> None:
    partials.exchange()


# ---------------------------------------------------------------------------
# host finalisation
# ---------------------------------------------------------------------------
@dataclass
class SweepResult:
    n_images: int
    n_voxels: int
    n_raters: int
    bin_total: np.ndarray            # (3, 21) int64, dataset level
    bin_true: np.ndarray             # (3, 21) int64
    bin_sums: np.ndarray             # (3, 21) float64
    rows_f64: np.ndarray             # (n_images, 80)
    rows_i64: np.ndarray             # (n_images, 156)
    maps: Dict[str, torch.Tensor] = field(default_factory=dict)   # local images only (keep_maps)
    labels: Optional[torch.Tensor] = None
    class_counts: Optional[np.ndarray] = None   # (n_images, R, C, 3) int64 tp / pred / gt (STAT_CLASS_COUNTS)

    def image_level(self, mean: bool = True) -> np.ndarray:
        s = self.rows_f64[:, F64["SUM"]:F64["SUM"] + 3]
        return s / self.n_voxels if mean else s

    def threshold_level(self, mean: bool = True) -> np.ndarray:
        s = self.rows_f64[:, F64["THR_SUM"]:F64["THR_SUM"] + 3]
        n = self.rows_i64[:, I64["THR_COUNT"]:I64["THR_COUNT"] + 3]
        return np.where(n > 0, s / np.maximum(n, 1), s) if mean else s

    def dice(self) -> np.ndarray:
        """Per-image Dice, mean over raters: the multi-class macro Dice of test_2D.py:901-918 when the sweep carried the class
        counts of a slab with more than two classes, else the binary Dice of test_2D.py:873-899."""
        if self.class_counts is not None and self.class_counts.shape[2] > 2:
            c = self.class_counts
            return _aurc.multiclass_dice_from_counts(c[..., 0], c[..., 1], c[..., 2])
        R = self.n_raters
        i = self.rows_i64
        return _aurc.binary_dice_from_counts(i[:, I64["DICE_TP"]:I64["DICE_TP"] + R], i[:, I64["DICE_PRED"]:I64["DICE_PRED"] + R],
                                             i[:, I64["DICE_GT"]:I64["DICE_GT"] + R])

    def calibration(self) -> Dict[str, Dict[str, float]]:
        """calibration.json-like summary (ace.py:523-531): mean of the per-image ACE / ECE
        and the dataset-level gace / gece, per uncertainty type."""
        out = {}
        B = self.n_images
        bs = self.rows_f64[:, F64["BIN_SUMS"]:F64["BIN_SUMS"] + 63].reshape(B, 3, 21)
        bt = self.rows_i64[:, I64["BIN_TRUE"]:I64["BIN_TRUE"] + 63].reshape(B, 3, 21)
        bn = self.rows_i64[:, I64["BIN_TOTAL"]:I64["BIN_TOTAL"] + 63].reshape(B, 3, 21)
        for k, name in enumerate(UNC):
            per = np.array([calibration.per_image_ace_ece(bs[b, k], bt[b, k], bn[b, k]) for b in range(B)])
            gace, gece = calibration.ace_ece_from_histogram(self.bin_sums[k], self.bin_true[k], self.bin_total[k])
            out[name] = {"ace": float(np.mean(per[:, 0])), "ece": float(np.mean(per[:, 1])), "gace": gace, "gece": gece}
        return out

    def failure_detection(self) -> Dict[str, Dict[str, float]]:
        """failure_detection.json-like summary (aurc.py:130-153): risk = 1 - Dice,
        confidence = -score, per uncertainty type and aggregation."""
        risks = 1.0 - self.dice().astype(np.float64)
        out = {}
        for agg, scores in (("image_level", self.image_level()), ("threshold", self.threshold_level())):
            for k, name in enumerate(UNC):
                out[f"{name}/{agg}"] = {"aurc": _aurc.aurc(risks, -scores[:, k]), "eaurc": _aurc.eaurc(risks, -scores[:, k])}
        return out


def unpack_result(partials: Partials, n_voxels: int, n_raters: int) -> SweepResult:
    return partials.result(n_voxels, n_raters)


# ---------------------------------------------------------------------------
# the sweep
# ---------------------------------------------------------------------------
Source = Callable[[int, int], Tuple[torch.Tensor, Optional[torch.Tensor]]]


class ShardedSweep:
    """One rank's part of a dataset sweep.

    source(first_image, n) -> (softmax_pred (P, n, C, *S) CUDA fp32, gt (n, R, *S) uint8/int64 or None);
    the default source is the on-device synthetic generator, so that any GPU
    count sees the same images."""

    def __init__(self, cfg: SweepConfig, rank: int = 0, world: int = 1, device=None, source: Optional[Source] = None):
        self.cfg, self.rank, self.world = cfg, rank, world
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lo, self.hi = shard_bounds(cfg.n_images, rank, world)
        self.source = source or self._synthetic
        self._slab = None
        self._calib = [calibration.platt_edges(a, b) for a, b in cfg.platt] if cfg.stats & _lib.STAT_CALIB else None

    def _synthetic(self, first: int, n: int):
        from . import synth
        cfg = self.cfg
        if self._slab is None or self._slab.shape[1] != n:
            self._slab = torch.empty((cfg.P, n, cfg.C) + tuple(cfg.spatial), dtype=torch.float32, device=self.device)
        x = synth.synth_slab(cfg.P, n, cfg.C, cfg.spatial, seed=cfg.seed, first_image=first, scale=cfg.scale, out=self._slab)
        gt = None
        if cfg.stats & (_lib.STAT_DICE | _lib.STAT_CALIB | _lib.STAT_NCC | _lib.STAT_CLASS_COUNTS):
            gt = synth.synth_gt(x, cfg.R, seed=cfg.seed, first_image=first, flip=cfg.flip, ignore_frac=cfg.ignore_frac,
                                ignore_value=255 if cfg.ignore_index is None else cfg.ignore_index)
        return x, gt

    def run(self) -> SweepResult:
        from .uncertainty import GroundTruth, fused_pass
        cfg = self.cfg
        n_local = self.hi - self.lo
        want_cls = bool(cfg.stats & _lib.STAT_CLASS_COUNTS)
        partials = Partials(cfg.n_images, self.device, (cfg.R, cfg.C) if want_cls else None)
        rows_f, rows_i = partials.local(self.lo, self.hi)  # the kernel accumulates straight into the exchange buffer
        kept = {k: [] for k in UNC}
        kept_labels = []
        for s in range(0, n_local, cfg.batch):
            n = min(cfg.batch, n_local - s)
            x, gt = self.source(self.lo + s, n)
            res = fused_pass(x, None if gt is None else GroundTruth(gt, cfg.ignore_index), stats=cfg.stats,
                             thresholds=cfg.thresholds, calib=self._calib, want_maps=cfg.keep_maps, want_labels=cfg.keep_maps,
                             stats_out=(rows_f[s:s + n], rows_i[s:s + n]),
                             class_counts_out=partials.rows_c[self.lo + s:self.lo + s + n] if want_cls else None)
            if cfg.keep_maps:
                for k in UNC:
                    kept[k].append(res.maps[k])
                kept_labels.append(res.labels)
        partials.exchange()  # once, at the end of the sweep
        out = partials.result(cfg.V, cfg.R)
        if cfg.keep_maps and kept_labels:
            out.maps = {k: torch.cat(v) for k, v in kept.items()}
            out.labels = torch.cat(kept_labels)
        return out
