"""Member-level scores on the slab the uncertainty pass reads:

    ged_binary_fast(output_softmax, ground_truth, ignore_index, additional_metrics)   evaluation/metrics/ged_fast.py:5-142
    compute_likelihood_stats(image_preds, gt_tensor, ignore_index, eps)               test_2D.py:1043-1083
    compute_expected_nll(pred_samples, gt_tensor, ignore_index, eps)                  test_2D.py:1085-1120
    member_scores(softmax_pred, gt, ...)        the batch form: one vu_member_scores launch for (P, B, C, *S)

The kernel (csrc/k5_members.cu) produces the integer pair counts and the float64 log-likelihood sums; the Dice / GED
arithmetic on P*G + P*P + G*G numbers and the means over P*G numbers run here, in the reference's float32.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .uncertainty import GroundTruth, _fill_gt, fill_slab, fused_pass


class MemberScores:
    """Result of one vu_member_scores launch.  The buffers stay on the device; ``nll_sum`` (B, R, P) float64 -- the sum over
    valid voxels of ln(max(p[member, gt], eps)) --, ``nll_count`` (B, R) int64 and ``ged_counts`` (B, vu_ged_cols(P, R)) int64
    are copied to the host on first use."""

    def __init__(self, P: int, R: int, nll_sum=None, nll_count=None, ged_counts=None, has_major: bool = False):
        self.P, self.R, self.has_major = P, R, has_major
        self.device_buffers = {"nll_sum": nll_sum, "nll_count": nll_count, "ged_counts": ged_counts}
        self._host = {}

    def _get(self, name: str) -> Optional[np.ndarray]:
        if name not in self._host:
            t = self.device_buffers[name]
            self._host[name] = None if t is None else t.cpu().numpy()
        return self._host[name]

    @property
    def nll_sum(self) -> Optional[np.ndarray]:
        return self._get("nll_sum")

    @property
    def nll_count(self) -> Optional[np.ndarray]:
        return self._get("nll_count")

    @property
    def ged_counts(self) -> Optional[np.ndarray]:
        return self._get("ged_counts")

    def ged_parts(self, b: int) -> Dict[str, np.ndarray]:
        """The count matrices of image b, by the names of valunc.h."""
        P, G = self.P, self.R
        row = self.ged_counts[b]
        o = 0
        out = {}
        for name, shape in (("pg_tp", (P, G)), ("pg_pred", (P, G)), ("g_sum", (G,)), ("pp_tp", (P, P)), ("pos", (P,)),
                            ("gg_tp", (G, G)), ("gg_sum", (G, G)), ("major", (3,))):
            n = int(np.prod(shape))
            out[name] = row[o:o + n].reshape(shape)
            o += n
        return out

    def ged(self, b: int, additional_metrics=("dice",)) -> Dict[str, float]:
        if "major_dice" in additional_metrics and not self.has_major:
            raise ValueError("major_dice needs the labels of the member mean (pass mean_labels)")
        return ged_from_counts(self.ged_parts(b), additional_metrics)

    def likelihood_stats(self, b: int):
        """(gt_model_nll [G][P], gt_nll [G], mean_nll) of image b (test_2D.py:1062-1083)."""
        n = self.nll_count[b]
        nll = np.where(n[:, None] > 0, -(self.nll_sum[b] / np.maximum(n, 1)[:, None]), 0.0)
        gt_model_nll = [[float(np.float32(v)) for v in row] for row in nll]
        gt_nll = [float(np.mean(np.array(row, np.float32), dtype=np.float32)) for row in gt_model_nll]
        flat = [v for row in gt_model_nll for v in row]
        return gt_model_nll, gt_nll, (float(np.mean(np.array(flat))) if flat else 0.0)

    def expected_nll(self, b: int) -> float:
        n = self.nll_count[b]
        nll = np.where(n[:, None] > 0, -(self.nll_sum[b] / np.maximum(n, 1)[:, None]), 0.0).astype(np.float32)
        return float(np.mean(nll, dtype=np.float32)) if nll.size else 0.0


class MemberScoreBuffers:
    """Device outputs of the member-level scores for a batch of B images: ``nll_sum`` (B, R, P) float64, ``nll_count`` (B, R),
    ``nll_bad`` (B), ``ged_counts`` (B, vu_ged_cols(P, R)) int64.  They ACCUMULATE like the statistics rows (zero them with
    ``zero_()`` to start over); ``scores()`` wraps them in a ``MemberScores``."""

    def __init__(self, P: int, B: int, R: int, device, nll: bool = True, ged: bool = True, eps: float = 1e-12):
        lib = _lib.load()
        self.P, self.B, self.R, self.eps = P, B, R, float(eps)
        self.flags = (_lib.MS_NLL if nll else 0) | (_lib.MS_GED if ged else 0)
        self.nll_sum = torch.zeros((B, R, P), dtype=torch.float64, device=device) if nll else None
        self.nll_count = torch.zeros((B, R), dtype=torch.int64, device=device) if nll else None
        self.nll_bad = torch.zeros((B,), dtype=torch.int64, device=device) if nll else None
        self.ged_counts = torch.zeros((B, int(lib.vu_ged_cols(P, R))), dtype=torch.int64, device=device) if ged else None

    def zero_(self) -> None:
        for t in (self.nll_sum, self.nll_count, self.nll_bad, self.ged_counts):
            if t is not None:
                t.zero_()

    def fill(self, m: "_lib.MemberOut", P: int, B: int, R: int, device) -> None:
        if (P, B, R) != (self.P, self.B, self.R) or (self.nll_sum if self.nll_sum is not None else self.ged_counts).device != device:
            raise ValueError(f"MemberScoreBuffers were made for P, B, R = {(self.P, self.B, self.R)}, this launch has {(P, B, R)}")
        m.flags, m.eps = self.flags, self.eps
        if self.nll_sum is not None:
            m.nll_sum, m.nll_count, m.nll_bad = self.nll_sum.data_ptr(), self.nll_count.data_ptr(), self.nll_bad.data_ptr()
        if self.ged_counts is not None:
            m.ged_counts = self.ged_counts.data_ptr()

    def scores(self, check: bool = True) -> MemberScores:
        if check and self.nll_bad is not None and int(self.nll_bad.sum()) != 0:
            # torch.gather raises "index ... is out of bounds" in the reference (test_2D.py:1067)
            raise RuntimeError("ground truth holds values that are neither a class index nor the ignore value")
        return MemberScores(P=self.P, R=self.R, nll_sum=self.nll_sum, nll_count=self.nll_count, ged_counts=self.ged_counts,
                            has_major=self.ged_counts is not None)


def fused_pass_with_member_scores(softmax_pred, gt: GroundTruth, *, nll: bool = True, ged: bool = True, eps: float = 1e-12,
                                  out: Optional[MemberScoreBuffers] = None, **fused_kwargs):
    """``fused_pass`` (maps, labels, statistics) AND the member-level scores of ``member_scores`` for the same batch: what
    ``Tester.process_output`` computes per image (test_2D.py:968-1120) in one read of the slab where the kernel supports it
    (binary slabs, at most 32 members and 4 uint8 references; see ``vu_member_out`` in valunc.h), in two passes otherwise.
    Returns (FusedResult, MemberScores); ``fused_kwargs`` are those of ``fused_pass`` (``stats`` must not be 0 for the
    single-pass form)."""
    if gt is None:
        raise ValueError("member scores need ground truth")
    probe = _lib.Slab()
    P, B, Cn, spatial, dev, _keep = fill_slab(probe, softmax_pred)
    seg = gt.seg if gt.seg.dim() > 1 + len(spatial) else gt.seg.unsqueeze(1)
    R = int(seg.shape[1])
    can_ged = ged and Cn == 2
    if ged and not can_ged:
        raise ValueError("ged_binary_fast expects (P, 2, H, W) softmax input for binary segmentation")  # ged_fast.py:33-34
    bufs = out if out is not None else MemberScoreBuffers(P, B, R, dev, nll=nll, ged=ged, eps=eps)
    try:
        res = fused_pass(softmax_pred, gt, members_out=bufs, **fused_kwargs)
        return res, bufs.scores()
    except NotImplementedError:
        pass
    res = fused_pass(softmax_pred, gt, **fused_kwargs)
    ms = member_scores(softmax_pred, gt, nll=nll, ged=ged, mean_labels=res.labels if ged else None, eps=eps)
    return res, ms


def ged_from_counts(c: Dict[str, np.ndarray], additional_metrics=("dice",)) -> Dict[str, float]:
    """ged_fast.py:60-140 on the integer counts, in float32 like the reference (counts above 2^24 round the same way)."""
    f = np.float32
    tp, ps = c["pg_tp"].astype(f), c["pg_pred"].astype(f)
    gs = np.broadcast_to(c["g_sum"].astype(f), tp.shape)
    denom = f(2) * tp + (ps - tp) + (gs - tp)                        # 2TP + FP + FN (:64)
    both_empty = (ps == 0) & (gs == 0)
    one_empty = (ps == 0) ^ (gs == 0)
    dice_pg = np.zeros(tp.shape, f)
    dice_pg[both_empty] = 1.0                                         # :68-69
    regular = ~(both_empty | one_empty) & (denom > 0)
    dice_pg[regular] = (f(2) * tp[regular]) / denom[regular]          # :74
    dist_gt_pred = float(np.mean(f(1) - dice_pg, dtype=f))            # :77
    pos = c["pos"].astype(f)
    denom_pp = pos[:, None] + pos[None, :]
    dice_pp = np.ones(denom_pp.shape, f)
    m = denom_pp > 0
    dice_pp[m] = (f(2) * c["pp_tp"].astype(f)[m]) / denom_pp[m]      # :88-91
    dist_pred_pred = float(np.mean(f(1) - dice_pp, dtype=f))
    G = c["g_sum"].shape[0]
    per_rater = []
    for j in range(G):                                                # :95-106
        denom_g = c["gg_sum"][:, j].astype(f) + f(c["g_sum"][j])
        dice_g = np.ones(G, f)
        mg = denom_g > 0
        dice_g[mg] = (f(2) * c["gg_tp"][:, j].astype(f)[mg]) / denom_g[mg]
        per_rater.append(f(1) - np.mean(dice_g, dtype=f))
    dist_gt_gt = float(np.mean(np.array(per_rater, f), dtype=f)) if per_rater else 0.0
    results = {"ged": float(2 * dist_gt_pred - dist_pred_pred - dist_gt_gt)}
    if "dice" in additional_metrics:
        results["dice"] = float(np.mean(dice_pg, dtype=f))
    if "max_dice_pred" in additional_metrics:
        results["max_dice_pred"] = float(np.mean(dice_pg.max(axis=1), dtype=f))
    if "max_dice_gt" in additional_metrics:
        results["max_dice_gt"] = float(np.mean(dice_pg.max(axis=0), dtype=f))
    if "major_dice" in additional_metrics:                            # :117-140
        tp_m, pred_m, gt_m = (f(v) for v in c["major"])
        if pred_m == 0 and gt_m == 0:
            results["major_dice"] = 1.0
        elif pred_m == 0 or gt_m == 0:
            results["major_dice"] = 0.0
        else:
            results["major_dice"] = float(f(2) * tp_m / (pred_m + gt_m))
    return results


def member_scores(softmax_pred, gt: GroundTruth, *, nll: bool = True, ged: bool = False, mean_labels: Optional[torch.Tensor] = None,
                  eps: float = 1e-12) -> MemberScores:
    """One launch over ``softmax_pred`` (P, B, C, *S) -- or a list of P member tensors -- and ``gt`` (B, R, *S).
    ``gt.ignore_index`` None means "every voxel counts" (test_2D.py:1058-1060 for ignore_index < 0).
    ``mean_labels``: (B, *S) uint8 labels of the member mean (``fused_pass(...).labels``) for the majority Dice."""
    if gt is None:
        raise ValueError("member scores need ground truth")
    if not (nll or ged):
        raise ValueError("nothing to compute")
    _lib.require_device()
    lib = _lib.load()
    a = _lib.MemberScoresArgs()
    a.struct_size = C.sizeof(_lib.MemberScoresArgs)
    a.flags = (_lib.MS_NLL if nll else 0) | (_lib.MS_GED if ged else 0)
    P, B, Cn, spatial, dev, keep_slab = fill_slab(a.slab, softmax_pred)
    if ged and Cn != 2:
        raise ValueError("ged_binary_fast expects (P, 2, H, W) softmax input for binary segmentation")  # ged_fast.py:33-34
    if ged and P > 32:
        raise NotImplementedError("GED counts are built for up to 32 members")
    with torch.cuda.device(dev):
        keep_gt = _fill_gt(a.gt, gt, B, spatial)
        R = int(a.gt.R)
        a.eps = float(eps)
        nll_sum = nll_cnt = nll_bad = ged_counts = None
        if nll:
            nll_sum = torch.zeros((B, R, P), dtype=torch.float64, device=dev)
            nll_cnt = torch.zeros((B, R), dtype=torch.int64, device=dev)
            nll_bad = torch.zeros((B,), dtype=torch.int64, device=dev)
            a.nll_sum, a.nll_count, a.nll_bad = nll_sum.data_ptr(), nll_cnt.data_ptr(), nll_bad.data_ptr()
        if ged:
            cols = int(lib.vu_ged_cols(P, R))
            ged_counts = torch.zeros((B, cols), dtype=torch.int64, device=dev)
            a.ged_counts = ged_counts.data_ptr()
            if mean_labels is not None:
                if mean_labels.dtype != torch.uint8 or tuple(mean_labels.shape) != (B,) + tuple(spatial) or mean_labels.device != dev:
                    raise ValueError(f"mean_labels must be a uint8 {(B,) + tuple(spatial)} tensor on {dev}")
                mean_labels = mean_labels.contiguous()
                a.labels = mean_labels.data_ptr()
        _lib.check(lib.vu_member_scores(C.byref(a), _lib.current_stream_ptr()), "vu_member_scores")
        if nll and int(nll_bad.sum()) != 0:
            # torch.gather raises "index ... is out of bounds" in the reference (test_2D.py:1067)
            raise RuntimeError("ground truth holds values that are neither a class index nor the ignore value")
    del keep_slab, keep_gt
    return MemberScores(P=P, R=R, nll_sum=nll_sum, nll_count=nll_cnt, ged_counts=ged_counts, has_major=mean_labels is not None and ged)


def _as_gt(ground_truth, n_spatial: int, device, ignore_index) -> GroundTruth:
    g = ground_truth if isinstance(ground_truth, torch.Tensor) else torch.as_tensor(np.asarray(ground_truth))
    if g.dim() == n_spatial:
        g = g.unsqueeze(0)  # test_2D.py:1044-1045: a single (H, W) reference
    g = g.to(device)
    if g.dtype != torch.uint8:
        g = g.long()
    return GroundTruth(g.unsqueeze(0), ignore_index)


def ged_binary_fast(output_softmax: torch.Tensor, ground_truth, ignore_index: Optional[int] = None,
                    additional_metrics: Optional[List[str]] = None) -> Dict[str, float]:
    """Drop-in for ged_fast.py:5-142: (P, 2, H, W) probabilities on the GPU, (G, H, W) references."""
    if additional_metrics is None:
        additional_metrics = ["dice"]
    if not isinstance(output_softmax, torch.Tensor) or output_softmax.ndim != 4 or output_softmax.shape[1] != 2:
        raise ValueError("ged_binary_fast expects (P, 2, H, W) softmax input for binary segmentation")
    g = ground_truth if isinstance(ground_truth, torch.Tensor) else torch.as_tensor(np.asarray(ground_truth))
    if g.ndim != 3:
        raise ValueError("ged_binary_fast expects ground_truth of shape (G, H, W)")
    slab = output_softmax.unsqueeze(1)
    gt = _as_gt(g, 2, output_softmax.device, ignore_index)
    labels = fused_pass(slab, want_maps=False).labels if "major_dice" in additional_metrics else None  # ged_fast.py:119
    return member_scores(slab, gt, nll=False, ged=True, mean_labels=labels).ged(0, additional_metrics)


def compute_likelihood_stats(image_preds: torch.Tensor, gt_tensor, ignore_index: int = -1, eps: float = 1e-12):
    """Drop-in for Tester._compute_likelihood_stats (test_2D.py:1043-1083): (P, C, H, W) probabilities, (G, H, W) or
    (H, W) references; ``ignore_index`` < 0 counts every voxel."""
    gt = _as_gt(gt_tensor, image_preds.dim() - 2, image_preds.device, ignore_index if ignore_index >= 0 else None)
    return member_scores(image_preds.unsqueeze(1), gt, nll=True, eps=eps).likelihood_stats(0)


def compute_expected_nll(pred_samples: torch.Tensor, gt_tensor, ignore_index: int = -1, eps: float = 1e-12) -> float:
    """Drop-in for Tester._compute_expected_nll (test_2D.py:1085-1120)."""
    gt = _as_gt(gt_tensor, pred_samples.dim() - 2, pred_samples.device, ignore_index if ignore_index >= 0 else None)
    return member_scores(pred_samples.unsqueeze(1), gt, nll=True, eps=eps).expected_nll(0)
