"""C2 uncertainty measures on the GPU, behind the reference's call surface.

``calculate_uncertainty`` / ``calculate_one_minus_msr`` are drop-ins for
uncertainty_modeling/unc_mod_utils/test_utils.py:833-864 (same argument, same
dict keys, fp32 maps on the input's device).  ``fused_pass`` is the batch form
that replaces the per-image loop of ``Tester.process_output``
(uncertainty_modeling/test_2D.py:968-1041): one launch for the whole
``softmax_pred`` of shape (P, B, C, *S), strided views accepted as they are.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import F64, I64

UNC_KEYS = ("TU", "AU", "EU")


def _spatial_strides_flat(t: torch.Tensor, first_spatial_dim: int) -> Optional[int]:
    """Element stride of the flattened spatial index, or None if the spatial
    dims cannot be addressed as one strided run (then a copy is unavoidable)."""
    shape, stride = t.shape[first_spatial_dim:], t.stride()[first_spatial_dim:]
    dims = [(n, s) for n, s in zip(shape, stride) if n != 1]
    if not dims:
        return 1
    for (n0, s0), (n1, s1) in zip(dims[:-1], dims[1:]):
        if s0 != n1 * s1:
            return None
    return dims[-1][1]


def _check_slab(x: torch.Tensor, what: str) -> None:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor, got {type(x)}")
    if not x.is_cuda:
        raise _lib.ValuncError(f"{what} must live on a CUDA device (no CPU fallback in diffuncertainty_b200)")


@dataclass
class GroundTruth:
    """batch["seg"] handed alongside the slab (test_2D.py:1122-1124): (B, R, *S)
    int64 or uint8, optional ignore value (ace.py:492-499, test_2D.py:880)."""
    seg: torch.Tensor
    ignore_index: Optional[int] = None


@dataclass
class FusedResult:
    """Outputs of one fused pass over a batch."""
    maps: Dict[str, torch.Tensor]        # "TU","AU","EU" (or "pred_entropy") -> (B, *S) fp32
    labels: Optional[torch.Tensor]       # (B, *S) uint8
    stats_f64: Optional[torch.Tensor]    # (B, 80) float64, see valunc.h
    stats_i64: Optional[torch.Tensor]    # (B, 156) int64
    n_voxels: int
    n_raters: int
    stat_flags: int
    member_labels: Optional[torch.Tensor] = None   # (P, B, *S) uint8: argmax of every member (test_2D.py:810-818)
    class_counts: Optional[torch.Tensor] = None    # (B, R, C, 3) int64: tp / pred / gt per rater and class (STAT_CLASS_COUNTS)

    # -- per-image scores, all computed on the host in float64 from the rows --
    def _rows(self):
        """Host copies of the two row blocks, read back once (the accessors below all go through here)."""
        if self.stats_f64 is None:
            raise ValueError("this pass was run without statistics (stats=0)")
        cached = self.__dict__.get("_host_rows")
        if cached is None:
            cached = (self.stats_f64.cpu().numpy(), self.stats_i64.cpu().numpy())
            self.__dict__["_host_rows"] = cached
        return cached

    def refresh(self) -> None:
        """Forget the host copies (the device rows accumulate: call this after another launch added to them)."""
        self.__dict__.pop("_host_rows", None)

    def image_level(self, mean: bool = True) -> np.ndarray:
        """(B, 3) image_level_aggregation scores (aggregate_uncertainties.py:37-39)."""
        f, _ = self._rows()
        s = f[:, F64["SUM"]:F64["SUM"] + 3]
        return s / self.n_voxels if mean else s

    def threshold_level(self, mean: bool = True) -> np.ndarray:
        """(B, 3) threshold_aggregation scores (aggregate_uncertainties.py:124-130):
        mean over the pixels >= t, or the (zero) sum when none qualifies."""
        f, i = self._rows()
        s = f[:, F64["THR_SUM"]:F64["THR_SUM"] + 3]
        n = i[:, I64["THR_COUNT"]:I64["THR_COUNT"] + 3]
        if not mean:
            return s
        return np.where(n > 0, s / np.maximum(n, 1), s)

    def area(self) -> np.ndarray:
        return self._rows()[1][:, I64["AREA"]].astype(np.float64)

    def border(self) -> np.ndarray:
        return self._rows()[1][:, I64["BORDER"]].astype(np.float64)

    def dice_counts(self):
        """(tp, pred_sum, gt_sum), each (B, R) int64 (test_2D.py:878-886)."""
        _, i = self._rows()
        R = self.n_raters
        return (i[:, I64["DICE_TP"]:I64["DICE_TP"] + R], i[:, I64["DICE_PRED"]:I64["DICE_PRED"] + R],
                i[:, I64["DICE_GT"]:I64["DICE_GT"] + R])

    def calib_histograms(self):
        """(bin_sums f64, bin_true i64, bin_total i64), each (B, 3, 21) (ace.py:352-356)."""
        f, i = self._rows()
        B = f.shape[0]
        return (f[:, F64["BIN_SUMS"]:F64["BIN_SUMS"] + 63].reshape(B, 3, 21),
                i[:, I64["BIN_TRUE"]:I64["BIN_TRUE"] + 63].reshape(B, 3, 21),
                i[:, I64["BIN_TOTAL"]:I64["BIN_TOTAL"] + 63].reshape(B, 3, 21))

    def class_count_arrays(self):
        """(tp, pred, gt), each (B, R, C) int64: the integers of the multi-class Dice of test_2D.py:901-918."""
        if self.class_counts is None:
            raise ValueError("this pass was run without STAT_CLASS_COUNTS")
        c = self.class_counts.cpu().numpy()
        return c[..., 0], c[..., 1], c[..., 2]

    def ncc_sums(self):
        f, _ = self._rows()
        return dict(g=f[:, F64["NCC_G"]], gg=f[:, F64["NCC_GG"]], u=f[:, F64["NCC_U"]:F64["NCC_U"] + 3],
                    uu=f[:, F64["NCC_UU"]:F64["NCC_UU"] + 3], gu=f[:, F64["NCC_GU"]:F64["NCC_GU"] + 3])


def _fill_gt(gt_struct: _lib.Gt, gt: Optional[GroundTruth], B: int, spatial) -> Optional[torch.Tensor]:
    if gt is None:
        gt_struct.data = None
        return None
    seg = gt.seg
    _check_slab(seg, "gt.seg")
    if seg.dim() == 1 + len(spatial):  # (B, *S) -> (B, 1, *S) as test_2D.py:1123-1124 does
        seg = seg.unsqueeze(1)
    if seg.shape[0] != B or tuple(seg.shape[2:]) != tuple(spatial):
        # ace.py:79-80 asserts the same contract
        raise AssertionError(f"gt must have shape (B, R, *spatial) = ({B}, R, {tuple(spatial)}), got {tuple(seg.shape)}")
    if seg.dtype == torch.uint8:
        gt_struct.dtype = _lib.GT_U8
    elif seg.dtype == torch.int64:
        gt_struct.dtype = _lib.GT_I64
    else:
        raise TypeError(f"gt dtype must be uint8 or int64, got {seg.dtype}")
    sv = _spatial_strides_flat(seg, 2)
    if sv is None:
        seg = seg.contiguous()
        sv = 1
    gt_struct.data = seg.data_ptr()
    gt_struct.R = seg.shape[1]
    gt_struct.stride_b, gt_struct.stride_r, gt_struct.stride_v = seg.stride(0), seg.stride(1), sv
    gt_struct.has_ignore = 0 if gt.ignore_index is None else 1
    gt_struct.ignore_index = 0 if gt.ignore_index is None else int(gt.ignore_index)
    return seg  # keep alive


def _members_view(members):
    """Sequence of P tensors (B, C, *S) with identical shape / strides / dtype / device -> (P, strides, host pointer list).
    This is the "stack without copying" form: torch.stack(groups) of test_2D.py:1277 is never materialised."""
    first = members[0]
    for m in members:
        _check_slab(m, "softmax_pred member")
        if m.shape != first.shape or m.stride() != first.stride() or m.dtype != first.dtype or m.device != first.device:
            raise ValueError("all members must share shape, strides, dtype and device")
    return first


@dataclass
class Groups:
    """``softmax_pred_groups`` of test_2D.py:1134-1136 handed over as they are: a list of G tensors ``(n_g, B, C, *S)`` -- one
    stochastic sample per group unless several generative draws share one (test_2D.py:1160) -- that ``fused_pass`` reads in
    place.  The kernel forms ``torch.stack(groups).mean(dim=1)`` (test_2D.py:1277) itself, in torch's order, optionally after
    ``_renormalize_probabilities`` (test_2D.py:188-194, the last step of the TTA inversion) and / or the ``--discretize`` one-hot
    (test_2D.py:1272-1275) of every draw: no stacked copy, no averaged copy, no one-hot copy of the slab exists."""
    groups: Sequence[torch.Tensor]
    renormalize: bool = False
    discretize: bool = False
    eps: float = 1e-12
    logits: bool = False  # the draws are network outputs before F.softmax (test_2D.py:1181-1256); see fused_pass(logits=True)


def group_members(groups: Sequence[torch.Tensor]) -> list:
    """test_2D.py:1277: ``torch.stack(groups).mean(dim=1)`` without the stack -- every group ``(n_g, B, C, *S)`` (one stochastic
    sample per group unless several generative draws share one, test_2D.py:1134-1136, 1160) becomes one member ``(B, C, *S)``
    that ``fused_pass`` / ``member_scores`` then read in place.  A single-sample group is a view, not a copy; the mean of a
    larger group is torch's own reduction on the group's device, as in the reference."""
    members = []
    for g in groups:
        if g.dim() < 4:
            raise ValueError(f"every group must be (n_g, B, C, *spatial), got shape {tuple(g.shape)}")
        members.append(g[0] if g.shape[0] == 1 else g.mean(dim=0))
    return members


_PTR_TABLES: Dict[tuple, tuple] = {}


def _pointer_table(members, dev):
    """(host array, device tensor) of the members' base pointers.  Cached by the pointers themselves: a loop that hands the
    same buffers over again (the usual case: the forward passes write into preallocated outputs) does no host -> device copy
    and no synchronisation after the first call, which also makes the call capturable in a CUDA graph."""
    ptrs = tuple(m.data_ptr() for m in members)
    key = (ptrs, dev.index)
    hit = _PTR_TABLES.get(key)
    if hit is None:
        if len(_PTR_TABLES) > 256:
            _PTR_TABLES.clear()
        host = (C.c_void_p * len(ptrs))(*ptrs)
        staging = torch.tensor(ptrs, dtype=torch.int64).pin_memory()
        table = torch.empty(len(ptrs), dtype=torch.int64, device=dev)
        table.copy_(staging, non_blocking=True)
        hit = (host, table, staging)
        _PTR_TABLES[key] = hit
    return hit[0], hit[1]


_HALF = {torch.bfloat16: _lib.SLAB_BF16, torch.float16: _lib.SLAB_F16}


def fill_slab(slab: _lib.Slab, softmax_pred, allow_half: bool = False):
    """Describe ``softmax_pred`` -- a (P, B, C, *S) tensor with any strides, or a list of P member tensors (B, C, *S)
    that are then read where they are (no torch.stack, test_2D.py:1277) -- in a vu_slab.  Returns
    (P, B, C, spatial, device, keep-alive objects).
    ``allow_half``: bfloat16 / float16 tensors are handed over as they are (vu_slab.dtype; the kernel widens the values as it
    reads them) instead of being upcast to float32 first."""
    members = None
    draws, sflags, eps = 1, 0, 1e-12
    sdtype = _lib.SLAB_F32
    if isinstance(softmax_pred, Groups):
        grp = softmax_pred
        if not grp.groups:
            raise ValueError("softmax_pred: empty group list")
        g0 = grp.groups[0]
        for g in grp.groups:
            _check_slab(g, "softmax_pred group")
            if g.dim() < 4 or g.shape != g0.shape or g.stride() != g0.stride() or g.device != g0.device:
                raise ValueError("every group must be (n_g, B, C, *spatial) with one shape / stride pattern (torch.stack needs that too)")
        draws = int(g0.shape[0])
        sflags = (_lib.SLAB_RENORMALIZE if grp.renormalize else 0) | (_lib.SLAB_DISCRETIZE if grp.discretize else 0)
        eps = float(grp.eps)
        softmax_pred = [g[d] for g in grp.groups for d in range(draws)]  # P * draws views, draw-minor
    if isinstance(softmax_pred, (list, tuple)):
        # P separate member tensors (B, C, *S): read where they are, no torch.stack
        if not softmax_pred:
            raise ValueError("softmax_pred: empty member list")
        d0 = softmax_pred[0].dtype
        if allow_half and draws == 1 and d0 in _HALF and all(m.dtype == d0 for m in softmax_pred):
            members, sdtype = list(softmax_pred), _HALF[d0]
        else:
            members = [m if m.dtype == torch.float32 else m.float() for m in softmax_pred]
        first = _members_view(members)
        if first.dim() < 3:
            raise ValueError(f"every member must be (B, C, *spatial), got shape {tuple(first.shape)}")
        if _spatial_strides_flat(first, 2) is None:
            members = [m.contiguous() for m in members]
            first = members[0]
        P, (B, Cn), spatial = len(members) // draws, first.shape[:2], tuple(first.shape[2:])
        sv, dev = _spatial_strides_flat(first, 2), first.device
        strides = (0, first.stride(0), first.stride(1))
    else:
        _check_slab(softmax_pred, "softmax_pred")
        if softmax_pred.dim() < 4:
            raise ValueError(f"softmax_pred must be (P, B, C, *spatial), got shape {tuple(softmax_pred.shape)}")
        if allow_half and softmax_pred.dtype in _HALF:
            sdtype = _HALF[softmax_pred.dtype]  # read as it is: half the HBM traffic, no upcast copy
        elif softmax_pred.dtype != torch.float32:
            # the reference computes fp32 maps whatever the input dtype (test_utils.py:836);
            # the kernels compute in fp32, so other dtypes are upcast once here
            softmax_pred = softmax_pred.float()
        P, B, Cn = softmax_pred.shape[:3]
        spatial = tuple(softmax_pred.shape[3:])
        sv = _spatial_strides_flat(softmax_pred, 3)
        if sv is None:
            softmax_pred = softmax_pred.contiguous()
            sv = 1
        dev = softmax_pred.device
        strides = (softmax_pred.stride(0), softmax_pred.stride(1), softmax_pred.stride(2))
    V = int(np.prod(spatial)) if spatial else 1
    slab.P, slab.B, slab.C, slab.V = P, B, Cn, V
    slab.stride_p, slab.stride_b, slab.stride_c = strides
    slab.stride_v = sv
    slab.draws, slab.flags, slab.renorm_eps, slab.stride_d = draws, sflags, eps, 0
    slab.dtype = sdtype
    if members is not None:
        host_ptrs, dev_ptrs = _pointer_table(members, dev)
        slab.member_ptrs = dev_ptrs.data_ptr()
        slab.member_ptrs_host = C.cast(host_ptrs, C.c_void_p)
        keep = (host_ptrs, dev_ptrs, members)
    else:
        slab.data = softmax_pred.data_ptr()
        keep = (softmax_pred,)
    return P, B, Cn, spatial, dev, keep


def fused_pass(softmax_pred, gt: Optional[GroundTruth] = None, *, stats: int = 0,
               thresholds: Optional[Sequence[float]] = None, calib=None, label_lut: Optional[torch.Tensor] = None,
               want_maps: bool = True, want_labels: bool = True,
               stats_out: Optional[tuple] = None, maps_out: Optional[Dict[str, torch.Tensor]] = None,
               labels_out: Optional[torch.Tensor] = None, platt_fit=None, want_member_labels: bool = False,
               members_out=None, class_counts_out: Optional[torch.Tensor] = None, logits: bool = False) -> FusedResult:
    """One launch over ``softmax_pred`` of shape (P, B, C, *S) (test_2D.py:1277) -- or over a list of P member tensors
    (B, C, *S), which are then read where they are (no torch.stack).

    stats      : OR of _lib.STAT_* flags
    thresholds : three floats (TU, AU, EU) for STAT_THRESHOLD
    calib      : three ``calibration.PlattEdges`` for STAT_CALIB
    stats_out  : optional (stats_f64, stats_i64) device tensors to accumulate
                 into (rows = images of this batch)
    platt_fit  : a ``calibration.PlattFitAccumulator`` for STAT_PLATT_FIT (dataset-level buffers)
    maps_out   : optional preallocated contiguous fp32 (B, *S) tensors keyed "TU","AU","EU"
                 (or "pred_entropy" when P == 1); labels_out: preallocated uint8 (B, *S)
    members_out: a ``members.MemberScoreBuffers``: the member-level scores (GED counts, likelihood sums) are computed in the
                 same pass -- one read of the slab.  Raises NotImplementedError when this launch cannot do that (see
                 ``vu_member_out`` in valunc.h); ``members.fused_pass_with_member_scores`` falls back to a second pass then.
    logits     : OPT-IN.  ``softmax_pred`` holds the network outputs BEFORE ``F.softmax(output, dim=1)``
                 (test_2D.py:1181, 1185, 1225, 1241, 1256); the kernel applies the softmax to every member / draw as it reads it
                 (vu_fused_pass_logits), so the probability slab is never written or re-read.  Relaxed contract: the device's
                 exponential differs from torch's by a few ulp -- maps keep the 1e-5 tolerance (+1e-6 absolute), labels can
                 differ where a voxel's two largest mean probabilities agree to ~1e-7 (rates in INTEGRATION.md section 5).
    """
    _lib.require_device()
    lib = _lib.load()
    a = _lib.FusedArgs()
    a.struct_size = C.sizeof(_lib.FusedArgs)
    a.stat_flags = int(stats)
    logits = bool(logits) or (isinstance(softmax_pred, Groups) and softmax_pred.logits)
    # bfloat16 / float16 slabs (autocast) are read as they are where the library can (plain slabs, 2..32 members, aligned rows)
    half_ok = not logits and members_out is None and not want_member_labels and not isinstance(softmax_pred, Groups)
    P, B, Cn, spatial, dev, ptr_keep = fill_slab(a.slab, softmax_pred, allow_half=half_ok)
    V = int(a.slab.V)
    if logits and members_out is not None:
        raise NotImplementedError("member-level scores are not available for slabs of logits")

    maps: Dict[str, torch.Tensor] = {}
    with torch.cuda.device(dev):
        if want_maps:
            names = UNC_KEYS if P > 1 else ("pred_entropy",)
            for name in names:
                if maps_out is not None:
                    m = maps_out[name]
                    if m.shape != (B,) + spatial or m.dtype != torch.float32 or not m.is_contiguous() or m.device != dev:
                        raise ValueError(f"maps_out[{name!r}] must be a contiguous float32 {(B,) + spatial} tensor on {dev}")
                    maps[name] = m
                else:
                    maps[name] = torch.empty((B,) + spatial, dtype=torch.float32, device=dev)
            a.tu = maps[names[0]].data_ptr()
            if P > 1:
                a.au, a.eu = maps["AU"].data_ptr(), maps["EU"].data_ptr()
        labels = None
        if want_labels:
            if labels_out is not None:
                if labels_out.shape != (B,) + spatial or labels_out.dtype != torch.uint8 or not labels_out.is_contiguous() \
                        or labels_out.device != dev:
                    raise ValueError(f"labels_out must be a contiguous uint8 {(B,) + spatial} tensor on {dev}")
                labels = labels_out
            else:
                labels = torch.empty((B,) + spatial, dtype=torch.uint8, device=dev)
        a.labels = labels.data_ptr() if labels is not None else None
        member_labels = None
        if want_member_labels:
            member_labels = torch.empty((P, B) + spatial, dtype=torch.uint8, device=dev)
            a.member_labels = member_labels.data_ptr()
        keep = _fill_gt(a.gt, gt, B, spatial)
        if thresholds is not None:
            for k in range(3):
                a.threshold[k] = float(thresholds[k])
        elif stats & _lib.STAT_THRESHOLD:
            # aggregate_uncertainties.py:112-115 raises a bare Exception in this situation
            raise Exception("A threshold needs to be provided for threshold aggregation!")
        if calib is not None:
            for k in range(min(3, len(calib))):
                a.calib[k] = calib[k].as_struct()
        elif stats & _lib.STAT_CALIB:
            raise ValueError("STAT_CALIB needs `calib` (three PlattEdges)")
        if label_lut is not None:
            if label_lut.dtype != torch.uint8 or label_lut.numel() != 256 or not label_lut.is_cuda:
                raise ValueError("label_lut must be a CUDA uint8 tensor with 256 entries")
            a.calib_label_lut = label_lut.data_ptr()
        if stats & _lib.STAT_PLATT_FIT:
            if platt_fit is None:
                raise ValueError("STAT_PLATT_FIT needs `platt_fit` (a calibration.PlattFitAccumulator)")
            a.platt_fit = C.pointer(platt_fit.edges)
            a.platt_i64, a.platt_f64 = platt_fit.counts.data_ptr(), platt_fit.sums.data_ptr()
        sf = si = None
        if stats:
            if stats_out is not None:
                sf, si = stats_out
                if sf.shape != (B, F64["COLS"]) or si.shape != (B, I64["COLS"]) or sf.dtype != torch.float64 \
                        or si.dtype != torch.int64 or not sf.is_contiguous() or not si.is_contiguous():
                    raise ValueError("stats_out must be contiguous (B, 80) float64 and (B, 156) int64 tensors")
            else:
                sf = torch.zeros((B, F64["COLS"]), dtype=torch.float64, device=dev)
                si = torch.zeros((B, I64["COLS"]), dtype=torch.int64, device=dev)
            a.stats_f64, a.stats_i64 = sf.data_ptr(), si.data_ptr()
        cls = None
        if stats & _lib.STAT_CLASS_COUNTS:
            if gt is None:
                raise ValueError("STAT_CLASS_COUNTS needs ground truth")
            shape = (B, int(a.gt.R), Cn, 3)
            cls = class_counts_out if class_counts_out is not None else torch.zeros(shape, dtype=torch.int64, device=dev)
            if tuple(cls.shape) != shape or cls.dtype != torch.int64 or not cls.is_contiguous() or cls.device != dev:
                raise ValueError(f"class_counts_out must be a contiguous int64 {shape} tensor on {dev}")
            a.class_counts = cls.data_ptr()
        if members_out is not None:
            members_out.fill(a.members, P, B, int(a.gt.R) if gt is not None else 0, dev)
            if not lib.vu_fused_members_supported(C.byref(a)):
                raise NotImplementedError("vu_fused_pass cannot compute the member-level scores in this launch: "
                                          + lib.vu_last_error().decode(errors="replace"))
        if logits:
            _lib.check(lib.vu_fused_pass_logits(C.byref(a), _lib.current_stream_ptr()), "vu_fused_pass_logits")
        else:
            rc = lib.vu_fused_pass(C.byref(a), _lib.current_stream_ptr())
            if rc == -2 and a.slab.dtype != _lib.SLAB_F32:
                # this 16-bit slab has no native form (unaligned rows, P > 32, ...): upcast it once, as for any other dtype
                P, B, Cn, spatial, dev, ptr_keep = fill_slab(a.slab, softmax_pred, allow_half=False)
                rc = lib.vu_fused_pass(C.byref(a), _lib.current_stream_ptr())
            _lib.check(rc, "vu_fused_pass")
    del keep, ptr_keep
    return FusedResult(maps=maps, labels=labels, stats_f64=sf, stats_i64=si, n_voxels=V,
                       n_raters=int(a.gt.R) if gt is not None else 0, stat_flags=int(stats), member_labels=member_labels,
                       class_counts=cls)


def calculate_uncertainty(softmax_preds: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Drop-in for test_utils.py:833-859: (P, C, *S) -> {"TU","AU","EU"} of shape S."""
    _check_slab(softmax_preds, "softmax_preds")
    if softmax_preds.dim() < 2:
        raise ValueError("softmax_preds must be (P, C, *spatial)")
    if softmax_preds.shape[0] == 1:
        # the reference only reaches this function with P > 1 (test_2D.py:1004-1007); its
        # formulas give TU = AU = H(p), EU = 0 for one member, which is what two identical
        # members produce: present the member twice through a stride-0 view (no copy)
        softmax_preds = softmax_preds.expand(2, *softmax_preds.shape[1:])
    res = fused_pass(softmax_preds.unsqueeze(1), want_labels=False)
    return {k: res.maps[k][0] for k in UNC_KEYS}


def calculate_uncertainty_from_logits(logits: torch.Tensor) -> Dict[str, torch.Tensor]:
    """``calculate_uncertainty(F.softmax(logits, dim=1))`` (test_2D.py:1181 -> test_utils.py:833-859) in one pass over the
    logits (P, C, *S); relaxed contract, see ``fused_pass(logits=True)``."""
    _check_slab(logits, "logits")
    if logits.dim() < 2:
        raise ValueError("logits must be (P, C, *spatial)")
    if logits.shape[0] == 1:
        logits = logits.expand(2, *logits.shape[1:])
    res = fused_pass(logits.unsqueeze(1), want_labels=False, logits=True)
    return {k: res.maps[k][0] for k in UNC_KEYS}


def calculate_one_minus_msr(softmax_pred: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Drop-in for test_utils.py:862-864: (C, *S) -> {"pred_entropy": 1 - max_c p}."""
    _check_slab(softmax_pred, "softmax_pred")
    res = fused_pass(softmax_pred.unsqueeze(0).unsqueeze(0), want_labels=False)
    return {"pred_entropy": res.maps["pred_entropy"][0]}


def mean_argmax_labels(softmax_pred: torch.Tensor) -> torch.Tensor:
    """argmax over classes of the member mean for a whole batch (test_2D.py:971, 871,
    815-818): (P, B, C, *S) -> (B, *S) uint8."""
    return fused_pass(softmax_pred, want_maps=False).labels


def map_stats(maps: Dict[str, torch.Tensor], labels: Optional[torch.Tensor] = None, gt: Optional[GroundTruth] = None, *,
              stats: int, thresholds: Optional[Sequence[float]] = None, calib=None, label_lut: Optional[torch.Tensor] = None,
              ncc_gt_map: Optional[torch.Tensor] = None, stats_out: Optional[tuple] = None, n_classes: Optional[int] = None,
              class_counts_out: Optional[torch.Tensor] = None) -> FusedResult:
    """The statistics of ``fused_pass`` on maps / labels that already exist in device memory -- what the reference's
    file-based evaluation works on (evaluation/eval_experiments.py:348-355).  ``maps``: {"TU","AU","EU"} -> (B, *S) contiguous
    fp32 CUDA tensors (a missing key skips that type); ``labels``: (B, *S) uint8; the other arguments as in ``fused_pass``."""
    _lib.require_device()
    lib = _lib.load()
    present = [maps.get(k) for k in UNC_KEYS]
    first = next((m for m in present if m is not None), None)
    if first is None:
        raise ValueError("map_stats needs at least one of the maps TU, AU, EU")
    _check_slab(first, "maps")
    B, spatial, dev = first.shape[0], tuple(first.shape[1:]), first.device
    V = int(np.prod(spatial)) if spatial else 1
    a = _lib.MapStatsArgs()
    a.struct_size = C.sizeof(_lib.MapStatsArgs)
    a.stat_flags = int(stats)
    a.B, a.V = B, V
    for k, m in enumerate(present):
        if m is None:
            continue
        if m.shape != first.shape or m.dtype != torch.float32 or not m.is_contiguous() or m.device != dev:
            raise ValueError("maps must be contiguous float32 CUDA tensors of one shape")
        a.maps[k] = m.data_ptr()
    if labels is not None:
        if labels.shape != first.shape or labels.dtype != torch.uint8 or not labels.is_contiguous() or labels.device != dev:
            raise ValueError("labels must be a contiguous uint8 tensor of the maps' shape")
        a.labels = labels.data_ptr()
    with torch.cuda.device(dev):
        keep = _fill_gt(a.gt, gt, B, spatial)
        if thresholds is not None:
            for k in range(3):
                a.threshold[k] = float(thresholds[k])
        elif stats & _lib.STAT_THRESHOLD:
            raise Exception("A threshold needs to be provided for threshold aggregation!")
        if calib is not None:
            for k in range(min(3, len(calib))):
                a.calib[k] = calib[k].as_struct()
        elif stats & _lib.STAT_CALIB:
            raise ValueError("STAT_CALIB needs `calib` (three PlattEdges)")
        if label_lut is not None:
            a.calib_label_lut = label_lut.data_ptr()
        if ncc_gt_map is not None:
            if ncc_gt_map.shape != first.shape or ncc_gt_map.dtype != torch.float64 or not ncc_gt_map.is_contiguous():
                raise ValueError("ncc_gt_map must be a contiguous float64 tensor of the maps' shape")
            a.ncc_gt_map = ncc_gt_map.data_ptr()
        if stats_out is not None:
            sf, si = stats_out
        else:
            sf = torch.zeros((B, F64["COLS"]), dtype=torch.float64, device=dev)
            si = torch.zeros((B, I64["COLS"]), dtype=torch.int64, device=dev)
        a.stats_f64, a.stats_i64 = sf.data_ptr(), si.data_ptr()
        cls = None
        if stats & _lib.STAT_CLASS_COUNTS:
            if gt is None or n_classes is None:
                raise ValueError("STAT_CLASS_COUNTS needs ground truth and n_classes")
            shape = (B, int(a.gt.R), int(n_classes), 3)
            cls = class_counts_out if class_counts_out is not None else torch.zeros(shape, dtype=torch.int64, device=dev)
            if tuple(cls.shape) != shape or cls.dtype != torch.int64 or not cls.is_contiguous():
                raise ValueError(f"class_counts_out must be a contiguous int64 {shape} tensor")
            a.class_counts, a.n_classes = cls.data_ptr(), int(n_classes)
        _lib.check(lib.vu_map_stats(C.byref(a), _lib.current_stream_ptr()), "vu_map_stats")
    del keep
    return FusedResult(maps={k: m for k, m in zip(UNC_KEYS, present) if m is not None}, labels=labels, stats_f64=sf, stats_i64=si,
                       n_voxels=V, n_raters=int(a.gt.R) if gt is not None else 0, stat_flags=int(stats), class_counts=cls)
