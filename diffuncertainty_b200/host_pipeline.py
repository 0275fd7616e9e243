"""The fused pass for slabs that live in HOST memory.

The reference's evaluation side works on arrays it loaded from disk
(evaluation/uncertainty_aggregation/aggregate_uncertainties.py:140-142,
evaluation/experiment_dataloader.py:305-312), and ``calculate_uncertainty``
(unc_mod_utils/test_utils.py:833) accepts CPU tensors.  This is the equivalent
entry: the caller hands pinned host buffers, the pipeline streams them through
the device in image chunks -- host->device copy of chunk i+1, kernel on chunk i
and device->host copy of the maps of chunk i-1 overlap on three streams -- and
returns host results.  All arithmetic still happens in libvalunc's kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, calibration
from ._lib import F64, I64
from .uncertainty import UNC_KEYS, GroundTruth, fused_pass


def h2d_ceiling_gbs(host: torch.Tensor, device, reps: int = 2) -> float:
    """What the box delivers for a plain pinned host -> device copy of ``host`` (one contiguous transfer, GB/s): the ceiling
    of any end-to-end number that feeds the GPU from host memory."""
    import time
    dst = torch.empty(host.shape, dtype=host.dtype, device=device)
    dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize(device)
    return host.numel() * host.element_size() * reps / (time.perf_counter() - t0) / 1e9


def bind_host_thread_to_device_node(device) -> Optional[int]:
    """Restrict the calling process to the CPU cores of the NUMA node the GPU hangs off, so that pinned buffers allocated
    afterwards (first touch) lie in that node's memory and the host->device copies of the ranks of a multi-GPU box do not
    cross the socket interconnect.  Returns the node, or None when the topology cannot be read (then nothing changes)."""
    import os
    try:
        import ctypes
        index = torch.device(device).index
        index = torch.cuda.current_device() if index is None else index
        buf = ctypes.create_string_buffer(32)
        cudart = ctypes.CDLL("libcudart.so.12")  # already loaded by torch: resolves to the same library
        if cudart.cudaDeviceGetPCIBusId(buf, 32, int(index)) != 0:
            return None
        bus = buf.value.decode().lower()  # "0000:1b:00.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


@dataclass
class HostResult:
    maps: Dict[str, torch.Tensor]      # pinned host fp32 (B, *S)
    labels: torch.Tensor               # pinned host uint8 (B, *S)
    stats_f64: Optional[np.ndarray]    # (B, 80)
    stats_i64: Optional[np.ndarray]    # (B, 156)
    h2d_bytes: int
    d2h_bytes: int


class HostPipeline:
    def __init__(self, P: int, C: int, spatial: Sequence[int], batch: int, R: int = 0, gt_dtype=torch.uint8,
                 chunk_images: int = 1, n_buffers: int = 3, stats: int = 0, thresholds: Optional[Sequence[float]] = None,
                 platt=None, ignore_index: Optional[int] = None, device=None, logits: bool = False, dtype=torch.float32):
        _lib.require_device()
        if dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise ValueError("dtype must be float32, bfloat16 or float16")
        self.dtype = dtype  # element type of the host slab: 16-bit slabs cross PCIe and HBM at half the bytes (vu_slab.dtype)
        self.logits = bool(logits)  # the host slab holds network outputs before F.softmax (fused_pass(logits=True): opt-in)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.P, self.C, self.spatial, self.B, self.R = P, C, tuple(spatial), batch, R
        self.chunk = max(1, min(chunk_images, batch))
        self.nb = n_buffers
        self.stats, self.thresholds, self.ignore_index = stats, thresholds, ignore_index
        self.calib = [calibration.platt_edges(a, b) for a, b in platt] if (platt is not None and stats & _lib.STAT_CALIB) else None
        S = self.spatial
        # P == 1 is the single-prediction path (calculate_one_minus_msr, test_2D.py:1006-1007): one map, key "pred_entropy"
        self.map_keys = UNC_KEYS if P > 1 else ("pred_entropy",)
        self._lib = _lib.load()
        with torch.cuda.device(self.dev):
            self.d_slab = [torch.empty((P, self.chunk, C) + S, dtype=dtype, device=self.dev) for _ in range(n_buffers)]
            self.d_gt = [torch.empty((self.chunk, R) + S, dtype=gt_dtype, device=self.dev) for _ in range(n_buffers)] if R else None
            self.d_maps = [{k: torch.empty((self.chunk,) + S, dtype=torch.float32, device=self.dev) for k in self.map_keys}
                           for _ in range(n_buffers)]
            self.d_labels = [torch.empty((self.chunk,) + S, dtype=torch.uint8, device=self.dev) for _ in range(n_buffers)]
            self.rows_f = torch.zeros((batch, F64["COLS"]), dtype=torch.float64, device=self.dev)
            self.rows_i = torch.zeros((batch, I64["COLS"]), dtype=torch.int64, device=self.dev)
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.h_maps = {k: torch.empty((batch,) + S, dtype=torch.float32).pin_memory() for k in self.map_keys}
        self.h_labels = torch.empty((batch,) + S, dtype=torch.uint8).pin_memory()
        self.h_rows_f = torch.empty((batch, F64["COLS"]), dtype=torch.float64).pin_memory()
        self.h_rows_i = torch.empty((batch, I64["COLS"]), dtype=torch.int64).pin_memory()

    def run(self, x_host: torch.Tensor, gt_host: Optional[torch.Tensor] = None) -> HostResult:
        """x_host: (P, B, C, *S) host tensor of the pipeline's dtype (pinned for full speed); gt_host: (B, R, *S)."""
        P, B, nb, ch = self.P, self.B, self.nb, self.chunk
        if tuple(x_host.shape) != (P, B, self.C) + self.spatial or x_host.dtype != self.dtype or x_host.is_cuda:
            raise ValueError(f"x_host must be a {self.dtype} host tensor of shape (P, B, C, *spatial)")
        if self.R and (gt_host is None or tuple(gt_host.shape) != (B, self.R) + self.spatial):
            raise ValueError("gt_host must have shape (B, R, *spatial)")
        h2d = d2h = 0
        img_bytes = self.C * int(np.prod(self.spatial)) * x_host.element_size()
        ev_in = [torch.cuda.Event() for _ in range(nb)]
        ev_run = [torch.cuda.Event() for _ in range(nb)]
        ev_out = [None] * nb
        n_chunks = (B + ch - 1) // ch
        with torch.cuda.device(self.dev):
            if self.stats:
                with torch.cuda.stream(self.s_run):
                    self.rows_f.zero_()
                    self.rows_i.zero_()
            for i in range(n_chunks):
                s, e = i * ch, min(B, (i + 1) * ch)
                n, j = e - s, i % nb
                with torch.cuda.stream(self.s_in):
                    self.s_in.wait_event(ev_run[j]) if i >= nb else None   # buffer j is free once its kernel ran
                    # the chunk is P runs of n images, one per member, B images apart in the host slab: ONE strided copy
                    # (P separate copies before: at 1 image per chunk the copy engine saw 16 small transfers per chunk)
                    run = n * img_bytes
                    src_pitch, dst_pitch = B * img_bytes, self.chunk * img_bytes
                    if x_host.is_contiguous() and src_pitch < 2 ** 31:
                        _lib.check(self._lib.vu_copy_2d_async(self.d_slab[j].data_ptr(), dst_pitch, x_host.data_ptr() + s * img_bytes,
                                                              src_pitch, run, P, 1, self.s_in.cuda_stream), "vu_copy_2d_async")
                    else:
                        for p in range(P):                                  # each source run is contiguous
                            self.d_slab[j][p, :n].copy_(x_host[p, s:e], non_blocking=True)
                    h2d += P * run
                    if self.R:
                        self.d_gt[j][:n].copy_(gt_host[s:e], non_blocking=True)
                        h2d += gt_host[s:e].numel() * gt_host.element_size()
                    ev_in[j].record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(ev_in[j])
                    if ev_out[j] is not None:
                        self.s_run.wait_event(ev_out[j])                    # maps of the previous user of buffer j left
                    maps_out = {k: v[:n] for k, v in self.d_maps[j].items()}
                    fused_pass(self.d_slab[j][:, :n], GroundTruth(self.d_gt[j][:n], self.ignore_index) if self.R else None,
                               stats=self.stats, thresholds=self.thresholds, calib=self.calib,
                               stats_out=(self.rows_f[s:e], self.rows_i[s:e]) if self.stats else None,
                               maps_out=maps_out, labels_out=self.d_labels[j][:n], logits=self.logits)
                    ev_run[j].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_run[j])
                    for k in self.map_keys:
                        self.h_maps[k][s:e].copy_(self.d_maps[j][k][:n], non_blocking=True)
                    self.h_labels[s:e].copy_(self.d_labels[j][:n], non_blocking=True)
                    d2h += n * int(np.prod(self.spatial)) * (4 * len(self.map_keys) + 1)
                    ev_out[j] = torch.cuda.Event()
                    ev_out[j].record(self.s_out)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_stream(self.s_run)
                if self.stats:
                    self.h_rows_f.copy_(self.rows_f, non_blocking=True)
                    self.h_rows_i.copy_(self.rows_i, non_blocking=True)
                    d2h += self.rows_f.numel() * 8 + self.rows_i.numel() * 8
            self.s_out.synchronize()
        return HostResult(maps=self.h_maps, labels=self.h_labels,
                          stats_f64=self.h_rows_f.numpy() if self.stats else None,
                          stats_i64=self.h_rows_i.numpy() if self.stats else None, h2d_bytes=h2d, d2h_bytes=d2h)
