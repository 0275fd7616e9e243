"""One fused pass on a BASELINE shape, for ncu (developer tool).
    python bench/prof_one.py <cfg> <flags> [variant] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffuncertainty_b200 as vu  # noqa: E402
from diffuncertainty_b200 import _lib, calibration, synth  # noqa: E402
from sweep_k1 import CONFIGS  # noqa: E402

cid, flags = int(sys.argv[1]), int(sys.argv[2], 0)
variant = int(sys.argv[3]) if len(sys.argv) > 3 else -1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = CONFIGS[cid]
P, C, B, R, spatial = cfg["P"], cfg["C"], cfg["B"], cfg["R"], cfg["spatial"]
x = synth.synth_slab(P, B, C, spatial, seed=cid, scale=3.0)
gt = None
if flags & (_lib.STAT_DICE | _lib.STAT_CALIB | _lib.STAT_NCC):
    gt = vu.GroundTruth(synth.synth_gt(x, max(R, 1), seed=cid, flip=0.2, ignore_frac=0.02), 255)
platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
_lib.set_option("k1_variant", variant)
sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
for _ in range(iters):
    vu.fused_pass(x, gt, stats=flags, thresholds=[0.3, 0.2, 0.02], calib=platt if flags & _lib.STAT_CALIB else None,
                  stats_out=(sf, si) if flags else None)
torch.cuda.synchronize()
print("ok", cid, hex(flags), variant)
