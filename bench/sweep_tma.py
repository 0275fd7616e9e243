"""Timing of the TMA-pipelined K1 variants (developer tool).
    python bench/sweep_tma.py [--configs 1,2,3,4,5] [--stages 0,2,3,4] [--flags 0x0,0x1f,0x3f]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffuncertainty_b200 as vu  # noqa: E402
from diffuncertainty_b200 import _lib, calibration, synth  # noqa: E402
from sweep_k1 import CONFIGS, time_call  # noqa: E402

# (C, VEC, LEVELS, CT, G, NCH) -- keep in sync with kTma in csrc/k1_tma.cu
TMA = {0: (2, 4, 1, 512, 2, 1), 1: (2, 4, 2, 512, 2, 1), 2: (19, 2, 1, 512, 1, 2), 3: (19, 2, 2, 512, 1, 2),
       4: (2, 4, 1, 256, 4, 1), 5: (2, 4, 2, 256, 4, 1), 6: (19, 2, 1, 256, 1, 1), 7: (19, 2, 2, 256, 1, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--stages", default="0")
    ap.add_argument("--flags", default="0x0,0x3f")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    for cid in [int(c) for c in args.configs.split(",")]:
        cfg = CONFIGS[cid]
        P, C, B, R, spatial = cfg["P"], cfg["C"], cfg["B"], cfg["R"], cfg["spatial"]
        x = synth.synth_slab(P, B, C, spatial, seed=cid, scale=3.0)
        V = x[0, 0, 0].numel()
        gt = vu.GroundTruth(synth.synth_gt(x, max(R, 1), seed=cid, flip=0.2, ignore_frac=0.02), 255)
        sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
        si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
        maps = {k: torch.empty((B,) + tuple(spatial), dtype=torch.float32, device="cuda") for k in ("TU", "AU", "EU")}
        labels = torch.empty((B,) + tuple(spatial), dtype=torch.uint8, device="cuda")
        levels = 1 if P <= 17 else 2
        for flags in [int(f, 0) for f in args.flags.split(",")]:
            if not R:
                flags &= 0x07
            for path, variants in ((1, [-1]), (2, [i for i, d in TMA.items() if d[0] == C and d[2] == levels])):
                for i in variants:
                    for stages in [int(s) for s in args.stages.split(",")] if path == 2 else [0]:
                        _lib.set_option("k1_path", path)
                        _lib.set_option("k1_tma_variant", i)
                        _lib.set_option("k1_tma_stages", stages)

                        def run():
                            vu.fused_pass(x, gt if flags & 0x38 else None, stats=flags, thresholds=[0.3, 0.2, 0.02],
                                          calib=platt if flags & _lib.STAT_CALIB else None, stats_out=(sf, si) if flags else None,
                                          maps_out=maps, labels_out=labels)
                        try:
                            ms = time_call(run, iters=args.iters)
                        except Exception as exc:
                            print(f"cfg{cid} flags={flags:#04x} path={path} var={i} stages={stages}: {exc}")
                            continue
                        bpv = 4 * P * C + 13 + (max(R, 1) if flags & 0x38 else 0)
                        print(f"cfg{cid} flags={flags:#04x} {'tma ' + str(TMA[i]) if path == 2 else 'regs      '} stages={stages}: "
                              f"{ms:8.3f} ms  {bpv * V * B / ms / 1e6:7.1f} GB/s  {P * V * B / ms / 1e6:7.1f} Gsv/s", flush=True)
        _lib.set_option("k1_path", 0)
        _lib.set_option("k1_tma_variant", -1)
        _lib.set_option("k1_tma_stages", 0)
        del x, gt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
