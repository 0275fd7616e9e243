"""Warp-split sweep of the TMA-pipelined K1 on the statistics-heavy few-class configs (developer tool).
    python bench/sweep_split.py [--configs 2,4] [--variants 0,12,14,...]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffuncertainty_b200 as vu  # noqa: E402
from diffuncertainty_b200 import _lib, calibration, synth  # noqa: E402
from sweep_k1 import time_call  # noqa: E402

CONFIGS = {
    2: dict(P=5, C=2, spatial=(64, 64, 64), B=128, R=4, ignore=None, flags=0x1d),
    4: dict(P=32, C=2, spatial=(128, 128), B=1024, R=4, ignore=None, flags=0x21),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2,4")
    ap.add_argument("--variants", default="-1")
    ap.add_argument("--stages", default="0")
    ap.add_argument("--shapes", default="0", help="k1_uni_shape values (CT * 100 + G * 10 + MINB), 0 = automatic")
    ap.add_argument("--flags", default=None)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    for cid in [int(c) for c in args.configs.split(",")]:
        cfg = CONFIGS[cid]
        P, C, B, R, spatial = cfg["P"], cfg["C"], cfg["B"], cfg["R"], cfg["spatial"]
        x = synth.synth_slab(P, B, C, spatial, seed=cid, scale=3.0)
        V = x[0, 0, 0].numel()
        gt = vu.GroundTruth(synth.synth_gt(x, R, seed=cid, flip=0.2, ignore_frac=0.0), cfg["ignore"])
        sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
        si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
        maps = {k: torch.empty((B,) + tuple(spatial), dtype=torch.float32, device="cuda") for k in ("TU", "AU", "EU")}
        labels = torch.empty((B,) + tuple(spatial), dtype=torch.uint8, device="cuda")
        for flags in ([cfg["flags"]] if args.flags is None else [int(f, 0) for f in args.flags.split(",")]):
            ref = None
            for var in [int(v) for v in args.variants.split(",")]:
                for stages, shape in [(int(s), int(h)) for s in args.stages.split(",") for h in (args.shapes.split(",") if var == -3 else ["0"])]:
                    _lib.set_option("k1_uni_shape", shape)
                    _lib.set_option("k1_path", 1 if var == -1 else (0 if var == -3 else 2))  # -1 registers, -3 automatic, >= 0 that TMA variant
                    _lib.set_option("k1_tma_variant", var if var >= 0 else -1)
                    _lib.set_option("k1_tma_stages", stages)

                    def run():
                        vu.fused_pass(x, gt if flags & 0x78 else None, stats=flags, thresholds=[0.3, 0.2, 0.02],
                                      calib=platt if flags & _lib.STAT_CALIB else None, stats_out=(sf, si) if flags else None,
                                      maps_out=maps, labels_out=labels)
                    try:
                        sf.zero_(); si.zero_()
                        run()
                        torch.cuda.synchronize()
                        cur = (si.clone(), sf.clone())
                        ms = time_call(run, iters=args.iters)
                    except Exception as exc:
                        print(f"cfg{cid} flags={flags:#04x} var={var} stages={stages} shape={shape}: {str(exc)[:150]}", flush=True)
                        continue
                    same = ""
                    if ref is None:
                        ref = cur
                    else:
                        same = " ints==" + str(bool(torch.equal(ref[0], cur[0]))) + \
                               f" fmaxrel={float(((ref[1] - cur[1]).abs() / ref[1].abs().clamp_min(1e-30)).max()):.1e}"
                    bpv = 4 * P * C + 13 + (R if flags & 0x78 else 0)
                    print(f"cfg{cid} flags={flags:#04x} var={var:3d} stages={stages} shape={shape:5d}: {ms:8.4f} ms  {bpv * V * B / ms / 1e6:7.1f} GB/s "
                          f"({bpv * V * B / ms / 1e6 / 6532.2:.3f}){same}", flush=True)
        _lib.set_option("k1_path", 0)
        _lib.set_option("k1_tma_variant", -1)
        _lib.set_option("k1_tma_stages", 0)
        _lib.set_option("k1_uni_shape", 0)
        del x, gt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
