"""Tuning sweep over the K1 fast-kernel variants (developer tool, not the bench
contract).  Prints achieved GB/s (algorithmic bytes / CUDA-event time) per
variant for the five BASELINE shapes.

    python bench/sweep_k1.py [--stats] [--configs 1,2,3,4,5] [--out gpurun_out/sweep.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffuncertainty_b200 as vu  # noqa: E402
from diffuncertainty_b200 import _lib, calibration, synth  # noqa: E402

CONFIGS = {
    1: dict(P=10, C=2, spatial=(256, 256), B=256, R=0),
    2: dict(P=5, C=2, spatial=(64, 64, 64), B=128, R=4),
    3: dict(P=10, C=19, spatial=(1024, 2048), B=2, R=5),
    4: dict(P=32, C=2, spatial=(128, 128), B=512, R=4),
    5: dict(P=16, C=19, spatial=(512, 1024), B=4, R=1),
}


def time_call(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--flags", type=lambda v: int(v, 0), default=None, help="explicit VU_STAT_* mask (implies --stats)")
    ap.add_argument("--variants", default=None, help="comma-separated variant indices to time")
    args = ap.parse_args()
    n_var = _lib.get_counter("k1_num_variants")
    variants = [[_lib.get_counter(f"k1_variant.{i}.{f}") for f in range(7)] for i in range(n_var)]
    results = []
    platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    for cid in [int(c) for c in args.configs.split(",")]:
        cfg = CONFIGS[cid]
        P, C, B, R, spatial = cfg["P"], cfg["C"], cfg["B"], cfg["R"], cfg["spatial"]
        x = synth.synth_slab(P, B, C, spatial, seed=cid, scale=3.0)
        V = x[0, 0, 0].numel()
        gt = None
        flags = 0
        if args.flags is not None:
            args.stats = True
        if args.stats:
            flags = _lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA
            if R:
                gt = vu.GroundTruth(synth.synth_gt(x, R, seed=cid, flip=0.2, ignore_frac=0.02), 255)
                flags |= _lib.STAT_DICE | _lib.STAT_CALIB | _lib.STAT_NCC
            if args.flags is not None:
                flags = args.flags
                if not R and flags & (_lib.STAT_DICE | _lib.STAT_CALIB | _lib.STAT_NCC):
                    gt = vu.GroundTruth(synth.synth_gt(x, 1, seed=cid, flip=0.2, ignore_frac=0.02), 255)
        bytes_per_voxel = 4 * P * C + 13 + (R if args.stats else 0)
        total_bytes = bytes_per_voxel * V * B
        levels = 1 if P <= 17 else 2
        sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
        si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
        # preallocate outputs once: time the kernel, not the allocator
        for i, d in enumerate(variants):
            if d[0] != C or d[2] != levels:
                continue
            if args.variants is not None and str(i) not in args.variants.split(","):
                continue
            _lib.set_option("k1_variant", i)

            def run():
                vu.fused_pass(x, gt, stats=flags, thresholds=[0.3, 0.2, 0.02], calib=platt if flags & _lib.STAT_CALIB else None,
                              stats_out=(sf, si) if flags else None)
            try:
                ms = time_call(run, iters=args.iters)
            except Exception as exc:  # variant not applicable
                print(f"cfg{cid} variant {i} {d}: {exc}")
                continue
            gbs = total_bytes / ms / 1e6
            svps = P * V * B / ms / 1e6
            row = dict(cfg=cid, variant=i, desc=dict(zip(("C", "VEC", "LEVELS", "THREADS", "MINB", "G", "DB"), d)),
                       ms=ms, gbs=gbs, gsv_per_s=svps, stats=bool(args.stats))
            results.append(row)
            print(f"cfg{cid} flags={flags:#04x} var {i:2d} VEC={d[1]} T={d[3]} MINB={d[4]} G={d[5]} DB={d[6]}: "
                  f"{ms:8.3f} ms  {gbs:7.1f} GB/s  {svps:7.1f} Gsv/s", flush=True)
        _lib.set_option("k1_variant", -1)
        del x, gt
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
