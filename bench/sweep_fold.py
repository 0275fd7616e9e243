"""Tuning sweep of the member-score fold (k1_uni with member scores) on configs[3] (developer tool).
    python bench/sweep_fold.py [--shapes 0,25620,38421,...]      shape = CT * 100 + 20 + shuffle steps, 0 = automatic"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import diffuncertainty_b200 as vu  # noqa: E402
from diffuncertainty_b200 import _lib, members, synth  # noqa: E402
from sweep_k1 import time_call  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="0")
    ap.add_argument("--stages", default="0")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    P, C, B, R, spatial = 32, 2, 1024, 4, (128, 128)
    x = synth.synth_slab(P, B, C, spatial, seed=4, scale=3.0)
    V = x[0, 0, 0].numel()
    gt = vu.GroundTruth(synth.synth_gt(x, R, seed=4, flip=0.2, ignore_frac=0.0), None)
    sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
    si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
    maps3 = torch.empty((3, B) + spatial, dtype=torch.float32, device="cuda")
    maps = {k: maps3[i] for i, k in enumerate(("TU", "AU", "EU"))}
    labels = torch.empty((B,) + spatial, dtype=torch.uint8, device="cuda")
    bufs = members.MemberScoreBuffers(P, B, R, x.device)
    ref = None
    for shape in [int(v) for v in args.shapes.split(",")]:
        for stages in [int(v) for v in args.stages.split(",")]:
            _lib.set_option("k1_uni_shape", shape)
            _lib.set_option("k1_tma_stages", stages)

            def run():
                vu.fused_pass(x, gt, stats=0x21, stats_out=(sf, si), maps_out=maps, labels_out=labels, members_out=bufs)
            try:
                bufs.zero_()
                run()
                torch.cuda.synchronize()
                cur = (bufs.ged_counts.clone(), bufs.nll_sum.clone())
                ms = time_call(run, iters=args.iters)
            except Exception as exc:
                print(f"shape={shape} stages={stages}: {str(exc)[:160]}", flush=True)
                continue
            same = ""
            if ref is None:
                ref = cur
            else:
                same = f" ged=={bool(torch.equal(ref[0], cur[0]))} nll maxrel={float(((ref[1] - cur[1]).abs() / ref[1].abs().clamp_min(1e-30)).max()):.1e}"
            bytes_alg = (4 * P * C + 13 + R) * V * B
            print(f"shape={shape:6d} stages={stages}: {ms:8.4f} ms  {bytes_alg / ms / 1e6:7.1f} GB/s ({bytes_alg / ms / 1e6 / 6532.2:.3f}){same}", flush=True)
    _lib.set_option("k1_uni_shape", 0)
    _lib.set_option("k1_tma_stages", 0)


if __name__ == "__main__":
    main()
