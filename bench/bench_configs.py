"""The named pipeline of every BASELINE.json config, timed stage by stage with CUDA events on resident synthetic inputs
(developer tool; bench.py is the contract benchmark and measures configs[4]).   python bench/bench_configs.py [--configs 1,2,3,4,5]

    cfg1  toy 2-D TTA            N=10, C=2, 256x256          PE/EE/MI maps + image-level mean
    cfg2  LIDC 3-D ensemble      N=5,  C=2, 64^3, R=4        maps + patch-level aggregation (10^3 box, 3 maps) + ACE histograms
    cfg3  GTA5 HRNet TTA         N=10, C=19, 1024x2048, R=5  maps + threshold aggregation + Dice counts (AURC inputs)
    cfg4  diffusion multi-rater  N=32, C=2, 128x128, R=4     maps + NCC sums; GED counts + likelihood sums (vu_member_scores)
    cfg5  sharded sweep          N=16, C=19, 512x1024, R=1   maps + image / threshold / area / Dice / ECE-ACE histograms

One JSON line per config: stage times, algorithmic bytes, GB/s of the whole pipeline against the
measured HBM peak, sample-voxels/s.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffuncertainty_b200 as vu  # noqa: E402
from diffuncertainty_b200 import _lib, aggregation, calibration, members, synth  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep_k1 import time_call  # noqa: E402

S = _lib
CONFIGS = {
    1: dict(P=10, C=2, spatial=(256, 256), B=256, R=0, ignore=None, flags=S.STAT_IMAGE_SUM, patch=None, members=False),
    2: dict(P=5, C=2, spatial=(64, 64, 64), B=128, R=4, ignore=None,
            flags=S.STAT_IMAGE_SUM | S.STAT_AREA | S.STAT_DICE | S.STAT_CALIB, patch=(10, 10, 10), members=False),
    3: dict(P=10, C=19, spatial=(1024, 2048), B=4, R=5, ignore=255,
            flags=S.STAT_IMAGE_SUM | S.STAT_THRESHOLD | S.STAT_AREA | S.STAT_DICE, patch=None, members=False),
    4: dict(P=32, C=2, spatial=(128, 128), B=1024, R=4, ignore=None, flags=S.STAT_IMAGE_SUM | S.STAT_NCC, patch=None, members=True),
    5: dict(P=16, C=19, spatial=(512, 1024), B=16, R=1, ignore=255,
            flags=S.STAT_IMAGE_SUM | S.STAT_THRESHOLD | S.STAT_AREA | S.STAT_DICE | S.STAT_CALIB, patch=None, members=False),
}


NAMES = {1: "cfg1 toy 2-D TTA: N=10, C=2, 256x256, B=256; maps + labels + image-level mean",
         2: "cfg2 LIDC 3-D ensemble: N=5, C=2, 64^3, R=4, B=128; maps + labels + image sums / area / Dice / ACE histograms, patch-level 10^3 box on TU, AU, EU",
         3: "cfg3 GTA5 HRNet TTA: N=10, C=19, 1024x2048, R=5 with 2% ignore(255), B=4; maps + labels + image / threshold / area / Dice counts (AURC inputs)",
         4: "cfg4 diffusion multi-rater: N=32, C=2, 128x128, R=4, B=1024; maps + labels + image sums + NCC sums; GED counts + likelihood sums",
         5: "cfg5 sharded sweep: N=16, C=19, 512x1024, R=1 with 2% ignore(255), B=16; maps + labels + image / threshold / area / Dice / ECE-ACE histograms"}


def time_config(cid: int, iters: int = 10, peak: float = 6532.2, logits: bool = False) -> dict:
    """The named pipeline of BASELINE config `cid` on resident synthetic inputs, stage by stage with CUDA events."""
    platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    cfg = CONFIGS[cid]
    P, C, B, R, spatial = cfg["P"], cfg["C"], cfg["B"], cfg["R"], cfg["spatial"]
    x = synth.synth_slab(P, B, C, spatial, seed=cid, scale=3.0)
    V = x[0, 0, 0].numel()
    gt = None
    if R:
        gt = vu.GroundTruth(synth.synth_gt(x, R, seed=cid, flip=0.2, ignore_frac=0.02 if cfg["ignore"] is not None else 0.0,
                                           ignore_value=cfg["ignore"] if cfg["ignore"] is not None else 255), cfg["ignore"])
    sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
    si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
    # the three maps are views of one buffer, so that patch-level aggregation takes all of them in one call (3 B "images")
    maps3 = torch.empty((3, B) + tuple(spatial), dtype=torch.float32, device="cuda")
    maps = {k: maps3[i] for i, k in enumerate(("TU", "AU", "EU"))}
    labels = torch.empty((B,) + tuple(spatial), dtype=torch.uint8, device="cuda")
    flags = cfg["flags"]
    ms_out = members.MemberScoreBuffers(P, B, R, x.device) if cfg["members"] and hasattr(members, "MemberScoreBuffers") else None

    def fused():
        vu.fused_pass(x, gt, stats=flags, thresholds=[0.3, 0.2, 0.02], calib=platt if flags & S.STAT_CALIB else None,
                      stats_out=(sf, si), maps_out=maps, labels_out=labels)

    stages = {}
    bytes_alg = (4 * P * C + 13 + R) * V * B
    if cfg["members"] and hasattr(vu, "fused_pass_with_member_scores"):
        def fused_ms():
            vu.fused_pass_with_member_scores(x, gt, stats=flags, stats_out=(sf, si), maps_out=maps, labels_out=labels, out=ms_out)

        stages["vu_fused_pass + member scores (one slab read)"] = time_call(fused_ms, iters=iters)
        stages["(vu_fused_pass alone)"] = time_call(fused, iters=iters)
    else:
        stages["vu_fused_pass"] = time_call(fused, iters=iters)
        if cfg["members"]:
            def member_scores():
                members.member_scores(x, gt, nll=True, ged=True, mean_labels=labels)

            stages["vu_member_scores"] = time_call(member_scores, iters=iters)
    if cfg["patch"]:
        dims = [1] * (3 - len(spatial)) + list(spatial)

        def patch():
            aggregation.patch_level_batched(maps3.reshape(3 * B, *dims), cfg["patch"])

        stages["vu_patch_max_ws (TU, AU, EU in one call)"] = time_call(patch, iters=iters)
    if logits:
        # the same launch with every member replaced by its one-hot argmax (--discretize, test_2D.py:1272-1275), read in place
        grp = vu.Groups([x[p:p + 1] for p in range(P)], discretize=True)

        def fused_onehot():
            vu.fused_pass(grp, gt, stats=flags, thresholds=[0.3, 0.2, 0.02], calib=platt if flags & S.STAT_CALIB else None,
                          stats_out=(sf, si), maps_out=maps, labels_out=labels)

        stages["(vu_fused_pass, members discretised in the read)"] = time_call(fused_onehot, iters=iters)
        del grp
        # the same launch over a slab of LOGITS (vu_fused_pass_logits: softmax folded into the read); log p are logits whose
        # softmax is p again, taken in place so that no second slab is resident
        x.clamp_(min=1e-30).log_()

        def fused_logits():
            vu.fused_pass(x, gt, stats=flags, thresholds=[0.3, 0.2, 0.02], calib=platt if flags & S.STAT_CALIB else None,
                          stats_out=(sf, si), maps_out=maps, labels_out=labels, logits=True)

        stages["(vu_fused_pass_logits, same launch over logits)"] = time_call(fused_logits, iters=iters)
    total = sum(v for k, v in stages.items() if not k.startswith("("))
    first = next(iter(stages.values()))
    fused_ms_only = stages.get("(vu_fused_pass alone)", stages.get("vu_fused_pass", first))
    line = {"workload": NAMES[cid], "config": f"cfg{cid}", "stat_flags": hex(flags),
            "stages_ms": {k: round(v, 4) for k, v in stages.items()}, "ms": round(total, 4),
            "algorithmic_bytes": bytes_alg, "GBps": round(bytes_alg / total / 1e6, 1), "frac": round(bytes_alg / total / 1e6 / peak, 3),
            "fused_pass_ms": round(fused_ms_only, 4), "fused_pass_GBps": round(bytes_alg / fused_ms_only / 1e6, 1),
            "fused_pass_frac": round(bytes_alg / fused_ms_only / 1e6 / peak, 3),
            "sample_voxels_per_s": round(P * V * B / total * 1e3, 1), "peak_GBps": peak}
    oh = stages.get("(vu_fused_pass, members discretised in the read)")
    if oh:
        line["discretize_pass_ms"] = round(oh, 4)
        line["discretize_pass_frac"] = round(bytes_alg / oh / 1e6 / peak, 3)
    lg = stages.get("(vu_fused_pass_logits, same launch over logits)")
    if lg:
        line["logits_pass_ms"] = round(lg, 4)
        line["logits_pass_frac"] = round(bytes_alg / lg / 1e6 / peak, 3)
    del x, gt, maps, maps3, labels
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--logits", action="store_true", help="also time vu_fused_pass_logits on every config")
    args = ap.parse_args()
    peak = 6532.2
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    for cid in [int(c) for c in args.configs.split(",")]:
        print(json.dumps(time_config(cid, args.iters, peak, args.logits)), flush=True)


if __name__ == "__main__":
    main()
