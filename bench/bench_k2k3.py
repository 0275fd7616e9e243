"""CUDA-event timing of K2 (patch-level aggregation, border) and K3 (statistics on stored maps) on the BASELINE shapes
(developer tool).   python bench/bench_k2k3.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffuncertainty_b200 as vu  # noqa: E402
from diffuncertainty_b200 import _lib, aggregation, calibration, synth  # noqa: E402
from sweep_k1 import time_call  # noqa: E402
import ctypes as C  # noqa: E402

SHAPES = {"cfg1 256x256": ((1, 256, 256), 256, 0), "cfg2 64^3": ((64, 64, 64), 128, 4), "cfg3 1024x2048": ((1, 1024, 2048), 8, 5),
          "cfg4 128x128": ((1, 128, 128), 512, 4), "cfg5 512x1024": ((1, 512, 1024), 16, 1)}


def main():
    lib = _lib.load()
    platt = [calibration.platt_edges(a, b).as_struct() for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    for name, (dims, B, R) in SHAPES.items():
        V = dims[0] * dims[1] * dims[2]
        maps = [torch.rand((B, V), device="cuda") ** 3 * 0.69 for _ in range(3)]
        labels = (torch.rand((B, V), device="cuda") < 0.3).to(torch.uint8)
        gt = (torch.rand((B, max(R, 1), V), device="cuda") < 0.3).to(torch.uint8)
        box = (10, 10, 10) if dims[0] > 1 else (1, 10, 10)
        out_max = torch.empty(B, dtype=torch.float64, device="cuda")
        out_first = torch.empty(B, dtype=torch.int64, device="cuda")
        si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
        sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
        st = _lib.current_stream_ptr()

        def patch():
            _lib.check(lib.vu_patch_max(maps[0].data_ptr(), B, dims[0], dims[1], dims[2], box[0], box[1], box[2], 0,
                                        out_max.data_ptr(), out_first.data_ptr(), st), "patch")

        ws_bytes = int(lib.vu_patch_workspace_bytes(B, dims[0], dims[1], dims[2], box[0], box[1], box[2]))
        ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.int64, device="cuda")

        def patch_ws():
            _lib.check(lib.vu_patch_max_ws(maps[0].data_ptr(), B, dims[0], dims[1], dims[2], box[0], box[1], box[2], 0,
                                           out_max.data_ptr(), out_first.data_ptr(), ws.data_ptr(), ws_bytes, st), "patch_ws")

        def border():
            _lib.check(lib.vu_border_count(labels.data_ptr(), B, dims[0], dims[1], dims[2], si.data_ptr(), st), "border")

        a = _lib.MapStatsArgs()
        a.struct_size = C.sizeof(_lib.MapStatsArgs)
        a.stat_flags = 0x1f if R else 0x07
        a.B, a.V = B, V
        for k in range(3):
            a.maps[k] = maps[k].data_ptr()
            a.calib[k] = platt[k]
            a.threshold[k] = 0.2
        a.labels = labels.data_ptr()
        if R:
            a.gt.data, a.gt.dtype, a.gt.R = gt.data_ptr(), _lib.GT_U8, R
            a.gt.stride_b, a.gt.stride_r, a.gt.stride_v = R * V, V, 1
        a.stats_f64, a.stats_i64 = sf.data_ptr(), si.data_ptr()

        def mapstats():
            _lib.check(lib.vu_map_stats(C.byref(a), st), "map_stats")

        def mapstats_general():
            _lib.set_option("stats_path", 1)
            mapstats()
            _lib.set_option("stats_path", 0)

        for label, fn, nbytes in (("K2 patch_max (one map, 10^d box)", patch, 4 * V * B),
                                  ("K2 patch_max_ws (with workspace)", patch_ws, 4 * V * B), ("K2 border", border, V * B),
                                  (f"K3 map_stats flags={a.stat_flags:#x} (lean form)", mapstats, (13 + R) * V * B),
                                  (f"K3 map_stats flags={a.stat_flags:#x} (general form)", mapstats_general, (13 + R) * V * B)):
            ms = time_call(fn, iters=10)
            print(f"{name:16s} B={B:4d} {label:44s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s  {V * B / ms / 1e6:8.1f} Gvox/s", flush=True)
        # the two forms must agree: integers exactly, float sums to rounding
        sf.zero_(); si.zero_(); mapstats(); torch.cuda.synchronize()
        f2, i2 = sf.clone(), si.clone()
        sf.zero_(); si.zero_(); mapstats_general(); torch.cuda.synchronize()
        rel = float(((f2 - sf).abs() / sf.abs().clamp_min(1e-300)).max())
        print(f"{name:16s} lean vs general form: integers equal = {bool(torch.equal(i2, si))}, max rel diff of the float64 sums = {rel:.2e}", flush=True)


def bench_k4():
    """K4: exact order statistics (radix select with the descent on the device) and what sits on them -- wall-clock per call
    (host work and the read-back included: these entry points return Python floats)."""
    import time
    import numpy as np
    from diffuncertainty_b200 import quantile
    torch.manual_seed(0)
    for name, shape, R in (("cfg5 512x1024", (512, 1024), 1), ("cfg2 64^3", (64, 64, 64), 4), ("cfg3 1024x2048", (1024, 2048), 5)):
        V = int(np.prod(shape))
        u = (torch.rand(shape, device="cuda") ** 3 * 0.69).contiguous()
        pred = (torch.rand(shape, device="cuda") < 0.3).to(torch.uint8)
        refs = (torch.rand((R,) + shape, device="cuda") < 0.3).to(torch.uint8)

        def wall(fn, iters=10):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / iters * 1e3

        sel = quantile.RadixSelect([u])
        qs = np.linspace(0.0, 1.0, 21)
        t_sel = wall(lambda: sel.select_quantile_stats(qs))
        t_old = wall(lambda: quantile.RadixSelect([u]).select(np.arange(0, V, V // 40)))
        t_q = wall(lambda: quantile.quantile([u], 0.9))
        t_eq = wall(lambda: calibration.eqace_from_maps(refs, pred, u, 3.5, -1.25))
        # device time of the three histogram passes alone
        lib = _lib.load()
        hist = torch.zeros((64, 2048), dtype=torch.int64, device="cuda")
        st = _lib.current_stream_ptr()

        def level0():
            _lib.check(lib.vu_radix_hist_state(u.data_ptr(), V, None, 0, None, hist.data_ptr(), st), "hist0")

        ms0 = time_call(level0, iters=10)
        print(f"{name:16s} K4 vu_radix_hist level 0 (device)              {ms0:8.3f} ms  {4 * V / ms0 / 1e6:8.1f} GB/s", flush=True)
        print(f"{name:16s} K4 21-quantile selection, descent on device  {t_sel:8.3f} ms wall (3 passes + 3 walks + 1 read-back)", flush=True)
        print(f"{name:16s} K4 same ranks, descent on the host (r01)      {t_old:8.3f} ms wall (3 read-backs)", flush=True)
        print(f"{name:16s} K4 quantile(map, 0.9)                          {t_q:8.3f} ms wall", flush=True)
        print(f"{name:16s} K4 eqace_from_maps (R={R}), one image            {t_eq:8.3f} ms wall", flush=True)
        # the batched form: B images x 3 uncertainty types in one launch sequence, two read-backs for all of them
        Bb = 16 if V <= (1 << 20) else 4
        ub = [(torch.rand((Bb,) + shape, device="cuda") ** (k + 2) * 0.69).contiguous() for k in range(3)]
        predb = (torch.rand((Bb,) + shape, device="cuda") < 0.3).to(torch.uint8)
        refsb = (torch.rand((Bb, R) + shape, device="cuda") < 0.3).to(torch.uint8)
        platt3 = [(3.5, -1.25), (6.0, -2.0), (-4.0, 0.5)]
        t_b = wall(lambda: calibration.eqace_from_maps_batch(refsb, predb, ub, platt3), iters=5)
        print(f"{name:16s} K4 eqace_from_maps_batch, {Bb} images x 3 types   {t_b:8.3f} ms wall = {t_b / (3 * Bb):6.3f} ms per image and type",
              flush=True)


def bench_generic():
    """Class counts without a compiled-in fast / TMA form (anything but C = 2, 3, 4, 19) run on the class-outer kernel
    (k1_co_tma / k1_classouter: run-time C, P <= 32; r02u: the generic kernel they used to fall to reached 0.5 TB/s), unaligned
    rows on the one-voxel-per-thread register form: throughput of the fused pass without statistics."""
    peak = 6532.2
    for name, P, C, shape, B in (("C=5  N=10 512x512", 10, 5, (512, 512), 8), ("C=7  N=10 512x512", 10, 7, (512, 512), 8),
                                 ("C=21 N=10 512x512", 10, 21, (512, 512), 4), ("C=19 N=10 511x513 (unaligned rows)", 10, 19, (511, 513), 4)):
        x = synth.synth_slab(P, B, C, shape, seed=3, scale=3.0)
        V = x[0, 0, 0].numel()
        maps = {k: torch.empty((B,) + shape, dtype=torch.float32, device="cuda") for k in ("TU", "AU", "EU")}
        labels = torch.empty((B,) + shape, dtype=torch.uint8, device="cuda")
        ms = time_call(lambda: vu.fused_pass(x, maps_out=maps, labels_out=labels), iters=5)
        nbytes = (4 * P * C + 13) * V * B
        print(f"other class counts {name:36s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s = {nbytes / ms / 1e6 / peak:5.3f} of the HBM peak", flush=True)
        if C == 21:  # ... and with reference-based statistics in the same pass (general statistics form)
            gt = vu.GroundTruth(synth.synth_gt(x, 1, seed=3, flip=0.2, ignore_frac=0.02, ignore_value=255), 255)
            sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
            si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
            platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
            for fl in (0x0f, 0x1f):
                ms = time_call(lambda: vu.fused_pass(x, gt, stats=fl, thresholds=[0.3, 0.2, 0.02], calib=platt if fl & 0x10 else None,
                                                     stats_out=(sf, si), maps_out=maps, labels_out=labels), iters=5)
                nb2 = (4 * P * C + 14) * V * B
                print(f"other class counts {name + f' + statistics {fl:#x}':36s} {ms:8.3f} ms  {nb2 / ms / 1e6:8.1f} GB/s = {nb2 / ms / 1e6 / peak:5.3f} of the HBM peak",
                      flush=True)
        del x


def bench_half():
    """A cfg5-shaped slab in bfloat16 (autocast output) read as it is (vu_slab.dtype, class-outer TMA form) against what the
    wrapper did before: upcast copy + the float32 pass."""
    peak = 6532.2
    P, C, shape, B = 16, 19, (512, 1024), 8
    x32 = synth.synth_slab(P, B, C, shape, seed=5, scale=3.0)
    x16 = x32.to(torch.bfloat16)
    V = shape[0] * shape[1]
    maps = {k: torch.empty((B,) + shape, dtype=torch.float32, device="cuda") for k in ("TU", "AU", "EU")}
    labels = torch.empty((B,) + shape, dtype=torch.uint8, device="cuda")
    ms16 = time_call(lambda: vu.fused_pass(x16, maps_out=maps, labels_out=labels), iters=5)
    ms32 = time_call(lambda: vu.fused_pass(x32, maps_out=maps, labels_out=labels), iters=5)
    msup = time_call(lambda: vu.fused_pass(x16.float(), maps_out=maps, labels_out=labels), iters=5)
    nb16, nb32 = (2 * P * C + 13) * V * B, (4 * P * C + 13) * V * B
    print(f"bf16 slab N=16 C=19 512x1024 B={B}: read as it is {ms16:7.3f} ms ({nb16 / ms16 / 1e6:7.1f} GB/s = {nb16 / ms16 / 1e6 / peak:5.3f} of the HBM peak, "
          f"{P * V * B / ms16 / 1e6:6.1f} G sample-voxels/s); upcast copy + float32 pass {msup:7.3f} ms; float32 slab {ms32:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
    bench_k4()
    bench_generic()
    bench_half()
