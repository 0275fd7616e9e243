#!/usr/bin/env python
"""Headline benchmark: sample-voxels/s of the fused uncertainty + aggregation +
calibration pass on BASELINE.json's sharded-sweep shape (configs[4]: N=16 members,
C=19 classes, 512x1024 images, dataset-level ECE/ACE histograms), per GPU and
aggregated over the GPUs of one box, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one resident batch of synthetic images:
one vu_fused_pass launch (maps, labels, per-image sums / threshold / area / Dice
counts / calibration histograms) plus, on N > 1 GPUs, the all-reduce of the packed
dataset-level partials.  Inputs (10 GB per GPU) are far larger than L2, so no
flush is needed between steps.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sample_voxels_per_s"
UNIT = "sample-voxels/s"
WORKLOAD = dict(P=16, C=19, spatial=(512, 1024), R=1, ignore_index=255, scale=3.0, flip=0.2, ignore_frac=0.02,
                thresholds=(0.3, 0.2, 0.02), platt=((3.5, -1.25), (6.0, -2.0), (40.0, -0.5)))
PEAKS_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def workload_name(images_per_step: int) -> str:
    return ("cfg5 sharded sweep: N=16 members, C=19, 512x1024, R=1 uint8 refs with 2% ignore(255); maps+labels+"
            f"image/threshold/area/Dice/ECE-ACE histograms; {images_per_step} images per GPU per step")


def algorithmic_bytes_per_voxel(P, C, R, gt_bytes=1):
    """slab read + TU/AU/EU fp32 + uint8 label + references."""
    return 4 * P * C + 12 + 1 + R * gt_bytes


# ---------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled by a thread through NVML every ~2 ms (nvidia-smi's 100 ms loop
    cannot see a 35 ms timed region); falls back to `nvidia-smi -lms 100` when NVML cannot be loaded.  `mark()` brackets
    the timed region: the reported median is over the samples inside it (the run also keeps the load on for a moment
    after the region, and says how many samples fell where)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []   # (t, sm_mhz, reasons bitmask)
        self.marks = []
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.backend = None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag:
            try:
                reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_loop(self):
        fields = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={fields}", "--format=csv,noheader,nounits", "-lms", "100",
                                      "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        bits = (0x8, 0x40, 0x20, 0x4)
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            try:
                mask = sum(bit for bit, val in zip(bits, parts[2:6]) if val.lower().startswith("active"))
                self.samples.append((time.perf_counter(), float(parts[0]), mask))
                self.max_mhz = float(parts[1])
            except (ValueError, IndexError):
                continue

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
                except Exception:
                    h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.backend = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
        except Exception:
            self.backend = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
        self.thread.start()

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self):
        self.stop_flag = True
        if self.backend == "nvidia-smi" and getattr(self, "proc", None) is not None:
            self.proc.terminate()
        if self.thread is not None:
            self.thread.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no clock samples"], "samples": 0}
        lo, hi = (self.marks + [None, None])[:2]
        inside = [s for s in self.samples if lo is not None and hi is not None and lo <= s[0] <= hi]
        used = inside if len(inside) >= 3 else self.samples
        sm = sorted(s[1] for s in used)
        mask = 0
        for s_ in used:
            mask |= s_[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(n for b_, n in self.REASONS.items() if mask & b_),
                "samples": len(used), "samples_inside_timed_region": len(inside), "source": self.backend}


# ---------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port)
# ---------------------------------------------------------------------------
def make_cpu_image(seed: int):
    import numpy as np
    import torch
    w = WORKLOAD
    g = torch.Generator().manual_seed(seed)
    x = torch.softmax(w["scale"] * torch.randn(w["P"], w["C"], *w["spatial"], generator=g), dim=1)
    lab0 = x[0].argmax(0).numpy()
    rng = np.random.default_rng(seed)
    gt = np.where(rng.random((w["R"],) + w["spatial"]) < w["flip"], rng.integers(0, w["C"], (w["R"],) + w["spatial"]), lab0[None])
    gt = np.where(rng.random(gt.shape) < w["ignore_frac"], w["ignore_index"], gt).astype(np.uint8)
    return x, gt


def reference_step(x, gt, acc):
    """One image through the reference's call sequence for this workload:
    mean + argmax + calculate_uncertainty, image-level and threshold aggregation, area / border,
    binary Dice counts, per-image ACE / ECE and the dataset accumulator."""
    from oracle import oracle
    w = WORKLOAD
    res = oracle.reference_pipeline_image(x, gt, thresholds=w["thresholds"], platt=w["platt"], ignore_value=w["ignore_index"])
    oracle.binary_dice_counts(res["label"], gt, w["ignore_index"])
    for k, name in enumerate(("TU", "AU", "EU")):
        s, t, n = res[f"{name}/hist"]
        acc[k].bin_sums += s
        acc[k].bin_true += t
        acc[k].bin_total += n
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import oracle
    torch.set_num_threads(os.cpu_count() or 1)  # all host cores (torchrun exports OMP_NUM_THREADS=1)
    w = WORKLOAD
    V = w["spatial"][0] * w["spatial"][1]
    x, gt = make_cpu_image(0)
    acc = [oracle.GlobalCalibAccumulator() for _ in range(3)]
    for _ in range(args.warmup):
        reference_step(x, gt, acc)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_step(x, gt, acc)
    dt = time.perf_counter() - t0
    value = w["P"] * V * args.steps / dt
    cores = torch.get_num_threads()
    sample = f"1 image of the workload per step ({args.steps} timed steps), oracle port of the reference's per-image CPU call sequence"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(1), "timing": "host wall clock, inputs resident in host memory"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def cpu_baseline_leg(n_images: int):
    import torch
    from oracle import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    w = WORKLOAD
    V = w["spatial"][0] * w["spatial"][1]
    x, gt = make_cpu_image(0)
    acc = [oracle.GlobalCalibAccumulator() for _ in range(3)]
    reference_step(x, gt, acc)  # warm-up (allocator, thread pool)
    t0 = time.perf_counter()
    for _ in range(n_images):
        reference_step(x, gt, acc)
    dt = time.perf_counter() - t0
    return {"value": w["P"] * V * n_images / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_images} images of the workload ({dt:.1f} s), oracle port of the reference's per-image CPU call sequence",
            "host_cpus": os.cpu_count()}


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def k1_source_digest() -> str:
    """Digest of the sources the dominant kernel is compiled from: `roofline.traffic` is an ncu measurement of one launch
    (profiles/k1_traffic.json) and is only reported while the kernel it was taken from is the kernel that runs."""
    import hashlib
    h = hashlib.sha256()
    for name in ("k1_tma.cu", "k1_core.cuh", "vu_common.cuh", "stats_v2.cuh", "tma_common.cuh"):
        with open(os.path.join(ROOT, "diffuncertainty_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback. Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None
    if world > 1:
        # every rank feeds its own GPU from host memory in the end-to-end leg: keep the pinned buffers on the GPU's NUMA node
        from diffuncertainty_b200.host_pipeline import bind_host_thread_to_device_node
        numa_node = bind_host_thread_to_device_node(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import diffuncertainty_b200 as vu
    from diffuncertainty_b200 import _lib, calibration, synth
    from diffuncertainty_b200.host_pipeline import HostPipeline, h2d_ceiling_gbs
    from diffuncertainty_b200.sweep import Partials

    w = WORKLOAD
    P, C, S, R = w["P"], w["C"], w["spatial"], w["R"]
    V = S[0] * S[1]
    B = args.images_per_step
    flags = _lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CALIB
    calib = [calibration.platt_edges(a, b) for a, b in w["platt"]]
    lo, hi = rank * B, (rank + 1) * B

    # resident inputs: every rank owns its own block of images (weak scaling)
    x = synth.synth_slab(P, B, C, S, seed=1234, first_image=lo, scale=w["scale"])
    gt_t = synth.synth_gt(x, R, seed=1234, first_image=lo, flip=w["flip"], ignore_frac=w["ignore_frac"],
                          ignore_value=w["ignore_index"])
    gt = vu.GroundTruth(gt_t, w["ignore_index"])
    maps = {k: torch.empty((B,) + S, dtype=torch.float32, device=dev) for k in ("TU", "AU", "EU")}
    labels = torch.empty((B,) + S, dtype=torch.uint8, device=dev)
    # The rows of ALL images of the job in one int64 buffer per step parity (sweep.Partials): the kernel accumulates straight
    # into the rows this rank owns, and ONE int64 all-reduce gathers them (exactly: every element has one non-zero
    # contributor).  Two buffers, so that the exchange of step i runs behind the kernel of step i + 1.
    parts = [Partials(world * B, dev) for _ in range(2)]
    works = [None, None]
    k1_events = []

    def step(i: int, timed: bool, collective: bool = True):
        part = parts[i % 2]
        if works[i % 2] is not None:
            works[i % 2].wait()  # the stream waits for the exchange that last used this buffer (no host block)
            works[i % 2] = None
        part.zero_()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        vu.fused_pass(x, gt, stats=flags, thresholds=w["thresholds"], calib=calib, stats_out=part.local(lo, hi),
                      maps_out=maps, labels_out=labels)
        if timed:
            e1.record()
            k1_events.append((e0, e1))
        if world > 1 and collective:
            works[i % 2] = part.exchange(async_op=True)

    def drain():
        for j in range(2):
            if works[j] is not None:
                works[j].wait()
                works[j] = None

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i, False)
    drain()
    fence()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    launches0 = _lib.get_counter("launches")
    tma0 = _lib.get_counter("launches.k1_tma")
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    sampler.mark()
    start.record()
    for i in range(args.steps):
        step(i, True)
    drain()  # every exchange of the timed steps has completed before the end event
    end.record()
    fence()
    sampler.mark()
    launches = _lib.get_counter("launches") - launches0
    kernel_name = ("k1_tma (vu_fused_pass, TMA-pipelined, producer / consumer / statistics warps)"
                   if _lib.get_counter("launches.k1_tma") - tma0 == launches else "k1_fast (vu_fused_pass, register-streaming)")
    ms_total = start.elapsed_time(end)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / max(1, len(k1_events))
    final = parts[(args.steps - 1) % 2]  # the gathered rows of the last timed step
    # keep the GPU busy a little longer for the fallback sampler (nvidia-smi, 100 ms period); rank 0 only, so without the
    # collective: the other ranks are not taking part.  Uses the buffer that does not hold the last step's rows.
    if rank == 0 and sampler.backend != "nvml":
        t_end = time.time() + 1.0
        while time.time() < t_end:
            step(args.steps, False, collective=False)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total, k1_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, k1_ms = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = P * V * B * world / (ms_per_step * 1e-3)

    # ---- end to end: host buffers through the public host API, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        Be = min(args.e2e_images, B)
        pipe = HostPipeline(P, C, S, Be, R=R, gt_dtype=torch.uint8, chunk_images=1, n_buffers=3, stats=flags,
                            thresholds=w["thresholds"], platt=w["platt"], ignore_index=w["ignore_index"], device=dev)
        xh = torch.empty((P, Be, C) + S, dtype=torch.float32).pin_memory()
        gh = torch.empty((Be, R) + S, dtype=torch.uint8).pin_memory()
        xh.copy_(x[:, :Be])
        gh.copy_(gt_t[:Be])
        torch.cuda.synchronize()
        r0 = pipe.run(xh, gh)  # warm-up
        if rank == 0:  # the host path must reproduce the resident path bit for bit
            assert torch.equal(r0.labels, labels[:Be].cpu()) and torch.equal(r0.maps["TU"], maps["TU"][:Be].cpu())
        fence()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r0 = pipe.run(xh, gh)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        fence()
        ceiling = h2d_ceiling_gbs(xh, dev)  # all ranks at once: what the box's host memory / PCIe delivers to N GPUs
        if world > 1:
            t = torch.tensor([dt, -ceiling], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, ceiling = float(t[0]), -float(t[1])
        e2e = {"value": P * V * Be * world * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": r0.h2d_bytes,
               "d2h_bytes_per_step": r0.d2h_bytes, "images_per_step": Be, "steps": args.e2e_steps, "seconds": dt,
               "h2d_GBps_per_gpu": r0.h2d_bytes * args.e2e_steps / dt / 1e9,
               "h2d_ceiling_GBps_per_gpu": ceiling,
               "ceiling_note": "plain pinned host -> device copy of the same slab, all ranks at the same time (slowest rank)",
               "api": "diffuncertainty_b200.host_pipeline.HostPipeline.run (pinned host slab -> host maps, labels, statistics rows; "
                      "one strided copy per image chunk, 3 chunks in flight)",
               "numa_node_of_rank0": numa_node}
        del pipe, xh, gh

    # ---- parity of the exchanged result with a single-GPU run of the whole job (outside every timed region): rank 0
    # regenerates the images of EVERY rank (the synthetic source is keyed by the image index), runs them on its own GPU into
    # a fresh buffer and compares with what the all-reduce of the last timed step delivered
    parity = None
    if world > 1:
        fence()
        if rank == 0:
            ref = Partials(world * B, dev)
            for r in range(world):
                synth.synth_slab(P, B, C, S, seed=1234, first_image=r * B, scale=w["scale"], out=x)
                g_r = synth.synth_gt(x, R, seed=1234, first_image=r * B, flip=w["flip"], ignore_frac=w["ignore_frac"],
                                     ignore_value=w["ignore_index"])
                vu.fused_pass(x, vu.GroundTruth(g_r, w["ignore_index"]), stats=flags, thresholds=w["thresholds"], calib=calib,
                              stats_out=ref.local(r * B, (r + 1) * B), want_maps=False, want_labels=False)
            torch.cuda.synchronize()
            ints_equal = bool(torch.equal(ref.rows_i, final.rows_i))
            fa, fb = ref.rows_f.cpu().numpy(), final.rows_f.cpu().numpy()
            rel = float(np.max(np.abs(fa - fb) / np.maximum(np.abs(fa), 1e-300)))
            res_a, res_b = ref.result(V, R), final.result(V, R)
            hist_equal = bool(np.array_equal(res_a.bin_total, res_b.bin_total) and np.array_equal(res_a.bin_true, res_b.bin_true))
            parity = {"verdict": "bit-exact" if ints_equal and hist_equal and rel <= 1e-12 else "MISMATCH",
                      "int64_rows_equal": ints_equal, "dataset_histograms_equal": hist_equal, "float64_rows_max_rel_diff": rel,
                      "images": world * B,
                      "how": f"rank 0 re-ran the images of all {world} ranks on one GPU; integer rows and dataset-level bin counts "
                             "compared bit for bit, float64 rows to 1e-12 (their atomics are unordered)"}
            assert parity["verdict"] == "bit-exact", parity
        fence()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = PEAKS_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    bpv = algorithmic_bytes_per_voxel(P, C, R)
    achieved = bpv * V * B / (k1_ms * 1e-3) / 1e9
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("k1_source_digest") != k1_source_digest():
                traffic_source = (f"profiles/k1_traffic.json was measured on another version of the kernel (digest "
                                  f"{tj.get('k1_source_digest')} != {k1_source_digest()}): not reported")
            else:
                traffic = tj.get("dram_bytes_per_launch_at_images", {}).get(str(B))
                if traffic is None and tj.get("dram_bytes_per_voxel") is not None:
                    traffic = tj["dram_bytes_per_voxel"] * V * B
                traffic_source = f"profiles/k1_traffic.json ({tj.get('source')}), kernel source digest {tj.get('k1_source_digest')}"
        except Exception:
            traffic = None
    del x, gt, gt_t, maps, labels
    torch.cuda.empty_cache()

    # ---- every BASELINE config, named pipeline, CUDA events, same process (bench/bench_configs.py)
    configs = None
    if world == 1 and not args.no_configs:
        sys.path.insert(0, os.path.join(ROOT, "bench"))
        from bench_configs import time_config
        configs = []
        for cid in (1, 2, 3, 4, 5):
            c = time_config(cid, iters=10, peak=peak, logits=True)
            configs.append(c)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(B), "parallelism": (f"images sharded over {world} GPU(s); one int64 all-reduce of the per-image rows per step "
                                                                      "(exact gather), running behind the next step's kernel") if world > 1 else "1 GPU",
                       "l2": "inputs (10.2 GB per GPU at 16 images) are larger than L2; no flush between steps",
                       "bytes_per_voxel": bpv},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_source, "kernel": kernel_name, "kernel_ms": k1_ms,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bpv * V * B},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e}
    if parity is not None:
        line["multi_gpu_parity"] = parity["verdict"]
        line["multi_gpu_parity_detail"] = parity
    if configs is not None:
        line["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(args.cpu_images)
    else:
        line["cpu_baseline"] = None
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the duration of the run and the line goes to the
    original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images-per-step", type=int, default=16)
    ap.add_argument("--e2e-images", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config pipelines (configs[0..4]) in the JSON line")
    ap.add_argument("--cpu-images", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    protect_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
