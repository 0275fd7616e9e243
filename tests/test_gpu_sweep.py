"""Dataset sweep on the GPU (diffuncertainty_b200.sweep): sharding invariance of the packed partials, parity of the
dataset-level calibration histograms with the oracle, and -- on a multi-GPU box -- the NCCL exchange itself."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(P=6, C=19, spatial=(32, 64), n_images=7, batch=3, R=2, ignore_index=255, seed=21, scale=4.0, flip=0.25, ignore_frac=0.04)


def _cfg():
    from diffuncertainty_b200 import sweep
    return sweep.SweepConfig(**CFG)


@pytest.fixture(scope="module")
def single():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffuncertainty_b200 import sweep
    return sweep.ShardedSweep(_cfg()).run()


def test_sweep_matches_oracle(single):
    from diffuncertainty_b200 import synth
    from oracle import oracle
    cfg = _cfg()
    x = synth.synth_slab(cfg.P, cfg.n_images, cfg.C, cfg.spatial, seed=cfg.seed, first_image=0, scale=cfg.scale)
    gt = synth.synth_gt(x, cfg.R, seed=cfg.seed, first_image=0, flip=cfg.flip, ignore_frac=cfg.ignore_frac, ignore_value=255)
    xc, gc = x.cpu(), gt.cpu().numpy()
    torch.set_num_threads(1)
    accs = [oracle.GlobalCalibAccumulator() for _ in range(3)]
    for b in range(cfg.n_images):
        r = oracle.reference_pipeline_image(xc[:, b], gc[b], thresholds=cfg.thresholds, platt=cfg.platt, ignore_value=255)
        for k, name in enumerate(("TU", "AU", "EU")):
            s, t, n = r[f"{name}/hist"]
            accs[k].bin_sums += s; accs[k].bin_true += t; accs[k].bin_total += n
            np.testing.assert_allclose(single.image_level()[b, k], r[f"{name}/image"], rtol=1e-5)
            np.testing.assert_allclose(single.threshold_level()[b, k], float(r[f"{name}/threshold"]), rtol=1e-5)
    cal = single.calibration()
    for k, name in enumerate(("TU", "AU", "EU")):
        assert np.array_equal(single.bin_total[k], accs[k].bin_total), name
        assert np.array_equal(single.bin_true[k], accs[k].bin_true.astype(np.int64)), name
        np.testing.assert_allclose(single.bin_sums[k], accs[k].bin_sums, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose([cal[name]["gace"], cal[name]["gece"]], [accs[k].compute_ace(), accs[k].compute_ece()], rtol=1e-5)
    fd = single.failure_detection()
    assert set(fd) == {f"{u}/{a}" for u in ("TU", "AU", "EU") for a in ("image_level", "threshold")}


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_invariance_on_one_device(single, world):
    """Ranks run one after the other on one GPU; their packed buffers are summed like the all-reduce would."""
    from diffuncertainty_b200 import sweep
    cfg = _cfg()
    total = None
    for rank in range(world):
        sh = sweep.ShardedSweep(cfg, rank=rank, world=world)
        n_local = sh.hi - sh.lo
        part = sweep.Partials(cfg.n_images, "cuda")
        rows_f, rows_i = part.local(sh.lo, sh.hi)
        from diffuncertainty_b200.uncertainty import GroundTruth, fused_pass
        for s in range(0, n_local, cfg.batch):
            n = min(cfg.batch, n_local - s)
            x, gt = sh.source(sh.lo + s, n)
            fused_pass(x, GroundTruth(gt, cfg.ignore_index), stats=cfg.stats, thresholds=cfg.thresholds, calib=sh._calib,
                       want_maps=False, want_labels=False, stats_out=(rows_f[s:s + n], rows_i[s:s + n]))
        if total is None:
            total = part
        else:
            total.buf += part.buf  # what the int64 all-reduce does
    got = total.result(cfg.V, cfg.R)
    assert np.array_equal(got.bin_total, single.bin_total) and np.array_equal(got.bin_true, single.bin_true)
    assert np.array_equal(got.rows_i64, single.rows_i64)
    np.testing.assert_allclose(got.rows_f64, single.rows_f64, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(got.bin_sums, single.bin_sums, rtol=1e-12)


def _quantile_maps():
    rng = np.random.default_rng(41)
    return [(rng.random(shape) ** 3 * 0.69).astype(np.float32) for shape in ((40, 64), (40, 64), (17, 23), (40, 64), (8, 8, 8))]


def _nccl_worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from diffuncertainty_b200 import sweep
        res = sweep.ShardedSweep(sweep.SweepConfig(**CFG), rank=rank, world=world).run()
        # dataset-level quantile over maps sharded by rank (find_threshold.py:98-105): one int64 all-reduce per histogram pass
        from diffuncertainty_b200 import quantile
        maps = _quantile_maps()
        mine = [torch.from_numpy(m).cuda() for i, m in enumerate(maps) if i % world == rank]
        q = np.array([float(quantile.quantile(mine, qq, reduce=quantile.all_reduce_sum)) for qq in (0.0, 0.37, 0.9, 1.0)])
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), bt=res.bin_total, bc=res.bin_true, bs=res.bin_sums, ri=res.rows_i64, rf=res.rows_f64, q=q)
    finally:
        dist.destroy_process_group()


def test_nccl_exchange_on_all_visible_gpus(single):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("single-GPU box")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_nccl_worker, args=(world, port, d), nprocs=world, join=True)
        for r in range(world):
            z = np.load(os.path.join(d, f"rank{r}.npz"))
            assert np.array_equal(z["bt"], single.bin_total) and np.array_equal(z["bc"], single.bin_true)   # bit-identical
            assert np.array_equal(z["ri"], single.rows_i64)
            np.testing.assert_allclose(z["rf"], single.rows_f64, rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(z["bs"], single.bin_sums, rtol=1e-12)
            allv = np.concatenate([m.ravel() for m in _quantile_maps()])
            assert np.array_equal(z["q"], np.array([float(np.quantile(allv, qq)) for qq in (0.0, 0.37, 0.9, 1.0)]))
