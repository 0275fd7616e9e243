"""Host side of the sharded sweep on CPU: image sharding, the
single int64 exchange buffer (float64 rows travel as bit patterns) and its all-reduce over gloo with two
and three processes, and the float64 finalisation.  The rows that the CUDA kernels would
produce are replaced by seeded fake rows -- no compute kernel is called here."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffuncertainty_b200 import sweep
from diffuncertainty_b200._lib import F64, I64


def fake_rows(n_images, seed=0):
    rng = np.random.default_rng(seed)
    f = np.zeros((n_images, F64["COLS"]))
    i = np.zeros((n_images, I64["COLS"]), np.int64)
    tot = rng.integers(0, 5000, (n_images, 3, 21))
    tot[:, :, 20] = 0
    tru = (tot * rng.random((n_images, 3, 21))).astype(np.int64)
    mids = (np.arange(21) + 0.5) / 20
    i[:, I64["BIN_TOTAL"]:I64["BIN_TOTAL"] + 63] = tot.reshape(n_images, 63)
    i[:, I64["BIN_TRUE"]:I64["BIN_TRUE"] + 63] = tru.reshape(n_images, 63)
    f[:, F64["BIN_SUMS"]:F64["BIN_SUMS"] + 63] = (tot * mids).reshape(n_images, 63)
    f[:, F64["SUM"]:F64["SUM"] + 3] = rng.random((n_images, 3)) * 1000
    f[:, F64["THR_SUM"]:F64["THR_SUM"] + 3] = rng.random((n_images, 3)) * 100
    i[:, I64["THR_COUNT"]:I64["THR_COUNT"] + 3] = rng.integers(0, 300, (n_images, 3))
    gs = rng.integers(1, 400, (n_images, 2))
    ps = rng.integers(1, 400, (n_images, 2))
    i[:, I64["DICE_GT"]:I64["DICE_GT"] + 2] = gs
    i[:, I64["DICE_PRED"]:I64["DICE_PRED"] + 2] = ps
    i[:, I64["DICE_TP"]:I64["DICE_TP"] + 2] = np.minimum(gs, ps) * rng.random((n_images, 2))
    return f, i


def single_process_result(n_images):
    f, i = fake_rows(n_images)
    return sweep.pack_partials(torch.from_numpy(f), torch.from_numpy(i), 0, n_images).result(4096, 2)


def test_shard_bounds_cover_every_image_once():
    for n in (0, 1, 7, 8, 10000):
        for world in (1, 2, 3, 8):
            blocks = [sweep.shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_and_finalise_single_process():
    n = 37
    res = single_process_result(n)
    f, i = fake_rows(n)
    assert np.array_equal(res.rows_i64, i) and np.array_equal(res.rows_f64, f)
    assert np.array_equal(res.bin_total.ravel(), i[:, I64["BIN_TOTAL"]:I64["BIN_TOTAL"] + 63].sum(0))
    assert np.array_equal(res.bin_true.ravel(), i[:, I64["BIN_TRUE"]:I64["BIN_TRUE"] + 63].sum(0))
    cal = res.calibration()
    from oracle import oracle
    for k, name in enumerate(("TU", "AU", "EU")):
        gace, gece = oracle.ace_ece_from_histogram(res.bin_sums[k], res.bin_true[k], res.bin_total[k])
        np.testing.assert_allclose([cal[name]["gace"], cal[name]["gece"]], [gace, gece], rtol=1e-12)
    fd = res.failure_detection()
    risks = 1 - res.dice().astype(np.float64)
    np.testing.assert_allclose(fd["EU/image_level"]["aurc"], oracle.aurc(risks, -res.image_level()[:, 2]), rtol=1e-10)
    np.testing.assert_allclose(fd["TU/threshold"]["eaurc"], oracle.eaurc(risks, -res.threshold_level()[:, 0]), rtol=1e-9, atol=1e-12)


def _worker(rank, world, port, n_images, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        f, i = fake_rows(n_images)
        lo, hi = sweep.shard_bounds(n_images, rank, world)
        part = sweep.pack_partials(torch.from_numpy(f[lo:hi].copy()), torch.from_numpy(i[lo:hi].copy()), lo, n_images)
        sweep.exchange(part)  # ONE int64 all-reduce
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), buf=part.buf.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_over_gloo_matches_single_process(world):
    n_images = 23
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, n_images, d), nprocs=world, join=True)
        want = single_process_result(n_images)
        for r in range(world):
            z = np.load(os.path.join(d, f"rank{r}.npz"))
            part = sweep.Partials(n_images, "cpu")
            part.buf.copy_(torch.from_numpy(z["buf"]))
            got = part.result(4096, 2)
            # everything is bit-identical at any world size: the all-reduce adds zeros to every element (float64 rows travel
            # as their bit patterns), and the dataset-level histograms are column sums of identical rows
            assert np.array_equal(got.bin_total, want.bin_total) and np.array_equal(got.bin_true, want.bin_true)
            assert np.array_equal(got.rows_i64, want.rows_i64) and np.array_equal(got.rows_f64, want.rows_f64)
            assert np.array_equal(got.bin_sums, want.bin_sums)
            assert got.calibration()["TU"]["ace"] == pytest.approx(want.calibration()["TU"]["ace"], rel=1e-12)
