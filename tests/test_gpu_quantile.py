"""GPU parity of the quantile-based pieces: exact rank selection (vu_radix_hist), np.quantile,
threshold discovery (find_threshold.py) and eqACE (ace.py:378-406), against the oracle and the golden vectors recorded
from the unmodified reference (tests/golden/quantile.npz)."""
import json
import os
import pathlib
import tempfile
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    with np.load(os.path.join(GOLDEN_DIR, "quantile.npz")) as z:
        return {k: z[k] for k in z.files}


def same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


def test_rank_selection_equals_sort():
    from diffuncertainty_b200 import quantile
    rng = np.random.default_rng(3)
    for n in (1, 2, 31, 257, 5000, 1 << 20):
        x = (rng.standard_normal(n) * 10.0 ** rng.integers(-20, 20, n)).astype(np.float32)
        x[rng.random(n) < 0.2] = 0.0
        x[rng.random(n) < 0.05] = np.float32(0.25)
        if n > 100:
            x[:3] = [np.inf, -np.inf, np.float32(1e-42)]  # a denormal
        ranks = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, 70)]))  # more than 64: two groups of passes
        got, total = quantile.order_statistics([x], ranks)
        assert total == n
        assert np.array_equal(got, np.sort(x)[ranks]), n
    x = rng.random(1000).astype(np.float32)
    x[rng.random(1000) < 0.1] = np.nan
    got, _ = quantile.order_statistics([x], np.arange(1000))
    assert np.array_equal(got, np.sort(x), equal_nan=True)  # NaN sorts last
    with pytest.raises(IndexError):
        quantile.order_statistics([x], [1000])
    got, total = quantile.order_statistics([np.zeros(0, np.float32)], [0])
    assert total == 0 and np.isnan(got).all()


def test_weighted_selection_counts_valid_raters():
    """one sample per (rater, voxel) pair with a reference that is not ignored (ace.py:492-499)"""
    from diffuncertainty_b200 import quantile
    rng = np.random.default_rng(4)
    for dtype in (np.uint8, np.int64):
        V, R = 3001, 5
        u = (rng.random(V) ** 2).astype(np.float32)
        refs = rng.integers(0, 19, (R, V)).astype(dtype)
        refs[rng.random((R, V)) < 0.3] = 255
        refs[:, :7] = 255  # voxels without a single valid rater
        flat = np.sort(np.repeat(u[None], R, 0)[refs != 255])
        ranks = np.unique(rng.integers(0, flat.size, 50))
        got, total = quantile.order_statistics([u], ranks, [refs], 255)
        assert total == flat.size and np.array_equal(got, flat[ranks])
        got, total = quantile.order_statistics([u], ranks, [refs], None)  # no ignore value: every voxel counts R times
        assert total == R * V and np.array_equal(got, np.sort(np.repeat(u, R))[ranks])


def test_quantile_matches_numpy_and_golden(g):
    from diffuncertainty_b200 import quantile
    maps = [g[f"thr/map{k}"] for k in range(int(g["thr/n_maps"]))]
    for qi, q in enumerate(g["thr/qs"]):
        for k, m in enumerate(maps):
            got = quantile.quantile([torch.from_numpy(m).cuda()], float(q))
            assert float(got) == g["thr/per_map"][qi, k], (q, k)
        assert float(quantile.quantile(maps, float(q))) == g["thr/all_maps"][qi]  # folded, never concatenated
    rng = np.random.default_rng(6)
    x = (rng.random((1024, 2048)) ** 3 * 0.69).astype(np.float32)  # one cfg-3 sized map
    for q in (0.0, 0.5, 0.9, 0.987654321, 1.0):
        assert same_bits(quantile.quantile([x], q), np.quantile(x, q))
    qs = np.linspace(0, 1, 21)
    assert same_bits(quantile.quantile([x], qs, dtype=np.float64), np.quantile(x.astype(np.float64), qs))
    assert same_bits(quantile.quantile([x], qs), np.quantile(x, qs))  # float32 data, float64 array q -> float64 lerp
    x[5, 5] = np.nan
    assert np.isnan(quantile.quantile([x], 0.5)) and np.isnan(np.quantile(x, 0.5))
    with pytest.raises(ValueError):
        quantile.quantile([x], 1.5)


def test_threshold_discovery_drivers(g):
    """get_foreground_quantile -> save_foreground_quantiles -> threshold_images_paths -> find_threshold
    (find_threshold.py:15-112) against the JSON files the reference's drivers wrote."""
    from diffuncertainty_b200 import tasks
    ids = [str(i) for i in g["drv/ids"]]
    with tempfile.TemporaryDirectory() as td:
        root = pathlib.Path(td)
        ds = root / "val"
        ds.mkdir()
        version = types.SimpleNamespace(unc_types=["TU", "AU", "EU"], exp_path=root, pred_model="Softmax", unc_ending=".tif",
                                        version_name="v0")
        loader = types.SimpleNamespace(exp_version=version, image_ids=ids, dataset_path=ds,
                                       unc_path_dict={u: ds / u for u in ("TU", "AU", "EU")},
                                       get_pred_segs=lambda i: list(g[f"drv/{i}/preds"]),
                                       load_unc_file=lambda u, i: g[f"drv/{i}/{u}"])
        with pytest.raises(FileNotFoundError):
            tasks.find_threshold(tasks.threshold_images_paths(loader))
        tasks.save_foreground_quantiles(tasks.get_foreground_quantile(loader))
        tasks.find_threshold(tasks.threshold_images_paths(loader))
        assert json.loads((root / "quantile_analysis.json").read_text()) == json.loads(str(g["drv/quantile_analysis.json"]))
        assert json.loads((root / "threshold_analysis.json").read_text()) == json.loads(str(g["drv/threshold_analysis.json"]))


def test_eqace_matches_golden(g):
    from diffuncertainty_b200 import calibration
    from oracle import oracle
    for name in g["eqace_cases"]:
        want = float(g[f"{name}/eqace"])
        got = calibration.calc_eqace(g[f"{name}/correct"], g[f"{name}/conf"])
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-9, err_msg=f"calc_eqace {name}")
        ign = int(g[f"{name}/ignore"])
        got = calibration.eqace_from_maps(g[f"{name}/refs"], g[f"{name}/pred"], g[f"{name}/unc"], float(g[f"{name}/a"]),
                                          float(g[f"{name}/b"]), None if ign == -999 else ign)
        # an error of mean confidences in [0, 1]: compared on that scale (device expf vs NumPy's exp in the float64 sums)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg=f"eqace_from_maps {name}")
    assert np.isnan(calibration.calc_eqace(np.zeros(0, int), np.zeros(0, np.float32)))
    # a bigger image with int64 references and an ignore value, against the oracle
    rng = np.random.default_rng(8)
    H, W, R = 256, 512, 4
    pred = rng.integers(0, 19, (H, W)).astype(np.uint8)
    refs = np.stack([np.where(rng.random((H, W)) < 0.85, pred, rng.integers(0, 19, (H, W))) for _ in range(R)]).astype(np.int64)
    refs[rng.random(refs.shape) < 0.02] = 255
    unc = (rng.random((H, W)) ** 2 * 2.9).astype(np.float32)
    unc[rng.random((H, W)) < 0.4] = 0.0
    for a, b in ((-3.0, 1.0), (2.0, -1.5)):
        correct, conf = oracle.calibration_inputs(refs, pred, unc, a, b, 255)
        want = oracle.calc_eqace(correct, conf)
        np.testing.assert_allclose(calibration.eqace_from_maps(refs, pred, unc, a, b, 255), want, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(calibration.calc_eqace(correct, conf), want, rtol=1e-5, atol=1e-9)


def test_eqace_batch_equals_per_image_and_oracle():
    """eqace_from_maps_batch (one selection / binning launch sequence for B images x 3 types) against eqace_from_maps image by
    image (identical) and against the oracle's calc_eqace (ace.py:378-406) on the per-(rater, pixel) arrays."""
    import torch
    from diffuncertainty_b200 import calibration
    from oracle import oracle
    rng = np.random.default_rng(21)
    B, H, W, R = 5, 96, 160, 3
    pred = rng.integers(0, 4, (B, H, W)).astype(np.uint8)
    refs = np.stack([np.where(rng.random((B, H, W)) < 0.8, pred, rng.integers(0, 4, (B, H, W))) for _ in range(R)], axis=1).astype(np.uint8)
    refs[rng.random(refs.shape) < 0.03] = 255
    refs[3] = 255  # an image without a single valid sample
    maps = [(rng.random((B, H, W)) ** (k + 1) * 1.3).astype(np.float32) for k in range(3)]
    maps[1][rng.random((B, H, W)) < 0.5] = 0.0  # heavy ties
    maps[2][1, :4] = np.nan
    platt = [(-3.0, 1.0), (2.5, -1.5), (0.0, 0.3)]
    got = calibration.eqace_from_maps_batch(torch.from_numpy(refs).cuda(), torch.from_numpy(pred).cuda(),
                                            [torch.from_numpy(m).cuda() for m in maps], platt, 255)
    assert got.shape == (3, B)
    for m, (a, b) in enumerate(platt):
        for i in range(B):
            one = calibration.eqace_from_maps(refs[i], pred[i], maps[m][i], a, b, 255)
            assert (np.isnan(one) and np.isnan(got[m, i])) or one == got[m, i], (m, i, one, got[m, i])
            if i == 3 or (m == 2 and i == 1):
                continue
            correct, conf = oracle.calibration_inputs(refs[i], pred[i], maps[m][i], a, b, 255)
            np.testing.assert_allclose(got[m, i], oracle.calc_eqace(correct, conf), rtol=1e-5, atol=1e-6)
    assert np.isnan(got[:, 3]).all()


def test_eqace_batch_slices_large_batches():
    """more than 96 segments: the batch goes through in slices (1 MB of histogram workspace per segment), same values"""
    import torch
    from diffuncertainty_b200 import calibration
    g = torch.Generator(device="cuda").manual_seed(9)
    B, shape, R = 40, (16, 32), 2
    pred = (torch.rand((B,) + shape, device="cuda", generator=g) < 0.4).to(torch.uint8)
    refs = (torch.rand((B, R) + shape, device="cuda", generator=g) < 0.4).to(torch.uint8)
    maps = [torch.rand((B,) + shape, device="cuda", generator=g) ** (k + 1) for k in range(3)]
    platt = [(3.5, -1.25), (-2.0, 0.5), (6.0, -2.0)]
    got = calibration.eqace_from_maps_batch(refs, pred, maps, platt)
    assert got.shape == (3, B)
    for m in range(3):
        for i in (0, 31, 32, 39):
            one = calibration.eqace_from_maps(refs[i], pred[i], maps[m][i], *platt[m])
            assert one == got[m, i], (m, i, one, got[m, i])
