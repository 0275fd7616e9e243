"""Seeded random combinations of what the parity tests cover one by one: shapes (1- to 3-D, aligned or not), member counts
across the cascade-sum boundaries, class counts with and without a specialised kernel, strided views, member lists, peaked
distributions with exact zeros and NaN, raters / ignore values / reference dtypes, every statistics mask, negative Platt
slopes -- against the oracle.  Labels and counts bit-exact, maps and sums within 1e-5."""
import numpy as np
import pytest
import torch

from test_gpu_parity import RTOL, assert_maps_close, oracle_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffuncertainty_b200 as pkg
    return pkg


def random_case(rng):
    ndim = int(rng.integers(1, 4))
    spatial = tuple(int(s) for s in rng.integers(1, 25, ndim))
    if rng.random() < 0.5:  # make the voxel count a multiple of 4: the vectorised kernels
        spatial = spatial[:-1] + (int(rng.choice([4, 8, 16, 32, 64])),)
    P = int(rng.choice([2, 3, 5, 10, 16, 17, 18, 31, 32, 33, 40]))
    C = int(rng.choice([2, 2, 3, 4, 5, 19, 19, 21]))
    B = int(rng.integers(1, 4))
    scale = float(rng.choice([0.5, 2.0, 6.0, 15.0, 40.0]))
    return P, B, C, spatial, scale


@pytest.mark.parametrize("seed", range(24))
def test_random_combination(vu, seed):
    from diffuncertainty_b200 import _lib, calibration, ncc as vncc
    from oracle import oracle
    rng = np.random.default_rng(1000 + seed)
    g = torch.Generator().manual_seed(1000 + seed)
    P, B, C, spatial, scale = random_case(rng)
    x = torch.softmax(scale * torch.randn(P, B, C, *spatial, generator=g), dim=2)
    if rng.random() < 0.3:
        x[torch.rand(x.shape, generator=g) < 0.02] = 0.0            # exact zeros: the p log p skip rule (test_utils.py:838-840)
    if rng.random() < 0.2:
        x[int(rng.integers(0, P)), 0, :, ...][..., 0] = float("nan")  # NaN: argmax's rule, NaN-poisoned means
    layout = rng.choice(["contiguous", "batch_slice", "member_list", "class_padded"])
    xd = x.cuda()
    if layout == "batch_slice":      # softmax_pred[:, i:j] of a bigger batch (test_2D.py:969): strided in B
        big = torch.empty((P, B + 2, C) + spatial, device="cuda")
        big[:, 1:B + 1] = xd
        arg = big[:, 1:B + 1]
    elif layout == "class_padded":   # a view with a larger class stride
        big = torch.empty((P, B, C + 3) + spatial, device="cuda")
        big[:, :, :C] = xd
        arg = big[:, :, :C]
    elif layout == "member_list":    # P separate tensors, read in place
        arg = [xd[p].clone() for p in range(P)]
    else:
        arg = xd
    R = int(rng.integers(1, 6))
    dtype = torch.int64 if rng.random() < 0.5 else torch.uint8
    ignore = rng.choice([None, 255, -1]) if dtype == torch.int64 else rng.choice([None, 255])
    ignore = None if ignore is None else int(ignore)
    member0 = torch.nan_to_num(x[0]).argmax(dim=1)
    noise = torch.randint(0, C, (B, R) + spatial, generator=g)
    gt = torch.where(torch.rand((B, R) + spatial, generator=g) < 0.7, member0.unsqueeze(1).expand((B, R) + spatial), noise)
    if ignore is not None:
        gt = torch.where(torch.rand(gt.shape, generator=g) < 0.1, torch.full_like(gt, ignore), gt)
    gt = gt.to(dtype)
    platt = [(float(rng.choice([3.5, -2.0, 25.0])), float(rng.choice([-1.25, 0.5, 0.0]))) for _ in range(3)]
    thr = [0.2, 0.1, 0.01]
    flags = int(rng.choice([0x01, 0x07, 0x0f, 0x1f, 0x3f, 0x21, 0x11, 0x19]))
    res = vu.fused_pass(arg, vu.GroundTruth(gt.cuda(), ignore) if flags & 0x38 else None, stats=flags, thresholds=thr,
                        calib=[calibration.platt_edges(a, b) for a, b in platt] if flags & _lib.STAT_CALIB else None,
                        want_member_labels=True)
    info = f"seed {seed}: P={P} B={B} C={C} S={spatial} scale={scale} layout={layout} R={R} {dtype} ignore={ignore} flags={flags:#x}"
    for b in range(B):
        ref, label = oracle_image(x[:, b])
        label = label.reshape(spatial)
        got = {k: res.maps[k][b].cpu().numpy().reshape(-1) for k in ("TU", "AU", "EU")}
        assert_maps_close(got, {k: v.reshape(-1) for k, v in ref.items()}, info)
        assert np.array_equal(res.labels[b].cpu().numpy(), label), info
        for p in range(P):
            want = oracle.argmax_first_nan_max(x[p, b].numpy()).astype(np.uint8).reshape(spatial)
            assert np.array_equal(res.member_labels[p, b].cpu().numpy(), want), info
        gnp = gt[b].numpy()
        if flags & _lib.STAT_AREA:
            assert res.area()[b] == oracle.compute_area(label), info
        if flags & _lib.STAT_DICE:
            tp, ps, gs = res.dice_counts()
            otp, ops, ogs = oracle.binary_dice_counts(label, gnp, -12345 if ignore is None else ignore)
            assert np.array_equal(tp[b], otp) and np.array_equal(ps[b], ops) and np.array_equal(gs[b], ogs), info
        for k, name in enumerate(("TU", "AU", "EU")):
            m = res.maps[name][b].cpu().numpy()
            if not np.isfinite(m).all():
                continue  # NaN maps: sums are NaN on both sides; the bin counts are covered by test_golden_calibration
            if flags & _lib.STAT_IMAGE_SUM:
                np.testing.assert_allclose(res.image_level()[b, k], oracle.image_level_aggregation(m)["max_score"], rtol=RTOL, atol=1e-9,
                                           err_msg=info)
            if flags & _lib.STAT_THRESHOLD:
                assert int(res.stats_i64[b, _lib.I64["THR_COUNT"] + k]) == int((m >= np.float32(thr[k])).sum()), info
            if flags & _lib.STAT_CALIB:
                bs, bt, bn = res.calib_histograms()
                correct, conf = oracle.calibration_inputs(gnp, label, m, platt[k][0], platt[k][1], ignore)
                s, t, n = oracle.calib_histogram(correct, conf, binarize=False)
                assert np.array_equal(bn[b, k], n) and np.array_equal(bt[b, k], t.astype(np.int64)), (info, name)
                np.testing.assert_allclose(bs[b, k], s, rtol=RTOL, atol=1e-9, err_msg=info)
            if flags & _lib.STAT_NCC:
                want_ncc = oracle.compute_ncc(oracle.rater_variance_map(gnp), m)
                np.testing.assert_allclose(vncc.ncc_from_result(res, k)[b], want_ncc, rtol=1e-4, atol=1e-6, err_msg=info)
