"""The lean statistics phase (csrc/stats_v2.cuh) -- used by the unified-warp TMA form of the fused pass (csrc/k1_uni.cu)
and by the lean form of vu_map_stats -- against the general statistics phase and against the oracle.

Integers (threshold counts, area, Dice counts, bin_total, bin_true, voxels seen) must be identical; the float64 sums agree
to rounding (different summation order; the NCC products are float32 in the lean form)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S = None
PLATT = [(3.5, -1.25), (6.0, -2.0), (-40.0, 0.5)]   # the last one falls with the uncertainty (conf = 1/(1+exp(-u a + b)), a < 0)
THR = [0.2, 0.15, 0.01]
MASKS = [0x0d, 0x0f, 0x1d, 0x1f, 0x21, 0x3f]


@pytest.fixture(scope="module")
def vu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffuncertainty_b200 as pkg
    from diffuncertainty_b200 import _lib
    _lib.require_device()
    return pkg


def make_case(P, B, spatial, R, ignore, seed, scale=3.0, bad=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.softmax(scale * torch.randn(P, B, 2, *spatial, generator=g), dim=2)
    if bad:  # NaN / inf / zero / negative inputs: the maps get NaN and inf values, which the calibration bins must survive
        flat = x.view(-1)
        idx = torch.randint(0, flat.numel(), (200,), generator=g)
        flat[idx[:50]] = float("nan")
        flat[idx[50:100]] = float("inf")
        flat[idx[100:150]] = 0.0
        flat[idx[150:]] = -0.25
    member0 = x[0].argmax(dim=1)
    noise = torch.randint(0, 2, (B, R, *spatial), generator=g)
    gt = torch.where(torch.rand(B, R, *spatial, generator=g) < 0.7, member0.unsqueeze(1).expand(B, R, *spatial), noise)
    if ignore is not None:
        gt = torch.where(torch.rand(B, R, *spatial, generator=g) < 0.05, torch.full_like(gt, ignore), gt)
    return x.cuda(), gt.to(torch.uint8).cuda()


def run(vu, x, gt, ignore, flags, lean, **kw):
    from diffuncertainty_b200 import _lib, calibration
    calib = [calibration.platt_edges(a, b) for a, b in PLATT]
    _lib.set_option("k1_path", 0 if lean else 1)
    _lib.set_option("stats_path", 0 if lean else 1)
    before = _lib.get_counter("launches.k1_uni")
    try:
        res = vu.fused_pass(x, vu.GroundTruth(gt, ignore), stats=flags, thresholds=THR, calib=calib, **kw)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("k1_path", 0)
        _lib.set_option("stats_path", 0)
    return res, _lib.get_counter("launches.k1_uni") - before


def assert_rows_agree(a, b, rtol=2e-6):
    assert torch.equal(a.stats_i64, b.stats_i64), (a.stats_i64 - b.stats_i64).nonzero()[:8]
    fa, fb = a.stats_f64.cpu().numpy(), b.stats_f64.cpu().numpy()
    assert np.array_equal(np.isnan(fa), np.isnan(fb))
    np.testing.assert_allclose(fa, fb, rtol=rtol, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("flags", MASKS)
@pytest.mark.parametrize("P,B,spatial,R,ignore", [
    (5, 3, (8, 16, 16), 4, None),     # configs[1]-like: one cascade level, no ignore value
    (20, 2, (40, 48), 4, 255),        # two cascade levels, ignore value present
    (32, 5, (128, 128), 4, None),     # configs[3] shape
    (3, 2, (36, 100), 7, 255),        # more than four raters (only the 0x3f mask has an eight-rater lean kernel), ragged last tile
    (9, 1, (8, 12), 1, 3),            # an image smaller than a tile, ignore value that never occurs
])
def test_unified_form_matches_general_form(vu, flags, P, B, spatial, R, ignore):
    x, gt = make_case(P, B, spatial, R, ignore, seed=P * 7 + R)
    lean, n_uni = run(vu, x, gt, ignore, flags, lean=True)
    general, n0 = run(vu, x, gt, ignore, flags, lean=False)
    assert n0 == 0
    if R <= 4 or flags == 0x3f:
        assert n_uni == 1, "the unified-warp kernel did not take this launch"
    assert torch.equal(lean.labels, general.labels)
    for k in ("TU", "AU", "EU"):
        assert torch.equal(lean.maps[k], general.maps[k])
    assert_rows_agree(lean, general)


@pytest.mark.parametrize("flags", [0x1d, 0x3f])
def test_unified_form_with_nan_and_inf_maps(vu, flags):
    x, gt = make_case(6, 2, (24, 64), 3, 255, seed=99, bad=True)
    lean, n_uni = run(vu, x, gt, 255, flags, lean=True)
    general, _ = run(vu, x, gt, 255, flags, lean=False)
    assert n_uni == 1
    # NaN probabilities are skipped by the entropy sums (test_utils.py:838-840), inf ones are not: TU = AU = -inf, EU = NaN
    assert bool(torch.isinf(lean.maps["TU"]).any()) and bool(torch.isnan(lean.maps["EU"]).any()), \
        "the case is meant to produce inf and NaN uncertainties"
    for k in ("TU", "AU", "EU"):
        assert torch.equal(torch.nan_to_num(lean.maps[k], nan=-7.0), torch.nan_to_num(general.maps[k], nan=-7.0))
    assert_rows_agree(lean, general)
    from diffuncertainty_b200._lib import I64
    nan_slot = sum(int(lean.stats_i64[:, I64["BIN_TOTAL"] + 21 * k + 20].sum()) for k in range(3))
    assert nan_slot > 0, "NaN samples belong to slot 20 (np.digitize)"


def test_unified_form_accumulates_and_member_labels_fall_back(vu):
    x, gt = make_case(5, 2, (16, 128), 2, None, seed=5)
    # per-member labels are written by the other two forms: the launch must not go to the unified kernel, and agree with it
    with_labels, n = run(vu, x, gt, None, 0x1d, lean=True, want_member_labels=True)
    assert n == 0
    assert torch.equal(with_labels.member_labels, x.argmax(dim=2).to(torch.uint8))
    once, n = run(vu, x, gt, None, 0x1d, lean=True)
    assert n == 1
    assert torch.equal(with_labels.stats_i64, once.stats_i64)
    torch.testing.assert_close(with_labels.stats_f64, once.stats_f64, rtol=1e-7, atol=0)  # float32 partial sums per tile, another tile size
    sf, si = torch.zeros_like(once.stats_f64), torch.zeros_like(once.stats_i64)
    for _ in range(3):
        run(vu, x, gt, None, 0x1d, lean=True, stats_out=(sf, si))
    assert torch.equal(si, 3 * once.stats_i64)
    torch.testing.assert_close(sf, 3 * once.stats_f64, rtol=1e-9, atol=0)


@pytest.mark.parametrize("flags", [0x07, 0x0f, 0x1d, 0x1f, 0x21, 0x3f])
@pytest.mark.parametrize("B,spatial,R,ignore", [(3, (8, 16, 16), 4, None), (2, (40, 52), 5, 255), (4, (128, 128), 1, 255)])
def test_map_stats_lean_form_matches_general_form(vu, flags, B, spatial, R, ignore):
    from diffuncertainty_b200 import _lib, calibration
    g = torch.Generator().manual_seed(B + R)
    maps = {k: (torch.rand(B, *spatial, generator=g) ** 3 * 0.69).cuda() for k in ("TU", "AU", "EU")}
    maps["AU"][0].view(-1)[5] = float("nan")
    maps["EU"][B - 1].view(-1)[9] = float("inf")
    labels = (torch.rand(B, *spatial, generator=g) < 0.3).to(torch.uint8).cuda()
    gt = (torch.rand(B, R, *spatial, generator=g) < 0.3).to(torch.uint8)
    if ignore is not None:
        gt = torch.where(torch.rand(B, R, *spatial, generator=g) < 0.05, torch.full_like(gt, ignore), gt)
    calib = [calibration.platt_edges(a, b) for a, b in PLATT]
    out = []
    for path in (0, 1):
        _lib.set_option("stats_path", path)
        before = _lib.get_counter("launches.k3_map_stats_v2")
        try:
            res = vu.map_stats(maps, labels, vu.GroundTruth(gt.cuda(), ignore), stats=flags, thresholds=THR, calib=calib)
            torch.cuda.synchronize()
        finally:
            _lib.set_option("stats_path", 0)
        out.append((res, _lib.get_counter("launches.k3_map_stats_v2") - before))
    assert out[0][1] == 1 and out[1][1] == 0
    assert_rows_agree(out[0][0], out[1][0])


def test_unified_form_vs_oracle_cfg2_like(vu):
    """configs[1] shape family straight against the oracle: labels, Dice counts and (given our maps) bin counts bit-exact."""
    from diffuncertainty_b200 import _lib, calibration
    from oracle import oracle
    P, B, spatial, R = 5, 2, (8, 16, 32), 4
    x, gt = make_case(P, B, spatial, R, None, seed=11, scale=4.0)
    res, n = run(vu, x, gt, None, 0x1d, lean=True)
    assert n == 1
    bs, bt, bn = res.calib_histograms()
    tp, ps, gs = res.dice_counts()
    xc, gc = x.cpu(), gt.cpu().numpy()
    torch.set_num_threads(1)
    for b in range(B):
        label = oracle.argmax_first_nan_max(oracle.mean_members_f32(xc[:, b].numpy())).astype(np.uint8)
        assert np.array_equal(res.labels[b].cpu().numpy(), label)
        otp, ops, ogs = oracle.binary_dice_counts(label, gc[b], -12345)
        assert np.array_equal(tp[b], otp) and np.array_equal(ps[b], ops) and np.array_equal(gs[b], ogs)
        assert res.area()[b] == oracle.compute_area(label)
        for k, name in enumerate(("TU", "AU", "EU")):
            m = res.maps[name][b].cpu().numpy()
            np.testing.assert_allclose(res.image_level()[b, k], oracle.image_level_aggregation(m)["max_score"], rtol=1e-5, atol=1e-9)
            correct, conf = oracle.calibration_inputs(gc[b], label, m, PLATT[k][0], PLATT[k][1], None)
            s, t, nn = oracle.calib_histogram(correct, conf, binarize=False)
            assert np.array_equal(bn[b, k], nn) and np.array_equal(bt[b, k], t.astype(np.int64)), name
            np.testing.assert_allclose(bs[b, k], s, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("lean", [True, False])
def test_zero_slope_platt_bins_like_numpy(vu, lean):
    """a == 0 (the reference's fallback when the Platt fit fails): the constant confidence puts every sample into ONE bin, the
    one np.digitize picks -- in both statistics forms."""
    from diffuncertainty_b200 import _lib, calibration
    x, gt = make_case(5, 2, (16, 64), 2, None, seed=2)
    platt = [(0.0, 0.3), (0.0, -2.0), (3.5, -1.25)]
    _lib.set_option("k1_path", 0 if lean else 1)
    _lib.set_option("stats_path", 0 if lean else 1)
    try:
        res = vu.fused_pass(x, vu.GroundTruth(gt, None), stats=0x1d, calib=[calibration.platt_edges(a, b) for a, b in platt])
    finally:
        _lib.set_option("k1_path", 0)
        _lib.set_option("stats_path", 0)
    _, _, bn = res.calib_histograms()
    n = 2 * 16 * 64
    for k, (a, b) in enumerate(platt[:2]):
        conf = np.clip(1 / (1 + np.exp(np.float32(0.0) * np.float32(a) + np.float32(b))), 0, 1)
        want = np.digitize(np.float32(conf), np.linspace(0, 1 + 1e-8, 21)) - 1
        for img in range(2):
            assert bn[img, k, want] == n and bn[img, k].sum() == n, (k, want, bn[img, k])
