"""GPU parity: the CUDA path (through the C ABI) against the oracle and the
golden vectors recorded from the unmodified reference.

Tolerances (BASELINE.json north_star):
  * labels, areas, borders, Dice counts, bin counts, threshold counts: bit-exact
  * TU, AU, scores, bin_sums, NCC: |d| <= RTOL * |ref| (+ ATOL for values that
    are ~0 in the reference: subnormal probabilities are flushed on the device)
  * EU = TU - AU cancels, so |d| <= RTOL * max(|TU|, |AU|)
"""
import numpy as np
import pytest
import torch

from conftest import case_names

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL = 1e-10


@pytest.fixture(scope="module")
def vu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffuncertainty_b200 as pkg
    from diffuncertainty_b200 import _lib
    _lib.require_device()
    return pkg


def assert_maps_close(got, ref, name=""):
    """got / ref: dicts with TU, AU, EU numpy arrays."""
    tu, au = ref["TU"].astype(np.float64), ref["AU"].astype(np.float64)
    for key in ("TU", "AU"):
        r = ref[key].astype(np.float64)
        g = got[key].astype(np.float64)
        fin = np.isfinite(r)
        assert np.array_equal(np.isfinite(g), fin), (name, key, "finite pattern")
        assert np.array_equal(g[~fin], r[~fin], equal_nan=True), (name, key)
        err = np.abs(g[fin] - r[fin])
        tol = RTOL * np.abs(r[fin]) + ATOL
        assert np.all(err <= tol), (name, key, float((err / np.maximum(np.abs(r[fin]), 1e-30)).max()))
    fin = np.isfinite(tu) & np.isfinite(au)
    err = np.abs(got["EU"].astype(np.float64)[fin] - ref["EU"].astype(np.float64)[fin])
    tol = RTOL * np.maximum(np.abs(tu[fin]), np.abs(au[fin])) + ATOL
    assert np.all(err <= tol), (name, "EU", float(err.max()))


def oracle_image(x_cpu):
    from oracle import oracle
    torch.set_num_threads(1)
    u = oracle.calculate_uncertainty(x_cpu)
    mean = oracle.mean_members_f32(x_cpu.numpy())
    label = oracle.argmax_first_nan_max(mean).astype(np.uint8)
    return {k: v.numpy() for k, v in u.items()}, label


# --------------------------------------------------------------------------
# golden vectors recorded from the reference
# --------------------------------------------------------------------------
def test_golden_uncertainty_maps_and_labels(vu, golden_unc):
    for name in case_names(golden_unc):
        x = torch.from_numpy(golden_unc[f"{name}/x"]).cuda()
        if name == "msr_p1":
            got = vu.calculate_one_minus_msr(x.squeeze(0))["pred_entropy"].cpu().numpy()
            np.testing.assert_allclose(got, golden_unc[f"{name}/pred_entropy"], rtol=1e-6, atol=1e-7)
            res = vu.fused_pass(x.unsqueeze(1))
            assert np.array_equal(res.labels[0].cpu().numpy(), golden_unc[f"{name}/label"])
            continue
        got = {k: v.cpu().numpy() for k, v in vu.calculate_uncertainty(x).items()}
        for k in got:
            assert got[k].dtype == np.float32 and got[k].shape == x.shape[2:]
        ref = {k: golden_unc[f"{name}/{k}"] for k in ("TU", "AU", "EU")}
        assert_maps_close(got, ref, name)
        res = vu.fused_pass(x.unsqueeze(1))
        lab = res.labels[0].cpu().numpy()
        # the reference's torch.mean uses an interleaved order in its SIMD tail (last numel % 32
        # elements, see oracle.cascade_sum_f32); labels there can only differ on 1-ulp ties
        from oracle import oracle
        canon = oracle.argmax_first_nan_max(oracle.mean_members_f32(golden_unc[f"{name}/x"])).astype(np.uint8)
        assert np.array_equal(lab, canon), name
        n = lab.size * x.shape[1]
        body = (n // 32) * 32 // x.shape[1]
        assert np.array_equal(lab.ravel()[: max(body - 32, 0)], golden_unc[f"{name}/label"].ravel()[: max(body - 32, 0)]), name


def test_plogp_accuracy_dense_sweep(vu):
    """-p log p against float64 over 30 decades and densely next to p = 1."""
    p = np.concatenate([np.logspace(-37, 0, 40000), 1 - np.logspace(-8, -0.31, 40000), np.linspace(0.4, 1.0, 40000),
                        1 + np.logspace(-7, 0, 4000)]).astype(np.float32)
    n = p.size
    x = torch.zeros(2, 1, 2, n)
    x[:, 0, 0] = torch.from_numpy(p)
    res = vu.fused_pass(x.cuda())
    tu = res.maps["TU"][0].cpu().numpy().astype(np.float64)
    p64 = p.astype(np.float64)
    want = -p64 * np.log(p64)
    err = np.abs(tu - want)
    # relative to the term itself, except right at p = 1 where the term vanishes
    assert np.all(err <= 2e-6 * np.abs(want) + 1e-9 * np.abs(p64 - 1) + 1e-37), float((err / np.maximum(np.abs(want), 1e-30)).max())


# --------------------------------------------------------------------------
# seeded random slabs against the oracle: shapes of all BASELINE configs (reduced)
# --------------------------------------------------------------------------
CASES = [
    # P, B, C, spatial, scale
    (10, 3, 2, (32, 32), 2.0),      # cfg1
    (5, 2, 2, (8, 16, 16), 2.0),    # cfg2 (3-D)
    (10, 1, 19, (16, 64), 6.0),     # cfg3
    (32, 2, 2, (32, 32), 10.0),     # cfg4, peaked
    (16, 2, 19, (8, 64), 3.0),      # cfg5
    (17, 1, 2, (16, 32), 2.0), (18, 1, 2, (16, 32), 2.0), (33, 1, 19, (4, 32), 2.0),
    (2, 1, 3, (8, 16), 2.0), (7, 2, 4, (8, 16), 2.0), (48, 1, 4, (8, 16), 4.0),
    (6, 1, 5, (8, 16), 2.0), (3, 1, 21, (4, 16), 2.0),           # generic kernel (C not specialised)
    (10, 2, 2, (7, 9), 2.0), (10, 1, 19, (5, 13), 2.0), (20, 1, 3, (3, 11), 2.0),   # unaligned V
    (272, 1, 2, (4, 16), 2.0), (300, 1, 3, (2, 16), 2.0),        # three cascade levels -> generic
]


@pytest.mark.parametrize("P,B,C,spatial,scale", CASES)
def test_random_slabs_vs_oracle(vu, P, B, C, spatial, scale):
    g = torch.Generator().manual_seed(P * 1000 + C * 10 + len(spatial))
    x = torch.softmax(scale * torch.randn(P, B, C, *spatial, generator=g), dim=2)
    res = vu.fused_pass(x.cuda())
    for b in range(B):
        ref, label = oracle_image(x[:, b])
        got = {k: res.maps[k][b].cpu().numpy() for k in ("TU", "AU", "EU")}
        assert_maps_close(got, ref, f"P{P} C{C} b{b}")
        assert np.array_equal(res.labels[b].cpu().numpy(), label.reshape(spatial))
        # create_TU.py:46-55 invariant: TU == AU + EU up to an ulp
        np.testing.assert_allclose(got["AU"] + got["EU"], got["TU"], rtol=3e-7, atol=1e-7)


def test_strided_views_are_read_in_place(vu):
    g = torch.Generator().manual_seed(9)
    big = torch.softmax(2 * torch.randn(6, 5, 2, 16, 48, generator=g), dim=2).cuda()
    views = {
        "image slice (test_2D.py:969)": big[:, 2:3],
        "every other image": big[:, ::2],
        "cropped columns (stride_v=1, rows not contiguous -> copy)": big[..., 8:40],
        "class-last memory layout": big.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3),
        "column step 2 (stride_v = 2)": big[..., ::2] if False else big[:, :, :, 0:1, ::2],
    }
    for name, v in views.items():
        res = vu.fused_pass(v)
        ref = vu.fused_pass(v.contiguous())
        for k in ("TU", "AU", "EU"):
            assert torch.equal(res.maps[k], ref.maps[k]), name
        assert torch.equal(res.labels, ref.labels), name
        rr, label = oracle_image(v[:, 0].cpu())
        assert_maps_close({k: res.maps[k][0].cpu().numpy() for k in ("TU", "AU", "EU")}, rr, name)


def test_input_contract(vu):
    with pytest.raises(Exception):
        vu.calculate_uncertainty(torch.rand(3, 2, 4, 4))  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        vu.fused_pass(torch.rand(3, 2, 4, device="cuda"))
    empty = vu.fused_pass(torch.rand(3, 0, 2, 4, 4, device="cuda"))
    assert empty.labels.shape == (0, 4, 4)
    x = torch.softmax(torch.randn(4, 2, 8, 8, dtype=torch.float64), 1).cuda()
    out = vu.calculate_uncertainty(x)  # fp64 in -> fp32 out (test_utils.py:836)
    assert out["TU"].dtype == torch.float32
    one = torch.softmax(torch.randn(1, 3, 8, 8), 1)
    from oracle import oracle
    ref = oracle.calculate_uncertainty(one)
    got = vu.calculate_uncertainty(one.cuda())
    np.testing.assert_allclose(got["TU"].cpu().numpy(), ref["TU"].numpy(), rtol=RTOL)
    assert float(got["EU"].abs().max()) <= 1e-6


# --------------------------------------------------------------------------
# fused per-image statistics
# --------------------------------------------------------------------------
@pytest.mark.parametrize("P,B,C,spatial,R,ignore,dtype", [
    (5, 3, 2, (8, 16, 16), 4, None, torch.uint8),       # cfg2-like
    (10, 2, 19, (16, 64), 5, 255, torch.uint8),         # cfg3-like with ignore
    (32, 3, 2, (32, 32), 4, None, torch.int64),         # cfg4-like, int64 gt as the loader hands it
    (16, 2, 19, (8, 64), 1, 255, torch.int64),          # cfg5-like
    (6, 2, 5, (9, 7), 3, -1, torch.int64),              # generic kernel, unaligned, ignore -1
])
def test_fused_statistics_vs_oracle(vu, P, B, C, spatial, R, ignore, dtype):
    from diffuncertainty_b200 import _lib, aurc as vaurc, calibration, ncc as vncc
    from oracle import oracle
    g = torch.Generator().manual_seed(P + C + R)
    x = torch.softmax(4 * torch.randn(P, B, C, *spatial, generator=g), dim=2)
    member0 = x[0].argmax(dim=1)
    noise = torch.randint(0, C, (B, R, *spatial), generator=g)
    gt = torch.where(torch.rand(B, R, *spatial, generator=g) < 0.75, member0.unsqueeze(1).expand(B, R, *spatial), noise)
    if ignore is not None:
        gt = torch.where(torch.rand(B, R, *spatial, generator=g) < 0.05, torch.full_like(gt, ignore), gt)
    gt = gt.to(dtype) if dtype == torch.int64 else (gt % 256).to(torch.uint8)
    ign = None if ignore is None else (ignore if dtype == torch.int64 else ignore % 256)
    platt = [(3.5, -1.25), (6.0, -2.0), (40.0, -0.5)]
    thr = [0.2, 0.15, 0.01]
    flags = (_lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CALIB | _lib.STAT_NCC)
    res = vu.fused_pass(x.cuda(), vu.GroundTruth(gt.cuda(), ign), stats=flags, thresholds=thr,
                        calib=[calibration.platt_edges(a, b) for a, b in platt])
    img_mean, thr_score = res.image_level(), res.threshold_level()
    tp, ps, gs = res.dice_counts()
    bs, bt, bn = res.calib_histograms()
    for b in range(B):
        _, label = oracle_image(x[:, b])
        label = label.reshape(spatial)
        assert np.array_equal(res.labels[b].cpu().numpy(), label)
        gnp = gt[b].numpy()
        assert res.area()[b] == oracle.compute_area(label)
        otp, ops, ogs = oracle.binary_dice_counts(label, gnp, -12345 if ign is None else ign)
        assert np.array_equal(tp[b], otp) and np.array_equal(ps[b], ops) and np.array_equal(gs[b], ogs)
        np.testing.assert_allclose(vaurc.binary_dice_from_counts(tp[b], ps[b], gs[b]),
                                   oracle.binary_dice_from_counts(otp, ops, ogs), rtol=1e-6)
        for k, name in enumerate(("TU", "AU", "EU")):
            m = res.maps[name][b].cpu().numpy()  # OUR map: counts must be exact given the map
            np.testing.assert_allclose(img_mean[b, k], oracle.image_level_aggregation(m)["max_score"], rtol=RTOL, atol=1e-9)
            o = oracle.threshold_aggregation(m, np.float32(thr[k]))
            np.testing.assert_allclose(thr_score[b, k], float(o["max_score"]), rtol=RTOL)
            assert int(res.stats_i64[b, _lib.I64["THR_COUNT"] + k]) == int((m >= np.float32(thr[k])).sum())
            correct, conf = oracle.calibration_inputs(gnp, label, m, platt[k][0], platt[k][1], ign)
            s, t, n = oracle.calib_histogram(correct, conf, binarize=False)
            assert np.array_equal(bn[b, k], n), (name, "bin_total")
            assert np.array_equal(bt[b, k], t.astype(np.int64)), (name, "bin_true")
            np.testing.assert_allclose(bs[b, k], s, rtol=RTOL, atol=1e-9)
            ace, ece = calibration.per_image_ace_ece(bs[b, k], bt[b, k], bn[b, k])
            np.testing.assert_allclose(ace, oracle.calc_ace(correct, conf), rtol=RTOL, atol=1e-9)
            np.testing.assert_allclose(ece, oracle.calc_ece(correct, conf), rtol=RTOL, atol=1e-9)
            want_ncc = oracle.compute_ncc(oracle.rater_variance_map(gnp), m)
            np.testing.assert_allclose(vncc.ncc_from_result(res, k)[b], want_ncc, rtol=1e-4, atol=1e-7)


def test_statistics_accumulate_across_calls(vu):
    from diffuncertainty_b200 import _lib
    x = torch.softmax(torch.randn(4, 2, 2, 16, 16), 2).cuda()
    once = vu.fused_pass(x, stats=_lib.STAT_IMAGE_SUM | _lib.STAT_AREA)
    sf = torch.zeros_like(once.stats_f64)
    si = torch.zeros_like(once.stats_i64)
    for _ in range(3):
        vu.fused_pass(x, stats=_lib.STAT_IMAGE_SUM | _lib.STAT_AREA, stats_out=(sf, si))
    assert torch.equal(si, 3 * once.stats_i64)
    torch.testing.assert_close(sf, 3 * once.stats_f64, rtol=1e-12, atol=0)


# --------------------------------------------------------------------------
# aggregation / calibration / NCC drop-ins against the golden vectors
# --------------------------------------------------------------------------
def test_golden_aggregations(vu, golden_agg):
    from diffuncertainty_b200 import aggregation as agg
    for name in ("img2d", "img3d", "plateau"):
        img = golden_agg[f"{name}/image"]
        np.testing.assert_allclose(agg.image_level_aggregation(img)["max_score"], golden_agg[f"{name}/image_level_mean"], rtol=RTOL)
        np.testing.assert_allclose(agg.image_level_aggregation(img, mean=False)["max_score"], golden_agg[f"{name}/image_level_sum"], rtol=RTOL)
        for ps in (10, 4):
            r = agg.patch_level_aggregation(img, ps)
            np.testing.assert_allclose(r["max_score"], golden_agg[f"{name}/patch{ps}_score"], rtol=RTOL)
            assert np.array_equal(np.asarray(r["bounding_box"]), golden_agg[f"{name}/patch{ps}_bbox"]), (name, ps)
            r = agg.patch_level_aggregation(torch.from_numpy(img).cuda(), ps, mean=True, image_id="x", unc_type="TU")
            np.testing.assert_allclose(r["max_score"], golden_agg[f"{name}/patch{ps}_mean_score"], rtol=RTOL)
        for t in ("mid", "above_max", "zero"):
            thr = float(golden_agg[f"{name}/thr_{t}_t"])
            np.testing.assert_allclose(float(agg.threshold_aggregation(img, threshold=thr)["max_score"]),
                                       golden_agg[f"{name}/thr_{t}_score"], rtol=RTOL)
            np.testing.assert_allclose(float(agg.threshold_aggregation(img, threshold=thr, mean=False)["max_score"]),
                                       golden_agg[f"{name}/thr_{t}_sum"], rtol=RTOL)
    for name in ("lab2d", "lab3d", "empty"):
        lab = golden_agg[f"{name}/label"]
        area, border = agg.prediction_shape_stats(lab)
        assert area == golden_agg[f"{name}/area"] and border == golden_agg[f"{name}/border"], name
        np.testing.assert_allclose(agg._normalize_uncertainty_sum(golden_agg["img2d/image"], area),
                                   golden_agg[f"{name}/norm_area"], rtol=RTOL)
    with pytest.raises(Exception):
        agg.threshold_aggregation(golden_agg["img2d/image"])


@pytest.mark.parametrize("shape,box", [
    ((40, 70, 75), 10),           # 3-D, several 32 x 32 windows per image, ragged edges
    ((64, 64, 64), 10),           # BASELINE configs[1]: the plane kernel, one 64 x 64 window per slice
    ((30, 130, 150), 10),         # plane kernel: 3 x 3 windows, rows that are not 16-byte aligned
    ((13, 17, 23), 10),           # plane kernel: a map smaller than one window, chunks of one output slice
    ((12, 33, 97), (3, 5, 7)),    # anisotropic box
    ((100, 130), 7),              # 2-D box without a specialised kernel
    ((300, 520), 10),             # 2-D strip kernel, several strips and row chunks per image
    ((75, 523), 10),              # 2-D strip kernel, rows that are not 16-byte aligned
    ((33, 41), 4),                # 2-D strip kernel, box edge 4
    ((9, 40, 40), (9, 33, 33)),   # the largest box of the sliding kernel, one output slice
    ((6, 50, 50), (2, 34, 34)),   # larger still: the tiled fallback kernel
])
def test_patch_level_multi_window(vu, shape, box):
    """patch_level_aggregation (aggregate_uncertainties.py:16-34) on maps that span several CTAs, with and without the
    per-CTA-maximum workspace, against the oracle (scipy's convolve, as the reference calls it)."""
    from diffuncertainty_b200 import _lib, aggregation as agg
    from oracle import oracle
    rng = np.random.default_rng(17)
    ks = [box] * len(shape) if isinstance(box, int) else list(box)
    imgs = []
    a = (rng.random(shape) ** 4 * 0.69).astype(np.float32)                   # one dominant blob somewhere
    imgs.append(a)
    b = np.zeros(shape, np.float32)                                          # plateau: many boxes tie with the maximum,
    b[tuple(slice(s // 3, s) for s in shape)] = 0.5                          # the first one in row-major order wins
    imgs.append(b)
    c = (rng.random(shape) * 1e-6).astype(np.float32)                        # two equal far-apart peaks in different windows
    lo = tuple(slice(0, k) for k in ks)
    hi = tuple(slice(s - k, s) for s, k in zip(shape, ks))
    c[lo] += 1.0
    c[hi] += 1.0
    imgs.append(c)
    for mean in (False, True):
        for img in imgs:
            want = oracle.patch_level_aggregation(img, ks, mean=mean)
            got = agg.patch_level_aggregation(img, ks, mean=mean)
            np.testing.assert_allclose(got["max_score"], want["max_score"], rtol=5e-7)  # scipy goes through an FFT (1e-7 of the largest value)
            assert got["bounding_box"] == want["bounding_box"], (shape, box, mean)
    # the plain entry point (no workspace) gives the same answer as the wrapper's vu_patch_max_ws
    lib = _lib.load()
    dims = [1] * (3 - len(shape)) + list(shape)
    k3 = [1] * (3 - len(shape)) + ks
    t = torch.from_numpy(np.stack(imgs)).cuda()
    batched = agg.patch_level_batched(t.reshape(len(imgs), *dims), k3)
    out_max = torch.empty(len(imgs), dtype=torch.float64, device="cuda")
    out_first = torch.empty(len(imgs), dtype=torch.int64, device="cuda")
    _lib.check(lib.vu_patch_max(t.data_ptr(), len(imgs), *dims, *k3, 0, out_max.data_ptr(), out_first.data_ptr(),
                                _lib.current_stream_ptr()), "vu_patch_max")
    assert np.array_equal(out_max.cpu().numpy(), batched["max_score"]) and np.array_equal(out_first.cpu().numpy(), batched["first_index"])


def test_patch_plane_kernel_equals_sliding_kernel(vu):
    """The r02 kernels behind vu_patch_max_ws (3-D plane kernel, 2-D strip kernel) against the r01 ones (option patch_path = 1) on
    batches large enough for one or two chunks of output slices per window: same maxima (float64 sums of float32 values: the
    order of the additions differs, 1e-12) and the same first-isclose indices."""
    from diffuncertainty_b200 import _lib, aggregation as agg
    g = torch.Generator(device="cuda").manual_seed(5)
    for B, dims in ((130, (64, 64, 64)), (40, (24, 70, 64)), (6, (1, 512, 1024)), (9, (1, 100, 301))):
        maps = torch.rand((B, *dims), device="cuda", generator=g) ** 3 * 0.69
        maps[1] = 0.25                      # every box ties
        maps[2, :, :, :] = 0.0
        maps[2, 40 % dims[0]:, 5:, 7:] = 0.5  # plateau
        box = (10, 10, 10) if dims[0] > 1 else (1, 10, 10)
        new = agg.patch_level_batched(maps, box)
        _lib.load().vu_set_option(b"patch_path", 1)
        try:
            old = agg.patch_level_batched(maps, box)
        finally:
            _lib.load().vu_set_option(b"patch_path", 0)
        np.testing.assert_allclose(new["max_score"], old["max_score"], rtol=1e-12)
        assert np.array_equal(new["first_index"], old["first_index"])


def test_golden_calibration(vu, golden_calib):
    from diffuncertainty_b200 import _lib, calibration
    from diffuncertainty_b200.uncertainty import GroundTruth
    import ctypes as C
    for name in case_names(golden_calib):
        g = {k.split("/", 1)[1]: v for k, v in golden_calib.items() if k.startswith(name + "/")}
        # (1) drop-in signature: per-pixel correct / confidence arrays
        np.testing.assert_allclose(calibration.calc_ace(g["correct"], g["conf"]), g["ace"], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(calibration.calc_ece(g["correct"], g["conf"]), g["ece"], rtol=RTOL, atol=1e-12)
        acc = calibration.GlobalCalibAccumulator()
        acc.accumulate(g["correct"], g["conf"])
        assert np.array_equal(acc.bin_total, g["g_bin_total"]), name
        assert np.array_equal(acc.bin_true, g["g_bin_true"]), name
        np.testing.assert_allclose(acc.bin_sums, g["g_bin_sums"], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(acc.compute_ace(), g["gace"], rtol=RTOL)
        np.testing.assert_allclose(acc.compute_ece(), g["gece"], rtol=RTOL)
        # (2) from the raw maps: refs, predicted labels, uncertainty map and Platt parameters
        lib = _lib.load()
        unc = torch.from_numpy(g["unc"]).cuda().contiguous()
        pred = torch.from_numpy(g["pred"]).cuda().contiguous()
        refs = torch.from_numpy(g["refs"]).cuda().contiguous()
        sf = torch.zeros((1, _lib.F64["COLS"]), dtype=torch.float64, device="cuda")
        si = torch.zeros((1, _lib.I64["COLS"]), dtype=torch.int64, device="cuda")
        a = _lib.MapStatsArgs()
        a.struct_size = C.sizeof(_lib.MapStatsArgs)
        a.stat_flags = _lib.STAT_CALIB
        a.B, a.V = 1, unc.numel()
        a.maps[0] = unc.data_ptr()
        a.labels = pred.data_ptr()
        a.gt.data, a.gt.dtype, a.gt.R = refs.data_ptr(), _lib.GT_U8, refs.shape[0]
        a.gt.stride_b, a.gt.stride_r, a.gt.stride_v = refs.numel(), unc.numel(), 1
        ign = int(g["ignore"])
        a.gt.has_ignore, a.gt.ignore_index = (0, 0) if ign == -999 else (1, ign)
        pe = calibration.platt_edges(float(g["a"]), float(g["b"])).as_struct()
        for k in range(3):
            a.calib[k] = pe
        a.stats_f64, a.stats_i64 = sf.data_ptr(), si.data_ptr()
        _lib.check(lib.vu_map_stats(C.byref(a), _lib.current_stream_ptr()), "vu_map_stats")
        i = si.cpu().numpy()[0]
        f = sf.cpu().numpy()[0]
        assert np.array_equal(i[_lib.I64["BIN_TOTAL"]:_lib.I64["BIN_TOTAL"] + 21], g["g_bin_total"]), name
        assert np.array_equal(i[_lib.I64["BIN_TRUE"]:_lib.I64["BIN_TRUE"] + 21], g["g_bin_true"].astype(np.int64)), name
        np.testing.assert_allclose(f[_lib.F64["BIN_SUMS"]:_lib.F64["BIN_SUMS"] + 21], g["g_bin_sums"], rtol=RTOL, atol=1e-12)
    with pytest.raises(ValueError):
        calibration.calc_ace(np.array([0, 1, 2]), np.array([0.1, 0.2, 0.3], np.float32))


def test_golden_ncc(vu, golden_ncc_aurc):
    from diffuncertainty_b200 import ncc as vncc
    g = golden_ncc_aurc
    np.testing.assert_allclose(vncc.compute_ncc(g["ncc/gt_map"], g["ncc/pred"]), g["ncc/value"], rtol=RTOL)
    np.testing.assert_allclose(vncc.compute_ncc(g["ncc/pred"], g["ncc/pred"]), g["ncc/self"], rtol=RTOL)
    assert vncc.compute_ncc(np.zeros_like(g["ncc/gt_map"]), g["ncc/pred"]) == 0.0
    assert vncc.compute_ncc(g["ncc/gt_map"], np.full(g["ncc/pred"].shape, 0.25, np.float32)) == 0.0


# --------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties
# --------------------------------------------------------------------------
@pytest.mark.parametrize("P,B,C,spatial", [
    (10, 4, 2, (256, 256)), (5, 4, 2, (64, 64, 64)), (10, 1, 19, (1024, 2048)), (32, 8, 2, (128, 128)),
    (16, 2, 19, (512, 1024)),
])
def test_full_size_properties(vu, P, B, C, spatial):
    from diffuncertainty_b200 import _lib, synth
    x = synth.synth_slab(P, B, C, spatial, seed=1, scale=3.0)
    torch.testing.assert_close(x.sum(dim=2), torch.ones_like(x[:, :, 0]), rtol=0, atol=1e-5)
    res = vu.fused_pass(x, stats=_lib.STAT_IMAGE_SUM | _lib.STAT_AREA)
    tu, au, eu = (res.maps[k] for k in ("TU", "AU", "EU"))
    assert float(tu.max()) <= np.log(C) * (1 + 1e-6) and float(au.min()) >= 0.0
    assert float(eu.min()) >= -2e-6  # Jensen: EU >= 0 up to rounding
    torch.testing.assert_close(au + eu, tu, rtol=3e-7, atol=1e-7)
    # image-level sums are the sums of the maps; area is the count of non-zero labels
    np.testing.assert_allclose(res.image_level(mean=False)[:, 0], tu.double().flatten(1).sum(1).cpu().numpy(), rtol=1e-7)
    assert np.array_equal(res.area(), (res.labels > 0).flatten(1).sum(1).cpu().numpy())
    # voxel-permutation equivariance: the result of a voxel does not depend on its position / tile
    perm = torch.randperm(x[0, 0, 0].numel(), device="cuda")
    xp = x.flatten(3)[..., perm].contiguous()
    rp = vu.fused_pass(xp)
    assert torch.equal(rp.maps["TU"], tu.flatten(1)[:, perm]) and torch.equal(rp.labels, res.labels.flatten(1)[:, perm])
    # one image checked voxel by voxel against the oracle on a 64k-voxel window
    flat = x.flatten(3)[:, 0, :, : 1 << 16].cpu()
    ref, label = oracle_image(flat)
    got = {k: res.maps[k].flatten(1)[0, : 1 << 16].cpu().numpy() for k in ("TU", "AU", "EU")}
    assert_maps_close(got, ref, "full-size window")
    assert np.array_equal(res.labels.flatten(1)[0, : 1 << 16].cpu().numpy(), label)


def test_all_fast_variants_agree(vu):
    """Every tuning variant of the fast kernel must give bit-identical maps."""
    from diffuncertainty_b200 import _lib, synth
    n = _lib.get_counter("k1_num_variants")
    assert n > 0
    try:
        for C, P in ((2, 10), (2, 32), (19, 10), (19, 32), (3, 7), (4, 20)):
            x = synth.synth_slab(P, 2, C, (64, 128), seed=3, scale=4.0)
            _lib.set_option("k1_variant", -2)  # generic kernel as the baseline
            base = vu.fused_pass(x)
            levels = 1 if P <= 17 else 2
            tested = 0
            for i in range(n):
                d = [_lib.get_counter(f"k1_variant.{i}.{f}") for f in range(7)]
                if d[0] != C or d[2] < levels:
                    continue
                _lib.set_option("k1_variant", i)
                r = vu.fused_pass(x)
                for k in ("TU", "AU", "EU"):
                    assert torch.equal(r.maps[k], base.maps[k]), (i, d, k)
                assert torch.equal(r.labels, base.labels), (i, d)
                tested += 1
            assert tested >= 2
    finally:
        _lib.set_option("k1_variant", -1)


def test_tma_variants_agree_with_the_other_kernels(vu):
    """The TMA-pipelined kernel (producer / consumer / statistics warps) must give bit-identical maps and labels, exact
    integer statistics and matching float statistics for every variant, ring depth, partial tiles and both GT dtypes."""
    from diffuncertainty_b200 import _lib, calibration, synth
    n = _lib.get_counter("k1_num_tma_variants")
    assert n >= 4
    platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    thr = [0.3, 0.2, 0.02]
    try:
        for C, P, spatial, R in ((2, 10, (96, 1000), 4), (2, 32, (64, 132), 2), (19, 10, (40, 500), 5), (19, 20, (33, 128), 1),
                                 (3, 7, (64, 128), 2), (4, 20, (64, 128), 1)):
            x = synth.synth_slab(P, 3, C, spatial, seed=11, scale=5.0)
            gt8 = synth.synth_gt(x, R, seed=11, flip=0.3, ignore_frac=0.05, ignore_value=255)
            for gt_t, ign in ((gt8, 255), (gt8.long(), 255)):
                gt = vu.GroundTruth(gt_t, ign)
                _lib.set_option("k1_path", 1)        # register-streaming / generic kernels as the baseline
                _lib.set_option("k1_variant", -2)
                base = vu.fused_pass(x)
                _lib.set_option("k1_variant", -1)
                tested = 0
                for i in range(n):
                    for stages in (0, 2):
                        for flags in (0, 0x07, 0x1f, 0x3f):
                            _lib.set_option("k1_path", 1)
                            ref = vu.fused_pass(x, gt if flags & 0x38 else None, stats=flags, thresholds=thr, calib=platt if flags & 0x10 else None)
                            _lib.set_option("k1_path", 2)
                            _lib.set_option("k1_tma_variant", i)
                            _lib.set_option("k1_tma_stages", stages)
                            try:
                                r = vu.fused_pass(x, gt if flags & 0x38 else None, stats=flags, thresholds=thr, calib=platt if flags & 0x10 else None)
                            except NotImplementedError:
                                break  # variant is for another class count / cascade depth
                            for k in ("TU", "AU", "EU"):
                                assert torch.equal(r.maps[k], base.maps[k]), (C, P, i, stages, flags, k)
                            assert torch.equal(r.labels, base.labels), (C, P, i, stages, flags)
                            if flags:
                                assert torch.equal(r.stats_i64, ref.stats_i64), (C, P, i, stages, flags)
                                torch.testing.assert_close(r.stats_f64, ref.stats_f64, rtol=1e-6, atol=1e-9)
                            tested += 1
                assert tested >= 8, (C, P, tested)
    finally:
        _lib.set_option("k1_path", 0)
        _lib.set_option("k1_variant", -1)
        _lib.set_option("k1_tma_variant", -1)
        _lib.set_option("k1_tma_stages", 0)


def test_every_visible_device(vu):
    """The library shares torch's CUDA runtime: a slab on cuda:1 must be processed on cuda:1."""
    from diffuncertainty_b200 import _lib, synth
    if torch.cuda.device_count() < 2:
        pytest.skip("single-GPU box")
    x0 = synth.synth_slab(6, 2, 19, (64, 128), seed=5, scale=3.0, device="cuda:0")
    want = vu.fused_pass(x0, stats=_lib.STAT_IMAGE_SUM)
    for d in range(1, torch.cuda.device_count()):
        with torch.cuda.device(d):
            xd = x0.to(f"cuda:{d}")
            got = vu.fused_pass(xd, stats=_lib.STAT_IMAGE_SUM)
            assert got.maps["TU"].device.index == d
            assert torch.equal(got.maps["TU"].cpu(), want.maps["TU"].cpu()) and torch.equal(got.labels.cpu(), want.labels.cpu())
            torch.testing.assert_close(got.stats_f64.cpu(), want.stats_f64.cpu(), rtol=1e-12, atol=0)


def test_golden_platt_fit(vu):
    """The 256-bin Platt-fit data (ace.py:14-285) from stored maps (vu_map_stats) and fused into the streaming pass:
    counts bit-exact, per-bin sums and the fitted (a, b) within tolerance of what the unmodified reference produced."""
    import os
    from conftest import GOLDEN_DIR
    from diffuncertainty_b200 import _lib, calibration
    from oracle import oracle
    with np.load(os.path.join(GOLDEN_DIR, "platt_fit.npz")) as z:
        g = {k: z[k] for k in z.files}
    for name in case_names(g):
        ign = int(g[f"{name}/ignore"])
        ign = None if ign == -999 else ign
        n_img = int(g[f"{name}/n_img"])
        acc = calibration.PlattFitAccumulator()
        want = [[np.zeros(256, np.int64), np.zeros(256, np.int64), np.zeros(256, np.int64), np.zeros(256)] for _ in range(3)]
        for i in range(n_img):
            maps = [g[f"{name}/{u}{i}"] for u in ("TU", "AU", "EU")]
            acc.accumulate_maps(g[f"{name}/refs{i}"], g[f"{name}/pred{i}"], maps, ign)
            for k in range(3):
                for dst, src in zip(want[k], oracle.platt_fit_histogram(g[f"{name}/refs{i}"], g[f"{name}/pred{i}"], maps[k], ign)):
                    dst += src
        total, pos, neg, sums = acc.histograms()
        for k, u in enumerate(("TU", "AU", "EU")):
            assert np.array_equal(total[k], want[k][0]) and np.array_equal(pos[k], want[k][1]) and np.array_equal(neg[k], want[k][2]), (name, u)
            inside = slice(1, 255)  # the clamped end bins hold out-of-range values whose sum is not representable in the bin's fixed point
            np.testing.assert_allclose(sums[k][inside], want[k][3][inside], rtol=2e-6, atol=1e-30)
            a, b = acc.fit(k)
            np.testing.assert_allclose([a, b], [g[f"{name}/{u}_a"], g[f"{name}/{u}_b"]], rtol=2e-4)

    # drop-in signature on an in-memory loader
    import pathlib, tempfile, types, json
    name = "lidc_like"
    n_img = int(g[f"{name}/n_img"])
    with tempfile.TemporaryDirectory() as td:
        loader = types.SimpleNamespace(
            exp_version=types.SimpleNamespace(unc_types=["TU", "AU", "EU"], exp_path=pathlib.Path(td)), image_ids=list(range(n_img)),
            get_reference_segs=lambda i: g[f"{name}/refs{i}"], get_mean_pred_seg=lambda i: g[f"{name}/pred{i}"],
            get_unc_map=lambda i, unc: g[f"{name}/{unc}{i}"])
        params = calibration.platt_scale_params(loader, ignore_value=None)
        assert json.load(open(pathlib.Path(td) / "platt_scale_params.json")) == params
    for u in ("TU", "AU", "EU"):
        np.testing.assert_allclose([params[u]["a"], params[u]["b"]], [g[f"{name}/{u}_a"], g[f"{name}/{u}_b"]], rtol=2e-4)


def test_platt_fit_fused_equals_stored_maps(vu):
    """STAT_PLATT_FIT fused into the streaming pass (both K1 forms) == the same statistic from the maps it wrote."""
    from diffuncertainty_b200 import _lib, calibration, synth
    try:
        for C, P, spatial, R in ((19, 10, (40, 500), 2), (2, 10, (64, 512), 4)):
            x = synth.synth_slab(P, 2, C, spatial, seed=3, scale=6.0)
            gt = synth.synth_gt(x, R, seed=3, flip=0.3, ignore_frac=0.05, ignore_value=255)
            ref = None
            for path in (1, 2):
                _lib.set_option("k1_path", path)
                acc = calibration.PlattFitAccumulator()
                res = vu.fused_pass(x, vu.GroundTruth(gt, 255), stats=_lib.STAT_PLATT_FIT | _lib.STAT_AREA, platt_fit=acc)
                stored = calibration.PlattFitAccumulator()
                for b in range(2):
                    stored.accumulate_maps(gt[b], res.labels[b], [res.maps[k][b] for k in ("TU", "AU", "EU")], 255)
                for got, want in zip(acc.histograms()[:3], stored.histograms()[:3]):
                    assert np.array_equal(got, want), (C, path)
                np.testing.assert_allclose(acc.histograms()[3], stored.histograms()[3], rtol=1e-9, atol=1e-30)
                if ref is not None:
                    assert np.array_equal(acc.histograms()[0], ref.histograms()[0])
                ref = acc
    finally:
        _lib.set_option("k1_path", 0)


def test_member_labels_match_torch_argmax(vu):
    """Per-member labels (save_prediction writes one per member, test_2D.py:810-818): bit-exact with torch.argmax on every
    kernel form, including ties (one-hot members, --discretize) and NaN (counts as the maximum)."""
    from diffuncertainty_b200 import _lib
    g = torch.Generator().manual_seed(77)
    try:
        for P, B, C, spatial in ((10, 2, 19, (40, 500)), (5, 3, 2, (16, 16, 64)), (20, 1, 19, (8, 512)), (7, 2, 5, (9, 7)), (1, 2, 3, (8, 16))):
            x = torch.softmax(3 * torch.randn(P, B, C, *spatial, generator=g), dim=2)
            x[0] = torch.nn.functional.one_hot(x[0].argmax(1), C).movedim(-1, 1).float()       # exact ties at 0 and a unique 1
            x[-1, :, :, ..., :3] = 0.25                                                             # all classes tie
            if C > 2:
                x[P // 2, 0, 1, ..., 5] = float("nan")
            want = x.argmax(dim=2).to(torch.uint8)
            for path in (0, 1, 2):
                _lib.set_option("k1_path", path)
                try:
                    r = vu.fused_pass(x.cuda(), want_member_labels=True)
                except NotImplementedError:
                    continue  # shape not eligible for the TMA form
                assert r.member_labels.shape == want.shape
                assert torch.equal(r.member_labels.cpu(), want), (P, C, path)
            _lib.set_option("k1_path", 1)
            _lib.set_option("k1_variant", -2)   # generic kernel
            assert torch.equal(vu.fused_pass(x.cuda(), want_member_labels=True).member_labels.cpu(), want), (P, C, "generic")
            _lib.set_option("k1_variant", -1)
    finally:
        _lib.set_option("k1_path", 0)
        _lib.set_option("k1_variant", -1)


def test_member_list_is_read_without_stacking(vu):
    """A list of P member tensors (the groups of test_2D.py:1134-1136 before torch.stack, :1277) gives bit-identical
    results to the stacked slab on every kernel form; members may live anywhere (different allocations, views)."""
    from diffuncertainty_b200 import _lib
    g = torch.Generator().manual_seed(5)
    try:
        for P, B, C, spatial in ((10, 2, 19, (40, 500)), (5, 2, 2, (16, 16, 64)), (6, 2, 5, (9, 7))):
            members = []
            for p in range(P):
                pad = torch.empty(1024 * (p % 3 + 1), device="cuda")  # scatter the allocations
                big = torch.softmax(3 * torch.randn(B + 1, C, *spatial, generator=g), dim=1).cuda()
                members.append(big[1:] if p % 2 else big[:B])               # views into larger tensors
                del pad
            gt = torch.randint(0, C, (B, 2, *spatial), generator=g).to(torch.uint8).cuda()
            stacked = torch.stack([m.contiguous() for m in members])
            for path in (0, 1, 2):
                _lib.set_option("k1_path", path)
                try:
                    want = vu.fused_pass(stacked, vu.GroundTruth(gt), stats=0x0f, thresholds=[0.3, 0.2, 0.02], want_member_labels=True)
                    got = vu.fused_pass(members, vu.GroundTruth(gt), stats=0x0f, thresholds=[0.3, 0.2, 0.02], want_member_labels=True)
                except NotImplementedError:
                    continue
                for k in ("TU", "AU", "EU"):
                    assert torch.equal(got.maps[k], want.maps[k]), (P, C, path, k)
                assert torch.equal(got.labels, want.labels) and torch.equal(got.member_labels, want.member_labels)
                assert torch.equal(got.stats_i64, want.stats_i64)
                torch.testing.assert_close(got.stats_f64, want.stats_f64, rtol=1e-9, atol=1e-12)
    finally:
        _lib.set_option("k1_path", 0)


def test_entry_points_can_be_captured_in_a_cuda_graph(vu):
    """The C ABI only enqueues work on the stream it is given (no hidden synchronisation, allocation or host read-back), so a
    launch-bound inner loop -- many small batches through vu_fused_pass + vu_patch_max_ws -- can be captured once and replayed."""
    from diffuncertainty_b200 import _lib, aggregation, calibration, synth
    P, B, C, spatial, R = 10, 4, 2, (64, 64), 3
    x = synth.synth_slab(P, B, C, spatial, seed=5, scale=3.0)
    gt = vu.GroundTruth(synth.synth_gt(x, R, seed=5, flip=0.2), None)
    flags = _lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CALIB | _lib.STAT_NCC
    calib = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    maps = {k: torch.empty((B,) + spatial, dtype=torch.float32, device="cuda") for k in ("TU", "AU", "EU")}
    labels = torch.empty((B,) + spatial, dtype=torch.uint8, device="cuda")
    sf = torch.zeros((B, 80), dtype=torch.float64, device="cuda")
    si = torch.zeros((B, 156), dtype=torch.int64, device="cuda")
    lib = _lib.load()
    out_max = torch.empty(B, dtype=torch.float64, device="cuda")
    out_first = torch.empty(B, dtype=torch.int64, device="cuda")
    ws_bytes = int(lib.vu_patch_workspace_bytes(B, 1, *spatial, 1, 10, 10))
    ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.int64, device="cuda")

    def work():
        sf.zero_(); si.zero_()
        vu.fused_pass(x, gt, stats=flags, thresholds=[0.3, 0.2, 0.02], calib=calib, stats_out=(sf, si), maps_out=maps, labels_out=labels)
        _lib.check(lib.vu_patch_max_ws(maps["TU"].data_ptr(), B, 1, *spatial, 1, 10, 10, 0, out_max.data_ptr(), out_first.data_ptr(),
                                       ws.data_ptr(), ws_bytes, _lib.current_stream_ptr()), "vu_patch_max_ws")

    work()  # eager: also warms up the per-kernel attribute calls
    torch.cuda.synchronize()
    want = [t.clone() for t in (maps["TU"], maps["EU"], labels, sf, si, out_max, out_first)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            work()
    for t in (maps["TU"], maps["EU"], labels, sf, si, out_max, out_first):
        t.zero_()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    for got, ref in zip((maps["TU"], maps["EU"], labels, sf, si, out_max, out_first), want):
        if got.dtype == torch.float64:  # sums folded with atomics: the order of the partials differs from run to run
            torch.testing.assert_close(got, ref, rtol=1e-12, atol=1e-300)
        else:
            assert torch.equal(got, ref)


def test_border_and_area_random_shapes(vu):
    """_compute_area / _compute_border (prediction_shape_stats.py:10-30) on random label maps: the word-wise kernel (rows of a
    multiple of 4 labels) and the scalar one, 1- to 3-D, several classes, batched through vu_border_count."""
    from diffuncertainty_b200 import _lib, aggregation as agg
    from oracle import oracle
    rng = np.random.default_rng(9)
    for shape in ((64, 64), (33, 40), (40, 33), (7, 9, 12), (5, 16, 20), (1, 8), (17,), (4, 4, 4), (3, 64, 128)):
        for n_cls in (2, 5):
            lab = rng.integers(0, n_cls, shape).astype(np.uint8)
            blob = tuple(slice(s // 4, s // 4 + max(1, s // 2)) for s in shape)
            lab[blob] = 1  # a compact region, so that not every neighbour pair differs
            area, border = agg.prediction_shape_stats(lab)
            assert area == oracle.compute_area(lab) and border == oracle.compute_border(lab), (shape, n_cls)
    lib = _lib.load()
    B, dims = 5, (6, 20, 24)
    labs = rng.integers(0, 3, (B,) + dims).astype(np.uint8)
    si = torch.zeros((B, _lib.I64["COLS"]), dtype=torch.int64, device="cuda")
    t = torch.from_numpy(labs).cuda()
    _lib.check(lib.vu_border_count(t.data_ptr(), B, *dims, si.data_ptr(), _lib.current_stream_ptr()), "vu_border_count")
    got = si[:, _lib.I64["BORDER"]].cpu().numpy()
    assert np.array_equal(got, [int(oracle.compute_border(l)) for l in labs])


# --------------------------------------------------------------------------
# statistics at one full BASELINE size per config (labels / counts bit-exact, sums 1e-5), one image each
# --------------------------------------------------------------------------
@pytest.mark.parametrize("name,P,C,spatial,R,ignore,flags", [
    ("cfg1", 10, 2, (256, 256), 0, None, 0x07),
    ("cfg2", 5, 2, (64, 64, 64), 4, None, 0x1d),
    ("cfg3", 10, 19, (1024, 2048), 5, 255, 0x0f),
    ("cfg4", 32, 2, (128, 128), 4, None, 0x21),
    ("cfg5", 16, 19, (512, 1024), 1, 255, 0x1f),
])
def test_full_size_statistics_vs_oracle(vu, name, P, C, spatial, R, ignore, flags):
    import os
    from diffuncertainty_b200 import _lib, calibration, ncc as vncc, synth
    from oracle import oracle
    x = synth.synth_slab(P, 1, C, spatial, seed=17, scale=3.0)
    gt_t = synth.synth_gt(x, max(R, 1), seed=17, flip=0.2, ignore_frac=0.02 if ignore is not None else 0.0,
                          ignore_value=255 if ignore is None else ignore)
    platt = [(3.5, -1.25), (6.0, -2.0), (40.0, -0.5)]
    thr = [0.3, 0.2, 0.02]
    res = vu.fused_pass(x, vu.GroundTruth(gt_t, ignore) if R else None, stats=flags, thresholds=thr,
                        calib=[calibration.platt_edges(a, b) for a, b in platt] if flags & _lib.STAT_CALIB else None)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    xc = x[:, 0].cpu()
    ref = oracle.calculate_uncertainty(xc)
    label = oracle.argmax_first_nan_max(oracle.mean_members_f32(xc.numpy())).astype(np.uint8)
    assert np.array_equal(res.labels[0].cpu().numpy(), label)
    got = {k: res.maps[k][0].cpu().numpy() for k in ("TU", "AU", "EU")}
    assert_maps_close(got, {k: v.numpy() for k, v in ref.items()}, name)
    gnp = gt_t[0].cpu().numpy()
    if flags & _lib.STAT_AREA:
        assert res.area()[0] == oracle.compute_area(label)
    if flags & _lib.STAT_DICE:
        tp, ps, gs = res.dice_counts()
        otp, ops, ogs = oracle.binary_dice_counts(label, gnp, -12345 if ignore is None else ignore)
        assert np.array_equal(tp[0], otp) and np.array_equal(ps[0], ops) and np.array_equal(gs[0], ogs)
    for k, key in enumerate(("TU", "AU", "EU")):
        m = got[key]
        if flags & _lib.STAT_IMAGE_SUM:
            np.testing.assert_allclose(res.image_level()[0, k], oracle.image_level_aggregation(m)["max_score"], rtol=RTOL, atol=1e-9)
        if flags & _lib.STAT_THRESHOLD:
            np.testing.assert_allclose(res.threshold_level()[0, k], float(oracle.threshold_aggregation(m, np.float32(thr[k]))["max_score"]), rtol=RTOL)
            assert int(res.stats_i64[0, _lib.I64["THR_COUNT"] + k]) == int((m >= np.float32(thr[k])).sum())
        if flags & _lib.STAT_CALIB:
            bs, bt, bn = res.calib_histograms()
            correct, conf = oracle.calibration_inputs(gnp, label, m, platt[k][0], platt[k][1], ignore)
            s, t, n = oracle.calib_histogram(correct, conf, binarize=False)
            assert np.array_equal(bn[0, k], n) and np.array_equal(bt[0, k], t.astype(np.int64)), key
            np.testing.assert_allclose(bs[0, k], s, rtol=RTOL, atol=1e-9)
        if flags & _lib.STAT_NCC:
            # NCC from the kernel's sums against the reference formula on the kernel's OWN map: 1e-5.  (Against the
            # reference's map the bound is the map error times the conditioning of the covariance)
            # NCC is a correlation in [-1, 1] formed from a covariance that cancels (synthetic references: ~3e-4 here), so the
            # error is bounded absolutely: 1e-5 relative or 1e-7 absolute, whichever is larger.
            want = oracle.compute_ncc(oracle.rater_variance_map(gnp), m)
            np.testing.assert_allclose(vncc.ncc_from_result(res, k)[0], want, rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("P,B,C,spatial", [(10, 2, 5, (16, 64)), (5, 1, 7, (8, 40)), (16, 1, 21, (8, 32)), (18, 2, 6, (4, 64)),
                                           (32, 1, 3, (3, 37)), (17, 1, 33, (5, 24)), (2, 3, 255, (2, 16))])
def test_classouter_kernel_equals_generic_and_oracle(vu, P, B, C, spatial):
    """Class counts without a compiled-in form (anything but 2, 3, 4, 19) take the class-outer kernel (k1_classouter: run-time
    C, members unrolled); it must give the bits of the generic kernel -- maps, labels, per-member labels, statistics -- and
    the oracle's labels."""
    from diffuncertainty_b200 import _lib
    from oracle import oracle
    g = torch.Generator().manual_seed(P * 100 + C)
    x = torch.softmax(3.0 * torch.randn(P, B, C, *spatial, generator=g), dim=2)
    x[0, 0, :, 0, :3] = 0.0
    x[1, 0, 0, 0, 5] = float("nan")
    gt = torch.randint(0, min(C, 250), (B, 2, *spatial), generator=g, dtype=torch.uint8)
    flags = _lib.STAT_IMAGE_SUM | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CLASS_COUNTS
    aligned = int(np.prod(spatial)) % 4 == 0 and C not in (2, 3, 4, 19)
    # the direct-load form (option k1_path = 1 keeps the TMA forms out)
    before = _lib.get_counter("launches.k1_classouter")
    _lib.load().vu_set_option(b"k1_path", 1)
    try:
        co = vu.fused_pass(x.cuda(), vu.GroundTruth(gt.cuda(), None), stats=flags)
    finally:
        _lib.load().vu_set_option(b"k1_path", 0)
    assert _lib.get_counter("launches.k1_classouter") == before + 1 or C in (2, 3, 4, 19)
    # with 16-byte aligned rows the rows go through the TMA ring (k1_co_tma) -- with statistics, without, and from a member
    # list: same bits
    before_tma = _lib.get_counter("launches.k1_co_tma")
    tma = vu.fused_pass(x.cuda(), vu.GroundTruth(gt.cuda(), None), stats=flags)
    plain = vu.fused_pass(x.cuda())
    lst = vu.fused_pass([x[p].cuda() for p in range(P)])
    if aligned:
        assert _lib.get_counter("launches.k1_co_tma") == before_tma + 3
    for r in (tma, plain, lst):
        assert torch.equal(r.labels, co.labels)
        for k in ("TU", "AU", "EU"):
            assert torch.equal(r.maps[k].view(torch.int32), co.maps[k].view(torch.int32)), k
    assert torch.equal(tma.stats_i64, co.stats_i64) and torch.equal(tma.class_counts, co.class_counts)
    np.testing.assert_allclose(tma.stats_f64.cpu().numpy(), co.stats_f64.cpu().numpy(), rtol=1e-7)
    _lib.load().vu_set_option(b"k1_variant", -2)
    try:
        gen = vu.fused_pass(x.cuda(), vu.GroundTruth(gt.cuda(), None), stats=flags, want_member_labels=True)
    finally:
        _lib.load().vu_set_option(b"k1_variant", -1)
    assert torch.equal(co.labels, gen.labels)
    for k in ("TU", "AU", "EU"):
        assert torch.equal(co.maps[k].view(torch.int32), gen.maps[k].view(torch.int32)), k
    assert torch.equal(co.stats_i64, gen.stats_i64) and torch.equal(co.class_counts, gen.class_counts)
    np.testing.assert_allclose(co.stats_f64.cpu().numpy(), gen.stats_f64.cpu().numpy(), rtol=1e-7)
    torch.set_num_threads(1)
    for b in range(B):
        label = oracle.argmax_first_nan_max(oracle.mean_members_f32(x[:, b].numpy())).astype(np.uint8)
        assert np.array_equal(co.labels[b].cpu().numpy(), label)
        assert torch.equal(gen.member_labels[:, b].cpu(), x[:, b].argmax(dim=1).to(torch.uint8))
