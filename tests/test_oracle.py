"""Pin oracle/oracle.py (the CPU restatement) against the golden vectors that
were recorded from the unmodified reference, and -- when /root/reference is
mounted -- against the reference functions themselves on fresh inputs."""
import json

import numpy as np
import pytest
import torch

from oracle import oracle, ref_shim
from conftest import case_names

RTOL = 1e-5  # north_star: floating-point maps and scores within 1e-5 relative


def same_bits(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def test_uncertainty_matches_golden_bitwise(golden_unc):
    torch.set_num_threads(1)
    for name in case_names(golden_unc):
        x = torch.from_numpy(golden_unc[f"{name}/x"])
        if name == "msr_p1":
            got = oracle.calculate_one_minus_msr(x.squeeze(0))["pred_entropy"].numpy()
            assert same_bits(got, golden_unc[f"{name}/pred_entropy"])
            continue
        got = oracle.calculate_uncertainty(x)
        for key in ("TU", "AU", "EU"):
            assert got[key].dtype == torch.float32
            assert same_bits(got[key].numpy(), golden_unc[f"{name}/{key}"]), (name, key)
        mean, label = oracle.mean_and_label(x)
        assert same_bits(label.numpy(), golden_unc[f"{name}/label"]), name


def test_canonical_mean_matches_torch_outside_simd_tail(golden_unc):
    """cascade_sum_f32 is bit-identical to torch.mean except in the last
    numel % 32 elements of the reduction row, where torch uses the interleaved
    order; there it must equal interleaved_tail_sum_f32."""
    for name in case_names(golden_unc):
        if name == "msr_p1":
            continue
        x = golden_unc[f"{name}/x"]
        ref = golden_unc[f"{name}/mean"].reshape(-1)
        P = x.shape[0]
        flat = x.reshape(P, -1)
        canon = (oracle.cascade_sum_f32(flat) / np.float32(P))
        tail = (oracle.interleaved_tail_sum_f32(flat) / np.float32(P))
        n = flat.shape[1]
        body = (n // 32) * 32
        cb, rb = canon.view(np.uint32), ref.view(np.uint32)
        nan_ok = np.isnan(canon) & np.isnan(ref)
        assert np.all((cb[:body] == rb[:body]) | nan_ok[:body]), name
        tb = tail.view(np.uint32)
        assert np.all((tb[body:] == rb[body:]) | nan_ok[body:]), name


@pytest.mark.parametrize("P", [2, 5, 10, 16, 17, 18, 31, 32, 33, 48, 100, 255, 256, 257, 300])
def test_cascade_sum_vs_torch_live(P):
    torch.manual_seed(P)
    torch.set_num_threads(1)
    x = torch.softmax(2 * torch.randn(P, 4, 8, 16), dim=1)  # 512 elements: no SIMD tail
    ref = torch.mean(x, dim=0).numpy()
    assert same_bits(oracle.mean_members_f32(x.numpy()), ref)
    xb = torch.softmax(2 * torch.randn(P, 3, 4, 8, 16), dim=2)[:, 1]  # strided per-image view
    assert same_bits(oracle.mean_members_f32(xb.numpy()), torch.mean(xb, dim=0).numpy())


def test_argmax_rule():
    m = np.array([[0.5, np.nan, 0.2, 1.0], [0.5, 0.1, np.nan, 1.0], [0.1, np.nan, 0.9, 1.0]], np.float32)
    ref = torch.from_numpy(m).argmax(dim=0).numpy()
    assert np.array_equal(oracle.argmax_first_nan_max(m), ref)
    assert ref.tolist() == [0, 0, 1, 0]


def test_aggregation_matches_golden(golden_agg):
    for name in ("img2d", "img3d", "plateau"):
        img = golden_agg[f"{name}/image"]
        assert oracle.image_level_aggregation(img)["max_score"] == golden_agg[f"{name}/image_level_mean"]
        assert oracle.image_level_aggregation(img, mean=False)["max_score"] == golden_agg[f"{name}/image_level_sum"]
        for ps in (10, 4):
            r = oracle.patch_level_aggregation(img, ps)
            assert r["max_score"] == golden_agg[f"{name}/patch{ps}_score"]
            assert np.array_equal(np.asarray(r["bounding_box"]), golden_agg[f"{name}/patch{ps}_bbox"])
            assert oracle.patch_level_aggregation(img, ps, mean=True)["max_score"] == \
                golden_agg[f"{name}/patch{ps}_mean_score"]
            direct = oracle.box_sum_direct_f64(img, img.ndim * [ps])
            # scipy runs the FFT in float32 for float32 images: ~5e-8 relative noise
            np.testing.assert_allclose(direct.max(), golden_agg[f"{name}/patch{ps}_score"], rtol=1e-6)
        for t in ("mid", "above_max", "zero"):
            thr = float(golden_agg[f"{name}/thr_{t}_t"])
            assert float(oracle.threshold_aggregation(img, thr)["max_score"]) == golden_agg[f"{name}/thr_{t}_score"]
            assert float(oracle.threshold_aggregation(img, thr, mean=False)["max_score"]) == \
                golden_agg[f"{name}/thr_{t}_sum"]
    for name in ("lab2d", "lab3d", "empty"):
        lab = golden_agg[f"{name}/label"]
        assert oracle.compute_area(lab) == golden_agg[f"{name}/area"]
        assert oracle.compute_border(lab) == golden_agg[f"{name}/border"]
        assert oracle.normalized_sum(golden_agg["img2d/image"], oracle.compute_area(lab)) == \
            golden_agg[f"{name}/norm_area"]


def test_calibration_matches_golden(golden_calib):
    for name in case_names(golden_calib):
        g = {k.split("/", 1)[1]: v for k, v in golden_calib.items() if k.startswith(name + "/")}
        ignore = None if int(g["ignore"]) == -999 else int(g["ignore"])
        correct, conf = oracle.calibration_inputs(g["refs"], g["pred"], g["unc"], float(g["a"]), float(g["b"]), ignore)
        assert conf.dtype == np.float32
        assert same_bits(conf, g["conf"]), name
        assert np.array_equal(correct.astype(np.uint8), g["correct"])
        assert oracle.calc_ace(correct, conf) == g["ace"], name
        assert oracle.calc_ece(correct, conf) == g["ece"], name
        acc = oracle.GlobalCalibAccumulator()
        acc.accumulate(correct, conf)
        assert np.array_equal(acc.bin_total, g["g_bin_total"])
        assert np.array_equal(acc.bin_true, g["g_bin_true"])
        assert np.array_equal(acc.bin_sums, g["g_bin_sums"])
        assert acc.compute_ace() == g["gace"] and acc.compute_ece() == g["gece"]
    # quirk Q8: an all-correct image is scored as if nothing were correct
    g = golden_calib
    assert g["all_correct/correct"].min() == 1
    s, t, n = oracle.calib_histogram(g["all_correct/correct"], g["all_correct/conf"], binarize=True)
    assert t.sum() == 0 and n.sum() == g["all_correct/correct"].size
    assert n[20] == 0 and len(n) == 21  # Q7: 21 slots, the last always empty


def test_ncc_aurc_match_golden(golden_ncc_aurc):
    g = golden_ncc_aurc
    assert np.array_equal(oracle.rater_variance_map(g["ncc/refs"]), g["ncc/gt_map"])
    assert oracle.compute_ncc(g["ncc/gt_map"], g["ncc/pred"]) == g["ncc/value"]
    assert oracle.compute_ncc(g["ncc/pred"], g["ncc/pred"]) == g["ncc/self"]
    n = g["ncc/pred"].size
    np.testing.assert_allclose(g["ncc/self"], (n - 1) / n, rtol=1e-6)  # quirk Q10
    assert oracle.compute_ncc(np.zeros_like(g["ncc/gt_map"]), g["ncc/pred"]) == 0.0 == g["ncc/const_gt"]
    assert oracle.compute_ncc(g["ncc/gt_map"], np.full(g["ncc/pred"].shape, 0.25, np.float32)) == g["ncc/const_pred"]
    assert oracle.aurc(g["aurc/risks"], g["aurc/confids"]) == g["aurc/aurc"]
    assert oracle.eaurc(g["aurc/risks"], g["aurc/confids"]) == g["aurc/eaurc"]
    cov, sel, w = oracle.rc_curve_stats(g["aurc/risks"], g["aurc/confids"])
    assert np.array_equal(np.asarray(sel), g["aurc/selective_risks"])
    assert np.array_equal(np.asarray(w), g["aurc/weights"])


def test_binary_dice_rule():
    label = np.array([[1, 1, 0, 0]], np.uint8)
    gt = np.array([[[1, 0, 0, 0]], [[0, 0, 0, 0]], [[1, 1, 255, 0]]], np.int64)
    tp, ps, gs = oracle.binary_dice_counts(label, gt, 255)
    assert tp.tolist() == [1, 0, 2] and ps.tolist() == [2, 2, 2] and gs.tolist() == [1, 0, 2]
    assert oracle.binary_dice_from_counts(tp, ps, gs) == pytest.approx((2 / 3 + 0 + 1) / 3, rel=1e-6)
    assert oracle.binary_dice_from_counts([0], [0], [0]) == 1.0


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_vs_live_reference():
    ref = ref_shim.load()
    torch.manual_seed(5)
    torch.set_num_threads(1)
    rng = np.random.default_rng(5)
    for P, C, S in ((10, 2, (32, 32)), (5, 2, (8, 8, 16)), (16, 19, (8, 32)), (33, 3, (16, 16))):
        x = torch.softmax(6 * torch.randn(P, C, *S), dim=1)
        a, b = ref.calculate_uncertainty(x), oracle.calculate_uncertainty(x)
        for k in ("TU", "AU", "EU"):
            assert same_bits(a[k].numpy(), b[k].numpy())
    img = rng.random((48, 64)).astype(np.float32)
    assert ref.patch_level_aggregation(img, 10) == oracle.patch_level_aggregation(img, 10)
    assert ref.image_level_aggregation(img) == oracle.image_level_aggregation(img)
    ra, rb = ref.threshold_aggregation(img, threshold=0.5), oracle.threshold_aggregation(img, 0.5)
    assert float(ra["max_score"]) == float(rb["max_score"])
    correct = rng.integers(0, 2, 5000)
    conf = rng.random(5000).astype(np.float32)
    assert ref.calc_ace(correct, conf) == oracle.calc_ace(correct, conf)
    assert ref.calc_ece(correct, conf) == oracle.calc_ece(correct, conf)
    lab = rng.integers(0, 3, (20, 30)).astype(np.uint8)
    assert ref.compute_border(lab) == oracle.compute_border(lab)
    risks, confids = rng.random(100), rng.random(100)
    assert ref.aurc(risks, confids) == oracle.aurc(risks, confids)
    assert ref.eaurc(risks, confids) == oracle.eaurc(risks, confids)


def test_platt_fit_matches_golden():
    """Oracle restatement of platt_scale_params (ace.py:14-285) against what the unmodified reference built and fitted."""
    import os
    from conftest import GOLDEN_DIR, case_names
    from oracle import oracle
    with np.load(os.path.join(GOLDEN_DIR, "platt_fit.npz")) as z:
        g = {k: z[k] for k in z.files}
    for name in case_names(g):
        ign = int(g[f"{name}/ignore"])
        ign = None if ign == -999 else ign
        for unc in ("TU", "AU", "EU"):
            tot = np.zeros(256, np.int64); pos = tot.copy(); neg = tot.copy(); sums = np.zeros(256)
            for i in range(int(g[f"{name}/n_img"])):
                t, p, n, s = oracle.platt_fit_histogram(g[f"{name}/refs{i}"], g[f"{name}/pred{i}"], g[f"{name}/{unc}{i}"], ign)
                tot += t; pos += p; neg += n; sums += s
            assert np.array_equal(tot, pos + neg)
            F, y, w = oracle.platt_fit_samples(tot, pos, neg, sums)
            assert np.array_equal(y, g[f"{name}/{unc}_y"]) and np.array_equal(w, g[f"{name}/{unc}_w"])
            np.testing.assert_allclose(F, g[f"{name}/{unc}_F"], rtol=1e-13)
            a, b = oracle.platt_fit(tot, pos, neg, sums)
            np.testing.assert_allclose([a, b], [g[f"{name}/{unc}_a"], g[f"{name}/{unc}_b"]], rtol=1e-9)


def test_quantile_restatement_matches_numpy():
    """oracle.quantile_linear against np.quantile itself (the third-party routine behind ace.py:388 and
    find_threshold.py:76), float32 and float64 data, ties, tiny inputs, NaN."""
    from oracle import oracle
    rng = np.random.default_rng(5)
    qs = np.concatenate([np.linspace(0, 1, 21), rng.random(40)])
    for dtype in (np.float32, np.float64):
        for n in (1, 2, 3, 20, 21, 1000, 4097):
            x = (rng.random(n) ** 3).astype(dtype)
            x[rng.random(n) < 0.3] = 0
            assert same_bits(oracle.quantile_linear(x, qs), np.quantile(x, qs))
            assert same_bits(np.asarray(oracle.quantile_linear(x, 0.37)), np.asarray(np.quantile(x, 0.37)))
    x = np.array([0.5, np.nan, 0.25], np.float32)
    assert np.isnan(oracle.quantile_linear(x, [0.0, 0.5])).all() and np.isnan(np.quantile(x, [0.0, 0.5])).all()


def test_eqace_and_thresholds_match_golden(golden_quantile):
    from oracle import oracle
    g = golden_quantile
    for name in g["eqace_cases"]:
        got = oracle.calc_eqace(g[f"{name}/correct"], g[f"{name}/conf"])
        assert got == float(g[f"{name}/eqace"]), name
        ign = int(g[f"{name}/ignore"])
        correct, conf = oracle.calibration_inputs(g[f"{name}/refs"], g[f"{name}/pred"], g[f"{name}/unc"], float(g[f"{name}/a"]),
                                                  float(g[f"{name}/b"]), None if ign == -999 else ign)
        assert oracle.calc_eqace(correct, conf) == float(g[f"{name}/eqace"]), name
    assert np.isnan(oracle.calc_eqace(np.zeros(0, int), np.zeros(0, np.float32))) and np.isnan(g["empty/eqace"])
    maps = [g[f"thr/map{k}"] for k in range(int(g["thr/n_maps"]))]
    for qi, q in enumerate(g["thr/qs"]):
        for k, m in enumerate(maps):
            assert oracle.uncertainty_threshold([m], float(q)) == g["thr/per_map"][qi, k]
        assert oracle.uncertainty_threshold(maps, float(q)) == g["thr/all_maps"][qi]
    # the drivers' JSON files (find_threshold.py:15-47, 80-112)
    ids = list(g["drv/ids"])
    want_q = json.loads(str(g["drv/quantile_analysis.json"]))["Softmax"]
    got_q = oracle.mean_foreground_quantile([p for i in ids for p in g[f"drv/{i}/preds"]])
    assert got_q == want_q
    want_t = json.loads(str(g["drv/threshold_analysis.json"]))["Softmax"]
    for u in ("TU", "AU", "EU"):
        assert oracle.uncertainty_threshold([g[f"drv/{i}/{u}"] for i in ids], got_q) == want_t[f"Mean {u} threshold"]


def test_ged_and_likelihood_match_golden(golden_ged_nll):
    """oracle.ged_binary_fast / compute_likelihood_stats / compute_expected_nll against what the unmodified
    ged_fast.ged_binary_fast and Tester._compute_likelihood_stats / _compute_expected_nll returned (float32 reductions in
    the reference: 2e-6 relative)."""
    from oracle import oracle
    g = golden_ged_nll
    am = ["dice", "max_dice_pred", "max_dice_gt", "major_dice"]
    for name in g["cases"]:
        x, gt, ign = g[f"{name}/x"], g[f"{name}/gt"], int(g[f"{name}/ignore"])
        if f"{name}/ged" in g:
            got = oracle.ged_binary_fast(torch.from_numpy(x), gt, None if ign == -999 else ign, am)
            for k in ["ged"] + am:
                np.testing.assert_allclose(got[k], float(g[f"{name}/{k}"]), rtol=2e-6, atol=2e-7, err_msg=f"{name}/{k}")
        if f"{name}/mean_nll" in g:
            gt_arg = gt[0] if name == "single_rater_2d" else gt
            gt_arg = gt_arg[None] if gt_arg.ndim == x.ndim - 2 else gt_arg
            m, r, mean = oracle.compute_likelihood_stats(x, gt_arg, -1 if ign == -999 else ign)
            np.testing.assert_allclose(np.array(m), g[f"{name}/gt_model_nll"], rtol=2e-6, atol=1e-7, err_msg=name)
            np.testing.assert_allclose(np.array(r), g[f"{name}/gt_nll"], rtol=2e-6, atol=1e-7, err_msg=name)
            np.testing.assert_allclose(mean, float(g[f"{name}/mean_nll"]), rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(oracle.compute_expected_nll(x, gt_arg, -1 if ign == -999 else ign),
                                       float(g[f"{name}/expected_nll"]), rtol=2e-6, atol=1e-7)
    with pytest.raises(ValueError):
        oracle.ged_binary_fast(torch.zeros(3, 3, 4, 4), np.zeros((2, 4, 4), np.int64))


def test_renormalize_canonical_order_matches_torch_and_reference():
    """The class-sum order the kernel uses for _renormalize_probabilities (cascade over the class axis) is torch's, and the
    oracle's expression is the reference's (test_2D.py:188-194)."""
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(3)
    for C, S in ((19, (8, 32)), (2, (16, 32)), (40, (4, 32))):
        p = torch.softmax(3 * torch.randn(2, C, *S, generator=g), dim=1) * (1 + 0.05 * torch.randn(2, 1, *S, generator=g))
        p[0, :, 0, :3] = 0.0
        a = oracle.renormalize_probabilities(p).numpy()
        b = oracle.renormalize_probabilities_canonical(p.numpy())
        assert same_bits(a, b)
        if ref_shim.available():
            ref_shim.load()
            import importlib
            backend = importlib.import_module("uncertainty_modeling.test_2D").AlbumentationsTTABackend
            assert same_bits(backend._renormalize_probabilities(p).numpy(), a)
