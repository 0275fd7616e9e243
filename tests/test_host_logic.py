"""Host-side logic of the product package against the oracle / golden vectors.
Runs without a GPU."""
import numpy as np
import pytest

from diffuncertainty_b200 import aurc as vaurc
from diffuncertainty_b200 import calibration, ncc as vncc
from oracle import oracle
from conftest import case_names


def test_platt_edges_reproduce_np_digitize(golden_calib):
    rng = np.random.default_rng(3)
    for name in case_names(golden_calib):
        a, b = float(golden_calib[f"{name}/a"]), float(golden_calib[f"{name}/b"])
        pe = calibration.platt_edges(a, b)
        u = np.concatenate([golden_calib[f"{name}/unc"].ravel(), (rng.random(20000) ** 3 * 3).astype(np.float32),
                            np.array([0.0, 1e-30, 0.6931472, 3.0, 1e10], np.float32)])
        fin = pe.edge_u[~np.isnan(pe.edge_u)]
        near = np.concatenate([np.nextafter(fin, np.float32(-np.inf)), fin, np.nextafter(fin, np.float32(np.inf))])
        u = np.concatenate([u, near.astype(np.float32)])
        conf = oracle.platt_scale_confid(-u, a, b)
        assert conf.dtype == np.float32
        want = np.digitize(np.clip(conf, 0, 1), oracle.calib_bin_edges()) - 1
        assert np.array_equal(pe.bin_of(u), want), name


def test_identity_edges():
    pe = calibration.identity_edges()
    rng = np.random.default_rng(0)
    conf = np.concatenate([rng.random(50000).astype(np.float32), np.array([0, 1, 0.05, 0.1, 0.95], np.float32),
                           pe.edge_u, np.nextafter(pe.edge_u, np.float32(0))])
    want = np.digitize(conf, oracle.calib_bin_edges()) - 1
    assert np.array_equal(pe.bin_of(conf), want)


def test_ace_ece_finalisers_match_oracle(golden_calib):
    for name in case_names(golden_calib):
        correct, conf = golden_calib[f"{name}/correct"], golden_calib[f"{name}/conf"]
        s, t, n = oracle.calib_histogram(correct, conf, binarize=False)
        ace, ece = calibration.per_image_ace_ece(s, t, n)
        assert ace == golden_calib[f"{name}/ace"], name
        assert ece == golden_calib[f"{name}/ece"], name
        g = calibration.GlobalCalibAccumulator()
        g.accumulate_histogram(s, t, n)
        assert g.compute_ace() == golden_calib[f"{name}/gace"]
        assert g.compute_ece() == golden_calib[f"{name}/gece"]
    assert np.isnan(calibration.GlobalCalibAccumulator().compute_ace())


def test_ncc_from_sums(golden_ncc_aurc):
    g = golden_ncc_aurc["ncc/gt_map"].astype(np.float64).ravel()
    u = golden_ncc_aurc["ncc/pred"].astype(np.float64).ravel()
    got = vncc.ncc_from_sums(g.size, g.sum(), (g * g).sum(), u.sum(), (u * u).sum(), (g * u).sum())
    np.testing.assert_allclose(got, golden_ncc_aurc["ncc/value"], rtol=1e-6)  # the reference centres the float32 map in float32
    assert vncc.ncc_from_sums(g.size, 0.0, 0.0, u.sum(), (u * u).sum(), 0.0) == 0.0
    c = np.full(u.size, 0.3, np.float32).astype(np.float64)
    assert vncc.ncc_from_sums(c.size, g.sum(), (g * g).sum(), c.sum(), (c * c).sum(), (g * c).sum()) == 0.0
    self_ncc = vncc.ncc_from_sums(u.size, u.sum(), (u * u).sum(), u.sum(), (u * u).sum(), (u * u).sum())
    np.testing.assert_allclose(self_ncc, (u.size - 1) / u.size, rtol=1e-9)


def test_aurc_matches_golden_and_oracle(golden_ncc_aurc):
    r, c = golden_ncc_aurc["aurc/risks"], golden_ncc_aurc["aurc/confids"]
    np.testing.assert_allclose(vaurc.aurc(r, c), golden_ncc_aurc["aurc/aurc"], rtol=1e-12)
    np.testing.assert_allclose(vaurc.eaurc(r, c), golden_ncc_aurc["aurc/eaurc"], rtol=1e-10)
    cov, sel, w = vaurc.rc_curve_stats(r, c)
    np.testing.assert_allclose(sel, golden_ncc_aurc["aurc/selective_risks"], rtol=1e-12)
    np.testing.assert_allclose(w, golden_ncc_aurc["aurc/weights"], rtol=1e-12)
    # tied confidences: points only on value changes (sorted stably here, so feed sorted ties)
    rng = np.random.default_rng(1)
    for n in (2, 3, 17, 100):
        c2 = np.sort(np.round(rng.random(n), 1))
        r2 = rng.random(n)
        np.testing.assert_allclose(vaurc.aurc(r2, c2), oracle.aurc(r2, c2), rtol=1e-12)
    assert vaurc.aurc(np.array([0.3]), np.array([0.1])) == oracle.aurc(np.array([0.3]), np.array([0.1]))


def test_dice_from_counts():
    tp = np.array([[1, 0, 2], [0, 0, 0]])
    ps = np.array([[2, 2, 2], [0, 0, 5]])
    gs = np.array([[1, 0, 2], [0, 3, 0]])
    got = vaurc.binary_dice_from_counts(tp, ps, gs)
    want = [oracle.binary_dice_from_counts(tp[i], ps[i], gs[i]) for i in range(2)]
    np.testing.assert_allclose(got, want, rtol=1e-7)


def test_group_members_equals_stack_mean():
    """test_2D.py:1277: torch.stack(groups).mean(dim=1) == one member per group; single-sample groups stay views."""
    import torch
    from diffuncertainty_b200 import group_members
    gen = torch.Generator().manual_seed(3)
    for n_g in (1, 3, 17):
        groups = [torch.softmax(torch.randn(n_g, 2, 4, 5, 6, generator=gen), dim=2) for _ in range(4)]
        want = torch.stack(groups).mean(dim=1)
        got = group_members(groups)
        assert len(got) == 4
        for i, m in enumerate(got):
            torch.testing.assert_close(m, want[i], rtol=1e-6, atol=1e-7)
            if n_g == 1:
                assert m.data_ptr() == groups[i].data_ptr() and torch.equal(m, want[i])
    try:
        group_members([torch.zeros(2, 3, 4)])
    except ValueError:
        pass
    else:
        raise AssertionError("a group without a spatial axis must be rejected")


def test_platt_edges_zero_slope_and_hints():
    """a == 0 is the reference's own fallback when the Platt fit fails (ace.py: a, b = 0.0, 0.0): the confidence is a constant
    and every sample falls into one bin.  Hints (used by eqACE to shorten the bisection) never change the thresholds."""
    import numpy as np
    from diffuncertainty_b200 import calibration as c
    rng = np.random.default_rng(0)
    u = np.concatenate([rng.random(500), [0.0, 7.5, -1.0]]).astype(np.float32)
    for b in (0.0, 0.3, -2.0, 5.0):
        pe = c.platt_edges(0.0, b)
        with np.errstate(over="ignore"):
            conf = np.clip(1 / (1 + np.exp((-u) * np.float32(0.0) + np.float32(b))), 0, 1)
        assert np.array_equal(pe.bin_of(u), np.digitize(conf, c.bin_edges()) - 1), b
    for a, b in ((3.5, -1.25), (-40.0, 0.5)):
        e = np.sort(rng.random(19)) * 0.9 + 0.05
        want = c.platt_edges(a, b, e).edge_u
        near = np.stack([np.nextafter(want, np.float32(-np.inf)), np.nextafter(want, np.float32(np.inf))])
        far = np.stack([np.sort(rng.random(19)).astype(np.float32), np.full(19, np.nan, np.float32)])
        for hints in (near, far):
            assert np.array_equal(c.platt_edges(a, b, e, hints=hints).edge_u, want, equal_nan=True)


def test_platt_edges_many_equals_platt_edges():
    """the row-vectorised inversion used by eqace_from_maps_batch is the scalar platt_edges, row by row"""
    from diffuncertainty_b200 import calibration
    rng = np.random.default_rng(5)
    for inc in (True, False):
        S = 7
        a = (rng.random(S) * 6 + 0.2) * (1 if inc else -1)
        b = rng.normal(size=S) * 2
        edges = np.sort(rng.random((S, 19)), axis=1)
        edges[0, 5:9] = edges[0, 5]  # tied quantile edges
        edges[1, 15:] = 1.0          # unreachable from below 1
        hints = (rng.random((2, S, 19)) * 3).astype(np.float32)
        hints[0, 2, 3] = np.nan
        thr = calibration._platt_edges_many(a.astype(np.float32), b.astype(np.float32), edges, hints, inc)
        for s in range(S):
            one = calibration.platt_edges(float(np.float32(a[s])), float(np.float32(b[s])), edges[s], hints=hints[:, s])
            np.testing.assert_array_equal(thr[s], one.edge_u)
            assert one.mode == (1 if inc else 0)
