"""GPU parity of the member-level scores: GED counts / ged_binary_fast (ged_fast.py:5-142) and
the likelihood statistics (test_2D.py:1043-1120), against the golden vectors recorded from the unmodified reference
(tests/golden/ged_nll.npz) and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

AM = ["dice", "max_dice_pred", "max_dice_gt", "major_dice"]


@pytest.fixture(scope="module")
def members():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffuncertainty_b200 import members
    return members


def test_golden_ged_and_likelihood(members, golden_ged_nll):
    g = golden_ged_nll
    for name in g["cases"]:
        x = torch.from_numpy(g[f"{name}/x"]).cuda()
        gt, ign = g[f"{name}/gt"], int(g[f"{name}/ignore"])
        if f"{name}/ged" in g:
            got = members.ged_binary_fast(x, gt, None if ign == -999 else ign, AM)
            for k in ["ged"] + AM:
                # float32 arithmetic on exact integer counts: the reference's own rounding
                np.testing.assert_allclose(got[k], float(g[f"{name}/{k}"]), rtol=2e-6, atol=2e-7, err_msg=f"{name}/{k}")
        if f"{name}/mean_nll" in g:
            gt_arg = gt[0] if name == "single_rater_2d" else gt
            m, r, mean = members.compute_likelihood_stats(x, gt_arg, -1 if ign == -999 else ign)
            np.testing.assert_allclose(np.array(m), g[f"{name}/gt_model_nll"], rtol=1e-5, atol=1e-7, err_msg=name)
            np.testing.assert_allclose(np.array(r), g[f"{name}/gt_nll"], rtol=1e-5, atol=1e-7, err_msg=name)
            np.testing.assert_allclose(mean, float(g[f"{name}/mean_nll"]), rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(members.compute_expected_nll(x, gt_arg, -1 if ign == -999 else ign),
                                       float(g[f"{name}/expected_nll"]), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("P,B,C,spatial,R,ignore,dtype", [
    (32, 5, 2, (128, 128), 4, None, torch.int64),   # configs[3]: diffusion samples, int64 references as the loader hands them
    (5, 3, 2, (20, 24, 28), 4, None, torch.uint8),  # configs[1]-like 3-D
    (7, 2, 2, (33, 47), 3, 255, torch.uint8),       # ragged sizes, ignore value
    (9, 2, 2, (16, 40), 2, 1, torch.int64),         # the ignore value is class 1 itself (gg counts take the raw labels)
    (10, 2, 19, (24, 40), 5, None, torch.uint8),    # multi-class: likelihood only
])
def test_member_scores_vs_oracle(members, P, B, C, spatial, R, ignore, dtype):
    from diffuncertainty_b200 import uncertainty as vu
    from oracle import oracle
    gen = torch.Generator().manual_seed(P * 100 + B)
    x = torch.softmax(3.0 * torch.randn(P, B, C, *spatial, generator=gen), dim=2)
    if C == 2 and ignore is None:
        x[0, 0, :, 0] = float("nan")  # argmax's NaN rule (ged_fast.py:44); the log-likelihood sums of member 0 become NaN
    gt = torch.randint(0, C, (B, R, *spatial), generator=gen)
    if ignore is not None:
        gt[torch.rand(gt.shape, generator=gen) < 0.1] = ignore
    gt = gt.to(dtype)
    xd = x.cuda()
    gtd = vu.GroundTruth(gt.cuda(), ignore)
    do_ged = C == 2
    labels = vu.fused_pass(xd, want_maps=False).labels if do_ged else None
    res = members.member_scores(xd, gtd, nll=True, ged=do_ged, mean_labels=labels)
    # the same through a list of members (read in place, no stack) and a strided view
    res2 = members.member_scores([xd[p] for p in range(P)], gtd, nll=True, ged=do_ged, mean_labels=labels)
    for b in range(B):
        xb, gb = x[:, b].numpy(), gt[b].numpy().astype(np.int64)
        if do_ged:
            lab = np.stack([oracle.argmax_first_nan_max(m) for m in xb])
            mean_lab = oracle.argmax_first_nan_max(oracle.mean_members_f32(xb))
            want = oracle.ged_counts(lab, gb, ignore, mean_lab)
            got = res.ged_parts(b)
            for k in want:
                assert np.array_equal(got[k], want[k]), (k, b)
            assert np.array_equal(res2.ged_counts[b], res.ged_counts[b])
            ged_want = oracle.ged_from_counts(want, AM)
            ged_got = res.ged(b, AM)
            for k in ged_want:
                np.testing.assert_allclose(ged_got[k], ged_want[k], rtol=1e-6, atol=1e-7)
        sums, counts = oracle.likelihood_sums(xb, gb, -1 if ignore is None else ignore)
        assert np.array_equal(res.nll_count[b], counts)
        np.testing.assert_allclose(res.nll_sum[b], sums, rtol=1e-5, atol=1e-6, equal_nan=True)
        np.testing.assert_allclose(res2.nll_sum[b], res.nll_sum[b], rtol=1e-12, equal_nan=True)


def test_member_scores_contract(members):
    from diffuncertainty_b200 import uncertainty as vu
    x = torch.softmax(torch.randn(3, 1, 3, 8, 8), 2).cuda()
    gt = vu.GroundTruth(torch.randint(0, 3, (1, 2, 8, 8)).cuda(), None)
    with pytest.raises(ValueError):
        members.member_scores(x, gt, nll=False, ged=True)            # GED is binary only (ged_fast.py:33)
    with pytest.raises(ValueError):
        members.ged_binary_fast(x[:, 0], gt.seg[0])
    with pytest.raises(ValueError):
        members.ged_binary_fast(torch.rand(3, 2, 8, 8).cuda(), torch.zeros(8, 8))
    bad = vu.GroundTruth(torch.full((1, 2, 8, 8), 7).cuda(), None)     # 7 is not a class: torch.gather raises in the reference
    with pytest.raises(RuntimeError):
        members.member_scores(x, bad, nll=True)
    with pytest.raises(Exception):
        members.member_scores(x.cpu(), gt)                             # no CPU fallback
    empty = members.member_scores(torch.rand(3, 0, 2, 4, 4).cuda(), vu.GroundTruth(torch.zeros(0, 1, 4, 4, dtype=torch.uint8).cuda()), ged=True)
    assert empty.nll_sum.shape == (0, 1, 3) and empty.ged_counts.shape[0] == 0


def test_fast_and_generic_kernels_agree(members):
    """The float4 kernel (4 consecutive voxels per lane) and the generic one (any strides) on the same slab: identical integer
    counts, likelihood sums equal up to float32 summation order."""
    from diffuncertainty_b200 import _lib, uncertainty as vu
    gen = torch.Generator().manual_seed(77)
    for P, B, spatial, R, ignore, dtype in ((32, 3, (64, 96), 4, None, torch.uint8), (6, 2, (12, 20, 8), 7, 255, torch.int64)):
        x = torch.softmax(4.0 * torch.randn(P, B, 2, *spatial, generator=gen), dim=2).cuda()
        gt = torch.randint(0, 2, (B, R, *spatial), generator=gen)
        if ignore is not None:
            gt[torch.rand(gt.shape, generator=gen) < 0.15] = ignore
        gtd = vu.GroundTruth(gt.to(dtype).cuda(), ignore)
        labels = vu.fused_pass(x, want_maps=False).labels
        out = []
        for path in (0, 1):
            _lib.set_option("k5_path", path)
            before = _lib.get_counter("launches.member_scores_c2v4")
            out.append(members.member_scores(x, gtd, nll=True, ged=True, mean_labels=labels))
            assert (_lib.get_counter("launches.member_scores_c2v4") > before) == (path == 0)
        _lib.set_option("k5_path", 0)
        assert np.array_equal(out[0].ged_counts, out[1].ged_counts)
        assert np.array_equal(out[0].nll_count, out[1].nll_count)
        np.testing.assert_allclose(out[0].nll_sum, out[1].nll_sum, rtol=2e-6)
