"""Member-level scores computed by the fused pass itself (vu_fused_args.members; csrc/members_fold.cuh in the unified-warp
kernel): GED counts bit-exact and likelihood sums within 1e-5 of the stand-alone pass (vu_member_scores) and of the oracle
(ged_fast.py:44-131, test_2D.py:1043-1120), with the maps / labels / statistics of the same launch unchanged."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import diffuncertainty_b200 as pkg
    from diffuncertainty_b200 import _lib
    _lib.require_device()
    return pkg


def make_case(P, B, spatial, R, ignore, seed, nan=False, bad_class=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.softmax(3.0 * torch.randn(P, B, 2, *spatial, generator=g), dim=2)
    if nan:
        x[0, 0, :, 0] = float("nan")       # argmax's NaN rule; the likelihood sums of member 0 become NaN
        x[min(2, P - 1), B - 1, 1].view(-1)[7] = float("nan")
    gt = torch.randint(0, 2, (B, R, *spatial), generator=g)
    if ignore is not None:
        gt[torch.rand(gt.shape, generator=g) < 0.1] = ignore
    if bad_class:
        gt[0, 0].view(-1)[3] = 7
    return x.cuda(), gt.to(torch.uint8).cuda()


def launches(names):
    from diffuncertainty_b200 import _lib
    return {n: _lib.get_counter("launches." + n) for n in names}


@pytest.mark.parametrize("flags", [0x21, 0x2d, 0x01])
@pytest.mark.parametrize("P,B,spatial,R,ignore,nan", [
    (32, 5, (128, 128), 4, None, False),   # configs[3]
    (5, 3, (20, 24, 28), 4, None, True),   # 3-D, one cascade level, NaN probabilities
    (7, 2, (36, 44), 3, 255, False),       # ignore value: the general form of the likelihood terms, ragged last tile
    (9, 2, (16, 40), 2, 1, True),          # the ignore value is class 1 itself (gg counts take the raw labels)
    (20, 1, (8, 12), 1, None, False),      # an image smaller than a tile, one rater
])
def test_fold_matches_standalone_pass(vu, flags, P, B, spatial, R, ignore, nan):
    from diffuncertainty_b200 import members
    x, gt = make_case(P, B, spatial, R, ignore, seed=P * 13 + R, nan=nan)
    gtd = vu.GroundTruth(gt, ignore)
    names = ("k1_uni", "member_scores", "member_scores_c2v4")
    before = launches(names)
    res, ms = members.fused_pass_with_member_scores(x, gtd, nll=True, ged=True, stats=flags)
    torch.cuda.synchronize()
    after = launches(names)
    assert after["k1_uni"] - before["k1_uni"] == 1, "the unified kernel did not take the launch"
    assert after["member_scores"] == before["member_scores"] and after["member_scores_c2v4"] == before["member_scores_c2v4"], \
        "a second pass over the slab was launched"
    plain = vu.fused_pass(x, gtd, stats=flags)
    assert torch.equal(res.labels, plain.labels)
    for k in ("TU", "AU", "EU"):
        assert torch.equal(torch.nan_to_num(res.maps[k], nan=-7.0), torch.nan_to_num(plain.maps[k], nan=-7.0))
    assert torch.equal(res.stats_i64, plain.stats_i64)
    np.testing.assert_allclose(res.stats_f64.cpu().numpy(), plain.stats_f64.cpu().numpy(), rtol=1e-7, atol=1e-12, equal_nan=True)  # other tile size, other summation order
    ref = members.member_scores(x, gtd, nll=True, ged=True, mean_labels=plain.labels)
    assert np.array_equal(ms.ged_counts, ref.ged_counts)
    assert np.array_equal(ms.nll_count, ref.nll_count)
    np.testing.assert_allclose(ms.nll_sum, ref.nll_sum, rtol=1e-5, atol=1e-6, equal_nan=True)
    for b in range(B):
        assert ms.ged(b, ["dice", "max_dice_pred", "max_dice_gt", "major_dice"]) == ref.ged(b, ["dice", "max_dice_pred", "max_dice_gt", "major_dice"])


def test_fold_vs_oracle(vu):
    from diffuncertainty_b200 import members
    from oracle import oracle
    P, B, spatial, R = 12, 2, (32, 48), 3
    x, gt = make_case(P, B, spatial, R, 255, seed=3)
    res, ms = members.fused_pass_with_member_scores(x, vu.GroundTruth(gt, 255), stats=0x21)
    xc, gc = x.cpu(), gt.cpu().numpy().astype(np.int64)
    for b in range(B):
        xb = xc[:, b].numpy()
        lab = np.stack([oracle.argmax_first_nan_max(m) for m in xb])
        mean_lab = oracle.argmax_first_nan_max(oracle.mean_members_f32(xb))
        want = oracle.ged_counts(lab, gc[b], 255, mean_lab)
        got = ms.ged_parts(b)
        for k in want:
            assert np.array_equal(got[k], want[k]), (k, b)
        sums, counts = oracle.likelihood_sums(xb, gc[b], 255)
        assert np.array_equal(ms.nll_count[b], counts)
        np.testing.assert_allclose(ms.nll_sum[b], sums, rtol=1e-5, atol=1e-6)


def test_fold_falls_back_to_two_passes_and_reports_bad_references(vu):
    from diffuncertainty_b200 import members
    # 33 members: the fold is built for 32 -> fused pass + vu_member_scores, same results object
    x = torch.softmax(torch.randn(33, 1, 2, 16, 16), 2).cuda()
    gt = vu.GroundTruth(torch.randint(0, 2, (1, 2, 16, 16), dtype=torch.uint8).cuda(), None)
    names = ("k1_uni", "member_scores", "member_scores_c2v4")
    before = launches(names)
    res, ms = members.fused_pass_with_member_scores(x, gt, nll=True, ged=False, stats=0x21)
    after = launches(names)
    assert after["member_scores"] + after["member_scores_c2v4"] - before["member_scores"] - before["member_scores_c2v4"] == 1
    assert ms.nll_sum.shape == (1, 2, 33)
    # a reference that is neither a class nor the ignore value: torch.gather raises in the reference (test_2D.py:1067)
    xb, gb = make_case(6, 1, (16, 16), 2, None, seed=1, bad_class=True)
    with pytest.raises(RuntimeError):
        members.fused_pass_with_member_scores(xb, vu.GroundTruth(gb, None), stats=0x21)


def test_fold_accumulates_into_given_buffers(vu):
    from diffuncertainty_b200 import members
    x, gt = make_case(8, 2, (32, 32), 2, None, seed=8)
    gtd = vu.GroundTruth(gt, None)
    _, once = members.fused_pass_with_member_scores(x, gtd, stats=0x21)
    bufs = members.MemberScoreBuffers(8, 2, 2, x.device)
    for _ in range(3):
        members.fused_pass_with_member_scores(x, gtd, stats=0x21, out=bufs)
    got = bufs.scores()
    assert np.array_equal(got.ged_counts, 3 * once.ged_counts) and np.array_equal(got.nll_count, 3 * once.nll_count)
    np.testing.assert_allclose(got.nll_sum, 3 * once.nll_sum, rtol=1e-9)
