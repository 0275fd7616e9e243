"""The upstream producers of the slab folded into the fused pass: groups of draws averaged
in the kernel (test_2D.py:1277), _renormalize_probabilities (test_2D.py:188-194) and the --discretize one-hot
(test_2D.py:1272-1275) -- against the oracle, which runs the reference's own torch expressions on the CPU and then the
reference's calculate_uncertainty.  Labels (mean and per member) bit-exact, maps within 1e-5."""
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


def reference_slab(groups, renormalize, discretize):
    from oracle import oracle
    torch.set_num_threads(1)
    if renormalize:
        groups = [torch.stack([oracle.renormalize_probabilities(d) for d in g]) for g in groups]
    return oracle.build_softmax_pred(groups, discretize)


@pytest.mark.parametrize("renormalize,discretize", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("G,n_g,B,C,spatial", [
    (5, 3, 2, 19, (8, 32)),     # 19 classes (class sum in two cascade chunks), three draws per group
    (4, 1, 2, 2, (16, 32)),     # single-draw groups: only the producers act
    (3, 18, 1, 4, (8, 32)),     # 18 draws: the group mean itself is a two-level cascade
    (20, 2, 1, 3, (4, 32)),     # 20 members: two cascade levels over the members
])
def test_groups_and_producers_vs_oracle(vu, renormalize, discretize, G, n_g, B, C, spatial):
    from oracle import oracle
    g = torch.Generator().manual_seed(G * 100 + n_g * 10 + C)
    groups = []
    for _ in range(G):
        t = torch.softmax(3.0 * torch.randn(n_g, B, C, *spatial, generator=g), dim=2)
        t = t * (1.0 + 0.05 * torch.randn(n_g, B, 1, *spatial, generator=g))  # like interpolated TTA output: sums drift off 1
        groups.append(t)
    groups[0][0, 0, :, 0, :4] = 0.0          # a voxel whose classes sum to 0: renormalisation must leave it alone
    if C > 2:
        groups[-1][0, 0, 1, 1] = groups[-1][0, 0, 2, 1]  # ties: first maximum wins in argmax
    want = reference_slab([t.clone() for t in groups], renormalize, discretize)   # (G, B, C, *S)
    res = vu.fused_pass(vu.Groups([t.cuda() for t in groups], renormalize=renormalize, discretize=discretize),
                        want_member_labels=True)
    for b in range(B):
        ref = oracle.calculate_uncertainty(want[:, b])
        label = want[:, b].mean(dim=0).argmax(dim=0).to(torch.uint8)
        assert torch.equal(res.labels[b].cpu(), label), "labels of the member mean"
        assert torch.equal(res.member_labels[:, b].cpu(), want[:, b].argmax(dim=1).to(torch.uint8)), "labels of the members"
        tu, au = ref["TU"].numpy().astype(np.float64), ref["AU"].numpy().astype(np.float64)
        for k in ("TU", "AU", "EU"):
            got = res.maps[k][b].cpu().numpy().astype(np.float64)
            # (the draws are deliberately NOT normalised -- sums drift off 1 like interpolated TTA output -- so positive and
            #  negative p log p terms cancel in TU / AU: the 1e-5 is relative to the size of the terms, ~0.03 at least)
            tol = 1e-5 * np.maximum(np.maximum(np.abs(tu), np.abs(au)), 0.03) + 1e-10
            assert np.all(np.abs(got - ref[k].numpy()) <= tol), k


def test_groups_equal_materialised_slab_bitwise(vu):
    """Grouped read == the same numbers stacked and averaged by torch on the device first (what group_members does)."""
    g = torch.Generator().manual_seed(1)
    groups = [torch.softmax(torch.randn(4, 2, 3, 16, 32, generator=g), dim=2).cuda() for _ in range(6)]
    a = vu.fused_pass(vu.Groups(groups))
    slab = torch.stack(groups).mean(dim=1)
    b = vu.fused_pass(slab)
    assert torch.equal(a.labels, b.labels)
    for k in ("TU", "AU", "EU"):
        assert torch.equal(a.maps[k], b.maps[k])


def test_groups_contract(vu):
    x = [torch.rand(2, 1, 3, 8, 8).cuda(), torch.rand(3, 1, 3, 8, 8).cuda()]
    with pytest.raises(ValueError):
        vu.fused_pass(vu.Groups(x))     # torch.stack would refuse groups of different sizes as well
    with pytest.raises(ValueError):
        vu.fused_pass(vu.Groups([]))


@pytest.mark.parametrize("P,B,C,spatial", [(10, 2, 19, (16, 64)), (32, 2, 2, (32, 32)), (5, 1, 3, (8, 64)), (18, 1, 4, (8, 32)),
                                           (6, 1, 7, (8, 32)), (4, 1, 19, (5, 7))])
def test_discretize_tma_form_equals_generic_and_oracle(vu, P, B, C, spatial):
    """--discretize on plain members (one draw per group, the usual case): C = 2, 3, 4, 19 with aligned rows take the one-hot TMA
    variants of k1_tma, everything else the generic kernel; same bits, and the reference's labels / maps."""
    from diffuncertainty_b200 import _lib
    from oracle import oracle
    g = torch.Generator().manual_seed(P * 100 + C)
    x = torch.softmax(3.0 * torch.randn(P, B, C, *spatial, generator=g), dim=2)
    x[0, 0, 0, 0, :4] = x[0, 0, 1, 0, :4]          # ties: first maximum
    x[1, 0, C - 1, 0, 4] = float("nan")            # NaN is maximal
    groups = [x[p:p + 1].cuda() for p in range(P)]
    before = _lib.get_counter("launches.k1_tma")
    res = vu.fused_pass(vu.Groups(groups, discretize=True), want_member_labels=True)
    fast = C in (2, 3, 4, 19) and int(np.prod(spatial)) % 4 == 0
    assert _lib.get_counter("launches.k1_tma") == before + (1 if fast else 0)
    _lib.load().vu_set_option(b"k1_variant", -2)
    try:
        gen = vu.fused_pass(vu.Groups(groups, discretize=True), want_member_labels=True)
    finally:
        _lib.load().vu_set_option(b"k1_variant", -1)
    assert torch.equal(res.labels, gen.labels) and torch.equal(res.member_labels, gen.member_labels)
    for k in ("TU", "AU", "EU"):
        assert torch.equal(res.maps[k], gen.maps[k]), k   # (value equality: AU is -0.0 / +0.0)
    torch.set_num_threads(1)
    want = oracle.build_softmax_pred([x[p:p + 1] for p in range(P)], True)
    for b in range(B):
        ref = oracle.calculate_uncertainty(want[:, b])
        assert torch.equal(res.labels[b].cpu(), want[:, b].mean(dim=0).argmax(dim=0).to(torch.uint8))
        assert torch.equal(res.member_labels[:, b].cpu(), want[:, b].argmax(dim=1).to(torch.uint8))
        for k in ("TU", "AU", "EU"):
            np.testing.assert_allclose(res.maps[k][b].cpu().numpy(), ref[k].numpy(), rtol=1e-5, atol=1e-7)
    # with statistics (Dice counts see the labels of the one-hot mean)
    gt = torch.randint(0, C, (B, 2, *spatial), generator=g, dtype=torch.uint8).cuda()
    fl = _lib.STAT_IMAGE_SUM | _lib.STAT_AREA | _lib.STAT_DICE
    a = vu.fused_pass(vu.Groups(groups, discretize=True), vu.GroundTruth(gt, None), stats=fl)
    _lib.load().vu_set_option(b"k1_variant", -2)
    try:
        c = vu.fused_pass(vu.Groups(groups, discretize=True), vu.GroundTruth(gt, None), stats=fl)
    finally:
        _lib.load().vu_set_option(b"k1_variant", -1)
    assert torch.equal(a.stats_i64, c.stats_i64) and torch.equal(a.labels, c.labels)
    np.testing.assert_allclose(a.stats_f64.cpu().numpy(), c.stats_f64.cpu().numpy(), rtol=1e-7, atol=1e-12)
