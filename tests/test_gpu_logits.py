"""The fused pass over LOGITS (vu_fused_pass_logits; SURVEY section 8f rank 3: F.softmax(output, dim=1) of test_2D.py:1181-1256
folded into the read) against the oracle: torch's CPU softmax followed by the reference's mean / argmax / calculate_uncertainty.

Relaxed contract (valunc.h, vu_slab): the device's exponential is not torch's, so
  * maps: |got - want| <= 1e-5 * max(|TU|, |AU|) + 1e-6 (the absolute term is the reference's own float32 rounding of a
    probability next to 1: an entropy of 1e-5 is only known to ~1e-7 from float32 probabilities);
  * labels: may differ only where the two largest mean probabilities of the voxel agree to 2e-6 relative; the measured rate is
    asserted below 1e-4 and written to gpurun_out/ (profiles/r02_logits_parity.json).
The TMA forms and the generic kernel agree bit for bit with each other."""
import json
import os

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


def reference(logits_cpu):
    """(P, B, C, *S) logits -> reference probabilities, labels, maps (oracle = the reference's own torch expressions)"""
    from oracle import oracle
    torch.set_num_threads(1)
    probs = oracle.softmax_logits(logits_cpu, dim=2)
    out = []
    for b in range(probs.shape[1]):
        ref = oracle.calculate_uncertainty(probs[:, b])
        mean = probs[:, b].mean(dim=0)
        out.append((mean, mean.argmax(dim=0).to(torch.uint8), ref))
    return probs, out


def check_against_reference(res, logits_cpu, what, stats=None):
    probs, refs = reference(logits_cpu)
    n_vox = n_flip = 0
    worst = 0.0
    for b, (mean, label, ref) in enumerate(refs):
        got_label = res.labels[b].cpu()
        flip = got_label != label
        n_vox += flip.numel(); n_flip += int(flip.sum())
        if flip.any():
            # a different label is only acceptable on a near-tie of the reference's own mean probabilities
            top2 = mean.flatten(1).topk(2, dim=0).values[:, flip.flatten()]
            gap = ((top2[0] - top2[1]) / top2[0]).max().item()
            assert gap <= 2e-6, f"{what}: label differs on a voxel whose top-2 mean probabilities are {gap:.2e} apart"
        tu, au = ref["TU"].numpy().astype(np.float64), ref["AU"].numpy().astype(np.float64)
        scale = np.maximum(np.abs(tu), np.abs(au))
        for k in ("TU", "AU", "EU"):
            got = res.maps[k][b].cpu().numpy().astype(np.float64)
            err = np.abs(got - ref[k].numpy().astype(np.float64))
            tol = 1e-5 * scale + 1e-6
            assert np.all(err <= tol), f"{what} {k}: max excess {np.max(err - tol):.3e}, max err {err.max():.3e}"
            worst = max(worst, float(np.max(err / (scale + 1e-6))))
    if stats is not None:
        stats.append({"case": what, "voxels": n_vox, "label_flips": n_flip, "max_err_over_scale": worst})
    return n_flip, n_vox


@pytest.mark.parametrize("P,B,C,spatial,scale", [
    (16, 2, 19, (32, 64), 3.0),    # cfg 5 shape family: TMA form, one voxel per thread
    (10, 1, 19, (16, 96), 8.0),    # peaked
    (20, 1, 19, (8, 64), 2.0),     # two cascade levels
    (10, 3, 2, (64, 64), 2.0),     # binary: TMA form, four voxels per thread
    (32, 2, 2, (32, 32), 6.0),
    (5, 2, 3, (16, 64), 3.0),
    (7, 1, 4, (16, 64), 3.0),
    (6, 2, 7, (8, 40), 3.0),       # no TMA form: generic kernel
    (5, 1, 19, (7, 9), 3.0),       # unaligned rows: generic kernel
])
def test_logits_vs_oracle(vu, P, B, C, spatial, scale):
    g = torch.Generator().manual_seed(P * 1000 + C * 10 + B)
    logits = scale * torch.randn(P, B, C, *spatial, generator=g) + 2.0 * torch.randn(P, B, 1, *spatial, generator=g)
    res = vu.fused_pass(logits.cuda(), logits=True)
    check_against_reference(res, logits, f"P{P} C{C} {spatial}")
    # the generic kernel gives the same bits
    from diffuncertainty_b200 import _lib
    _lib.load().vu_set_option(b"k1_variant", -2)
    _lib.load().vu_set_option(b"k1_path", 1)
    try:
        gen = vu.fused_pass(logits.cuda(), logits=True)
    finally:
        _lib.load().vu_set_option(b"k1_variant", -1)
        _lib.load().vu_set_option(b"k1_path", 0)
    assert torch.equal(gen.labels, res.labels)
    for k in ("TU", "AU", "EU"):
        assert torch.equal(gen.maps[k].view(torch.int32), res.maps[k].view(torch.int32)), f"{k}: TMA form != generic kernel"


def test_logits_special_values(vu):
    """torch's softmax rules: a NaN or +inf logit (or all -inf) makes the whole draw NaN (its terms are skipped, the mean is NaN
    and NaN is the argmax); a -inf logit next to finite ones has probability exactly 0."""
    from oracle import oracle
    for C, spatial in ((19, (8, 64)), (2, (16, 32)), (5, (4, 16))):
        g = torch.Generator().manual_seed(C)
        logits = 3.0 * torch.randn(6, 1, C, *spatial, generator=g)
        logits[1, 0, 0, 0, 0] = float("nan")
        logits[2, 0, C - 1, 0, 1] = float("inf")
        logits[3, 0, :, 0, 2] = float("-inf")
        logits[4, 0, 0, 0, 3] = float("-inf")          # masked-out class
        logits[0, 0, 1, 0, 4] = float("-inf")
        logits[5, 0, 1, 0, 4] = float("-inf")
        logits[:, 0, 0, 1, 0] = float("-inf")           # a class masked out in every member
        logits[2, 0, 0, 1, 1] = -3e38                   # difference overflows
        logits[2, 0, 1, 1, 1] = 3e38
        res = vu.fused_pass(logits.cuda(), logits=True, want_member_labels=True)
        probs, refs = reference(logits)
        mean, label, ref = refs[0]
        assert torch.equal(res.labels[0].cpu(), label)
        assert torch.equal(res.member_labels[:, 0].cpu(), probs[:, 0].argmax(dim=1).to(torch.uint8))
        for k in ("TU", "AU", "EU"):
            np.testing.assert_allclose(res.maps[k][0].cpu().numpy(), ref[k].numpy(), rtol=2e-5, atol=2e-6, err_msg=f"C={C} {k}")


def test_logits_with_statistics_members_and_groups(vu):
    """logits through the other forms of the slab: a member list, strided views, grouped draws with renormalise / discretise,
    P == 1, and the statistics of the same pass (which only see maps and labels)."""
    from diffuncertainty_b200 import _lib
    from oracle import oracle
    g = torch.Generator().manual_seed(3)
    P, B, C, spatial = 6, 2, 19, (16, 64)
    logits = 3.0 * torch.randn(P, B, C, *spatial, generator=g)
    gt = torch.randint(0, C, (B, 2, *spatial), generator=g, dtype=torch.uint8)
    flags = _lib.STAT_IMAGE_SUM | _lib.STAT_AREA | _lib.STAT_DICE
    whole = vu.fused_pass(logits.cuda(), vu.GroundTruth(gt.cuda(), None), stats=flags, logits=True)
    check_against_reference(whole, logits, "statistics launch")
    sums = whole.stats_f64[:, :3].cpu().numpy()
    for b in range(B):
        for k, name in enumerate(("TU", "AU", "EU")):
            np.testing.assert_allclose(sums[b, k], whole.maps[name][b].double().sum().item(), rtol=1e-7)
    # binary slab with reference-based statistics: the unified-warp TMA form (k1_uni) against the generic kernel
    from diffuncertainty_b200 import calibration
    lb = 4.0 * torch.randn(5, 3, 2, 16, 16, 16, generator=g)
    gb = torch.randint(0, 2, (3, 4, 16, 16, 16), generator=g, dtype=torch.uint8)
    platt = [calibration.platt_edges(a_, b_) for a_, b_ in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    for fl in (_lib.STAT_IMAGE_SUM | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CALIB, _lib.STAT_IMAGE_SUM | _lib.STAT_NCC):
        kw = dict(stats=fl, calib=platt if fl & _lib.STAT_CALIB else None, logits=True)
        before = _lib.get_counter("launches.k1_uni")
        uni = vu.fused_pass(lb.cuda(), vu.GroundTruth(gb.cuda(), None), **kw)
        assert _lib.get_counter("launches.k1_uni") == before + 1, "this launch should take the unified-warp kernel"
        _lib.load().vu_set_option(b"k1_variant", -2)
        try:
            gen = vu.fused_pass(lb.cuda(), vu.GroundTruth(gb.cuda(), None), **kw)
        finally:
            _lib.load().vu_set_option(b"k1_variant", -1)
        assert torch.equal(uni.labels, gen.labels)
        for k in ("TU", "AU", "EU"):
            assert torch.equal(uni.maps[k].view(torch.int32), gen.maps[k].view(torch.int32)), k
        assert torch.equal(uni.stats_i64, gen.stats_i64)
        np.testing.assert_allclose(uni.stats_f64.cpu().numpy(), gen.stats_f64.cpu().numpy(), rtol=1e-7, atol=1e-9)
        check_against_reference(uni, lb, f"binary logits, statistics {fl:#x}")
    # member list == stacked tensor, strided batch view == contiguous
    lst = vu.fused_pass([logits[p].cuda() for p in range(P)], logits=True)
    assert torch.equal(lst.labels, whole.labels) and torch.equal(lst.maps["TU"], whole.maps["TU"])
    big = torch.zeros(P, B + 1, C, *spatial)
    big[:, 1:] = logits
    view = vu.fused_pass(big.cuda()[:, 1:], logits=True)
    assert torch.equal(view.labels, whole.labels) and torch.equal(view.maps["AU"], whole.maps["AU"])
    # grouped draws of logits, renormalised and one-hot: softmax comes first (it is what the network returns)
    groups = [3.0 * torch.randn(2, B, 4, 8, 32, generator=g) for _ in range(5)]
    for renorm, onehot in ((False, False), (True, False), (False, True)):
        probs = [torch.stack([oracle.softmax_logits(d, dim=1) for d in grp]) for grp in groups]
        if renorm:
            probs = [torch.stack([oracle.renormalize_probabilities(d) for d in grp]) for grp in probs]
        want = oracle.build_softmax_pred(probs, onehot)
        res = vu.fused_pass(vu.Groups([t.cuda() for t in groups], renormalize=renorm, discretize=onehot, logits=True))
        for b in range(B):
            ref = oracle.calculate_uncertainty(want[:, b])
            label = want[:, b].mean(dim=0).argmax(dim=0).to(torch.uint8)
            flips = (res.labels[b].cpu() != label).float().mean().item()
            assert flips <= (0.02 if onehot else 1e-3), (renorm, onehot, flips)  # one-hot votes flip with a member's near-tie
            if not onehot:
                for k in ("TU", "AU"):
                    np.testing.assert_allclose(res.maps[k][b].cpu().numpy(), ref[k].numpy(), rtol=2e-5, atol=2e-6)
    # P == 1: 1 - max softmax (test_utils.py:862-864)
    one = vu.fused_pass(logits[:1].cuda(), logits=True)
    want = 1.0 - oracle.softmax_logits(logits[0], dim=1).max(dim=1).values
    np.testing.assert_allclose(one.maps["pred_entropy"].cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    # drop-in
    d = vu.calculate_uncertainty_from_logits(logits[:, 0].cuda())
    assert torch.equal(d["TU"], whole.maps["TU"][0])
    # the plain entry point refuses the flag
    a = _lib.FusedArgs()
    import ctypes as C_
    a.struct_size = C_.sizeof(_lib.FusedArgs)
    a.slab.flags = _lib.SLAB_LOGITS
    assert _lib.load().vu_fused_pass(C_.byref(a), None) == -1


def test_logits_label_flip_rate_and_map_error(vu):
    """The measured numbers of the relaxed contract on BASELINE-shaped inputs (cfg 5: N = 16, C = 19; cfg 1: N = 10, C = 2;
    peaked variant): label flips against the reference's labels and the largest map error relative to max(|TU|, |AU|)."""
    stats = []
    for name, P, C, spatial, scale, B in (("cfg5-shaped", 16, 19, (128, 256), 3.0, 2), ("cfg5 peaked", 16, 19, (128, 256), 8.0, 2),
                                          ("cfg1-shaped", 10, 2, (256, 256), 2.0, 2), ("cfg4-shaped", 32, 2, (128, 128), 4.0, 2),
                                          ("cfg5 full size (one 512x1024 image)", 16, 19, (512, 1024), 4.0, 1)):
        g = torch.Generator().manual_seed(len(name))
        logits = scale * torch.randn(P, B, C, *spatial, generator=g)
        res = vu.fused_pass(logits.cuda(), logits=True)
        check_against_reference(res, logits, name, stats)
    total = sum(s["voxels"] for s in stats)
    flips = sum(s["label_flips"] for s in stats)
    assert flips <= 1e-4 * total, stats
    out = {"test": "tests/test_gpu_logits.py::test_logits_label_flip_rate_and_map_error",
           "reference": "torch CPU F.softmax(dim=1) -> mean -> argmax / calculate_uncertainty (oracle)",
           "voxels": total, "label_flips": flips, "cases": stats}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r02_logits_parity.json", "w") as f:
        json.dump(out, f, indent=1)


@pytest.mark.parametrize("seed", range(12))
def test_logits_random_combination(vu, seed):
    """Seeded random shapes (1- to 3-D, aligned or not), member counts across the cascade boundaries, class counts with and
    without a TMA form, strided views and member lists, masked (-inf) classes: logits=True against torch's CPU softmax + the
    reference pipeline, under the relaxed contract."""
    rng = np.random.default_rng(4000 + seed)
    g = torch.Generator().manual_seed(4000 + seed)
    ndim = int(rng.integers(1, 4))
    spatial = tuple(int(s) for s in rng.integers(1, 20, ndim))
    if rng.random() < 0.6:
        spatial = spatial[:-1] + (int(rng.choice([4, 8, 16, 32, 64])),)
    P = int(rng.choice([2, 3, 5, 10, 16, 17, 18, 32, 33]))
    C = int(rng.choice([2, 2, 3, 4, 5, 19, 19, 21]))
    B = int(rng.integers(1, 4))
    scale = float(rng.choice([0.5, 2.0, 6.0, 15.0]))
    logits = scale * torch.randn(P, B, C, *spatial, generator=g) + 3.0 * torch.randn(P, B, 1, *spatial, generator=g)
    if rng.random() < 0.3 and C > 2:
        logits[:, :, 1] = float("-inf")  # a class masked out everywhere: probability exactly 0
    xd = logits.cuda()
    layout = rng.choice(["contiguous", "batch_slice", "member_list", "class_padded"])
    if layout == "batch_slice":
        big = torch.zeros((P, B + 2, C) + spatial, device="cuda")
        big[:, 1:B + 1] = xd
        arg = big[:, 1:B + 1]
    elif layout == "class_padded":
        big = torch.zeros((P, B, C + 3) + spatial, device="cuda")
        big[:, :, :C] = xd
        arg = big[:, :, :C]
    elif layout == "member_list":
        arg = [xd[p].clone() for p in range(P)]
    else:
        arg = xd
    res = vu.fused_pass(arg, logits=True)
    flips, n = check_against_reference(res, logits, f"seed {seed}: P{P} B{B} C{C} {spatial} {layout}")
    assert flips <= max(1, n // 1000)
