"""VU_STAT_CLASS_COUNTS: the per-rater, per-class tp / pred / gt counts of the multi-class Dice (test_2D.py:901-918), fused
into the streaming pass and from stored labels, bit-exact against NumPy; the host macro Dice against the restated
dice_wrapped / DiceScore semantics; and the failure-detection summary of a multi-class sweep fed from the kernel."""
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


@pytest.mark.parametrize("P,B,C,spatial,R,ignore,dtype", [
    (10, 2, 19, (64, 128), 5, 255, torch.uint8),    # configs[2]: GTA-like, 5 raters, ignore 255
    (4, 3, 19, (33, 47), 2, 255, torch.int64),      # ragged sizes (register-streaming / generic kernels), int64 references
    (6, 2, 5, (24, 40), 3, None, torch.uint8),      # generic class count, no ignore value
    (5, 2, 2, (16, 16, 16), 4, None, torch.uint8),  # binary 3-D
])
def test_class_counts_fused_and_stored(vu, P, B, C, spatial, R, ignore, dtype):
    from diffuncertainty_b200 import _lib, aurc
    from oracle import oracle
    g = torch.Generator().manual_seed(P * C + R)
    x = torch.softmax(3.0 * torch.randn(P, B, C, *spatial, generator=g), dim=2)
    member0 = x[0].argmax(dim=1)
    gt = torch.where(torch.rand(B, R, *spatial, generator=g) < 0.7, member0.unsqueeze(1).expand(B, R, *spatial),
                     torch.randint(0, C, (B, R, *spatial), generator=g))
    if ignore is not None:
        gt = torch.where(torch.rand(B, R, *spatial, generator=g) < 0.05, torch.full_like(gt, ignore), gt)
    gt = gt.to(dtype)
    gtd = vu.GroundTruth(gt.cuda(), ignore)
    flags = _lib.STAT_IMAGE_SUM | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CLASS_COUNTS
    res = vu.fused_pass(x.cuda(), gtd, stats=flags)
    tp, ps, gs = res.class_count_arrays()
    labels = res.labels.cpu().numpy()
    for b in range(B):
        otp, ops, ogs = oracle.class_counts(labels[b], gt[b].numpy(), C, ignore)
        assert np.array_equal(tp[b], otp) and np.array_equal(ps[b], ops) and np.array_equal(gs[b], ogs), b
        # class 1 of the class counts is the binary Dice statistic (test_2D.py:878-886)
        btp, bps, bgs = res.dice_counts()
        assert np.array_equal(btp[b], otp[:, 1]) and np.array_equal(bps[b], ops[:, 1]) and np.array_equal(bgs[b], ogs[:, 1])
        want = np.mean([oracle.macro_dice_reference_semantics(labels[b], gt[b, r].numpy(), C, ignore) for r in range(R)])
        got = aurc.multiclass_dice_from_counts(tp[b], ps[b], gs[b])
        np.testing.assert_allclose(got, want, rtol=1e-12)
    # the same from stored labels (vu_map_stats), accumulated twice into one buffer
    maps = {k: res.maps[k] for k in ("TU", "AU", "EU")}
    buf = torch.zeros((B, R, C, 3), dtype=torch.int64, device="cuda")
    for _ in range(2):
        vu.map_stats(maps, res.labels, gtd, stats=_lib.STAT_CLASS_COUNTS, n_classes=C, class_counts_out=buf)
    assert torch.equal(buf, 2 * res.class_counts)


def test_macro_dice_edge_cases():
    from diffuncertainty_b200 import aurc
    z = np.zeros((1, 4), np.int64)
    assert aurc.multiclass_dice_from_counts(z, z, z) == 1.0                                          # every pixel ignored
    only_bg = np.array([[50, 0, 0, 0]])
    assert aurc.multiclass_dice_from_counts(only_bg, only_bg, only_bg) == 1.0                        # background on both sides
    tp, ps, gs = np.array([[40, 5, 0, 0]]), np.array([[45, 5, 10, 0]]), np.array([[50, 10, 0, 0]])   # class 2 predicted, never true
    np.testing.assert_allclose(aurc.multiclass_dice_from_counts(tp, ps, gs), np.mean([2 * 5 / 15, 0.0]))


def test_multiclass_sweep_failure_detection_uses_class_counts(vu):
    from diffuncertainty_b200 import _lib, aurc, sweep
    cfg = sweep.SweepConfig(P=4, C=19, spatial=(32, 64), n_images=6, batch=4, R=2, ignore_index=255, seed=3, scale=3.0, flip=0.3,
                            ignore_frac=0.03, stats=_lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA | _lib.STAT_DICE |
                            _lib.STAT_CLASS_COUNTS)
    res = sweep.ShardedSweep(cfg).run()
    assert res.class_counts is not None and res.class_counts.shape == (6, 2, 19, 3)
    dice = res.dice()
    want = aurc.multiclass_dice_from_counts(res.class_counts[..., 0], res.class_counts[..., 1], res.class_counts[..., 2])
    np.testing.assert_allclose(dice, want)
    fd = res.failure_detection()
    np.testing.assert_allclose(fd["EU/image_level"]["aurc"], aurc.aurc(1.0 - want, -res.image_level()[:, 2]))
