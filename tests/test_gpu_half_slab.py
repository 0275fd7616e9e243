"""bfloat16 / float16 slabs (what a network under autocast returns) read as they are (vu_slab.dtype): the kernel widens every
value to float32 as it reads it, so maps, labels and statistics must equal those of the upcast slab bit for bit -- the
reference computes float32 maps whatever the input dtype (test_utils.py:836).  Slabs without a native form
(unaligned rows, one member, more than 32 members) are upcast by the wrapper as before."""
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


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("P,B,C,spatial,native", [
    (16, 2, 19, (16, 64), True),    # cfg 5 shape family
    (10, 3, 2, (32, 32), True),     # binary
    (5, 1, 7, (8, 24), True),       # a class count without a compiled-in form
    (20, 1, 3, (4, 64), True),      # two cascade levels
    (4, 2, 5, (5, 7), False),       # rows that are not 16-byte aligned: upcast by the wrapper
    (1, 2, 4, (8, 16), False),      # single prediction
    (40, 1, 2, (8, 16), False),     # more than 32 members
])
def test_half_slab_equals_upcast(vu, dtype, P, B, C, spatial, native):
    from diffuncertainty_b200 import _lib, calibration
    g = torch.Generator().manual_seed(P * 10 + C)
    x = torch.softmax(3.0 * torch.randn(P, B, C, *spatial, generator=g), dim=2).to(dtype)
    x[0, 0, :, 0, :2] = 0
    if P > 2:
        x[1, 0, 0, 0, 3] = float("nan")
    gt = torch.randint(0, C, (B, 2, *spatial), generator=g, dtype=torch.uint8)
    flags = (_lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA | _lib.STAT_DICE | _lib.STAT_CALIB) if P > 1 else 0
    platt = [calibration.platt_edges(a, b) for a, b in ((3.5, -1.25), (6.0, -2.0), (40.0, -0.5))]
    kw = dict(stats=flags, thresholds=[0.3, 0.2, 0.02], calib=platt if flags else None)
    xd = x.cuda()
    before = _lib.get_counter("launches.k1_co_tma")
    half = vu.fused_pass(xd, vu.GroundTruth(gt.cuda(), None), **kw)
    half_plain = vu.fused_pass(xd)
    half_list = vu.fused_pass([xd[p] for p in range(P)]) if P > 1 else half_plain
    took_native = _lib.get_counter("launches.k1_co_tma") - before
    assert took_native == (3 if native else 0), took_native
    ref = vu.fused_pass(xd.float(), vu.GroundTruth(gt.cuda(), None), **kw)
    for r in (half, half_plain, half_list):
        assert set(r.maps) == set(ref.maps)
        for k in ref.maps:
            assert torch.equal(r.maps[k].view(torch.int32), ref.maps[k].view(torch.int32)), k
        assert torch.equal(r.labels, ref.labels)
    if flags:
        assert torch.equal(half.stats_i64, ref.stats_i64)
        np.testing.assert_allclose(half.stats_f64.cpu().numpy(), ref.stats_f64.cpu().numpy(), rtol=1e-7, atol=1e-12)
    # drop-in: calculate_uncertainty on a half tensor
    if P > 1:
        d = vu.calculate_uncertainty(xd[:, 0])
        assert torch.equal(d["TU"].view(torch.int32), ref.maps["TU"][0].view(torch.int32))


def test_half_slab_strided_batch_view(vu):
    g = torch.Generator().manual_seed(2)
    x = torch.softmax(2.0 * torch.randn(6, 5, 19, 8, 32, generator=g), dim=2).to(torch.bfloat16).cuda()
    view = x[:, 1:4]
    a = vu.fused_pass(view)
    b = vu.fused_pass(view.float())
    for k in ("TU", "AU", "EU"):
        assert torch.equal(a.maps[k].view(torch.int32), b.maps[k].view(torch.int32))
    assert torch.equal(a.labels, b.labels)
