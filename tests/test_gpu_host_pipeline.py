"""HostPipeline (the end-to-end API bench.py's `e2e` leg measures: pinned host slab -> chunks through the device -> host maps,
labels and statistics rows) against one fused_pass over the whole slab on the device: same bits, for every chunking, with and
without statistics, for a single prediction (P == 1, test_2D.py:1006-1007) and for a slab of logits."""
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


@pytest.mark.parametrize("P,B,C,spatial,R,chunk,logits", [
    (6, 5, 19, (16, 64), 2, 2, False),     # ragged last chunk, statistics with references
    (4, 3, 2, (8, 8, 16), 0, 1, False),    # 3-D, no references
    (1, 4, 3, (16, 32), 0, 3, False),      # single prediction: "pred_entropy"
    (5, 4, 19, (16, 64), 1, 4, True),      # logits, one chunk
    (3, 7, 5, (8, 32), 0, 2, True),        # logits, class count without a compiled-in form
])
def test_host_pipeline_equals_device_pass(vu, P, B, C, spatial, R, chunk, logits):
    from diffuncertainty_b200 import _lib, calibration
    from diffuncertainty_b200.host_pipeline import HostPipeline
    g = torch.Generator().manual_seed(P * 100 + B * 10 + C)
    x = 3.0 * torch.randn(P, B, C, *spatial, generator=g)
    if not logits:
        x = torch.softmax(x, dim=2)
    x = x.contiguous().pin_memory()
    gt = torch.randint(0, C, (B, max(R, 1), *spatial), generator=g, dtype=torch.uint8).pin_memory() if R else None
    platt = [(3.5, -1.25), (6.0, -2.0), (40.0, -0.5)]
    stats = (_lib.STAT_IMAGE_SUM | _lib.STAT_THRESHOLD | _lib.STAT_AREA | (_lib.STAT_DICE | _lib.STAT_CALIB if R else 0)) if P > 1 else 0
    pipe = HostPipeline(P, C, spatial, B, R=R, chunk_images=chunk, stats=stats, thresholds=[0.3, 0.2, 0.02],
                        platt=platt if stats & _lib.STAT_CALIB else None, logits=logits)
    res = pipe.run(x, gt)
    ref = vu.fused_pass(x.cuda(), vu.GroundTruth(gt.cuda(), None) if R else None, stats=stats, thresholds=[0.3, 0.2, 0.02],
                        calib=[calibration.platt_edges(a, b) for a, b in platt] if stats & _lib.STAT_CALIB else None, logits=logits)
    torch.cuda.synchronize()
    assert set(res.maps) == set(ref.maps)
    for k in ref.maps:
        assert torch.equal(res.maps[k].view(torch.int32), ref.maps[k].cpu().view(torch.int32)), k
    assert torch.equal(res.labels, ref.labels.cpu())
    if stats:
        assert np.array_equal(res.stats_i64, ref.stats_i64.cpu().numpy())
        np.testing.assert_allclose(res.stats_f64, ref.stats_f64.cpu().numpy(), rtol=1e-9, atol=1e-12)
    assert res.h2d_bytes == x.numel() * 4 + (gt.numel() if R else 0)
    assert res.d2h_bytes >= B * int(np.prod(spatial)) * (4 * len(ref.maps) + 1)


def test_host_pipeline_half_slab(vu):
    """a bfloat16 host slab crosses PCIe and HBM at half the bytes and gives the maps of the upcast slab bit for bit"""
    from diffuncertainty_b200.host_pipeline import HostPipeline
    g = torch.Generator().manual_seed(11)
    P, B, C, spatial = 8, 5, 19, (16, 64)
    x = torch.softmax(3.0 * torch.randn(P, B, C, *spatial, generator=g), dim=2).to(torch.bfloat16).contiguous().pin_memory()
    res = HostPipeline(P, C, spatial, B, chunk_images=2, dtype=torch.bfloat16).run(x)
    ref = vu.fused_pass(x.cuda().float())
    torch.cuda.synchronize()
    for k in ref.maps:
        assert torch.equal(res.maps[k].view(torch.int32), ref.maps[k].cpu().view(torch.int32)), k
    assert torch.equal(res.labels, ref.labels.cpu())
    assert res.h2d_bytes == x.numel() * 2
