"""td_region_composite (csrc/td_region.cu) against the reference's tensor expressions (helpers.region_composite_reference:
multidiffusion.py:187-216, mixtureofdiffusers.py:145-175 evaluated by torch on the CPU): bit patterns, all dtypes,
overlapping BACKGROUND and FOREGROUND regions, with and without the weight-canvas normalisation."""
import pytest
import torch

from helpers import DTYPES, assert_bit_equal, region_composite_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dn", list(DTYPES))
@pytest.mark.parametrize("divide", [True, False])
def test_region_composite_bit_exact(dn, divide):
    from multidiffusion_upscaler_for_automatic1111_b200 import engine
    g = torch.Generator().manual_seed(11)
    N, C, H, W = 2, 4, 40, 56
    dt = DTYPES[dn]
    xb = (torch.randn((N, C, H, W), generator=g) * 2).to(dt)
    weights = torch.randint(0, 5, (H, W), generator=g).float()
    rects = [(3, 2, 20, 17, 0), (10, 8, 30, 25, 1), (0, 0, 56, 40, 0), (25, 5, 24, 30, 1), (40, 20, 16, 20, 1), (8, 30, 9, 7, 0)]
    regions = []
    for (x, y, w, h, mode) in rects:
        out = (torch.randn((N, C, h, w), generator=g) * 1.5).to(dt)
        if mode == 1:
            aux = torch.rand((h, w), generator=g)
        else:
            aux = None if divide else torch.rand((1, 1, h, w), generator=g) * 3     # Mixture of Diffusers: pre-rescaled custom weights
        regions.append((x, y, w, h, mode, out, aux))
    want = region_composite_reference(xb, weights if divide else None, regions)
    dev_regions = [(x, y, w, h, m, o.cuda(), None if a is None else a.cuda()) for (x, y, w, h, m, o, a) in regions]
    got = engine.region_composite(xb.cuda(), weights.cuda() if divide else None, dev_regions)
    assert_bit_equal(got.cpu(), want, f"region composite {dn} divide={divide}")
    # no regions at all: plain normalisation
    got0 = engine.region_composite(xb.cuda(), weights.cuda() if divide else None, [])
    assert_bit_equal(got0.cpu(), region_composite_reference(xb, weights if divide else None, []), "no regions")
