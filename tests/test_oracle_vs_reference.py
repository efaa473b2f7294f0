"""Live check of the oracle against the unmodified reference (build container only)."""
import numpy as np
import pytest
import torch

from helpers import DTYPES, assert_bit_equal
from oracle import blend, ref_shim, synth, tiling

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


@pytest.mark.parametrize("method", ["md", "mod"])
@pytest.mark.parametrize("dn", list(DTYPES))
def test_step_matches_reference(ref, method, dn):
    from oracle.make_golden import run_reference_step
    N, C, W, H, tw, th, ov, bs = 2, 4, 88, 56, 32, 24, 10, 3
    x = synth.latent(11, (N, C, H, W), DTYPES[dn])
    d, want = run_reference_step(ref, method, x, W, H, tw, th, ov, bs)
    plan = tiling.GridPlan(W, H, tw, th, ov, bs, method == "mod")
    assert [(b.x, b.y, b.w, b.h) for bb in d.batched_bboxes for b in bb] == plan.bboxes
    den = lambda t, bb: synth.fake_denoise(t, bb, N)
    if method == "md":
        got = blend.multidiffusion_step(x, plan.batched_bboxes, plan.weights, den)
    else:
        got = blend.mixture_step(x, plan.batched_bboxes, plan.tile_weights, plan.rescale_factor, den)
    assert_bit_equal(got, want, f"{method}/{dn}")


def test_scatter_matches_reference_cat(ref):
    x = synth.latent(5, (2, 4, 64, 80), torch.float16)
    bbs, _ = ref.utils.split_bboxes(80, 64, 24, 16, 6, 1.0)
    want = torch.cat([x[b.slicer] for b in bbs], dim=0)
    got = blend.scatter_tiles(x, [(b.x, b.y, b.w, b.h) for b in bbs])
    assert torch.equal(got, want)
