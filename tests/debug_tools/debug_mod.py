# Debug helper (test infrastructure, run by hand on a GPU box): may use oracle/ as the checker.
"""Debug helper (GPU box): where does the MoD blend differ from the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from multidiffusion_upscaler_for_automatic1111_b200 import engine
from oracle import blend, synth, tiling
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
CASES = [(2, 4, 72, 72, 24, 16, 6, 3), (2, 4, 64, 48, 16, 16, 8, 4), (3, 4, 96, 64, 40, 24, 4, 2), (2, 4, 128, 128, 96, 96, 48, 4), (2, 4, 512, 512, 96, 96, 48, 4)]
for (N, C, W, H, tw, th, ov, bs) in CASES:
    x = synth.latent(3, (N, C, H, W), torch.float16)
    g = engine.make_grid(W, H, tw, th, ov, bs)
    plan = tiling.GridPlan(W, H, tw, th, ov, bs, True)
    xd = x.cuda()
    tiles = engine.scatter_tiles(g, xd, flags=2)
    outs, off = [], 0
    for bbs in plan.batched_bboxes:
        outs.append(synth.fake_denoise(tiles[off * N:(off + len(bbs)) * N], bbs, N)); off += len(bbs)
    xb = torch.empty_like(xd)
    engine.blend_mixture(g, outs, N, C, g.tile_bs, torch.from_numpy(plan.tile_weights).cuda(), torch.from_numpy(plan.rescale_factor).cuda(), xb, flags=flags)
    got = xb.cpu()
    want = blend.mixture_step(x, plan.batched_bboxes, plan.tile_weights, plan.rescale_factor, lambda t, bb: synth.fake_denoise(t, bb, N))
    bad = torch.from_numpy(got.view(torch.int16).numpy() != want.view(torch.int16).numpy())   # bit compare (+-0 matter)
    print((N, C, W, H, tw, th, ov, bs), "mismatches", int(bad.sum()), "of", bad.numel(), "maxdiff", float((got.float() - want.float()).abs().max()))
    if bad.any():
        idx = bad.nonzero()
        print("  rows", sorted(set(idx[:, 2].tolist()))[:24], "cols", sorted(set(idx[:, 3].tolist()))[:48])
        i = idx[0].tolist(); print("  first", i, float(got[tuple(i)]), float(want[tuple(i)]))
