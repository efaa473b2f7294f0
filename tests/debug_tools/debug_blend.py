# Debug helper (test infrastructure, run by hand on a GPU box): may use oracle/ as the checker.
"""Debug helper (GPU box): where does a blend path differ from the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from multidiffusion_upscaler_for_automatic1111_b200 import engine
from oracle import blend, synth, tiling
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
CASES = [(2, 4, 72, 72, 24, 16, 6, 3), (2, 4, 64, 48, 16, 16, 8, 4), (3, 4, 96, 64, 40, 24, 4, 2), (2, 4, 128, 128, 96, 96, 48, 4), (2, 4, 512, 512, 96, 96, 48, 4)]
for (N, C, W, H, tw, th, ov, bs) in CASES:
    for use_rcp in (False, True):
        x = synth.latent(3, (N, C, H, W), torch.float16)
        g = engine.make_grid(W, H, tw, th, ov, bs)
        plan = tiling.GridPlan(W, H, tw, th, ov, bs, False)
        xd = x.cuda()
        tiles = engine.scatter_tiles(g, xd, flags=2)
        outs, off = [], 0
        for bbs in plan.batched_bboxes:
            outs.append(synth.fake_denoise(tiles[off * N:(off + len(bbs)) * N], bbs, N)); off += len(bbs)
        w = torch.from_numpy(plan.weights).cuda()
        rcp = torch.from_numpy(engine.exact_reciprocals(plan.weights)).cuda() if use_rcp else None
        got = engine.blend_multidiffusion(g, outs, N, C, g.tile_bs, w, xd.dtype, flags=flags, rcp_weights=rcp).cpu()
        want = blend.multidiffusion_step(x, plan.batched_bboxes, plan.weights, lambda t, bb: synth.fake_denoise(t, bb, N))
        bad = (got != want)
        print((N, C, W, H, tw, th, ov, bs), "rcp" if use_rcp else "ieee", "mismatches", int(bad.sum()), "of", bad.numel())
        if bad.any():
            idx = bad.nonzero()
            print("  planes", sorted(set((int(a), int(b)) for a, b in idx[:, :2].tolist()))[:8], "rows", sorted(set(idx[:, 2].tolist()))[:20], "cols", sorted(set(idx[:, 3].tolist()))[:40])
            i = idx[0].tolist(); print("  first", i, float(got[tuple(i)]), float(want[tuple(i)]))
