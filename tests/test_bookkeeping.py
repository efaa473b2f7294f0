"""C-ABI host bookkeeping vs the oracle and vs the reference-generated fixtures (bit-exact)."""
import itertools
import os

import numpy as np
import pytest

from oracle import tiling


@pytest.fixture(scope="module")
def pkg():
    import multidiffusion_upscaler_for_automatic1111_b200 as m
    return m


def test_split_bboxes_matches_golden(pkg, golden_dir):
    from multidiffusion_upscaler_for_automatic1111_b200.tile_utils import utils
    g = np.load(os.path.join(golden_dir, "bboxes.npz"))
    off = 0
    for (w, h, tw, th, ov), n in zip(g["cases"], g["counts"]):
        got = utils.split_bboxes_xywh(int(w), int(h), int(tw), int(th), int(ov))
        assert got.shape == (n, 4)
        assert np.array_equal(got, g["xywh"][off:off + n]), (w, h, tw, th, ov)
        off += n
    assert off == len(g["xywh"])


def test_splitable_matches_golden(pkg, golden_dir):
    from multidiffusion_upscaler_for_automatic1111_b200.tile_utils import utils
    g = np.load(os.path.join(golden_dir, "bboxes.npz"))
    for (w, h, tw, th, ov, want) in g["splitable"]:
        assert utils.splitable(int(w), int(h), int(tw), int(th), int(ov)) == bool(want)


def test_split_bboxes_matches_oracle_dense_sweep(pkg):
    from multidiffusion_upscaler_for_automatic1111_b200.tile_utils import utils
    n = 0
    for (w, h) in [(97, 96), (96, 97), (511, 513), (1000, 40), (33, 33), (4096 // 8, 4096 // 8), (768, 768)]:
        for (tw, th, ov) in itertools.product([16, 31, 96, 128], [16, 40, 96], [0, 3, 12, 48, 64]):
            tw_, th_ = min(tw, w), min(th, h)
            ov_ = max(0, min(ov, min(tw, th) - 4))
            if tw_ <= ov_ or th_ <= ov_:
                continue
            want = np.array(tiling.split_bboxes(w, h, tw_, th_, ov_), np.int32)
            assert np.array_equal(utils.split_bboxes_xywh(w, h, tw_, th_, ov_), want)
            n += 1
    assert n > 200


def test_gaussian_weights_bit_exact(pkg, golden_dir):
    from multidiffusion_upscaler_for_automatic1111_b200.tile_utils import utils
    g = np.load(os.path.join(golden_dir, "gaussian.npz"))
    for key in g.files:
        tw, th = map(int, key[2:].split("x"))
        got = utils.gaussian_weights_np(tw, th)
        assert got.shape == (th, tw)
        assert np.array_equal(got.view(np.uint32), g[key].view(np.uint32)), key
        assert np.array_equal(tiling.gaussian_weights(tw, th).view(np.uint32), g[key].view(np.uint32)), key


def test_grid_plan_matches_golden(pkg, golden_dir):
    """init_grid_bbox state: clamps, tile list, re-balanced batches, weight canvas, MoD rescale."""
    import hashlib

    from multidiffusion_upscaler_for_automatic1111_b200 import engine
    from multidiffusion_upscaler_for_automatic1111_b200.tile_utils import utils
    g = np.load(os.path.join(golden_dir, "grid_plans.npz"))
    for i, (W, H, tw, th, ov, bs) in enumerate(g["cases"]):
        for method in ("md", "mod"):
            key = f"{i}_{method}"
            grid = engine.make_grid(int(W), int(H), int(tw), int(th), int(ov), int(bs))
            scal = np.array([grid.tile_w, grid.tile_h, grid.num_tiles, grid.num_batches, grid.tile_bs], np.int32)
            assert np.array_equal(scal, g[key + "_scalars"]), key
            assert np.array_equal(engine.grid_bboxes_xywh(grid), g[key + "_bboxes"]), key
            tile_w = utils.gaussian_weights_np(grid.tile_w, grid.tile_h) if method == "mod" else None
            weights = engine.grid_weights(grid, tile_w)
            plan = tiling.GridPlan(int(W), int(H), int(tw), int(th), int(ov), int(bs), method == "mod")
            assert np.array_equal(weights.view(np.uint32), plan.weights.view(np.uint32)), key
            assert [len(b) for b in plan.batched_bboxes] == list(g[key + "_batch_sizes"]), key
            if key + "_weights" in g.files:
                assert np.array_equal(weights.view(np.uint32), g[key + "_weights"].view(np.uint32)), key
                if method == "mod":
                    rf = engine.rescale_factor(weights)
                    assert np.array_equal(rf.view(np.uint32), g[key + "_rescale"].view(np.uint32)), key
                    assert np.array_equal(plan.rescale_factor.view(np.uint32), g[key + "_rescale"].view(np.uint32)), key
            else:
                want = bytes(g[key + "_weights_sha"]).decode()
                assert hashlib.sha256(weights.view(np.uint32).tobytes()).hexdigest() == want, key


def test_grid_init_error_behaviour(pkg):
    from multidiffusion_upscaler_for_automatic1111_b200 import _cabi, engine
    with pytest.raises(_cabi.TdError) as e:
        engine.make_grid(64, 64, 16, 16, 4, 0)  # tile_bs 0 (Python: ZeroDivisionError)
    assert e.value.status == _cabi.TD_ERR_INVALID_ARG
    with pytest.raises(_cabi.TdError) as e:
        engine.make_grid(4096, 4096, 16, 16, 12, 4)  # 1021 cols > TD_MAX_GRID_DIM
    assert e.value.status == _cabi.TD_ERR_UNSUPPORTED
    g = engine.make_grid(40, 30, 96, 96, 48, 4)  # tile larger than canvas: clamps to one tile
    assert (g.tile_w, g.tile_h, g.num_tiles, g.rows, g.cols) == (40, 30, 1, 1, 1)
