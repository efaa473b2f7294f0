"""ControlNet / StableSR tile caches: the per-batch hint tiles our delegate caches and
hands to the extension objects must equal the reference's, for k-diffusion and DDIM samplers, grid and custom bboxes.

CPU: the UNMODIFIED reference (oracle/ref_shim.py) next to our delegate, with the scatter kernel swapped for the
oracle's scatter (the kernel itself is pinned by tests/test_gpu_diffusion.py).  GPU: the same caches through
td_scatter_tiles on the x8-scaled tile plan, against plain slicing.
"""
import types

import pytest
import torch

from oracle import blend, ref_shim

W, H = 64, 48
ROWS = [(True, 0.1, 0.2, 0.5, 0.4, "a cat", "", "Background", 0.2, -1),
        (True, 0.4, 0.3, 0.45, 0.6, "a dog", "ugly", "Foreground", 0.3, 5)]


def _p():
    return types.SimpleNamespace(width=W * 8, height=H * 8, sampler_name="Euler a", disable_extra_networks=True,
                                 batch_size=1, steps=20, styles=None, all_prompts=["a photo"], all_negative_prompts=["blurry"])


def _hints(device="cpu"):
    a = (torch.arange(1 * 3 * H * 8 * W * 8, dtype=torch.float32, device=device).view(1, 3, H * 8, W * 8) % 1021) / 1021.0
    b = (torch.arange(3 * H * 8 * W * 8, dtype=torch.float32, device=device).view(3, H * 8, W * 8) % 509) / 509.0   # 3-d: unsqueezed in place
    return [a, b]


def _controlnet_script(device="cpu"):
    params = [types.SimpleNamespace(hint_cond=t) for t in _hints(device)]
    return types.SimpleNamespace(latest_network=types.SimpleNamespace(control_params=params))


def _oracle_scatter(monkeypatch):
    from multidiffusion_upscaler_for_automatic1111_b200 import engine

    def scatter_tiles(g, x, out=None, tile_begin=0, tile_end=None, flags=0):
        bbs = [tuple(int(v) for v in r) for r in engine.grid_bboxes_xywh(g)]
        return blend.scatter_tiles(x, bbs[tile_begin:tile_end])
    monkeypatch.setattr(engine, "scatter_tiles", scatter_tiles)


def _delegate(cls, sampler, settings, with_regions):
    d = cls(_p(), sampler)
    d.init_grid_bbox(16, 16, 8, 4)
    if with_regions:
        d.init_custom_bbox(settings, True, False)
    return d


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
@pytest.mark.parametrize("kdiff", [True, False], ids=["kdiff", "ddim"])
@pytest.mark.parametrize("with_regions", [False, True], ids=["grid", "grid+regions"])
@pytest.mark.parametrize("tensor_cpu", [False, True], ids=["dev", "cpu_cache"])
def test_controlnet_tile_caches_like_the_reference(monkeypatch, kdiff, with_regions, tensor_cpu):
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion, host
    ref = ref_shim.load()
    host._a1111_cache.clear()
    _oracle_scatter(monkeypatch)

    def sampler():
        if kdiff:
            return ref_shim.make_kdiff_sampler(lambda x, s, cond=None: x)

        class _S(ref.CompVisSampler):
            pass
        s = _S()
        s.model_wrap_cfg = types.SimpleNamespace(inner_model=types.SimpleNamespace(forward=None), image_cfg_scale=None, step=0)
        return s

    d_ref = _delegate(ref.multidiffusion.MultiDiffusion, sampler(), {i: ref.utils.BBoxSettings(*r) for i, r in enumerate(ROWS)}, with_regions)
    d_our = _delegate(MultiDiffusion, sampler(), {i: r for i, r in enumerate(ROWS)}, with_regions)
    assert d_our.is_kdiff == kdiff == d_ref.is_kdiff
    scripts = []
    for d in (d_ref, d_our):
        cs = _controlnet_script()
        scripts.append(cs)
        d.init_controlnet(cs, tensor_cpu)
        d.init_done()
        if getattr(d, "pbar", None) is not None:
            d.pbar.disable = True
    cs_ref, cs_our = scripts

    assert len(d_our.control_tensor_batch) == len(d_ref.control_tensor_batch) == 2
    for pr, po in zip(d_ref.control_tensor_batch, d_our.control_tensor_batch):
        assert len(pr) == len(po) == d_ref.num_batches
        for tr, to in zip(pr, po):
            assert tr.shape == to.shape and torch.equal(tr, to)
    if with_regions:
        for pr, po in zip(d_ref.control_tensor_custom, d_our.control_tensor_custom):
            for tr, to in zip(pr, po):
                assert torch.equal(tr, to)

    for batch_id in (0, d_ref.num_batches - 1):
        n_tiles = len(d_ref.batched_bboxes[batch_id])
        for is_denoise in (False, True):
            for d in (d_ref, d_our):
                d.switch_controlnet_tensors(batch_id, 2, n_tiles, is_denoise=is_denoise)
            for a, b in zip(cs_ref.latest_network.control_params, cs_our.latest_network.control_params):
                assert a.hint_cond.shape == b.hint_cond.shape and torch.equal(a.hint_cond, b.hint_cond)
    if with_regions:
        for d in (d_ref, d_our):
            d.set_custom_controlnet_tensors(1, 3)
        for a, b in zip(cs_ref.latest_network.control_params, cs_our.latest_network.control_params):
            assert torch.equal(a.hint_cond, b.hint_cond)
    for d in (d_ref, d_our):
        d.reset_controlnet_tensors()
    for a, b in zip(cs_ref.latest_network.control_params, cs_our.latest_network.control_params):
        assert a.hint_cond.shape == (1, 3, H * 8, W * 8) and torch.equal(a.hint_cond, b.hint_cond)
    host._a1111_cache.clear()


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_stablesr_tile_caches_like_the_reference(monkeypatch):
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion, host
    ref = ref_shim.load()
    host._a1111_cache.clear()
    _oracle_scatter(monkeypatch)
    latent = torch.arange(2 * 4 * H * W, dtype=torch.float32).view(2, 4, H, W)
    models = []
    for cls, settings in ((ref.multidiffusion.MultiDiffusion, {i: ref.utils.BBoxSettings(*r) for i, r in enumerate(ROWS)}),
                          (MultiDiffusion, {i: r for i, r in enumerate(ROWS)})):
        d = _delegate(cls, ref_shim.make_kdiff_sampler(lambda x, s, cond=None: x), settings, True)
        model = types.SimpleNamespace(set_image_hooks={}, latent_image=None)
        d.init_stablesr(types.SimpleNamespace(stablesr_model=model))
        d.init_done()
        if getattr(d, "pbar", None) is not None:
            d.pbar.disable = True
        model.set_image_hooks["TiledDiffusion"](latent)
        models.append((d, model))
    (d_ref, m_ref), (d_our, m_our) = models
    assert d_our.enable_stablesr and len(d_our.stablesr_tensor_batch) == len(d_ref.stablesr_tensor_batch)
    for b in range(d_ref.num_batches):
        for d in (d_ref, d_our):
            d.switch_stablesr_tensors(b)
        assert m_ref.latent_image.shape == m_our.latent_image.shape and torch.equal(m_ref.latent_image, m_our.latent_image)
    for d in (d_ref, d_our):
        d.set_custom_stablesr_tensors(1)
    assert torch.equal(m_ref.latent_image, m_our.latent_image)
    for d in (d_ref, d_our):
        d.reset_stablesr_tensors()
    assert m_our.latent_image is latent and m_ref.latent_image is latent
    host._a1111_cache.clear()


def test_scaled_grid_is_the_tile_plan_times_opt_f():
    from multidiffusion_upscaler_for_automatic1111_b200 import engine
    g = engine.make_grid(W, H, 16, 16, 8, 4)
    s = engine.scaled_grid(g, 8)
    assert (s.H, s.W, s.tile_h, s.tile_w, s.rows, s.cols, s.num_tiles, s.tile_bs) == (H * 8, W * 8, 128, 128, g.rows, g.cols, g.num_tiles, g.tile_bs)
    assert (engine.grid_bboxes_xywh(s) == engine.grid_bboxes_xywh(g) * 8).all()
