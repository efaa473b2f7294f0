"""Host-side delegate logic that needs no GPU: construction, bookkeeping fields, error behaviour."""
import types

import pytest
import torch


def _p(w=1024, h=1024, sampler_name="Euler a"):
    return types.SimpleNamespace(width=w, height=h, sampler_name=sampler_name)


def _sampler():
    inner = types.SimpleNamespace(forward=lambda x, s, cond=None: x)
    return types.SimpleNamespace(model_wrap_cfg=types.SimpleNamespace(inner_model=inner, image_cfg_scale=None))


def test_multidiffusion_fields_like_reference():
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    d = MultiDiffusion(_p(4096, 4096), _sampler())
    d.init_grid_bbox(96, 96, 48, 4)
    d.init_done()
    assert (d.w, d.h, d.tile_w, d.tile_h) == (512, 512, 96, 96)
    assert (d.num_tiles, d.num_batches, d.tile_bs) == (100, 25, 4)
    assert len(d.batched_bboxes) == 25 and all(len(b) == 4 for b in d.batched_bboxes)
    assert d.weights.shape == (1, 1, 512, 512) and d.weights.dtype == torch.float32
    assert float(d.weights.max()) == 9.0 and float(d.weights.min()) == 1.0
    b = d.batched_bboxes[0][1]
    assert b.box == [46, 0, 142, 96] and b.slicer[3] == slice(46, 142)
    assert d.sampler_raw is d.sampler and d.method == "MultiDiffusion"


def test_tile_bs_is_rebalanced():
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    d = MultiDiffusion(_p(4096, 4096), _sampler())
    d.init_grid_bbox(96, 96, 48, 8)  # 100 tiles / 8 -> 13 batches -> tile_bs 8 (ceil(100/13))
    assert (d.num_batches, d.tile_bs) == (13, 8)
    assert [len(b) for b in d.batched_bboxes] == [8] * 12 + [4]


def test_unipc_rejected_and_nothing_to_paint():
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    with pytest.raises(AssertionError):
        MultiDiffusion(_p(sampler_name="UniPC"), _sampler())
    d = MultiDiffusion(_p(), _sampler())
    with pytest.raises(AssertionError, match="Nothing to paint"):
        d.init_done()


def test_hook_patches_inner_model_forward():
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    s = _sampler()
    orig = s.model_wrap_cfg.inner_model.forward
    d = MultiDiffusion(_p(), s)
    d.init_grid_bbox(96, 96, 48, 4)
    d.init_done()
    d.hook()
    assert d.sampler_forward is orig
    assert s.model_wrap_cfg.inner_model.forward == d.kdiff_forward


def test_mixture_hook_unhook_roundtrip():
    from multidiffusion_upscaler_for_automatic1111_b200 import MixtureOfDiffusers, host
    model = types.SimpleNamespace(apply_model=lambda x, t, c: x, model=types.SimpleNamespace(conditioning_key="crossattn"),
                                  cond_stage_key="txt")
    orig = model.apply_model
    host.use_shared(types.SimpleNamespace(state=types.SimpleNamespace(interrupted=False, sampling_step=0, sampling_steps=1),
                                          sd_model=model))
    try:
        d = MixtureOfDiffusers(_p(), _sampler())
        d.init_grid_bbox(96, 96, 48, 4)
        d.init_done()
        assert d.rescale_factor.shape == (1, 1, 128, 128)
        d.hook()
        assert model.apply_model == d.apply_model_hijack and model.apply_model_original_md is orig
        MixtureOfDiffusers.unhook()
        assert model.apply_model is orig and not hasattr(model, "apply_model_original_md")
    finally:
        host.use_shared(None)


def test_hires_pass_is_not_tiled():
    """(H, W) != (self.h, self.w): the original forward runs untiled (multidiffusion.py:141-144)."""
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    d = MultiDiffusion(_p(), _sampler())
    d.init_grid_bbox(96, 96, 48, 4)
    d.init_done()
    x = torch.zeros(2, 4, 64, 64)
    out = d.sample_one_step(x, lambda t: t + 1, None, None)
    assert torch.equal(out, x + 1)


def test_cpu_tensor_is_refused_not_emulated():
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    d = MultiDiffusion(_p(), _sampler())
    d.init_grid_bbox(96, 96, 48, 4)
    d.init_done()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.sample_one_step(torch.zeros(2, 4, 128, 128), None, lambda t, b: t, None)


def test_empty_region_table_switches_region_control_off_again():
    from multidiffusion_upscaler_for_automatic1111_b200 import AbstractDiffusion, MultiDiffusion
    d = MultiDiffusion(_p(), _sampler())
    d.init_custom_bbox({}, True, False)
    assert d.enable_custom_bbox is False and d.custom_bboxes == []
    with pytest.raises(NotImplementedError):      # abstract in the base class, like the reference (abstractdiffusion.py:746)
        AbstractDiffusion.get_noise(d, None, None, None, 0)


def test_repeat_tensor_semantics():
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    d = MultiDiffusion(_p(), _sampler())
    a = torch.arange(6.0).view(1, 2, 3)
    r = d.repeat_tensor(a, 4)
    assert r.shape == (4, 2, 3) and r.data_ptr() == a.data_ptr()  # expand, not copy
    b = torch.arange(4.0).view(2, 2)
    assert torch.equal(d.repeat_tensor(b, 3), b.repeat(3, 1))
    assert d.repeat_tensor(b, 1) is b


def test_cond_memo_follows_the_source_tensors():
    """repeat_cond_dict is memoised per (cond dict, repeat count): every batch gets its own dict object; the memo is
    re-validated once per sampler step and dropped when a source tensor was modified in place or replaced."""
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    from multidiffusion_upscaler_for_automatic1111_b200.tile_utils.utils import BBox
    d = MultiDiffusion(_p(), _sampler())
    d.init_grid_bbox(96, 96, 48, 4)
    d.init_done()
    bbs = [BBox(0, 0, 96, 96)] * 3
    t = torch.arange(2 * 3 * 4, dtype=torch.float32).view(2, 3, 4)
    cond = {"c_crossattn": [t], "c_concat": [torch.zeros(2, 5, 1, 1)]}
    a, b = d.repeat_cond_dict(cond, bbs), d.repeat_cond_dict(cond, bbs)
    assert a is not b and a["c_crossattn"][0] is b["c_crossattn"][0]          # fresh dict, shared (read-only) tensors
    assert torch.equal(a["c_crossattn"][0], t.repeat(3, 1, 1)) and a["c_concat"][0].shape[0] == 6
    t.add_(1)                                                                  # in-place update between sampler steps
    d._step_token = d.__dict__.get("_step_token", 0) + 1                       # (sample_one_step bumps the token once per step)
    c = d.repeat_cond_dict(cond, bbs)
    assert torch.equal(c["c_crossattn"][0], t.repeat(3, 1, 1)) and c["c_crossattn"][0] is not a["c_crossattn"][0]
    cond["c_crossattn"] = [t * 2]                                              # replaced tensor in the same dict
    d._step_token += 1
    e = d.repeat_cond_dict(cond, bbs)
    assert torch.equal(e["c_crossattn"][0], (t * 2).repeat(3, 1, 1))
    assert d.repeat_cond_dict(cond, bbs[:2])["c_crossattn"][0].shape[0] == 4   # another repeat count


def test_cat_repeat_memo():
    from multidiffusion_upscaler_for_automatic1111_b200 import MixtureOfDiffusers, host
    model = types.SimpleNamespace(apply_model=lambda x, t, c: x, model=types.SimpleNamespace(conditioning_key="crossattn"), cond_stage_key="txt")
    host.use_shared(types.SimpleNamespace(state=types.SimpleNamespace(interrupted=False, sampling_step=0, sampling_steps=1), sd_model=model))
    try:
        d = MixtureOfDiffusers(_p(), _sampler())
        x = torch.arange(6.0).view(2, 3)
        r = d.cat_repeat(x, 3)
        assert torch.equal(r, torch.cat([x] * 3)) and d.cat_repeat(x, 3) is r and d.cat_repeat(x, 1) is x
        x.mul_(2)
        assert torch.equal(d.cat_repeat(x, 3), torch.cat([x] * 3))
    finally:
        host.use_shared(None)


def test_more_batches_than_pointer_slots_are_merged_in_order():
    """441 tiles with tile_bs 3 = 147 UNet output tensors > TD_MAX_BATCH_PTRS: consecutive batches are concatenated
    (same tile order) instead of failing mid-sampling; the reference accepts any batch count."""
    import torch
    from multidiffusion_upscaler_for_automatic1111_b200 import _cabi, engine
    T, bs, N = 441, 3, 2
    outs = [torch.full((min(bs, T - b * bs) * N, 1, 2, 2), float(b)) for b in range(-(-T // bs))]
    merged, new_bs = engine._fit_batch_table(None, outs, bs)
    assert len(merged) <= _cabi.TD_MAX_BATCH_PTRS and new_bs % bs == 0
    assert sum(m.shape[0] for m in merged) == T * N
    assert all(m.shape[0] == new_bs * N for m in merged[:-1])
    assert torch.equal(torch.cat(merged), torch.cat(outs))
    same, same_bs = engine._fit_batch_table(None, outs[:10], bs)
    assert same_bs == bs and len(same) == 10
