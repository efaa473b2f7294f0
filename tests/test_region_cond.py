"""Region prompt control, conditioning side: `kdiff_custom_forward` / `ddim_custom_forward` must hand the UNet exactly
the rows, sigmas and cond tensors the reference hands it, for every way the k-diffusion CFG wrapper slices a batch.

Runs the UNMODIFIED reference (oracle/ref_shim.py, with its deterministic stand-in for the WebUI prompt parser) next
to our delegate and compares the recorded UNet calls; needs /root/reference, so it is skipped on the GPU box.
"""
import itertools
import types

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")

W, H = 64, 48
ROWS = [(True, 0.1, 0.2, 0.5, 0.4, "a cat", "", "Background", 0.2, -1),
        (True, 0.4, 0.3, 0.45, 0.6, "a very long region prompt that spills into a second chunk of tokens", "ugly", "Foreground", 0.3, 5)]
SHORT, LONG = "a photo", "a photo of something described with so many words that the encoder needs two chunks"


@pytest.fixture()
def hosts():
    from multidiffusion_upscaler_for_automatic1111_b200 import host
    ref = ref_shim.load()
    host._a1111_cache.clear()          # pick up the stub `modules.*` the shim just installed
    keep = ref.shared.batch_cond_uncond
    yield ref, host
    ref.shared.batch_cond_uncond = keep
    host._a1111_cache.clear()


def _p(prompt, neg, batch_size):
    return types.SimpleNamespace(width=W * 8, height=H * 8, sampler_name="Euler a", disable_extra_networks=True,
                                 batch_size=batch_size, steps=20, styles=None,
                                 all_prompts=[f"{prompt} {i}" for i in range(batch_size)],
                                 all_negative_prompts=[f"{neg} {i}" for i in range(batch_size)])


def _pair(ref, prompt, neg, batch_size, edit):
    """(reference delegate, our delegate) with the same regions; both on CPU."""
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion
    out = []
    for cls, settings in ((ref.multidiffusion.MultiDiffusion, {i: ref.utils.BBoxSettings(*r) for i, r in enumerate(ROWS)}),
                          (MultiDiffusion, {i: r for i, r in enumerate(ROWS)})):
        sampler = ref_shim.make_kdiff_sampler(lambda x, s, cond=None: x)
        d = cls(_p(prompt, neg, batch_size), sampler)
        d.init_grid_bbox(16, 16, 8, 4)
        d.init_custom_bbox(settings, True, False)
        d.init_done()
        if getattr(d, "pbar", None) is not None:
            d.pbar.disable = True
        d.is_edit_model = edit
        out.append(d)
    return out


class _Recorder:
    def __init__(self):
        self.calls = []

    def __call__(self, x, sigma, cond=None):
        tc = cond["c_crossattn"][0]
        ic = cond["c_concat"][0]
        self.calls.append((x.clone(), sigma.clone(), tc.clone(), ic.clone()))
        return x * 2 + tc.mean() + sigma.view(-1, 1, 1, 1)


def _drive(d, chunks, steps=(0, 1)):
    """Feed every region the virtual batch in `chunks` (row counts), for two sampler steps."""
    rec = _Recorder()
    outs = []
    rows = sum(chunks)
    for step in steps:
        d.sampler.model_wrap_cfg.step = step
        for bbox_id, bbox in enumerate(d.custom_bboxes):
            x_full = torch.arange(rows * 4 * bbox.h * bbox.w, dtype=torch.float32).view(rows, 4, bbox.h, bbox.w) / 1000.0 + step
            sigma = torch.arange(rows, dtype=torch.float32) + 1
            cond = {"c_crossattn": [torch.zeros(rows, 77, 8)], "c_concat": [torch.full((rows, 5, 1, 1), 0.5)]}
            lo = 0
            for n in chunks:
                outs.append(d.kdiff_custom_forward(x_full[lo:lo + n], sigma[lo:lo + n], cond, bbox_id, bbox, rec).clone())
                lo += n
    return rec.calls, outs


def _same(a, b):
    assert len(a) == len(b)
    for ta, tb in zip(a, b):
        if isinstance(ta, tuple):
            _same(ta, tb)
        else:
            assert ta.shape == tb.shape and torch.equal(ta, tb)


SCENARIOS = []
for batch_size, (prompt, neg), edit, bcu in itertools.product([1, 2], [(SHORT, SHORT), (LONG, SHORT)], [False, True], [True, False]):
    rows = batch_size * (3 if edit else 2)
    splits = {(rows,), tuple([1] * rows), (batch_size,) * (rows // batch_size)}
    if rows >= 3:
        splits.add((rows - 1, 1))
        splits.add((1, rows - 1))
    for chunks in sorted(splits):
        if bcu and len(chunks) > 1:
            continue                      # with batch_cond_uncond the wrapper always sends the whole batch at once
        SCENARIOS.append((batch_size, prompt, neg, edit, bcu, chunks))


@pytest.mark.parametrize("sc", SCENARIOS, ids=[f"bs{s[0]}_{'long' if s[1] is LONG else 'short'}_{'edit' if s[3] else 'std'}_"
                                               f"{'bcu' if s[4] else 'seq'}_{'-'.join(map(str, s[5]))}" for s in SCENARIOS])
def test_kdiff_custom_forward_feeds_the_unet_like_the_reference(hosts, sc):
    ref, host = hosts
    batch_size, prompt, neg, edit, bcu, chunks = sc
    ref.shared.batch_cond_uncond = bcu
    d_ref, d_ours = _pair(ref, prompt, neg, batch_size, edit)
    try:
        want_calls, want_outs = _drive(d_ref, chunks)
    except Exception as e:           # slicings the reference itself cannot serve (it raises): nothing to compare
        pytest.skip(f"reference raises {type(e).__name__} here")
    got_calls, got_outs = _drive(d_ours, chunks)
    _same(got_calls, want_calls)
    _same(got_outs, want_outs)


@pytest.mark.parametrize("prompt,neg", [(SHORT, SHORT), (LONG, SHORT), (SHORT, LONG)])
def test_ddim_custom_forward_like_the_reference(hosts, prompt, neg):
    ref, host = hosts
    d_ref, d_ours = _pair(ref, prompt, neg, 1, False)
    res = []
    for d in (d_ref, d_ours):
        d.sampler.model_wrap_cfg.step = 3
        calls = []

        def fwd(x, cond, ts, unconditional_conditioning=None):
            calls.append((x.clone(), ts.clone(), cond["c_crossattn"][0].clone(), unconditional_conditioning["c_crossattn"][0].clone(),
                          cond["c_concat"][0].clone()))
            return x + 1
        outs = []
        for bbox in d.custom_bboxes:
            x = torch.ones(1, 4, bbox.h, bbox.w)
            cond_in = {"c_crossattn": [torch.zeros(1, 77, 8)], "c_concat": [torch.arange(5 * H * W, dtype=torch.float32).view(1, 5, H, W)]}
            outs.append(d.ddim_custom_forward(x, cond_in, bbox, torch.tensor([7]), fwd))
        res.append((calls, outs))
    _same(res[1][0], res[0][0])
    _same(res[1][1], res[0][1])
