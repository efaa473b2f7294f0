"""Tiled noise inversion on the device: our MultiDiffusion delegate with the REAL kernels
(scatter, fused blend, region composite) replays the job of tests/noise_inverse_job.py; the inverted latent and the
combined noise must match the fixture the unmodified reference produced on CPU (oracle/make_noise_inverse_golden.py).
fp32 latents; the only non-bit-exact ingredient is the bilinear resize of the retouch mask (ATen CPU vs CUDA): 1e-5."""
import os
import types

import numpy as np
import pytest
import torch

import noise_inverse_job as job
from oracle import ref_shim

pytestmark = pytest.mark.gpu


@pytest.fixture()
def prompt_parser_stub():
    """The same deterministic stand-in for the WebUI's prompt parser that produced the fixture (oracle/ref_shim.py),
    installed as `modules.prompt_parser` / `modules.extra_networks` for the duration of a test."""
    import sys
    from multidiffusion_upscaler_for_automatic1111_b200 import host
    saved = {k: sys.modules.get(k) for k in ("modules", "modules.prompt_parser", "modules.extra_networks")}
    root = sys.modules.get("modules") or types.ModuleType("modules")
    pp = types.ModuleType("modules.prompt_parser")
    pp.get_multicond_learned_conditioning, pp.get_learned_conditioning = ref_shim.fake_multicond, ref_shim.fake_learned
    pp.reconstruct_multicond_batch, pp.reconstruct_cond_batch = ref_shim.fake_reconstruct_multicond, ref_shim.fake_reconstruct_cond
    en = types.ModuleType("modules.extra_networks")
    en.parse_prompts = lambda prompts: (list(prompts), {})
    en.activate = lambda p, data: None
    en.deactivate = lambda p, data: None
    root.prompt_parser, root.extra_networks = pp, en
    sys.modules.update({"modules": root, "modules.prompt_parser": pp, "modules.extra_networks": en})
    host._a1111_cache.clear()
    yield
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    host._a1111_cache.clear()


class KDiffusionSampler:      # duck-typed by host.is_kdiff_sampler (class name in the MRO)
    pass


@pytest.mark.parametrize("mode", job.MODES)
def test_noise_inversion_on_device_matches_reference_fixture(golden_dir, prompt_parser_stub, mode):
    from multidiffusion_upscaler_for_automatic1111_b200 import MultiDiffusion, host
    from multidiffusion_upscaler_for_automatic1111_b200.tile_utils import utils
    gold = np.load(os.path.join(golden_dir, "noise_inverse.npz"))
    dev = torch.device("cuda")
    state = types.SimpleNamespace(interrupted=False, sampling_step=0, sampling_steps=1, job_count=0, nextjob=lambda: None)
    model = types.SimpleNamespace(apply_model=job.fake_apply_model, cond_stage_key="txt", parameterization="eps",
                                  model=types.SimpleNamespace(conditioning_key="crossattn"))
    host.use_shared(types.SimpleNamespace(state=state, sd_model=model, batch_cond_uncond=True))
    try:
        with_regions, bg = mode != "grid", mode != "regions_only"
        settings = {i: r for i, r in enumerate(job.ROWS)}
        d, sampler, p, cache = job.make_job(MultiDiffusion, settings, bg, with_regions, KDiffusionSampler, utils.NoiseInverseCache, dev)
        res = sampler.sample_img2img(p, job.x0().to(dev), job.noise().to(dev), None, None)
        torch.cuda.synchronize()
        assert res[0] == "sampled" and cache["sets"] == 1
        xt, nz = cache["v"].xt.float().cpu().numpy(), res[2].float().cpu().numpy()
        assert np.allclose(xt, gold[f"{mode}_xt"], rtol=1e-5, atol=1e-5), f"{mode}: inverted latent differs, max {np.abs(xt - gold[f'{mode}_xt']).max()}"
        assert np.allclose(nz, gold[f"{mode}_noise"], rtol=1e-5, atol=1e-5), f"{mode}: combined noise differs, max {np.abs(nz - gold[f'{mode}_noise']).max()}"
    finally:
        host.use_shared(None)
