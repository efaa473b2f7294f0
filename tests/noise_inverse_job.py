"""The synthetic noise-inversion job shared by oracle/make_noise_inverse_golden.py (unmodified reference, CPU) and
tests/test_gpu_noise_inverse.py (our delegate, real kernels): stand-in k-diffusion wrapper, sampler, processing object."""
import types

import numpy as np
import torch

W, H = 64, 48
MODES = ["grid", "grid+regions", "regions_only"]
ROWS = [(True, 0.1, 0.2, 0.5, 0.4, "a cat", "", "Background", 0.2, -1),
        (True, 0.4, 0.3, 0.45, 0.6, "a dog", "ugly", "Foreground", 0.3, 5)]


class DNW:
    """CompVisDenoiser look-alike: the three calls find_noise_for_image_sigma_adjustment makes."""

    def get_sigmas(self, n):
        return torch.cat([torch.linspace(14.0, 0.5, n), torch.zeros(1)])

    def get_scalings(self, sigma):
        return -sigma, 1 / (sigma ** 2 + 1) ** 0.5

    def sigma_to_t(self, sigma):
        return sigma * 7 + 1


def fake_apply_model(x, t, cond=None):
    return x * 0.5 + t.view(-1, 1, 1, 1) * 1e-3 + cond["c_crossattn"][0].mean() * 0.1


def image():
    from PIL import Image
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=(H * 8, W * 8, 3)).astype(np.uint8)
    a[: H * 4] = (a[: H * 4] // 64) * 64          # a flat half and a busy half
    return Image.fromarray(a)


def x0():
    return torch.zeros(1, 4, H, W)


def noise():
    return torch.randn(1, 4, H, W, generator=torch.Generator().manual_seed(9))


def make_job(cls, settings, draw_background, with_regions, kdiff_base, cache_type, device):
    sd_model = types.SimpleNamespace(sd_model_hash="abc", get_learned_conditioning=lambda prompts: torch.full((len(prompts), 77, 8), 0.25, device=device))
    g = torch.Generator().manual_seed(5)
    p = types.SimpleNamespace(width=W * 8, height=H * 8, sampler_name="Euler", disable_extra_networks=True, batch_size=1, steps=6,
                              styles=None, all_prompts=["a photo"], all_negative_prompts=["blurry"], sd_model=sd_model,
                              init_latent=torch.randn(1, 4, H, W, generator=g).to(device), image_conditioning=torch.zeros(1, 5, 1, 1, device=device),
                              init_images=[image()], show_tile_progress=False)

    class _Sampler(kdiff_base):
        pass
    sampler = _Sampler()
    sampler.model_wrap = DNW()
    sampler.model_wrap_cfg = types.SimpleNamespace(inner_model=types.SimpleNamespace(forward=None), image_cfg_scale=None, step=0)
    sampler.get_sigmas = lambda p_, steps: torch.linspace(9.0, 0.0, steps + 1).to(device)
    sampler.sample_img2img = lambda p_, x, noise_, c, uc, steps=None, image_conditioning=None: ("sampled", x, noise_, steps)
    cache = {"v": None, "sets": 0}

    def get_cache():
        return cache["v"]

    def set_cache(x0_, xt, prompts):
        cache["sets"] += 1
        cache["v"] = cache_type("abc", x0_, xt, 4, 1.5, prompts)

    d = cls(p, sampler)
    d.init_grid_bbox(16, 16, 8, 4)
    if with_regions:
        d.init_custom_bbox(settings, draw_background, False)
    d.init_noise_inverse(4, 1.5, get_cache, set_cache, 1.0, 8)
    d.init_done()
    if getattr(d, "pbar", None) is not None:
        d.pbar.disable = True
    return d, sampler, p, cache
