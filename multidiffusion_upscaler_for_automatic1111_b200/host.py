"""Adapter to the host application (AUTOMATIC1111 WebUI) -- or its absence.

Inside the WebUI the reference reads three globals: `modules.shared.state`
(cooperative cancel, tile_methods/multidiffusion.py:152), `modules.shared.sd_model`
(patched by Mixture of Diffusers, mixtureofdiffusers.py:18-27) and
`modules.devices.device`.  This module resolves them lazily from A1111 when it is
importable and otherwise from a small standalone namespace that a caller (tests,
bench, another host) fills in: `host.shared.sd_model = my_model`.
"""
from __future__ import annotations

import types

import torch

opt_f = 8  # latent down-scale factor (modules.processing.opt_f)


class _StandaloneState:
    interrupted = False
    sampling_step = 0
    sampling_steps = 1


_standalone_shared = types.SimpleNamespace(state=_StandaloneState(), sd_model=None)
_forced_shared = None


def use_shared(ns) -> None:
    """Force a specific `shared`-like namespace (tests use this to share the stub host)."""
    global _forced_shared
    _forced_shared = ns


_a1111_cache = {}


def _a1111(name: str):
    """`modules.<name>` of the WebUI if importable; looked up once (a failing import is slow)."""
    if name not in _a1111_cache:
        try:
            import importlib
            _a1111_cache[name] = importlib.import_module(f"modules.{name}")
        except Exception:
            _a1111_cache[name] = None
    return _a1111_cache[name]


def a1111_module(name: str):
    """Public form of `_a1111` for the region-prompt helpers (prompt_parser, extra_networks)."""
    return _a1111(name)


def _a1111_shared():
    m = _a1111("shared")
    return m if m is not None and hasattr(m, "state") else None


def get_shared():
    if _forced_shared is not None:
        return _forced_shared
    return _a1111_shared() or _standalone_shared


class _SharedProxy:
    def __getattr__(self, name):
        return getattr(get_shared(), name)

    def __setattr__(self, name, value):
        setattr(get_shared(), name, value)


shared = _SharedProxy()


def interrupted() -> bool:
    st = getattr(get_shared(), "state", None)
    return bool(getattr(st, "interrupted", False))


def device() -> torch.device:
    """Device the delegate keeps its persistent buffers on (CUDA whenever one exists)."""
    a1111_devices = _a1111("devices")
    if a1111_devices is not None:
        try:
            d = torch.device(a1111_devices.device)
            if d.type == "cuda":
                return d
        except Exception:
            pass
    if not torch.cuda.is_available():
        # bookkeeping tensors (weight canvases) may live on the host; every compute
        # entry point still refuses CPU tensors -- there is no CPU fallback.
        return torch.device("cpu")
    return torch.device("cuda", torch.cuda.current_device())


def _mro_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


def is_kdiff_sampler(sampler) -> bool:
    """isinstance(sampler, KDiffusionSampler) (abstractdiffusion.py:77-79), duck-typed without A1111."""
    m = _a1111("sd_samplers_kdiffusion")
    if m is not None and hasattr(m, "KDiffusionSampler") and isinstance(sampler, m.KDiffusionSampler):
        return True
    names = _mro_names(sampler)
    if "KDiffusionSampler" in names:
        return True
    if names & {"CompVisSampler", "VanillaStableDiffusionSampler"}:
        return False
    return hasattr(sampler, "model_wrap_cfg")


def is_ddim_sampler(sampler) -> bool:
    m = _a1111("sd_samplers_timesteps")
    if m is not None and hasattr(m, "CompVisSampler") and isinstance(sampler, m.CompVisSampler):
        return True
    return bool(_mro_names(sampler) & {"CompVisSampler", "VanillaStableDiffusionSampler"})


def batch_cond_uncond() -> bool:
    """`shared.batch_cond_uncond` (abstractdiffusion.py:276); True outside the WebUI."""
    return bool(getattr(get_shared(), "batch_cond_uncond", True))


def extra_networks_activate(p, data) -> None:
    """extra_networks.activate under autocast (multidiffusion.py:178-180); no-op outside the WebUI."""
    m, d = _a1111("extra_networks"), _a1111("devices")
    if m is None or not hasattr(m, "activate"):
        return
    if d is not None and hasattr(d, "autocast"):
        with d.autocast():
            m.activate(p, data)
    else:
        m.activate(p, data)


def extra_networks_deactivate(p, data) -> None:
    m, d = _a1111("extra_networks"), _a1111("devices")
    if m is None or not hasattr(m, "deactivate"):
        return
    if d is not None and hasattr(d, "autocast"):
        with d.autocast():
            m.deactivate(p, data)
    else:
        m.deactivate(p, data)


def setup_img2img_steps(p, steps=None):
    """modules.sd_samplers_common.setup_img2img_steps; without the WebUI: its plain rule (steps, t_enc)."""
    m = _a1111("sd_samplers_common")
    if m is not None and hasattr(m, "setup_img2img_steps"):
        return m.setup_img2img_steps(p, steps)
    steps = steps if steps is not None else p.steps
    return steps, int(min(getattr(p, "denoising_strength", 1.0), 0.999) * steps)


def store_latent(x) -> None:
    """Live-preview hook of the WebUI (sd_samplers_common.store_latent); no-op elsewhere."""
    m = _a1111("sd_samplers_common")
    if m is not None and hasattr(m, "store_latent"):
        m.store_latent(x)


cheap_approximation = None   # a caller outside the WebUI may set this (latent [4,h,w] -> RGB [3,h,w])


def get_cheap_approximation():
    """modules.sd_vae_approx.cheap_approximation (tilevae.py:575), the override above, or None."""
    if cheap_approximation is not None:
        return cheap_approximation
    m = _a1111("sd_vae_approx")
    return getattr(m, "cheap_approximation", None) if m is not None else None
