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


def _a1111_shared():
    try:
        from modules import shared as a1111_shared  # type: ignore
        if hasattr(a1111_shared, "state"):
            return a1111_shared
    except Exception:
        pass
    return None


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
    try:
        from modules import devices as a1111_devices  # type: ignore
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
    try:
        from modules.sd_samplers_kdiffusion import KDiffusionSampler  # type: ignore
        if isinstance(sampler, KDiffusionSampler):
            return True
    except Exception:
        pass
    names = _mro_names(sampler)
    if "KDiffusionSampler" in names:
        return True
    if names & {"CompVisSampler", "VanillaStableDiffusionSampler"}:
        return False
    return hasattr(sampler, "model_wrap_cfg")


def is_ddim_sampler(sampler) -> bool:
    try:
        from modules.sd_samplers_timesteps import CompVisSampler  # type: ignore
        if isinstance(sampler, CompVisSampler):
            return True
    except Exception:
        pass
    return bool(_mro_names(sampler) & {"CompVisSampler", "VanillaStableDiffusionSampler"})
