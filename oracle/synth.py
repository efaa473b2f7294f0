"""TEST INFRASTRUCTURE -- platform-stable synthetic inputs.

`torch.randn` on CPU goes through vectorised libm code whose last bits may differ
between hosts, so fixtures and full-size hash goldens are built from numpy PCG64
*integers* mapped onto exactly representable values: the same bytes on the build
container and on the GPU box.
"""
from __future__ import annotations

import numpy as np
import torch


def latent(seed: int, shape, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """Roughly N(0,1)-looking values k/256, |k| <= 1024 (exact in fp16/bf16? bf16 keeps 8 bits -> rounded once, deterministically)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    # sum of three uniforms -> bell-shaped, exactly representable in fp16 (11-bit significand)
    k = rng.integers(-341, 342, size=(3,) + tuple(shape), dtype=np.int32).sum(axis=0)
    x = torch.from_numpy((k.astype(np.float32) / 256.0))
    return x.to(dtype)


def case_seed(name: str, dtype_name: str) -> int:
    """Seed of a named fixture case (stable across processes, unlike hash())."""
    return sum(ord(ch) for ch in name) * 7 + len(dtype_name)


def tile_scale(x: int, y: int) -> float:
    """Per-tile power-of-two factor: makes each tile's contribution distinct, exact in every dtype."""
    return (0.5, 1.0, -1.0, 2.0, -0.5)[(x + 3 * y) % 5]


def fake_denoise(x_tile: torch.Tensor, bboxes, n_per_tile: int) -> torch.Tensor:
    """Deterministic stand-in for the UNet: tile i is multiplied by tile_scale(x_i, y_i).

    Exact on CPU and GPU alike (multiplication by +-2^k then one RN rounding), so the
    blend parity tests are bit-exact without shipping tile outputs in the fixtures.
    `bboxes` items are (x, y, w, h) tuples or objects with .x/.y.
    """
    scales = []
    for b in bboxes:
        bx, by = (b.x, b.y) if hasattr(b, "x") else (b[0], b[1])
        scales += [tile_scale(int(bx), int(by))] * n_per_tile
    s = torch.tensor(scales, dtype=torch.float32, device=x_tile.device).view(-1, 1, 1, 1)
    return (x_tile.float() * s).to(x_tile.dtype)


def fake_region_denoise(x_tile: torch.Tensor, region_id: int) -> torch.Tensor:
    """Deterministic stand-in for the UNet on a custom region (region prompt control): a per-region power-of-two
    factor, exact in every dtype on CPU and GPU alike."""
    scale = (0.5, -1.0, 2.0, 1.0, -0.5)[region_id % 5]
    return (x_tile.float() * scale).to(x_tile.dtype)
