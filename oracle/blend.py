"""TEST INFRASTRUCTURE -- oracle for the per-step scatter / blend / normalise.

torch-CPU restatement (explicit per-tile loops, explicit dtype round trips) of

  scatter                 tile_methods/multidiffusion.py:155, mixtureofdiffusers.py:82-104
  blend-accumulate (MD)   tile_methods/multidiffusion.py:166-167
  normalise (MD)          tile_methods/multidiffusion.py:208
  blend (MoD)             tile_methods/mixtureofdiffusers.py:122-126, :169/:179
  reset_buffer            tile_methods/abstractdiffusion.py:97-102

Numerics that matter: the canvas accumulates in
`x_in.dtype` with one rounding per tile add, in tile-list order; MultiDiffusion's
divide promotes to fp32 (`weights` is fp32) and only fires where weights > 1;
MoD multiplies by `tile_weights * rescale_factor[slicer]` (fp32 product, its own
rounding), then adds in fp32 and rounds to `x_in.dtype`; MoD returns the buffer.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch

BBox = Tuple[int, int, int, int]  # x, y, w, h


def scatter_tiles(x_in: torch.Tensor, bboxes: Sequence[BBox]) -> torch.Tensor:
    """[N,C,H,W] -> [len(bboxes)*N, C, th, tw], tile-major (multidiffusion.py:155)."""
    return torch.cat([x_in[:, :, y:y + h, x:x + w] for (x, y, w, h) in bboxes], dim=0)


def accumulate_md(x_buffer: torch.Tensor, tile_out: torch.Tensor, bboxes: Sequence[BBox], N: int) -> None:
    """multidiffusion.py:166-167, op for op: `x_buffer[slicer] += tile` (in-place add in the buffer dtype; torch
    rounds the exact sum once, i.e. half(float(a) + float(b)) for fp16 / bf16)."""
    for i, (x, y, w, h) in enumerate(bboxes):
        x_buffer[:, :, y:y + h, x:x + w] += tile_out[i * N:(i + 1) * N, :, :, :]


def normalise_md(x_buffer: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """multidiffusion.py:208.  weights: fp32 [1,1,H,W].  Result is fp32."""
    return torch.where(weights > 1, x_buffer / weights, x_buffer)


def accumulate_mod(x_buffer: torch.Tensor, tile_out: torch.Tensor, bboxes: Sequence[BBox], N: int,
                   tile_weights: torch.Tensor, rescale_factor: torch.Tensor) -> None:
    """mixtureofdiffusers.py:125-126, op for op."""
    for i, (x, y, w, h) in enumerate(bboxes):
        wgt = tile_weights * rescale_factor[:, :, y:y + h, x:x + w]          # fp32 product (own rounding)
        x_buffer[:, :, y:y + h, x:x + w] += tile_out[i * N:(i + 1) * N, :, :, :] * wgt   # fp32 product, add, round to buffer dtype


def multidiffusion_step(x_in: torch.Tensor, batched_bboxes: List[List[BBox]], weights: np.ndarray,
                        denoise: Callable[[torch.Tensor, Sequence[BBox]], torch.Tensor]) -> torch.Tensor:
    """multidiffusion.py:131-218 (grid part): reset, scatter, denoise, accumulate, normalise."""
    N = x_in.shape[0]
    x_buffer = torch.zeros_like(x_in)
    for bboxes in batched_bboxes:
        x_tile = scatter_tiles(x_in, bboxes)
        out = denoise(x_tile, bboxes)
        accumulate_md(x_buffer, out, bboxes, N)
    w = torch.from_numpy(weights).view(1, 1, *weights.shape).to(x_in.device)
    return normalise_md(x_buffer, w)


def mixture_step(x_in: torch.Tensor, batched_bboxes: List[List[BBox]], tile_weights: np.ndarray,
                 rescale_factor: np.ndarray,
                 denoise: Callable[[torch.Tensor, Sequence[BBox]], torch.Tensor]) -> torch.Tensor:
    """mixtureofdiffusers.py:61-179 (grid part).  Returns x_buffer (x_in.dtype)."""
    N = x_in.shape[0]
    x_buffer = torch.zeros_like(x_in)
    tw = torch.from_numpy(tile_weights).to(x_in.device)
    rf = torch.from_numpy(rescale_factor).view(1, 1, *rescale_factor.shape).to(x_in.device)
    for bboxes in batched_bboxes:
        x_tile = scatter_tiles(x_in, bboxes)
        out = denoise(x_tile, bboxes)
        accumulate_mod(x_buffer, out, bboxes, N, tw, rf)
    return x_buffer
