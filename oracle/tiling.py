"""TEST INFRASTRUCTURE -- oracle for tile/bbox bookkeeping (integer-exact).

Restates, with Python's own float64/int semantics (which is what the reference
executes), the reference functions:

  split_bboxes            tile_utils/utils.py:160-177
  splitable               tile_utils/utils.py:151-158
  gaussian_weights        tile_utils/utils.py:180-194
  init_grid_bbox          tile_methods/abstractdiffusion.py:172-186
  MoD init_done rescale   tile_methods/mixtureofdiffusers.py:29-36
  DemoFusion get_views    tile_methods/demofusion.py:101-162 (jitter off, and random jitter with Python's `random`)
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np


def split_bboxes(w: int, h: int, tile_w: int, tile_h: int, overlap: int = 16) -> List[Tuple[int, int, int, int]]:
    """utils.py:160-177.  Returns [(x, y, w, h)] in row-major order (row outer, col inner)."""
    cols = math.ceil((w - overlap) / (tile_w - overlap))
    rows = math.ceil((h - overlap) / (tile_h - overlap))
    dx = (w - tile_w) / (cols - 1) if cols > 1 else 0
    dy = (h - tile_h) / (rows - 1) if rows > 1 else 0
    out = []
    for row in range(rows):
        y = min(int(row * dy), h - tile_h)
        for col in range(cols):
            x = min(int(col * dx), w - tile_w)
            out.append((x, y, tile_w, tile_h))
    return out


def splitable(w: int, h: int, tile_w: int, tile_h: int, overlap: int = 16) -> bool:
    """utils.py:151-158 (w, h in image pixels; opt_f = 8)."""
    w, h = w // 8, h // 8
    min_tile_size = min(tile_w, tile_h)
    if overlap >= min_tile_size:
        overlap = min_tile_size - 4
    cols = math.ceil((w - overlap) / (tile_w - overlap))
    rows = math.ceil((h - overlap) / (tile_h - overlap))
    return cols > 1 or rows > 1


def gaussian_weights(tile_w: int, tile_h: int) -> np.ndarray:
    """utils.py:180-194.  float64 maths, fp32 result [tile_h, tile_w].

    Quirks kept: x midpoint (tile_w-1)/2, y midpoint tile_h/2, and the y
    profile also divides by tile_w**2.
    """
    var = 0.01

    def f(x, midpoint):
        return np.exp(-(x - midpoint) * (x - midpoint) / (tile_w * tile_w) / (2 * var)) / np.sqrt(2 * np.pi * var)

    x_probs = [f(x, (tile_w - 1) / 2) for x in range(tile_w)]
    y_probs = [f(y, tile_h / 2) for y in range(tile_h)]
    return np.outer(y_probs, x_probs).astype(np.float32)


class GridPlan:
    """State that `init_grid_bbox` (+ MoD `init_done`) leaves on the delegate."""

    def __init__(self, w: int, h: int, tile_w: int, tile_h: int, overlap: int, tile_bs: int, gaussian: bool):
        # abstractdiffusion.py:176-178 -- overlap clamp uses the UNclamped tile args
        self.w, self.h = w, h
        self.tile_w = min(tile_w, w)
        self.tile_h = min(tile_h, h)
        self.overlap = max(0, min(overlap, min(tile_w, tile_h) - 4))
        self.bboxes = split_bboxes(w, h, self.tile_w, self.tile_h, self.overlap)
        # utils.py:167,175 -- fp32 canvas, `+= init_weight` per tile in list order
        if gaussian:
            self.tile_weights = gaussian_weights(self.tile_w, self.tile_h)
        else:
            self.tile_weights = None
        weights = np.zeros((h, w), dtype=np.float32)
        for (x, y, tw, th) in self.bboxes:
            if gaussian:
                weights[y:y + th, x:x + tw] += self.tile_weights
            else:
                weights[y:y + th, x:x + tw] += np.float32(1.0)
        self.weights = weights
        # abstractdiffusion.py:183-186
        self.num_tiles = len(self.bboxes)
        self.num_batches = math.ceil(self.num_tiles / tile_bs)
        self.tile_bs = math.ceil(self.num_tiles / self.num_batches)
        self.batched_bboxes = [self.bboxes[i * self.tile_bs:(i + 1) * self.tile_bs] for i in range(self.num_batches)]
        # mixtureofdiffusers.py:32 (fp32 IEEE divide; inf where uncovered)
        with np.errstate(divide="ignore"):
            self.rescale_factor = (np.float32(1.0) / weights).astype(np.float32) if gaussian else None


def demofusion_views(w: int, h: int, window_size: int, overlap: int):
    """demofusion.py:101-162 with random_jitter off: local windows + stride."""
    overlap = max(0, min(overlap, window_size - 4))
    stride = max(4, window_size - overlap)
    tile_w = tile_h = window_size
    cols = math.ceil((w - overlap) / (tile_w - overlap))
    rows = math.ceil((h - overlap) / (tile_h - overlap))
    rows = rows or 1
    cols = cols or 1
    dx = (w - tile_w) / (cols - 1) if cols > 1 else 0
    dy = (h - tile_h) / (rows - 1) if rows > 1 else 0
    out = []
    for row in range(rows):
        for col in range(cols):
            y = min(int(row * dy), h - tile_h)
            x = min(int(col * dx), w - tile_w)
            out.append((x, y, tile_w, tile_h))
    return out, overlap, stride


def demofusion_views_jitter(w: int, h: int, window_size: int, overlap: int, rng):
    """demofusion.py:101-139 with random_jitter on.  `rng` is Python's `random` module (or a `random.Random`): the
    reference draws `randint` per window, x before y, row-major, only for the cases listed below.
    Windows are returned in PADDED canvas coordinates (+ jitter_range); also returns (overlap, stride, jitter_range)."""
    overlap = max(0, min(overlap, window_size - 4))
    stride = max(4, window_size - overlap)
    tile_w = tile_h = window_size
    cols = math.ceil((w - overlap) / (tile_w - overlap)) or 1
    rows = math.ceil((h - overlap) / (tile_h - overlap)) or 1
    dx = (w - tile_w) / (cols - 1) if cols > 1 else 0
    dy = (h - tile_h) / (rows - 1) if rows > 1 else 0
    jr = min(max((min(w, h) - stride) // 4, 0), min(int(window_size / 2), int(overlap / 2)))
    out = []
    for row in range(rows):
        for col in range(cols):
            y = min(int(row * dy), h - tile_h)
            x = min(int(col * dx), w - tile_w)
            xj = yj = 0
            if x != 0 and x + tile_w != w:
                xj = rng.randint(-jr, jr)
            elif x == 0 and x + tile_w != w:
                xj = rng.randint(-jr, 0)
            elif x != 0 and x + tile_w == w:
                xj = rng.randint(0, jr)
            if y != 0 and y + tile_h != h:
                yj = rng.randint(-jr, jr)
            elif y == 0 and y + tile_h != h:
                yj = rng.randint(-jr, 0)
            elif y != 0 and y + tile_h == h:
                yj = rng.randint(0, jr)
            out.append((x + xj + jr, y + yj + jr, tile_w, tile_h))
    return out, overlap, stride, jr
