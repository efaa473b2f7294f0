"""Tile bookkeeping with the reference's names (tile_utils/utils.py), computed by
the C-ABI host functions in csrc/td_host.cpp -- integer results are bit-exact with
the reference's Python float64 arithmetic.

In scope: `BBox`, `split_bboxes`, `splitable`,
`gaussian_weights`, the `Method` / `BlendMode` enums.  Region prompts, feather
masks and the retouch mask are later rows (section 8(f)) and are not here.
"""
from __future__ import annotations

import ctypes
from enum import Enum
from typing import Any, List, Tuple, Union

import numpy as np
import torch

from .. import _cabi, host


class ComparableEnum(Enum):
    def __eq__(self, other: Any) -> bool:
        if isinstance(other, str):
            return self.value == other
        if isinstance(other, ComparableEnum):
            return self.value == other.value
        raise TypeError(f"unsupported type: {type(other)}")

    def __hash__(self):
        return hash(self.value)


class Method(ComparableEnum):
    MULTI_DIFF = "MultiDiffusion"
    MIX_DIFF = "Mixture of Diffusers"


class Method_2(ComparableEnum):
    DEMO_FU = "DemoFusion"


class BlendMode(Enum):
    FOREGROUND = "Foreground"
    BACKGROUND = "Background"


class BBox:
    """Grid bbox: `box = [x, y, x+w, y+h]`, `slicer` crops [N,C,H,W] (utils.py:69-82)."""

    __slots__ = ("x", "y", "w", "h", "box", "slicer")

    def __init__(self, x: int, y: int, w: int, h: int):
        self.x, self.y, self.w, self.h = x, y, w, h
        self.box = [x, y, x + w, y + h]
        self.slicer = (slice(None), slice(None), slice(y, y + h), slice(x, x + w))

    def __getitem__(self, idx: int) -> int:
        return self.box[idx]

    def __repr__(self):
        return f"BBox(x={self.x}, y={self.y}, w={self.w}, h={self.h})"


def _float_ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def splitable(w: int, h: int, tile_w: int, tile_h: int, overlap: int = 16) -> bool:
    """utils.py:151-158; w, h in image pixels."""
    return bool(_cabi.check(_cabi.lib.td_splitable(int(w), int(h), int(tile_w), int(tile_h), int(overlap))))


def split_bboxes_xywh(w: int, h: int, tile_w: int, tile_h: int, overlap: int = 16) -> np.ndarray:
    """int32 [T,4] (x, y, w, h), row-major tile order."""
    cols, rows = ctypes.c_int(0), ctypes.c_int(0)
    n = _cabi.check(_cabi.lib.td_split_bboxes(int(w), int(h), int(tile_w), int(tile_h), int(overlap), None, 0,
                                               ctypes.byref(cols), ctypes.byref(rows)))
    out = np.empty((n, 4), dtype=np.int32)
    _cabi.check(_cabi.lib.td_split_bboxes(int(w), int(h), int(tile_w), int(tile_h), int(overlap),
                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n, None, None))
    return out


def split_bboxes(w: int, h: int, tile_w: int, tile_h: int, overlap: int = 16,
                 init_weight: Union[torch.Tensor, float] = 1.0) -> Tuple[List[BBox], torch.Tensor]:
    """utils.py:160-177: tile list + fp32 weight canvas [1,1,h,w] (`+= init_weight` per tile, list order)."""
    xywh = split_bboxes_xywh(w, h, tile_w, tile_h, overlap)
    bboxes = [BBox(int(x), int(y), int(bw), int(bh)) for x, y, bw, bh in xywh]
    weight = np.zeros((h, w), dtype=np.float32)
    if isinstance(init_weight, torch.Tensor):
        iw = init_weight.detach().to("cpu", torch.float32).numpy()
    else:
        iw = np.float32(init_weight)
    for b in bboxes:
        weight[b.y:b.y + b.h, b.x:b.x + b.w] += iw
    return bboxes, torch.from_numpy(weight).view(1, 1, h, w).to(host.device())


def gaussian_weights_np(tile_w: int, tile_h: int) -> np.ndarray:
    out = np.empty((tile_h, tile_w), dtype=np.float32)
    _cabi.check(_cabi.lib.td_gaussian_weights(int(tile_w), int(tile_h), _float_ptr(out)))
    return out


def gaussian_weights(tile_w: int, tile_h: int) -> torch.Tensor:
    """utils.py:180-194: fp32 [tile_h, tile_w] on the device (asymmetric y midpoint kept)."""
    return torch.from_numpy(gaussian_weights_np(tile_w, tile_h)).to(host.device())


def null_decorator(fn):
    return fn


keep_signature = null_decorator
controlnet = null_decorator
stablesr = null_decorator
grid_bbox = null_decorator
custom_bbox = null_decorator
noise_inverse = null_decorator
