"""Shared helpers for the parity tests (bit views, oracle plans)."""
import hashlib

import numpy as np
import torch

DTYPES = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float32:
        return t.numpy().view(np.uint32)
    return t.view(torch.int16).numpy().view(np.uint16)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(bits(t).tobytes()).hexdigest()


def assert_bit_equal(a: torch.Tensor, b: torch.Tensor, what: str = ""):
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} vs {b.dtype}"
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        # +0.0 / -0.0 are the same value to torch.equal; report genuine mismatches only
        fa, fb = a.detach().cpu().float(), b.detach().cpu().float()
        bad = ~((fa == fb) | (fa.isnan() & fb.isnan()))
        n = int(bad.sum())
        assert n == 0, f"{what}: {n} mismatching elements, max abs diff {(fa - fb).abs().max().item()}"
