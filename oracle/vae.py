"""TEST INFRASTRUCTURE -- oracle for the tiled VAE (torch-CPU restatement).

Restates scripts/tilevae.py of the reference:

  get_best_tile_size      :390-403      split_tiles            :405-462
  crop_valid_region       :248-259      get_var_mean           :207-215
  custom_group_norm       :218-245      GroupNormParam.summary :320-335
  build_task_queue        :107-195      estimate_group_norm    :464-505
  fast-mode prelude       :542-563      vae_tile_forward       :509-656

Semantics kept on purpose: GroupNorm group count 32 and eps 1e-6;
per-tile statistics merged as a pixel-count weighted AVERAGE of variances and means (no
between-tile term); unbiased std in the fast-mode re-standardisation; decoder pad 11 /
encoder pad 32; bbox order [x1, x2, y1, y2]; fp32 result canvas cast at the end.
CPU offload, progress bars and the cheap-approximation fallback are execution details of
the reference and have no numeric effect; they are not restated.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

NUM_GROUPS = 32
GN_EPS = 1e-6


# ------------------------------------------------------------------ tile geometry
def get_best_tile_size(lowerbound: int, upperbound: int) -> int:
    """tilevae.py:390-403: round the tile up to a multiple of 32/16/8/4/2 if that still fits."""
    divider = 32
    while divider >= 2:
        remainer = lowerbound % divider
        if remainer == 0:
            return lowerbound
        candidate = lowerbound - remainer + divider
        if candidate <= upperbound:
            return candidate
        divider //= 2
    return lowerbound


def split_tiles(h: int, w: int, tile_size: int, pad: int, is_decoder: bool):
    """tilevae.py:405-462 -> (input bboxes, output bboxes), each [x1, x2, y1, y2]."""
    n_h = max(math.ceil((h - 2 * pad) / tile_size), 1)
    n_w = max(math.ceil((w - 2 * pad) / tile_size), 1)
    real_h = get_best_tile_size(math.ceil((h - 2 * pad) / n_h), tile_size)
    real_w = get_best_tile_size(math.ceil((w - 2 * pad) / n_w), tile_size)
    in_bboxes, out_bboxes = [], []
    for i in range(n_h):
        for j in range(n_w):
            ib = [pad + j * real_w, min(pad + (j + 1) * real_w, w), pad + i * real_h, min(pad + (i + 1) * real_h, h)]
            ob = [ib[0] if ib[0] > pad else 0, ib[1] if ib[1] < w - pad else w,
                  ib[2] if ib[2] > pad else 0, ib[3] if ib[3] < h - pad else h]
            out_bboxes.append([x * 8 if is_decoder else x // 8 for x in ob])
            in_bboxes.append([max(0, ib[0] - pad), min(w, ib[1] + pad), max(0, ib[2] - pad), min(h, ib[3] + pad)])
    return in_bboxes, out_bboxes


def crop_margins(in_bbox, out_bbox, is_decoder: bool):
    """tilevae.py:257-258: margin[i] = target - padded (x1, x2, y1, y2)."""
    padded = [i * 8 if is_decoder else i // 8 for i in in_bbox]
    return [out_bbox[i] - padded[i] for i in range(4)]


def crop_valid_region(x: torch.Tensor, in_bbox, out_bbox, is_decoder: bool) -> torch.Tensor:
    m = crop_margins(in_bbox, out_bbox, is_decoder)
    return x[:, :, m[2]:x.size(2) + m[3], m[0]:x.size(3) + m[1]]


# ------------------------------------------------------------------ group norm with given statistics
def get_var_mean(x: torch.Tensor, num_groups: int = NUM_GROUPS):
    """tilevae.py:207-215: biased var and mean per (batch, group) -> [b*groups] each."""
    b, c = x.size(0), x.size(1)
    cpg = int(c / num_groups)
    xr = x.contiguous().view(1, int(b * num_groups), cpg, *x.size()[2:])
    var, mean = torch.var_mean(xr, dim=[0, 2, 3, 4], unbiased=False)
    return var, mean


def custom_group_norm(x: torch.Tensor, mean, var, weight=None, bias=None, num_groups: int = NUM_GROUPS, eps: float = GN_EPS):
    """tilevae.py:218-245: (x - mean) / sqrt(var + eps) per (batch, group), then * gamma[c] + beta[c]."""
    b, c = x.size(0), x.size(1)
    cpg = int(c / num_groups)
    xr = x.contiguous().view(1, int(b * num_groups), cpg, *x.size()[2:])
    out = F.batch_norm(xr, mean.to(x), var.to(x), weight=None, bias=None, training=False, momentum=0, eps=eps)
    out = out.view(b, c, *x.size()[2:])
    if weight is not None:
        out = out * weight.view(1, -1, 1, 1)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


def merge_tile_stats(vars_: Sequence[torch.Tensor], means: Sequence[torch.Tensor], pixels: Sequence[int]):
    """tilevae.py:320-335: pixel-count weighted average of the per-tile variances and means."""
    var = torch.vstack(list(vars_))
    mean = torch.vstack(list(means))
    px = torch.tensor(list(pixels), dtype=torch.float32, device=var.device) / max(pixels)
    px = (px / px.sum()).unsqueeze(1)
    return (var * px).sum(dim=0), (mean * px).sum(dim=0)


# ------------------------------------------------------------------ op list (task queue)
def attention(net, h: torch.Tensor) -> torch.Tensor:
    """tile_utils/attn.py:49-72: single-head softmax(QK^T / sqrt(C)) V + proj_out, no norm / residual."""
    q, k, v = net.q(h), net.k(h), net.v(h)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    return net.proj_out(torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww))


def _resblock_ops(ops, block):
    if block.in_channels != block.out_channels:
        ops.append(["store_res", block.conv_shortcut if block.use_conv_shortcut else block.nin_shortcut])
    else:
        ops.append(["store_res", lambda x: x])
    ops += [["pre_norm", block.norm1], ["silu", F.silu], ["conv1", block.conv1],
            ["pre_norm", block.norm2], ["silu", F.silu], ["conv2", block.conv2], ["add_res", None]]


def _attn_ops(ops, net):
    ops += [["store_res", lambda x: x], ["pre_norm", net.norm], ["attn", lambda x, net=net: attention(net, x)], ["add_res", None]]


def build_task_queue(net, is_decoder: bool):
    """tilevae.py:139-195: the Encoder / Decoder flattened into a list of [kind, callable]."""
    ops = [["conv_in", net.conv_in]]
    if is_decoder:
        _resblock_ops(ops, net.mid.block_1); _attn_ops(ops, net.mid.attn_1); _resblock_ops(ops, net.mid.block_2)
        for i_level in reversed(range(net.num_resolutions)):
            for i_block in range(net.num_res_blocks + 1):
                _resblock_ops(ops, net.up[i_level].block[i_block])
            if i_level != 0:
                ops.append(["upsample", net.up[i_level].upsample])
    else:
        for i_level in range(net.num_resolutions):
            for i_block in range(net.num_res_blocks):
                _resblock_ops(ops, net.down[i_level].block[i_block])
            if i_level != net.num_resolutions - 1:
                ops.append(["downsample", net.down[i_level].downsample])
        _resblock_ops(ops, net.mid.block_1); _attn_ops(ops, net.mid.attn_1); _resblock_ops(ops, net.mid.block_2)
    if not is_decoder or not net.give_pre_end:
        ops += [["pre_norm", net.norm_out], ["silu", F.silu], ["conv_out", net.conv_out]]
        if is_decoder and net.tanh_out:
            ops.append(["tanh", torch.tanh])
    return ops


def _clone_ops(ops):
    return [[k, f] for k, f in ops]


def _frozen_norm(tile: torch.Tensor, layer) -> Callable:
    """GroupNormParam.from_tile (tilevae.py:337-361): statistics of ONE tensor frozen into a closure."""
    var, mean = get_var_mean(tile)
    weight = getattr(layer, "weight", None)
    bias = getattr(layer, "bias", None)
    return lambda x, mean=mean, var=var, weight=weight, bias=bias: custom_group_norm(x, mean, var, weight, bias)


def estimate_group_norm(z: torch.Tensor, ops, color_fix: bool) -> bool:
    """tilevae.py:464-505: run the op list on the down-sampled input; every pre_norm met on the way
    becomes an ('apply_norm', closure) with that tensor's statistics.  Edits `ops` in place."""
    tile = z
    last = len(ops) - 1
    while last >= 0 and ops[last][0] != "pre_norm":
        last -= 1
    if last <= 0:
        raise ValueError("No group norm found in the task queue")
    for i in range(last + 1):
        kind, fn = ops[i]
        if kind == "pre_norm":
            norm = _frozen_norm(tile, fn)
            ops[i] = ["apply_norm", norm]
            if i == last:
                return True
            tile = norm(tile)
        elif kind == "store_res":
            j = i + 1
            while j < last and ops[j][0] != "add_res":
                j += 1
            if j >= last:
                continue
            ops[j][1] = fn(tile)
        elif kind == "add_res":
            tile = tile + fn
            ops[i][1] = None
        elif color_fix and kind == "downsample":
            return True   # later norms stay 'pre_norm' (exact, slow-mode statistics)
        else:
            tile = fn(tile)
        if torch.isnan(tile).any():
            return False
    raise IndexError("Should not reach here")


def fast_mode_estimator_input(z: torch.Tensor, tile_size: int) -> torch.Tensor:
    """tilevae.py:545-559: nearest-exact resample to ~tile_size, re-standardise per channel, clamp."""
    height, width = z.shape[2], z.shape[3]
    scale_factor = tile_size / max(height, width)
    ds = F.interpolate(z, scale_factor=scale_factor, mode="nearest-exact")
    std_old, mean_old = torch.std_mean(z, dim=[0, 2, 3], keepdim=True)
    std_new, mean_new = torch.std_mean(ds, dim=[0, 2, 3], keepdim=True)
    ds = (ds - mean_new) / std_new * std_old + mean_old
    return torch.clamp(ds, min=z.min(), max=z.max())


# ------------------------------------------------------------------ executor
def vae_tile_forward(net, z: torch.Tensor, tile_size: int, is_decoder: bool, fast_mode: bool, color_fix: bool = False) -> torch.Tensor:
    """tilevae.py:509-656 (numerics only)."""
    dtype = next(net.parameters()).dtype
    pad = 11 if is_decoder else 32
    color_fix = color_fix and not is_decoder
    z = z.detach()
    N, height, width = z.shape[0], z.shape[2], z.shape[3]
    in_bboxes, out_bboxes = split_tiles(height, width, tile_size, pad, is_decoder)
    tiles: List[Optional[torch.Tensor]] = [z[:, :, b[2]:b[3], b[0]:b[1]].clone() for b in in_bboxes]
    T = len(tiles)

    ops = build_task_queue(net, is_decoder)
    if fast_mode:
        est_ops = _clone_ops(ops)
        if estimate_group_norm(fast_mode_estimator_input(z, tile_size), est_ops, color_fix):
            ops = est_ops
    queues = [_clone_ops(ops) for _ in range(T)]

    result = None
    done = 0
    forward = True
    while True:
        stats = ([], [], [])   # vars, means, pixels of this GroupNorm round
        layer = None
        order = range(T) if forward else reversed(range(T))
        for i in order:
            tile, q = tiles[i], queues[i]
            while q:
                kind, fn = q.pop(0)
                if kind == "pre_norm":
                    var, mean = get_var_mean(tile)
                    stats[0].append(var); stats[1].append(mean); stats[2].append(tile.shape[2] * tile.shape[3])
                    layer = fn
                    break
                elif kind in ("store_res", "store_res_cpu"):
                    j = 0
                    while q[j][0] != "add_res":
                        j += 1
                    q[j][1] = fn(tile)
                elif kind == "add_res":
                    tile = tile + fn
                else:
                    tile = fn(tile)
            if torch.isnan(tile).any():
                raise RuntimeError("NaN in vae tile")
            if not q:
                tiles[i] = None
                done += 1
                if result is None:
                    result = torch.zeros((N, tile.shape[1], height * 8 if is_decoder else height // 8,
                                          width * 8 if is_decoder else width // 8), dtype=torch.float32)
                ob = out_bboxes[i]
                result[:, :, ob[2]:ob[3], ob[0]:ob[1]] = crop_valid_region(tile, in_bboxes[i], ob, is_decoder)
            else:
                tiles[i] = tile
                if i == T - 1 and forward:
                    forward = False
                elif i == 0 and not forward:
                    forward = True
        if done == T:
            break
        if stats[0]:
            var, mean = merge_tile_stats(*stats)
            weight, bias = getattr(layer, "weight", None), getattr(layer, "bias", None)
            norm = lambda x, mean=mean, var=var, weight=weight, bias=bias: custom_group_norm(x, mean, var, weight, bias)
            for q in queues:
                if q:
                    q.insert(0, ["apply_norm", norm])
    return result.to(dtype)


def vae_hook_call(net, x: torch.Tensor, tile_size: int, is_decoder: bool, fast_mode: bool, color_fix: bool = False):
    """VAEHook.__call__ (tilevae.py:375-388): tiny inputs bypass tiling."""
    pad = 11 if is_decoder else 32
    if max(x.shape[2], x.shape[3]) <= pad * 2 + tile_size:
        return net.original_forward(x) if hasattr(net, "original_forward") else net(x)
    return vae_tile_forward(net, x, tile_size, is_decoder, fast_mode, color_fix)
