"""Tiled VAE with the reference's surface (scripts/tilevae.py: `VAEHook`, `GroupNormParam`,
`get_var_mean`, `custom_group_norm`, `crop_valid_region`, `build_task_queue`), B200-native
underneath.

What is different from the reference by design (180 GB of HBM, hand-written sm_100a kernels):
  * tiles, residuals and parked activations stay in HBM -- the reference stages every tile
    through host RAM (`.cpu()` at tilevae.py:534,610,641) and re-uploads it each round;
  * GroupNorm statistics: one streaming read (`td_gn_stats`, fp32 Chan/Welford merge with
    warp shuffles) instead of `var_mean` on a reshaped copy;
  * GroupNorm apply + affine + SiLU: ONE fused pass (`td_gn_apply`) instead of
    batch_norm, `*= weight`, `+= bias`, `silu_` (4 read+write passes, tilevae.py:237-244,:104);
  * tile crop and valid-region paste are `td_copy_region` launches; the result canvas is
    allocated once in the network dtype (the reference's fp32 canvas + final cast is a
    lossless round trip for fp16/bf16 tile outputs);
  * no per-tile NaN test (each one is a full read + host sync, tilevae.py:625): the result
    is checked once at the end and the estimator keeps its NaN fallback.
The convolutions / attention GEMMs stay the host application's modules (cuDNN / SDPA).
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import host
from ._cabi import check, current_stream_ptr, dtype_code, lib

NUM_GROUPS = 32
GN_EPS = 1e-6


# --------------------------------------------------------------------------- kernels (tensor wrappers)
def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: the tiled-VAE path has no CPU fallback (got {t.device})")


def _segment_stats(x: torch.Tensor, nseg: int, seg_len: int, unbiased: bool, want_minmax: bool = False):
    _require_cuda(x, "activation")
    dev = x.device
    ws_bytes = lib.td_gn_stats_workspace_bytes(nseg, seg_len, dtype_code(x.dtype))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    mean = torch.empty(nseg, dtype=torch.float32, device=dev)
    var = torch.empty(nseg, dtype=torch.float32, device=dev)
    lo = torch.empty(nseg, dtype=torch.float32, device=dev) if want_minmax else None
    hi = torch.empty(nseg, dtype=torch.float32, device=dev) if want_minmax else None
    with torch.cuda.device(dev):
        check(lib.td_gn_stats(x.data_ptr(), nseg, seg_len, dtype_code(x.dtype), int(unbiased), ws.data_ptr(), ws_bytes,
                              mean.data_ptr(), var.data_ptr(), lo.data_ptr() if want_minmax else None,
                              hi.data_ptr() if want_minmax else None, current_stream_ptr(dev)))
    return (var, mean, lo, hi) if want_minmax else (var, mean)


def get_var_mean(input: torch.Tensor, num_groups: int, eps: float = 1e-6):
    """tilevae.py:207-215: biased variance and mean per (batch, group), fp32 [b*num_groups] each."""
    b, c = input.size(0), input.size(1)
    cpg = int(c / num_groups)
    x = input.contiguous()
    seg_len = cpg * int(np.prod(x.shape[2:]))
    return _segment_stats(x, int(b * num_groups), seg_len, unbiased=False)


def _affine32(layer_weight, layer_bias, device):
    g = layer_weight.detach().to(device=device, dtype=torch.float32).contiguous() if layer_weight is not None else None
    b = layer_bias.detach().to(device=device, dtype=torch.float32).contiguous() if layer_bias is not None else None
    return g, b


def custom_group_norm(input: torch.Tensor, num_groups: int, mean: torch.Tensor, var: torch.Tensor, weight=None, bias=None,
                      eps: float = 1e-6, act: bool = False, inplace: bool = False) -> torch.Tensor:
    """tilevae.py:218-245 (+ optional fused SiLU): group norm with externally supplied statistics.

    mean / var: [num_groups] (shared by the batch, as the reference's merged statistics are)
    or [b*num_groups]."""
    _require_cuda(input, "activation")
    b, c = input.size(0), input.size(1)
    x = input.contiguous()
    hw = int(np.prod(x.shape[2:]))
    per_batch = mean.numel() == b * num_groups and b > 1
    if mean.numel() not in (num_groups, b * num_groups):
        raise ValueError(f"statistics have {mean.numel()} entries; expected {num_groups} or {b * num_groups}")
    mean32 = mean.to(device=x.device, dtype=torch.float32).contiguous()
    var32 = var.to(device=x.device, dtype=torch.float32).contiguous()
    gamma, beta = weight, bias
    if gamma is not None and gamma.dtype != torch.float32:
        gamma, beta = _affine32(gamma, beta, x.device)
    out = x if inplace else torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.td_gn_apply(x.data_ptr(), out.data_ptr(), b, c, hw, dtype_code(x.dtype), num_groups, mean32.data_ptr(),
                              var32.data_ptr(), int(per_batch), gamma.data_ptr() if gamma is not None else None,
                              beta.data_ptr() if beta is not None else None, float(eps), int(act), current_stream_ptr(x.device)))
    return out


def copy_region(src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[...] = src[...] for two 4-D views [N, C, rows, cols] whose last dim has stride 1 and whose
    (N, C) dims collapse to one plane stride (contiguous parents).  One td_copy_region launch."""
    _require_cuda(src, "src"); _require_cuda(dst, "dst")
    assert src.shape == dst.shape and src.dtype == dst.dtype and src.dim() == 4
    n, c, rows, cols = src.shape
    for t in (src, dst):
        assert t.stride(3) == 1 and (n == 1 or t.stride(0) == c * t.stride(1)), "unsupported view layout"
    with torch.cuda.device(dst.device):
        check(lib.td_copy_region(src.data_ptr(), dst.data_ptr(), n * c, rows, cols, src.stride(1), src.stride(2), dst.stride(1),
                                 dst.stride(2), dtype_code(src.dtype), current_stream_ptr(dst.device)))


def crop_valid_region(x: torch.Tensor, input_bbox, target_bbox, is_decoder: bool) -> torch.Tensor:
    """tilevae.py:248-259 (a view; the paste that follows is the copy)."""
    padded = [i * 8 if is_decoder else i // 8 for i in input_bbox]
    margin = [target_bbox[i] - padded[i] for i in range(4)]
    return x[:, :, margin[2]:x.size(2) + margin[3], margin[0]:x.size(3) + margin[1]]


def nearest_exact_indices(in_size: int, scale_factor: float):
    """Output size and source indices of F.interpolate(mode='nearest-exact', scale_factor=s):
    out = floor(in * s); src = min(floor((dst + 0.5) * float32(1 / s)), in - 1) (ATen UpSample)."""
    out = int(math.floor(float(in_size) * scale_factor))
    inv = np.float32(1.0 / scale_factor)
    idx = np.floor((np.arange(out, dtype=np.float32) + np.float32(0.5)) * inv).astype(np.int64)
    return out, np.minimum(idx, in_size - 1).astype(np.int32)


def fast_mode_estimator_input(z: torch.Tensor, tile_size: int) -> torch.Tensor:
    """tilevae.py:545-559 on the device: nearest-exact resample, per-channel re-standardisation
    (unbiased std) and clamp to z's range; z is read once for its statistics."""
    _require_cuda(z, "z")
    z = z.contiguous()
    B, C, H, W = z.shape
    scale = tile_size / max(H, W)
    oh, sy = nearest_exact_indices(H, scale)
    ow, sx = nearest_exact_indices(W, scale)
    dev = z.device
    sy_d, sx_d = torch.from_numpy(sy).to(dev), torch.from_numpy(sx).to(dev)
    ds = torch.empty((B, C, oh, ow), dtype=z.dtype, device=dev)
    with torch.cuda.device(dev):
        check(lib.td_resample_nearest(z.data_ptr(), ds.data_ptr(), B * C, H, W, oh, ow, sy_d.data_ptr(), sx_d.data_ptr(),
                                      dtype_code(z.dtype), current_stream_ptr(dev)))

    def channel_stats(t, hw, want_minmax):
        out = _segment_stats(t, B * C, hw, unbiased=True, want_minmax=want_minmax)
        var, mean = out[0].view(B, C), out[1].view(B, C)
        if B > 1:   # torch.std_mean over dims [0, 2, 3]: merge the per-(b, c) moments (equal counts)
            m = mean.mean(dim=0)
            m2 = (var * (hw - 1)).sum(dim=0) + hw * ((mean - m) ** 2).sum(dim=0)
            var, mean = m2 / (B * hw - 1), m
        else:
            var, mean = var[0], mean[0]
        return (var.sqrt(), mean) + ((out[2].min().view(1), out[3].max().view(1)) if want_minmax else ())

    std_old, mean_old, lo, hi = channel_stats(z, H * W, True)
    std_new, mean_new = channel_stats(ds, oh * ow, False)
    # the reference's statistics are tensors of z.dtype: round them the same way
    rt = lambda t: t.to(z.dtype).float().contiguous()
    mean_new, std_new, mean_old, std_old = rt(mean_new), rt(std_new), rt(mean_old), rt(std_old)
    with torch.cuda.device(dev):
        check(lib.td_affine_clamp(ds.data_ptr(), B, C, oh * ow, dtype_code(ds.dtype), mean_new.data_ptr(), std_new.data_ptr(),
                                  mean_old.data_ptr(), std_old.data_ptr(), lo.data_ptr(), hi.data_ptr(), current_stream_ptr(dev)))
    return ds


# --------------------------------------------------------------------------- task queue
def inplace_nonlinearity(x):
    return F.silu(x, inplace=True)


def attn_forward(net, h_: torch.Tensor) -> torch.Tensor:
    """VAE mid-block attention without norm / residual (tile_utils/attn.py:49-72): one
    single-head softmax(QK^T / sqrt(C)) V + proj_out.  The reference's six back-end variants
    are memory work-arounds of this same function."""
    q, k, v = net.q(h_), net.k(h_), net.v(h_)
    b, c, h, w = q.shape
    q, k, v = (t.reshape(b, 1, c, h * w).transpose(2, 3) for t in (q, k, v))   # [b, 1, hw, c]
    o = F.scaled_dot_product_attention(q, k, v)                                 # scale = c ** -0.5
    return net.proj_out(o.transpose(2, 3).reshape(b, c, h, w))


def attn2task(task_queue, net):
    task_queue.append(('store_res', lambda x: x))
    task_queue.append(('pre_norm', net.norm))
    task_queue.append(('attn', lambda x, net=net: attn_forward(net, x)))
    task_queue.append(['add_res', None])


def resblock2task(queue, block):
    if block.in_channels != block.out_channels:
        queue.append(('store_res', block.conv_shortcut if block.use_conv_shortcut else block.nin_shortcut))
    else:
        queue.append(('store_res', lambda x: x))
    queue.append(('pre_norm', block.norm1))
    queue.append(('silu', inplace_nonlinearity))
    queue.append(('conv1', block.conv1))
    queue.append(('pre_norm', block.norm2))
    queue.append(('silu', inplace_nonlinearity))
    queue.append(('conv2', block.conv2))
    queue.append(['add_res', None])


def build_sampling(task_queue, net, is_decoder):
    if is_decoder:
        resblock2task(task_queue, net.mid.block_1)
        attn2task(task_queue, net.mid.attn_1)
        resblock2task(task_queue, net.mid.block_2)
        for i_level in reversed(range(net.num_resolutions)):
            for i_block in range(net.num_res_blocks + 1):
                resblock2task(task_queue, net.up[i_level].block[i_block])
            if i_level != 0:
                task_queue.append(('upsample', net.up[i_level].upsample))
    else:
        for i_level in range(net.num_resolutions):
            for i_block in range(net.num_res_blocks):
                resblock2task(task_queue, net.down[i_level].block[i_block])
            if i_level != net.num_resolutions - 1:
                task_queue.append(('downsample', net.down[i_level].downsample))
        resblock2task(task_queue, net.mid.block_1)
        attn2task(task_queue, net.mid.attn_1)
        resblock2task(task_queue, net.mid.block_2)


def build_task_queue(net, is_decoder):
    """tilevae.py:174-195: the Encoder / Decoder as a flat op list."""
    task_queue = [('conv_in', net.conv_in)]
    build_sampling(task_queue, net, is_decoder)
    if not is_decoder or not net.give_pre_end:
        task_queue.append(('pre_norm', net.norm_out))
        task_queue.append(('silu', inplace_nonlinearity))
        task_queue.append(('conv_out', net.conv_out))
        if is_decoder and net.tanh_out:
            task_queue.append(('tanh', torch.tanh))
    return task_queue


def clone_task_queue(task_queue):
    return [[item for item in task] for task in task_queue]


class _Norm:
    """('apply_norm', _Norm): frozen statistics + affine of one GroupNorm site; fuses the SiLU that follows."""

    def __init__(self, mean, var, layer):
        self.mean, self.var = mean, var
        w, b = getattr(layer, "weight", None), getattr(layer, "bias", None)
        self.gamma, self.beta = _affine32(w, b, mean.device) if w is not None else (None, None)

    def __call__(self, x, act: bool = False):
        return custom_group_norm(x, NUM_GROUPS, self.mean, self.var, self.gamma, self.beta, GN_EPS, act=act)


class GroupNormParam:
    """tilevae.py:289-361: collects per-tile statistics of one GroupNorm round and merges them.

    With `group` set (tile shard over ranks) the pixel-weighted sums are all-reduced, so every rank applies
    the statistics of ALL tiles (same weights as the reference: p_i = pixels_i / sum(pixels))."""

    def __init__(self, group=None, sharded: bool = False):
        self.group, self.sharded = group, sharded
        self.var_list = []
        self.mean_list = []
        self.pixel_list = []
        self.weight = None
        self.bias = None
        self.layer = None

    def add_tile(self, tile, layer):
        var, mean = get_var_mean(tile, NUM_GROUPS)   # fp32 statistics: no fp16 overflow branch needed (tilevae.py:300-304)
        self.var_list.append(var)
        self.mean_list.append(mean)
        self.pixel_list.append(tile.shape[2] * tile.shape[3])
        self.layer = layer
        self.weight = getattr(layer, 'weight', None)
        self.bias = getattr(layer, 'bias', None)

    def summary(self):
        """Pixel-count weighted average of the tile variances and means (tilevae.py:320-335)."""
        if len(self.var_list) == 0 and not self.sharded:
            return None
        if self.sharded:
            import torch.distributed as dist
            dev = host.device()
            if self.var_list:
                px = torch.tensor(self.pixel_list, dtype=torch.float32, device=self.var_list[0].device).unsqueeze(1)
                packed = torch.cat([(torch.vstack(self.var_list) * px).sum(0), (torch.vstack(self.mean_list) * px).sum(0), px.sum().view(1)])
            else:
                packed = None
            # ranks may own no tile in this round: agree on the vector length first
            n = torch.tensor([0 if packed is None else packed.numel()], device=dev)
            dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.group)
            if int(n.item()) == 0:
                return None
            if packed is None:
                packed = torch.zeros(int(n.item()), dtype=torch.float32, device=dev)
            dist.all_reduce(packed, group=self.group)
            k = (packed.numel() - 1) // 2
            var, mean = packed[:k] / packed[-1], packed[k:2 * k] / packed[-1]
            return _Norm(mean, var, self.layer) if self.layer is not None else _Norm(mean, var, None)
        var = torch.vstack(self.var_list)
        mean = torch.vstack(self.mean_list)
        max_value = max(self.pixel_list)
        pixels = torch.tensor(self.pixel_list, dtype=torch.float32, device=var.device) / max_value
        pixels = (pixels / torch.sum(pixels)).unsqueeze(1)
        return _Norm(torch.sum(mean * pixels, dim=0), torch.sum(var * pixels, dim=0), self.layer)

    @staticmethod
    def from_tile(tile, norm):
        """Statistics of a single tensor frozen into a norm function (tilevae.py:337-361)."""
        var, mean = get_var_mean(tile, NUM_GROUPS)
        return _Norm(mean, var, norm)


# --------------------------------------------------------------------------- the hook
class VAEHook:

    def __init__(self, net, tile_size, is_decoder: bool, fast_decoder: bool, fast_encoder: bool, color_fix: bool, to_gpu: bool = False):
        self.net = net
        self.tile_size = tile_size
        self.is_decoder = is_decoder
        self.fast_mode = (fast_encoder and not is_decoder) or (fast_decoder and is_decoder)
        self.color_fix = color_fix and not is_decoder
        self.to_gpu = to_gpu
        self.pad = 11 if is_decoder else 32
        self.verbose = False
        self._shard_group = None
        self._shard_world = 1
        self._shard_rank = 0

    def init_tile_shard(self, group=None):
        """Shard the VAE tiles over the ranks of `group` (one process per GPU): rank r runs tiles i with
        i % world == r; GroupNorm statistics (slow mode) are all-reduced per round and the disjoint output
        regions are combined with one all-reduce of the canvas (x + 0 is exact).  New: the reference is
        single-device."""
        import torch.distributed as dist
        self._shard_group = group
        self._shard_world = dist.get_world_size(group)
        self._shard_rank = dist.get_rank(group)

    def __call__(self, x):
        original_device = next(self.net.parameters()).device
        try:
            if self.to_gpu:
                self.net = self.net.to(host.device())
            B, C, H, W = x.shape
            if max(H, W) <= self.pad * 2 + self.tile_size:
                if self.verbose:
                    print("[Tiled VAE]: the input size is tiny and unnecessary to tile.")
                return self.net.original_forward(x)
            return self.vae_tile_forward(x)
        finally:
            self.net = self.net.to(original_device)

    def get_best_tile_size(self, lowerbound, upperbound):
        return lib.td_vae_best_tile_size(int(lowerbound), int(upperbound))

    def split_tiles(self, h, w):
        """tilevae.py:405-462 -> (tile_input_bboxes, tile_output_bboxes), each [x1, x2, y1, y2]."""
        n = check(lib.td_vae_split_tiles(int(h), int(w), int(self.tile_size), int(self.pad), int(self.is_decoder), None, None, 0))
        ib = np.empty((n, 4), dtype=np.int32)
        ob = np.empty((n, 4), dtype=np.int32)
        p = ctypes.POINTER(ctypes.c_int32)
        check(lib.td_vae_split_tiles(int(h), int(w), int(self.tile_size), int(self.pad), int(self.is_decoder),
                                     ib.ctypes.data_as(p), ob.ctypes.data_as(p), n))
        return ib.tolist(), ob.tolist()

    @torch.no_grad()
    def estimate_group_norm(self, z, task_queue, color_fix):
        """tilevae.py:464-505: run the queue on the down-sampled input and freeze every GroupNorm met."""
        tile = z
        last_id = len(task_queue) - 1
        while last_id >= 0 and task_queue[last_id][0] != 'pre_norm':
            last_id -= 1
        if last_id <= 0 or task_queue[last_id][0] != 'pre_norm':
            raise ValueError('No group norm found in the task queue')
        i = 0
        while i <= last_id:
            task = task_queue[i]
            if task[0] == 'pre_norm':
                norm = GroupNormParam.from_tile(tile, task[1])
                task_queue[i] = ('apply_norm', norm)
                if i == last_id:
                    return True
                fuse = task_queue[i + 1][0] == 'silu'
                tile = norm(tile, act=fuse)
                if fuse:
                    i += 1
            elif task[0] == 'store_res':
                task_id = i + 1
                while task_id < last_id and task_queue[task_id][0] != 'add_res':
                    task_id += 1
                if task_id < last_id:
                    task_queue[task_id][1] = task[1](tile)
            elif task[0] == 'add_res':
                tile = tile + task[1]
                task[1] = None
            elif color_fix and task[0] == 'downsample':
                return True
            else:
                tile = task[1](tile)
            if torch.isnan(tile).any():
                print('Nan detected in fast mode estimation. Fast mode disabled.')
                return False
            i += 1
        raise IndexError('Should not reach here')

    @torch.no_grad()
    def vae_tile_forward(self, z):
        """tilevae.py:509-656, device-resident."""
        param = next(self.net.parameters())
        device, dtype = param.device, param.dtype
        if device.type != "cuda":
            raise RuntimeError(f"VAEHook: network is on {device}; the B200 tiled-VAE path has no CPU fallback")
        net, tile_size, is_decoder = self.net, self.tile_size, self.is_decoder

        z = z.detach().to(device=device, dtype=dtype).contiguous()
        N, height, width = z.shape[0], z.shape[2], z.shape[3]
        net.last_z_shape = z.shape
        in_bboxes, out_bboxes = self.split_tiles(height, width)

        # tile crop: straight HBM -> HBM (the reference goes through host RAM)
        sharded = self._shard_world > 1
        if sharded:   # this rank's tiles only
            mine = [i for i in range(len(in_bboxes)) if i % self._shard_world == self._shard_rank]
            all_out_shape = None
            in_bboxes, out_bboxes = [in_bboxes[i] for i in mine], [out_bboxes[i] for i in mine]
        tiles: List[Optional[torch.Tensor]] = []
        for b in in_bboxes:
            t = torch.empty((N, z.shape[1], b[3] - b[2], b[1] - b[0]), dtype=dtype, device=device)
            copy_region(z[:, :, b[2]:b[3], b[0]:b[1]], t)
            tiles.append(t)
        num_tiles = len(tiles)

        single_task_queue = build_task_queue(net, is_decoder)
        if self.fast_mode:
            estimate_task_queue = clone_task_queue(single_task_queue)
            if self.estimate_group_norm(fast_mode_estimator_input(z, tile_size), estimate_task_queue, color_fix=self.color_fix):
                single_task_queue = estimate_task_queue
        del z
        task_queues = [clone_task_queue(single_task_queue) for _ in range(num_tiles)]

        result = None
        num_completed = 0
        forward = True
        while True:
            if host.interrupted():
                break
            group_norm_param = GroupNormParam(self._shard_group, sharded)
            for i in (range(num_tiles) if forward else reversed(range(num_tiles))):
                if host.interrupted():
                    break
                tile = tiles[i]
                task_queue = task_queues[i]
                while len(task_queue) > 0:
                    task = task_queue.pop(0)
                    kind = task[0]
                    if kind == 'pre_norm':
                        group_norm_param.add_tile(tile, task[1])
                        break
                    elif kind == 'apply_norm':
                        fuse = len(task_queue) > 0 and task_queue[0][0] == 'silu'
                        if fuse:
                            task_queue.pop(0)
                        tile = task[1](tile, act=fuse)
                    elif kind == 'store_res' or kind == 'store_res_cpu':
                        task_id = 0
                        res = task[1](tile)
                        while task_queue[task_id][0] != 'add_res':
                            task_id += 1
                        task_queue[task_id][1] = res
                    elif kind == 'add_res':
                        tile = tile + task[1] if tile.data_ptr() == task[1].data_ptr() else tile.add_(task[1])
                        task[1] = None
                    else:
                        tile = task[1](tile)

                if len(task_queue) == 0:
                    tiles[i] = None
                    num_completed += 1
                    if result is None:
                        alloc = torch.zeros if sharded else torch.empty   # sharded: the other ranks' regions must read 0
                        result = alloc((N, tile.shape[1], height * 8 if is_decoder else height // 8,
                                        width * 8 if is_decoder else width // 8), dtype=dtype, device=device)
                    ob = out_bboxes[i]
                    valid = crop_valid_region(tile.to(dtype).contiguous(), in_bboxes[i], ob, is_decoder)
                    copy_region(valid, result[:, :, ob[2]:ob[3], ob[0]:ob[1]])
                    del tile
                else:
                    tiles[i] = tile
                    if i == num_tiles - 1 and forward:
                        forward = False
                    elif i == 0 and not forward:
                        forward = True

            if (num_completed == num_tiles and not sharded) or host.interrupted():
                break
            norm = group_norm_param.summary()   # sharded: a collective -- every rank takes part in every round
            if sharded and norm is None:
                break
            if norm is not None:
                for q in task_queues:
                    if len(q) > 0:
                        q.insert(0, ('apply_norm', norm))

        if sharded:
            import torch.distributed as dist
            out_c = self.net.conv_out.out_channels
            if result is None:
                result = torch.zeros((N, out_c, height * 8 if is_decoder else height // 8, width * 8 if is_decoder else width // 8),
                                     dtype=dtype, device=device)
            dist.all_reduce(result, group=self._shard_group)   # disjoint regions + zeros: exact
        if result is None or num_completed != num_tiles:
            raise RuntimeError("[Tiled VAE]: interrupted before any tile finished")
        if torch.isnan(result).any():
            raise RuntimeError("[Tiled VAE]: NaN in the result (the reference's test_for_nans, tilevae.py:625)")
        return result
