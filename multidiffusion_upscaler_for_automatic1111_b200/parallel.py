"""Tile shard across the GPUs of one box (one process per GPU, `torch.distributed` for the plumbing).

The reference is single-device; this is new work behind the same delegate API:
`delegate.init_tile_shard(group, fused=True)` after `init_grid_bbox`.  Rank r scatters and denoises
the contiguous chunk `[r*chunk, (r+1)*chunk)` of the row-major tile list; then either

  * fused (default): the tile outputs are placed in an exchange buffer that every peer has mapped
    through CUDA IPC, `td_peer_signal` publishes the step counter into the peers' flag arrays, and
    `td_blend_multidiffusion_peer` waits for all ranks and blends while reading the peers' tiles over
    NVLink -- one kernel does the "all-gather" and the blend, no NCCL call on the data path;
  * NCCL: `all_gather_into_tensor` of the padded chunks, then the ordinary pointer-table blend.

Either way every rank runs the identical deterministic blend, so the latent is bit-identical on all
ranks and to a single-GPU run.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _cabi
from ._cabi import check, lib


class TileShard:
    """Contiguous partition of T tiles over `world` ranks (chunk = ceil(T / world))."""

    def __init__(self, num_tiles: int, rank: int, world: int):
        if not (0 <= rank < world):
            raise ValueError(f"rank {rank} outside world {world}")
        self.num_tiles, self.rank, self.world = num_tiles, rank, world
        self.chunk = -(-num_tiles // world)
        self.begin = min(rank * self.chunk, num_tiles)
        self.end = min(self.begin + self.chunk, num_tiles)
        self.num_chunks = -(-num_tiles // self.chunk)     # ranks >= num_chunks own no tile

    @property
    def num_local(self) -> int:
        return self.end - self.begin

    def owner(self, tile: int) -> int:
        return tile // self.chunk


def gather_tile_outputs(local_padded: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather of the per-rank padded chunks -> [world*chunk*N, C, th, tw] (any backend / device)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * local_padded.shape[0],) + tuple(local_padded.shape[1:]), dtype=local_padded.dtype,
                      device=local_padded.device)
    try:
        dist.all_gather_into_tensor(out, local_padded.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):   # backends without the tensor variant
        parts = list(out.chunk(world, dim=0))
        dist.all_gather(parts, local_padded.contiguous(), group=group)
    return out


class _RawCudaBuffer:
    """A cudaMalloc'd (IPC-exportable) buffer exposed to torch without a copy."""

    def __init__(self, nbytes: int):
        p = ctypes.c_void_p()
        check(lib.td_dev_alloc(int(nbytes), ctypes.byref(p)))
        self.ptr, self.nbytes = p.value, int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def tensor(self, device) -> torch.Tensor:
        return torch.as_tensor(self, device=device)

    def handle(self) -> bytes:
        h = ctypes.create_string_buffer(_cabi.TD_IPC_HANDLE_BYTES)
        check(lib.td_ipc_get_handle(ctypes.c_void_p(self.ptr), h))
        return h.raw

    def free(self):
        if self.ptr:
            lib.td_dev_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


class PeerExchange:
    """Double-buffered tile-output exchange + step flags, mapped into every rank through CUDA IPC."""

    def __init__(self, nbytes: int, device: torch.device, group=None):
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _cabi.TD_MAX_PEERS:
            raise ValueError(f"world size {self.world} exceeds TD_MAX_PEERS={_cabi.TD_MAX_PEERS}")
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        with torch.cuda.device(device):
            self._bufs = [_RawCudaBuffer(self.nbytes), _RawCudaBuffer(self.nbytes)]
            self._flags = _RawCudaBuffer(256)
            torch.cuda.synchronize(device)
            mine = [b.handle() for b in self._bufs] + [self._flags.handle()]
            everyone: List[Optional[list]] = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            self._opened: List[int] = []
            self.buf_ptrs = [[0] * self.world, [0] * self.world]   # [parity][rank] -> device pointer valid HERE
            self.flag_ptrs = [0] * self.world
            for r in range(self.world):
                if r == self.rank:
                    self.buf_ptrs[0][r], self.buf_ptrs[1][r] = self._bufs[0].ptr, self._bufs[1].ptr
                    self.flag_ptrs[r] = self._flags.ptr
                    continue
                ptrs = []
                for h in everyone[r]:
                    p = ctypes.c_void_p()
                    check(lib.td_ipc_open(ctypes.create_string_buffer(h, len(h)), ctypes.byref(p)))
                    ptrs.append(p.value)
                    self._opened.append(p.value)
                self.buf_ptrs[0][r], self.buf_ptrs[1][r], self.flag_ptrs[r] = ptrs
        self._flag_table = (ctypes.c_void_p * self.world)(*self.flag_ptrs)
        self.step = 0
        dist.barrier(group=group)

    def local_buffer(self, parity: int, dtype: torch.dtype) -> torch.Tensor:
        return self._bufs[parity].tensor(self.device).view(dtype)

    @property
    def counter_ptr(self) -> int:
        """This rank's device-side step counter (lives behind the flag slots of the same allocation)."""
        return self._flags.ptr + 128

    def signal(self):
        """Bump the device-side step counter and publish it to every rank (replayable in a CUDA graph)."""
        with torch.cuda.device(self.device):
            check(lib.td_peer_signal(self._flag_table, self.world, self.rank, ctypes.c_void_p(self.counter_ptr),
                                     _cabi.current_stream_ptr(self.device)))

    def close(self):
        for p in self._opened:
            lib.td_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        for b in self._bufs + [self._flags]:
            b.free()


def blend_multidiffusion_peer(g, exchange: PeerExchange, parity: int, shard: TileShard, N: int, C: int, weights: torch.Tensor,
                              dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Wait for every rank's step counter, then blend reading the peers' chunks over NVLink (one kernel)."""
    dev = exchange.device
    nb = shard.num_chunks
    ptrs = (ctypes.c_void_p * nb)(*exchange.buf_ptrs[parity][:nb])
    x_out = out if out is not None else torch.empty((N, C, g.H, g.W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.td_blend_multidiffusion_peer(ctypes.byref(g), ptrs, nb, shard.chunk, N, C, _cabi.dtype_code(dtype),
                                               _cabi.dtype_code(dtype), weights.data_ptr(), x_out.data_ptr(), None,
                                               ctypes.c_void_p(exchange.flag_ptrs[exchange.rank]), exchange.world,
                                               ctypes.c_void_p(exchange.counter_ptr), 0, _cabi.current_stream_ptr(dev)))
    return x_out
