"""Base tile-method delegate with the reference's surface
(tile_methods/abstractdiffusion.py), B200-native underneath.

What differs from the reference by design:
  * `init_grid_bbox` also builds a `td_grid` plan (C-ABI) and a persistent
    tile-batch buffer: the whole step's scatter is ONE kernel launch and the
    UNet receives views of that buffer (180 GB of HBM: no need to re-cat).
  * `reset_buffer` only (re)allocates: the fused gather-form blend writes every
    canvas pixel, so the per-step `zero_()` pass of the reference is not needed.
  * Region prompt control, ControlNet / StableSR tile caches and noise inversion
    are later rows of the scope table: their `init_*`
    raise NotImplementedError instead of silently doing something else.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Union

import torch
from torch import Tensor

from .. import engine, host
from ..host import opt_f
from ..tile_utils.utils import BBox, custom_bbox, grid_bbox, noise_inverse, controlnet, stablesr

CondDict = Dict[str, Union[Tensor, List[Tensor]]]


class AbstractDiffusion:

    def __init__(self, p, sampler):
        self.method = self.__class__.__name__
        self.p = p
        self.pbar = None

        # sampler (abstractdiffusion.py:11-14)
        self.sampler_name = p.sampler_name
        self.sampler_raw = sampler
        self.sampler = sampler

        # image-editing (ip2p) models add one more cond slot (abstractdiffusion.py:16-20)
        if self.is_kdiff and not hasattr(self, "is_edit_model"):
            sd_model = self._sd_model()
            cfg = getattr(self.sampler, "model_wrap_cfg", None)
            scale = getattr(cfg, "image_cfg_scale", None)
            self.is_edit_model = (getattr(sd_model, "cond_stage_key", None) == "edit"
                                  and scale is not None and scale != 1.0)

        # latent canvas + persistent state (abstractdiffusion.py:22-28)
        self.x_buffer: Optional[Tensor] = None
        self.w: int = int(self.p.width // opt_f)
        self.h: int = int(self.p.height // opt_f)
        self.weights: Tensor = torch.zeros((1, 1, self.h, self.w), device=host.device(), dtype=torch.float32)

        self.step_count = 0
        self.inner_loop_count = 0
        self.kdiff_step = -1

        # grid tiling (abstractdiffusion.py:35-42)
        self.enable_grid_bbox: bool = False
        self.tile_w: Optional[int] = None
        self.tile_h: Optional[int] = None
        self.tile_bs: Optional[int] = None
        self.num_tiles: Optional[int] = None
        self.num_batches: Optional[int] = None
        self.batched_bboxes: List[List[BBox]] = []

        # region prompt control (not on this path yet)
        self.enable_custom_bbox: bool = False
        self.custom_bboxes: list = []
        self.draw_background: bool = True
        self.causal_layers: Optional[bool] = None

        # noise inversion / controlnet / stablesr (not on this path yet)
        self.noise_inverse_enabled: bool = False
        self.enable_controlnet: bool = False
        self.enable_stablesr: bool = False

        # B200 engine state
        self._grid = None                     # td_grid plan
        self._tiles: Optional[Tensor] = None  # persistent [T*N, C, th, tw] scatter target
        self._icond_tiles: Optional[Tensor] = None
        self._blend_flags = 0                 # tests flip TD_FLAG_FORCE_GENERIC here
        self._rcp_weights: Optional[Tensor] = None   # RN(1/weights) when all weights are small integers
        self._shard = None                    # parallel.TileShard when tiles are sharded over ranks
        self._shard_group = None
        self._shard_fused = False
        self._exchange = None                 # parallel.PeerExchange (fused path)
        self._shard_step = 0

    # ----------------------------------------------------------------- helpers
    def _sd_model(self):
        m = getattr(host.get_shared(), "sd_model", None)
        return m if m is not None else getattr(self.p, "sd_model", None)

    @property
    def is_kdiff(self) -> bool:
        return host.is_kdiff_sampler(self.sampler_raw)

    @property
    def is_ddim(self) -> bool:
        return host.is_ddim_sampler(self.sampler_raw)

    def update_pbar(self):
        if self.pbar is None:
            return
        if self.pbar.n >= self.pbar.total:
            self.pbar.close()
            return
        st = host.get_shared().state
        if self.step_count == st.sampling_step:
            self.inner_loop_count += 1
            if self.inner_loop_count < self.total_bboxes:
                self.pbar.update()
        else:
            self.step_count = st.sampling_step
            self.inner_loop_count = 0

    def reset_buffer(self, x_in: Tensor):
        """abstractdiffusion.py:97-102.  Allocation only -- the blend kernel overwrites every pixel."""
        if self.x_buffer is None or self.x_buffer.shape != x_in.shape or self.x_buffer.dtype != x_in.dtype \
                or self.x_buffer.device != x_in.device:
            self.x_buffer = torch.zeros_like(x_in)

    def init_done(self):
        """abstractdiffusion.py:104-117: sanity check + progress accounting."""
        rcp = engine.exact_reciprocals(self.weights.detach().to("cpu", torch.float32).numpy().reshape(self.h, self.w))
        self._rcp_weights = None if rcp is None else torch.from_numpy(rcp).to(self.weights.device)
        self.total_bboxes = 0
        if self.enable_grid_bbox:
            self.total_bboxes += self.num_batches
        if self.enable_custom_bbox:
            self.total_bboxes += len(self.custom_bboxes)
        assert self.total_bboxes > 0, "Nothing to paint! No background to draw and no custom bboxes were provided."
        self.pbar = None
        if getattr(self.p, "show_tile_progress", False):
            from tqdm import tqdm
            steps = getattr(host.get_shared().state, "sampling_steps", 1)
            self.pbar = tqdm(total=self.total_bboxes * steps, desc=f"{self.method} Sampling: ")

    # ------------------------------------------------------- cond-dict access
    def _tcond_key(self, cond_dict: CondDict) -> str:
        return "crossattn" if "crossattn" in cond_dict else "c_crossattn"

    def get_tcond(self, cond_dict: CondDict) -> Tensor:
        tcond = cond_dict[self._tcond_key(cond_dict)]
        return tcond[0] if isinstance(tcond, list) else tcond

    def set_tcond(self, cond_dict: CondDict, tcond: Tensor):
        key = self._tcond_key(cond_dict)
        cond_dict[key] = [tcond] if isinstance(cond_dict[key], list) else tcond

    def _icond_key(self, cond_dict: CondDict) -> str:
        model = getattr(self._sd_model(), "model", None)
        ck = getattr(model, "conditioning_key", None)
        return "c_adm" if ck in ("crossattn-adm", "adm") else "c_concat"

    def get_icond(self, cond_dict: CondDict) -> Tensor:
        icond = cond_dict[self._icond_key(cond_dict)]
        return icond[0] if isinstance(icond, list) else icond

    def set_icond(self, cond_dict: CondDict, icond: Tensor):
        key = self._icond_key(cond_dict)
        cond_dict[key] = [icond] if isinstance(cond_dict[key], list) else icond

    def _vcond_key(self, cond_dict: CondDict) -> Optional[str]:
        return "vector" if "vector" in cond_dict else None

    def get_vcond(self, cond_dict: CondDict) -> Optional[Tensor]:
        return cond_dict.get(self._vcond_key(cond_dict))

    def set_vcond(self, cond_dict: CondDict, vcond: Optional[Tensor]):
        key = self._vcond_key(cond_dict)
        if key is not None:
            cond_dict[key] = vcond

    def make_cond_dict(self, cond_in: CondDict, tcond: Tensor, icond: Tensor, vcond: Tensor = None) -> CondDict:
        cond_out = cond_in.copy()
        self.set_tcond(cond_out, tcond)
        self.set_icond(cond_out, icond)
        self.set_vcond(cond_out, vcond)
        return cond_out

    # ------------------------------------------------------------ grid tiling
    @grid_bbox
    def init_grid_bbox(self, tile_w: int, tile_h: int, overlap: int, tile_bs: int):
        """abstractdiffusion.py:172-186, bookkeeping done by td_grid_init (C++, bit-exact)."""
        self.enable_grid_bbox = True
        g = engine.make_grid(self.w, self.h, tile_w, tile_h, overlap, tile_bs)
        self._grid = g
        self.tile_w, self.tile_h = int(g.tile_w), int(g.tile_h)
        tile_weights = self.get_tile_weights()
        tw_np = None
        if isinstance(tile_weights, Tensor):
            tw_np = tile_weights.detach().to("cpu", torch.float32).numpy()
        elif float(tile_weights) != 1.0:
            raise ValueError("scalar tile weights other than 1.0 are not part of the reference")
        weights = engine.grid_weights(g, tw_np)
        self.weights += torch.from_numpy(weights).view(1, 1, self.h, self.w).to(self.weights.device)
        bboxes = [BBox(int(x), int(y), int(w), int(h)) for x, y, w, h in engine.grid_bboxes_xywh(g)]
        self.num_tiles = len(bboxes)
        self.num_batches = int(g.num_batches)
        self.tile_bs = int(g.tile_bs)
        self.batched_bboxes = [bboxes[i * self.tile_bs:(i + 1) * self.tile_bs] for i in range(self.num_batches)]

    @grid_bbox
    def get_tile_weights(self) -> Union[Tensor, float]:
        return 1.0

    # ------------------------------------------------------------- multi-GPU
    def init_tile_shard(self, group=None, fused: bool = True):
        """Shard the tile list over the ranks of `group` (one process per GPU).  Call after init_grid_bbox.

        Each rank denoises `tiles[begin:end)` only; per step the tile outputs are exchanged (fused: peer
        reads over NVLink inside the blend kernel; else NCCL all-gather) and every rank blends the full
        latent deterministically (bit-identical across ranks and to a single-GPU run)."""
        import torch.distributed as dist
        from .. import parallel
        if self._grid is None:
            raise RuntimeError("init_tile_shard() must follow init_grid_bbox()")
        self._shard = parallel.TileShard(self.num_tiles, dist.get_rank(group), dist.get_world_size(group))
        self._shard_group, self._shard_fused = group, fused
        bboxes = [b for batch in self.batched_bboxes for b in batch]
        local = bboxes[self._shard.begin:self._shard.end]
        self.local_batched_bboxes = [local[i:i + self.tile_bs] for i in range(0, len(local), self.tile_bs)]
        return self._shard

    def _exchange_and_blend_md(self, outs, x: Tensor, N: int, C: int) -> Tensor:
        """Tile-shard tail of MultiDiffusion.sample_one_step."""
        from .. import parallel
        sh, g = self._shard, self._grid
        plane = N * C * g.tile_h * g.tile_w
        dt = outs[0].dtype if outs else x.dtype
        self._shard_step += 1
        if self._shard_fused:
            if self._exchange is None:
                self._exchange = parallel.PeerExchange(sh.chunk * plane * x.element_size(), x.device, self._shard_group)
            parity = self._shard_step & 1
            if outs:
                buf = self._exchange.local_buffer(parity, dt)[:sh.num_local * plane].view(sh.num_local * N, C, g.tile_h, g.tile_w)
                torch.cat([o.to(dt) for o in outs], dim=0, out=buf)
            self._exchange.signal()
            return parallel.blend_multidiffusion_peer(g, self._exchange, parity, sh, N, C, self.weights, dt)
        local = torch.zeros((sh.chunk * N, C, g.tile_h, g.tile_w), dtype=dt, device=x.device)
        if outs:
            torch.cat(outs, dim=0, out=local[:sh.num_local * N])
        gathered = parallel.gather_tile_outputs(local, self._shard_group)
        chunks = []
        for b in range(sh.num_chunks):
            nt = min(sh.chunk, sh.num_tiles - b * sh.chunk)
            chunks.append(gathered[b * sh.chunk * N: (b * sh.chunk + nt) * N])
        return engine.blend_multidiffusion(g, chunks, N, C, sh.chunk, self.weights, x.dtype, x_buffer=None, flags=self._blend_flags,
                                           rcp_weights=self._rcp_weights)

    # ------------------------------------------- later rows of the scope table
    @custom_bbox
    def init_custom_bbox(self, bbox_settings, draw_background: bool, causal_layers: bool):
        raise NotImplementedError("Region prompt control is not on the B200 hot path yet (SURVEY.md section 8(f)-1)")

    @noise_inverse
    def init_noise_inverse(self, *args, **kwargs):
        raise NotImplementedError("Tiled noise inversion is not on the B200 hot path yet (SURVEY.md section 8(f)-3)")

    @controlnet
    def init_controlnet(self, *args, **kwargs):
        raise NotImplementedError("ControlNet tile caches are not on the B200 hot path yet (SURVEY.md section 8(f)-2)")

    @stablesr
    def init_stablesr(self, *args, **kwargs):
        raise NotImplementedError("StableSR tile caches are not on the B200 hot path yet (SURVEY.md section 8(f)-2)")

    def reset_controlnet_tensors(self):
        pass

    def switch_controlnet_tensors(self, batch_id: int, x_batch_size: int, tile_batch_size: int, is_denoise=False):
        pass

    def switch_stablesr_tensors(self, batch_id: int):
        pass

    # ----------------------------------------------------------- engine glue
    def _check_input(self, x_in: Tensor) -> Tensor:
        if not x_in.is_cuda:
            raise RuntimeError(f"{self.method}: latent is on {x_in.device}; the B200 path has no CPU fallback")
        if self._grid is None:
            raise RuntimeError(f"{self.method}: init_grid_bbox() has not been called")
        if self.weights.device != x_in.device:
            self.weights = self.weights.to(x_in.device)
        if self._rcp_weights is not None and self._rcp_weights.device != x_in.device:
            self._rcp_weights = self._rcp_weights.to(x_in.device)
        return x_in.contiguous()

    def _scatter_all(self, x_in: Tensor) -> Tensor:
        """One launch: every grid tile of this step, tile-major (multidiffusion.py:155 for all batches)."""
        self._tiles = engine.scatter_tiles(self._grid, x_in, out=self._tiles, flags=self._blend_flags)
        return self._tiles

    def _tile_batch(self, tiles: Tensor, batch_id: int, N: int) -> Tensor:
        lo = batch_id * self.tile_bs * N
        hi = min((batch_id + 1) * self.tile_bs, self.num_tiles) * N
        return tiles[lo:hi]

    def _icond_tile_batches(self, icond: Tensor):
        """img2img: spatial icond is cropped per tile like the latent (multidiffusion.py:121-122)."""
        self._icond_tiles = engine.scatter_tiles(self._grid, icond, out=self._icond_tiles, flags=self._blend_flags)
        return self._icond_tiles

    def repeat_tensor(self, x: Tensor, n: int) -> Tensor:
        """Repeat on dim 0 (multidiffusion.py:100-110): expand when B == 1, else tile.

        The tiled copy is memoised per source tensor OBJECT (+ version): all T/tile_bs batches of a step
        repeat the same cond, so the reference's per-batch `repeat` becomes one launch per step."""
        if n == 1:
            return x
        r_dims = x.dim() - 1
        if x.shape[0] == 1:
            return x.expand([n] + [-1] * r_dims)
        cache = self.__dict__.setdefault("_repeat_cache", {})
        key = (id(x), n)
        hit = cache.get(key)
        if hit is not None and hit[0] is x and hit[1] == x._version:
            return hit[2]
        out = x.repeat([n] + [1] * r_dims)
        if len(cache) > 16:
            cache.clear()
        cache[key] = (x, x._version, out)   # the strong reference keeps id(x) from being recycled
        return out
