"""MultiDiffusion (arXiv 2302.08113) tile method with the reference's surface
(tile_methods/multidiffusion.py), executed by two sm_100a kernels per sampler step:

    td_scatter_tiles          <- `torch.cat([x_in[bbox.slicer] ...])` for every batch   (:155)
    td_blend_multidiffusion   <- `x_buffer[slicer] += tile` x T  +  `torch.where(...)`   (:166-167, :208)

The UNet (`repeat_func`) stays the host application's; it receives views of the
persistent tile buffer and its output tensors are read in place by the blend
kernel (pointer table, no concatenation copy).
"""
from __future__ import annotations

from typing import Callable, List, Union

import torch
from torch import Tensor

from .. import engine, host
from ..tile_utils.utils import BBox, custom_bbox, keep_signature
from .abstractdiffusion import AbstractDiffusion, CondDict


class MultiDiffusion(AbstractDiffusion):

    def __init__(self, p, *args, **kwargs):
        super().__init__(p, *args, **kwargs)
        assert p.sampler_name != "UniPC", "MultiDiffusion is not compatible with UniPC!"

    # ---------------------------------------------------------------- hooks
    def hook(self):
        """Patch `sampler.model_wrap_cfg.inner_model.forward` (multidiffusion.py:15-29)."""
        inner = self.sampler.model_wrap_cfg.inner_model
        self.sampler_forward = inner.forward
        inner.forward = self.kdiff_forward if self.is_kdiff else self.ddim_forward

    @staticmethod
    def unhook():
        # the patched sampler object dies with the generation job (multidiffusion.py:31-35)
        pass

    def reset_buffer(self, x_in: Tensor):
        super().reset_buffer(x_in)

    @custom_bbox
    def init_custom_bbox(self, *args):
        super().init_custom_bbox(*args)

    # ------------------------------------------------------- kernel hijacks
    @torch.no_grad()
    @keep_signature
    def kdiff_forward(self, x_in: Tensor, sigma_in: Tensor, cond: CondDict) -> Tensor:
        def org_func(x: Tensor) -> Tensor:
            return self.sampler_forward(x, sigma_in, cond=cond)

        def repeat_func(x_tile: Tensor, bboxes: List[BBox]) -> Tensor:
            sigma_tile = self.repeat_tensor(sigma_in, len(bboxes))
            cond_tile = self.repeat_cond_dict(cond, bboxes)
            return self.sampler_forward(x_tile, sigma_tile, cond=cond_tile)

        return self.sample_one_step(x_in, org_func, repeat_func, None)

    @torch.no_grad()
    @keep_signature
    def ddim_forward(self, x_in: Tensor, ts_in: Tensor, cond: Union[CondDict, Tensor]) -> Tensor:
        def org_func(x: Tensor) -> Tensor:
            return self.sampler_forward(x, ts_in, cond=cond)

        def repeat_func(x_tile: Tensor, bboxes: List[BBox]) -> Tensor:
            n_rep = len(bboxes)
            ts_tile = self.repeat_tensor(ts_in, n_rep)
            if isinstance(cond, dict):
                cond_tile = self.repeat_cond_dict(cond, bboxes)
            else:
                cond_tile = self.repeat_tensor(cond, n_rep)
            return self.sampler_forward(x_tile, ts_tile, cond=cond_tile)

        return self.sample_one_step(x_in, org_func, repeat_func, None)

    def repeat_cond_dict(self, cond_in: CondDict, bboxes: List[BBox]) -> CondDict:
        """Per-batch cond (multidiffusion.py:112-129): text/vector cond repeated, spatial icond cropped per tile."""
        n_rep = len(bboxes)
        tcond = self.repeat_tensor(self.get_tcond(cond_in), n_rep)
        icond = self.get_icond(cond_in)
        if tuple(icond.shape[2:]) == (self.h, self.w):
            icond = self._crop_icond(icond, bboxes)
        else:
            icond = self.repeat_tensor(icond, n_rep)
        vcond = self.get_vcond(cond_in)
        if vcond is not None:
            vcond = self.repeat_tensor(vcond, n_rep)
        return self.make_cond_dict(cond_in, tcond, icond, vcond)

    def _crop_icond(self, icond: Tensor, bboxes: List[BBox]) -> Tensor:
        """Spatial icond -> tile batch with the same scatter kernel as the latent."""
        first = self._bbox_index(bboxes[0])
        return engine.scatter_tiles(self._grid, icond, tile_begin=first, tile_end=first + len(bboxes),
                                    flags=self._blend_flags)

    def _bbox_index(self, bbox: BBox) -> int:
        g = self._grid
        r = list(g.ys[:g.rows]).index(bbox.y)
        c = list(g.xs[:g.cols]).index(bbox.x)
        return r * g.cols + c

    def sample_one_step(self, x_in: Tensor, org_func: Callable, repeat_func: Callable, custom_func: Callable) -> Tensor:
        """One denoiser call over the whole latent, tile by tile (multidiffusion.py:131-218, grid part).

        Returns a fresh fp32 tensor, as the reference's `torch.where` against the
        fp32 weights does even for fp16 latents.
        """
        N, C, H, W = x_in.shape
        if (H, W) != (self.h, self.w):
            # hires-fix second pass is not tiled (multidiffusion.py:141-144)
            self.reset_controlnet_tensors()
            return org_func(x_in)

        x = self._check_input(x_in)
        if not self.draw_background:
            raise NotImplementedError("draw_background=False needs region prompt control (SURVEY.md section 8(f)-1)")

        if self._shard is not None:
            return self._sample_one_step_sharded(x_in, x, repeat_func, N, C)

        tiles = self._scatter_all(x)
        outs = []
        for batch_id, bboxes in enumerate(self.batched_bboxes):
            if host.interrupted():
                return x_in
            x_tile = self._tile_batch(tiles, batch_id, N)
            self.switch_controlnet_tensors(batch_id, N, len(bboxes))
            self.switch_stablesr_tensors(batch_id)
            outs.append(repeat_func(x_tile, bboxes))
            self.update_pbar()

        return engine.blend_multidiffusion(self._grid, outs, N, C, self.tile_bs, self.weights, x.dtype,
                                           x_buffer=None, flags=self._blend_flags, rcp_weights=self._rcp_weights)

    def _sample_one_step_sharded(self, x_in: Tensor, x: Tensor, repeat_func: Callable, N: int, C: int) -> Tensor:
        """This rank's chunk of the tile list, then exchange + deterministic blend (init_tile_shard)."""
        sh = self._shard
        outs = []
        if sh.num_local > 0:
            self._tiles = engine.scatter_tiles(self._grid, x, out=self._tiles, tile_begin=sh.begin, tile_end=sh.end,
                                               flags=self._blend_flags)
            off = 0
            for batch_id, bboxes in enumerate(self.local_batched_bboxes):
                if host.interrupted():
                    return x_in
                x_tile = self._tiles[off * N:(off + len(bboxes)) * N]
                off += len(bboxes)
                outs.append(repeat_func(x_tile, bboxes))
                self.update_pbar()
        return self._exchange_and_blend_md(outs, x, N, C)

    def get_noise(self, x_in: Tensor, sigma_in: Tensor, cond_in: CondDict, step: int) -> Tensor:
        """Tiled eps prediction used by noise inversion (multidiffusion.py:220-243, grid part)."""
        cond_orig = cond_in.copy()
        sd_model = self._sd_model()

        def org_func(x: Tensor):
            return sd_model.apply_model(x, sigma_in, cond=cond_orig)

        def repeat_func(x_tile: Tensor, bboxes: List[BBox]):
            sigma_tile = sigma_in.repeat(len(bboxes))
            cond_out = self.repeat_cond_dict(cond_orig, bboxes)
            return sd_model.apply_model(x_tile, sigma_tile, cond=cond_out)

        return self.sample_one_step(x_in, org_func, repeat_func, None)
