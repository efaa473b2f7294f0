"""Host-side pieces of the tensor-core VAE path that need no GPU: the folded taps of `upsample2x -> conv3x3`
(vae_ops.fold_upsample_weight) and argument checks of the C-ABI entry points that run before any device work."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from multidiffusion_upscaler_for_automatic1111_b200 import _cabi, vae_ops


def _folded_conv(x, w16, cout):
    """What td_upconv2x_nhwc computes, in plain torch: four 2x2 parity convolutions of the low-resolution image."""
    n, ci, h, w = x.shape
    out = torch.zeros((n, cout, 2 * h, 2 * w), dtype=x.dtype)
    xp = F.pad(x, (1, 1, 1, 1))
    for py in range(2):
        for px in range(2):
            acc = 0
            for ty in range(2):
                for tx in range(2):
                    tap = w16[(py * 2 + px) * 4 + ty * 2 + tx][:, :ci].to(x.dtype)          # [Cout, Cin]
                    patch = xp[:, :, py + ty:py + ty + h, px + tx:px + tx + w]               # rows i-1+py+ty, cols j-1+px+tx
                    acc = acc + torch.einsum("oc,nchw->nohw", tap, patch)
            out[:, :, py::2, px::2] = acc
    return out


@pytest.mark.parametrize("shape", [(1, 3, 5, 6, 4), (2, 8, 7, 3, 5)])
def test_folded_taps_reproduce_upsample_then_conv(shape):
    n, ci, h, w, co = shape
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, ci, h, w), generator=g, dtype=torch.float64)
    wt = torch.randn((co, ci, 3, 3), generator=g, dtype=torch.float64)
    want = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), wt, padding=1)
    w16 = vae_ops.fold_upsample_weight(wt, torch.float64, cin_pad=ci)
    assert w16.shape == (16, co, ci)
    got = _folded_conv(x, w16, co)
    assert (got - want).abs().max().item() < 1e-5 * want.abs().max().item()      # taps are summed in fp32 by design


def test_folded_taps_are_summed_in_fp32_and_rounded_once():
    g = torch.Generator().manual_seed(2)
    wt = (torch.randn((4, 8, 3, 3), generator=g) * 0.1).half()
    w16 = vae_ops.fold_upsample_weight(wt, torch.float16)
    assert w16.dtype == torch.float16 and w16.shape == (16, 4, 64) and (w16[:, :, 8:] == 0).all()
    # parity (0, 0), tap (1, 1) = rows {1, 2} x cols {1, 2} of the 3x3 kernel
    want = (wt[:, :, 1:, 1:].float().sum(dim=(2, 3))).half()
    assert torch.equal(w16[0 * 4 + 1 * 2 + 1][:, :8], want)
    # parity (1, 1), tap (1, 1) = the single corner tap (2, 2)
    assert torch.equal(w16[3 * 4 + 3][:, :8], wt[:, :, 2, 2])


def _desc(**kw):
    base = dict(N=1, H=8, W=8, Cin=64, Cout=64, kh=3, kw=3, stride=1, pad_top=1, pad_left=1, OH=16, OW=16, dtype=_cabi.TD_F16,
                bias_per_row=0, alpha=1.0, x_pitch=64, w_pitch=64, y_pitch=64, res_pitch=0)
    base.update(kw)
    return _cabi.TdConvDesc(**base)


@pytest.mark.parametrize("bad", [dict(OH=8), dict(stride=2), dict(pad_top=0), dict(kh=1, kw=1), dict(bias_per_row=1)])
def test_upconv_rejects_descriptors_that_are_not_an_upsample_block(bad):
    buf = (ctypes.c_char * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    d = _desc(**bad)
    rc = _cabi.lib.td_upconv2x_nhwc(ctypes.byref(d), p, p, None, p, None)
    assert rc in (_cabi.TD_ERR_INVALID_ARG, _cabi.TD_ERR_UNSUPPORTED), (bad, rc)


def test_conv_second_output_needs_the_post_stage():
    buf = (ctypes.c_char * 4096)()
    addr = (ctypes.addressof(buf) + 255) // 256 * 256
    p = ctypes.c_void_p(addr)
    d = _desc(OH=8, OW=8)
    d.y2, d.y2_pitch = addr, 64          # y2 without post_scale / post_shift
    rc = _cabi.lib.td_conv2d_nhwc(ctypes.byref(d), p, p, None, None, p, None)
    assert rc != _cabi.TD_OK and b"y2" in _cabi.lib.td_last_error()


def test_push_regions_checks_its_tables():
    lib = _cabi.lib
    assert lib.td_push_regions(None, 1, None, 0, None, None, 0, None, None) == _cabi.TD_ERR_INVALID_ARG          # regions missing
    assert lib.td_push_regions(None, _cabi.TD_MAX_PUSH_REGIONS + 1, None, 0, None, None, 0, None, None) == _cabi.TD_ERR_INVALID_ARG
    assert lib.td_push_regions(None, 0, None, 1, None, None, 0, None, None) == _cabi.TD_ERR_INVALID_ARG          # targets missing
    reg = (_cabi.TdPushRegion * 1)()
    reg[0] = _cabi.TdPushRegion(0x1008, 0x2000, 1, 1, 64, 64, 64, 64, 64)      # source not 16-byte aligned
    assert lib.td_push_regions(reg, 1, None, 0, None, None, 0, None, None) == _cabi.TD_ERR_INVALID_ARG
    assert b"aligned" in lib.td_last_error()
    flags = (ctypes.c_uint32 * 4)()
    assert lib.td_push_regions(None, 0, None, 0, None, flags, 1, None, None) == _cabi.TD_ERR_INVALID_ARG         # wait without an expect counter
