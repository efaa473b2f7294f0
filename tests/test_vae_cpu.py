"""Tiled-VAE pieces that need no GPU: the oracle against the reference-generated fixtures,
and the C-ABI bookkeeping (split_tiles / get_best_tile_size) against both."""
import os

import numpy as np
import pytest
import torch

from oracle import ldm_vae, vae
from oracle.make_golden import VAE_CASES, VAE_SUBSAMPLE, vae_case_inputs


def test_split_tiles_and_best_tile_size_match_reference(golden_dir):
    from multidiffusion_upscaler_for_automatic1111_b200 import VAEHook
    g = np.load(os.path.join(golden_dir, "vae_geometry.npz"))
    for i, (h, w, ts, dec) in enumerate(g["cases"]):
        hook = VAEHook(None, int(ts), bool(dec), True, True, False)
        ib, ob = hook.split_tiles(int(h), int(w))
        assert np.array_equal(np.array(ib, np.int32), g[f"in_{i}"]), (h, w, ts, dec)
        assert np.array_equal(np.array(ob, np.int32), g[f"out_{i}"]), (h, w, ts, dec)
        oi, oo = vae.split_tiles(int(h), int(w), int(ts), hook.pad, bool(dec))
        assert oi == ib and oo == ob
    hook = VAEHook(None, 64, True, True, True, False)
    for lo, up, want in g["best_tile"]:
        assert hook.get_best_tile_size(int(lo), int(up)) == int(want)
        assert vae.get_best_tile_size(int(lo), int(up)) == int(want)


def test_baseline_cfg4_geometry(golden_dir):
    """BASELINE config 4: 8192^2 RGB, decoder tile 96 (121 tiles <= 118x118), encoder tile 1536 (36 tiles)."""
    from multidiffusion_upscaler_for_automatic1111_b200 import VAEHook
    g = np.load(os.path.join(golden_dir, "vae_geometry.npz"))
    ib, ob = VAEHook(None, 96, True, True, True, False).split_tiles(1024, 1024)
    assert len(ib) == 121 and np.array_equal(np.array(ib, np.int32), g["cfg4_dec"])
    assert ob[0] == [0, 856, 0, 856] and ib[-1] == [960, 1024, 960, 1024]
    ib, _ = VAEHook(None, 1536, False, True, True, False).split_tiles(8192, 8192)
    assert len(ib) == 36 and ib[0] == [0, 1440, 0, 1440] and np.array_equal(np.array(ib, np.int32), g["cfg4_enc"])


@pytest.mark.parametrize("case", VAE_CASES, ids=[c[0] for c in VAE_CASES])
def test_oracle_matches_reference_fixture(golden_dir, case):
    name, is_dec, fast, cf, H, W, tile = case
    g = np.load(os.path.join(golden_dir, "vae_small.npz"))
    net, z = vae_case_inputs(is_dec, H, W)
    with torch.no_grad():
        y = vae.vae_hook_call(net, z, tile, is_dec, fast, cf)
    assert list(y.shape) == list(g[name + "_shape"])
    sub = y[:, :, ::VAE_SUBSAMPLE, ::VAE_SUBSAMPLE].numpy()
    scale = float(g[name + "_absmax"])
    # same algorithm, possibly another CPU's conv kernels: fp32 round-off only
    assert np.abs(sub - g[name]).max() <= 2e-5 * max(scale, 1.0)


def test_tiled_differs_from_untiled_but_oracle_tracks_reference(golden_dir):
    """Parity target is the reference's TILED output (fast mode deviates visibly from the plain VAE)."""
    g = np.load(os.path.join(golden_dir, "vae_small.npz"))
    net, z = vae_case_inputs(True, 40, 52)
    with torch.no_grad():
        full = net(z)[:, :, ::VAE_SUBSAMPLE, ::VAE_SUBSAMPLE].numpy()
    assert np.abs(full - g["dec_fast"]).max() > 0.05
    assert np.abs(full - g["dec_tiny_bypass"][:, :, :1, :1]).size > 0


def test_merge_is_weighted_average_of_variances():
    v = [torch.tensor([1.0, 4.0]), torch.tensor([3.0, 0.0])]
    m = [torch.tensor([0.0, 1.0]), torch.tensor([2.0, 3.0])]
    var, mean = vae.merge_tile_stats(v, m, [100, 300])
    assert torch.allclose(var, torch.tensor([2.5, 1.0])) and torch.allclose(mean, torch.tensor([1.5, 2.5]))


def test_task_queue_shape_of_sd_vae():
    """SD autoencoder: decoder = 123 tasks / 30 GroupNorm sites, encoder = 22 sites."""
    from multidiffusion_upscaler_for_automatic1111_b200 import tilevae
    dec = ldm_vae.Decoder(ch=32)
    enc = ldm_vae.Encoder(ch=32)
    for build in (tilevae.build_task_queue, vae.build_task_queue):
        qd, qe = build(dec, True), build(enc, False)
        assert len(qd) == 123 and sum(t[0] == "pre_norm" for t in qd) == 30
        assert sum(t[0] == "pre_norm" for t in qe) == 22
        assert [t[0] for t in qd[:4]] == ["conv_in", "store_res", "pre_norm", "silu"]


def test_vaehook_refuses_cpu_network():
    from multidiffusion_upscaler_for_automatic1111_b200 import VAEHook
    net, z = vae_case_inputs(True, 40, 52)
    hook = VAEHook(net, 16, True, True, True, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hook(z)
    small = torch.zeros(1, 4, 20, 30)
    assert hook(small).shape == (1, 3, 160, 240)   # tiny input: original_forward, untiled (tilevae.py:382-384)
