"""The executor's epilogue-fusion peepholes (vae_engine.Executor.run), checked on the host with a plain-torch backend that
offers the tensor-core backend's fusing interface (conv(..., post, dual), attention(..., post, dual), norm_affine): the
fused walk must produce what the op-by-op walk produces, and in fast mode no stand-alone GroupNorm pass may be left."""
import pytest
import torch
import torch.nn.functional as F

from multidiffusion_upscaler_for_automatic1111_b200 import vae_engine as ve
from oracle import ldm_vae

GROUPS, EPS = 32, 1e-6


def _gn(a, mean, var, module, act):
    n, c, h, w = a.shape
    cpg = c // GROUPS
    y = (a - mean.repeat_interleave(cpg).view(1, c, 1, 1)) / torch.sqrt(var.repeat_interleave(cpg).view(1, c, 1, 1) + EPS)
    y = y * module.weight.view(1, c, 1, 1) + module.bias.view(1, c, 1, 1)
    return F.silu(y) if act else y


class PlainBackend:
    """NCHW fp64 torch ops, one op at a time (no fusion)."""
    fuses_norm = False

    def __init__(self):
        self.norm_calls = 0

    def conv(self, a, op, skip):
        if op.upsample_first:
            a = F.interpolate(a, scale_factor=2.0, mode="nearest")
        if op.downsample:
            a = F.pad(a, (0, 1, 0, 1))
        y = op.module(a)
        return y + skip if skip is not None else y

    def shortcut(self, a, op):
        return a if op.module is None else op.module(a)

    def norm(self, a, op, mean, var):
        self.norm_calls += 1
        return _gn(a, mean, var, op.module, op.act)

    def attention(self, a, op, skip):
        m = op.module
        n, c, h, w = a.shape
        q, k, v = m.q(a).reshape(n, c, h * w), m.k(a).reshape(n, c, h * w), m.v(a).reshape(n, c, h * w)
        s = torch.softmax(torch.bmm(q.transpose(1, 2), k) * (c ** -0.5), dim=2)
        o = torch.bmm(v, s.transpose(1, 2)).reshape(n, c, h, w)
        y = m.proj_out(o)
        return y + skip if skip is not None else y

    def tanh(self, a):
        return torch.tanh(a)


class FusingBackend(PlainBackend):
    """Same arithmetic behind the fusing interface of TensorCoreBackend."""
    fuses_norm = True

    def __init__(self):
        super().__init__()
        self.fused_single, self.fused_dual = 0, 0

    def norm_affine(self, op, mean, var):
        c = op.module.num_channels
        cpg = c // GROUPS
        rstd = (1.0 / torch.sqrt(var + EPS)).repeat_interleave(cpg)
        scale = rstd * op.module.weight
        return scale, op.module.bias - mean.repeat_interleave(cpg) * scale, bool(op.act)

    def _post(self, y, post, dual):
        if post is None:
            return y
        z = y * post[0].view(1, -1, 1, 1) + post[1].view(1, -1, 1, 1)
        z = F.silu(z) if post[2] else z
        if dual:
            self.fused_dual += 1
            return y, z
        self.fused_single += 1
        return z

    def conv(self, a, op, skip, post=None, dual=False):
        return self._post(super().conv(a, op, skip), post, dual)

    def attention(self, a, op, skip, post=None, dual=False):
        return self._post(super().attention(a, op, skip), post, dual)


def _walk(program, backend, frozen, x):
    ex = ve.Executor(program, backend)
    ex.frozen = list(frozen)
    st = ve.TileState(x.clone())
    assert ex.run(st) is None
    return st.act


@pytest.mark.parametrize("is_decoder", [True, False])
def test_fused_walk_equals_op_by_op_walk(is_decoder, monkeypatch):
    torch.manual_seed(0)
    cls = ldm_vae.Decoder if is_decoder else ldm_vae.Encoder
    net = ldm_vae.seeded_init(cls(ch=32, ch_mult=(1, 2, 4, 4), num_res_blocks=2), 7).double().eval()
    program = ve.compile_program(net, is_decoder)
    g = torch.Generator().manual_seed(3)
    frozen = [(torch.randn(GROUPS, generator=g, dtype=torch.float64) * 0.2, torch.rand(GROUPS, generator=g, dtype=torch.float64) + 0.5)
              for _ in range(program.num_sites)]
    x = torch.randn((1, 4, 6, 7) if is_decoder else (1, 3, 24, 32), generator=g, dtype=torch.float64)
    with torch.no_grad():
        plain = PlainBackend()
        want = _walk(program, plain, frozen, x)
        assert plain.norm_calls == program.num_sites
        fused = FusingBackend()
        got = _walk(program, fused, frozen, x)
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-10, atol=1e-10)
    # every GroupNorm site of a fast-mode tile rides in a producer's epilogue
    assert fused.norm_calls == 0 and fused.fused_single + fused.fused_dual == program.num_sites
    # boundaries between blocks (producer -> Skip -> Norm) are the two-output form: every norm1 / attention norm
    n_boundaries = sum(1 for i, op in enumerate(program.ops[:-2]) if isinstance(op, (ve.Conv, ve.Attention))
                       and isinstance(program.ops[i + 1], ve.Skip) and isinstance(program.ops[i + 2], ve.Norm))
    assert fused.fused_dual == n_boundaries > 0

    # with the two-output form switched off the boundaries fall back to a stand-alone pass, same result
    monkeypatch.setattr(ve, "DUAL_OUTPUT", False)
    with torch.no_grad():
        half = FusingBackend()
        got2 = _walk(program, half, frozen, x)
    assert torch.allclose(got2, want, rtol=1e-10, atol=1e-10)
    assert half.fused_dual == 0 and half.norm_calls == n_boundaries


def test_barrier_sites_are_never_fused():
    """Slow mode: a site without frozen statistics stops the walk ON the Norm op, raw activation and shortcut intact."""
    net = ldm_vae.seeded_init(ldm_vae.Decoder(ch=32, ch_mult=(1, 2), num_res_blocks=1), 5).double().eval()
    program = ve.compile_program(net, True)
    ex = ve.Executor(program, FusingBackend())
    st = ve.TileState(torch.randn((1, 4, 5, 5), dtype=torch.float64))
    seen = 0
    with torch.no_grad():
        while True:
            op = ex.run(st)
            if op is None:
                break
            assert isinstance(op, ve.Norm) and program.ops[st.pc] is op
            var, mean = torch.var_mean(st.act.view(1, GROUPS, -1), dim=2, unbiased=False)
            ex.apply_barrier(st, op, mean.view(-1), var.view(-1))
            seen += 1
    assert seen == program.num_sites and ex.be.fused_single == 0 and ex.be.fused_dual == 0
