"""Where one full-size decoder tile (latent 118 x 118 -> 944 x 944 px) spends its time on the tensor-core backend:
every backend call is bracketed by CUDA events (synchronised: the numbers are per-op device times, not a pipeline trace).

    python tests/debug_tools/vae_breakdown.py [latent_edge]
"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multidiffusion_upscaler_for_automatic1111_b200 import tilevae, vae_engine as ve, vae_ops  # noqa: E402


def main():
    edge = int(sys.argv[1]) if len(sys.argv) > 1 else 118
    dev = torch.device("cuda")
    net = bench._sd_vae_half(True, 1).to(dev).half()
    hook = tilevae.VAEHook(net, 96, True, fast_decoder=True, fast_encoder=True, color_fix=False)
    z = torch.randn((1, 4, 160, 160), generator=torch.Generator().manual_seed(7)).half().to(dev)
    hook(z)                                             # weights packed, statistics path warmed
    program = hook._program
    be = ve.pick_backend(program, dev, torch.float16)
    ex = ve.Executor(program, be)
    # frozen statistics for every site (values are irrelevant for timing)
    for i in range(program.num_sites):
        ex.frozen[i] = (torch.zeros(32, device=dev), torch.ones(32, device=dev))
    times = collections.defaultdict(float)
    counts = collections.Counter()

    def wrap(obj, name, label=None):
        fn = getattr(obj, name)

        def timed(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **k)
            e.record(); e.synchronize()
            key = label(a, k, out) if label else name
            times[key] += s.elapsed_time(e); counts[key] += 1
            return out
        setattr(obj, name, timed)

    def conv_label(a, k, out):
        x = a[0]
        kind = "upconv" if "w16" in k or (len(a) > 1 and a[1].shape[0] == 16) else "conv"
        o = out[0] if isinstance(out, tuple) else out
        return (f"{kind} {tuple(x.shape[1:3])} {x.shape[3]}->{o.shape[3]} k{k.get('ksize', 3)}{' +gn' if k.get('post') else ''}"
                f"{' +res' if k.get('residual') is not None else ''}{' dual' if k.get('dual') else ''}")
    wrap(vae_ops, "conv2d_nhwc", conv_label)
    wrap(vae_ops, "upconv2x_nhwc", conv_label)
    wrap(vae_ops, "upsample2x_nhwc", lambda a, k, o: f"upsample {tuple(a[0].shape[1:])}")
    wrap(vae_ops, "gn_apply_nhwc", lambda a, k, o: f"gn_apply {tuple(a[0].shape[1:])}")
    wrap(vae_ops, "gemm_nt", lambda a, k, o: "attn gemm")
    wrap(vae_ops, "softmax_rows", lambda a, k, o: "attn softmax")
    wrap(vae_ops, "nchw_to_nhwc", lambda a, k, o: "load (nchw->nhwc)")
    wrap(vae_ops, "nhwc_to_nchw_region", lambda a, k, o: "paste (nhwc->nchw)")
    zt = torch.randn((1, 4, edge, edge), device=dev).half()
    result = torch.zeros((1, 3, edge * 8, edge * 8), device=dev, dtype=torch.float16)
    for rep in range(3):
        if rep == 2:
            times.clear(); counts.clear()
        st = ve.TileState(be.load(zt))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        assert ex.run(st) is None
        be.paste(st.act, result, (0, edge, 0, edge), (0, edge * 8, 0, edge * 8), True)
        e.record(); e.synchronize()
    total = sum(times.values())
    rows = sorted(times.items(), key=lambda kv: -kv[1])
    cat = collections.defaultdict(float)
    for k, v in rows:
        cat[k.split(" ")[0]] += v
    print(json.dumps({"latent_edge": edge, "sum_of_ops_ms": total, "wall_ms_with_event_syncs": s.elapsed_time(e),
                      "by_kind_ms": dict(sorted(cat.items(), key=lambda kv: -kv[1])),
                      "ops": [{"op": k, "ms": round(v, 4), "calls": counts[k]} for k, v in rows]}, indent=1))


if __name__ == "__main__":
    main()
