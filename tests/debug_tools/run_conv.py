"""Hand-run launcher for ncu captures of the tcgen05 convolution (GPU box).
Usage: python tests/debug_tools/run_conv.py H W Cin Cout k [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from multidiffusion_upscaler_for_automatic1111_b200 import vae_ops as ops  # noqa: E402

H, W, Cin, Cout, k = (int(v) for v in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
x = (torch.randn((1, H, W, Cin), device="cuda") * 0.5).half()
w = (torch.randn((k * k, Cout, Cin), device="cuda") * 0.02).half()
b = torch.zeros(Cout, device="cuda")
y = torch.empty((1, H, W, Cout), device="cuda", dtype=torch.float16)
for _ in range(reps):
    ops.conv2d_nhwc(x, w, b, ksize=k, pad=(k // 2, k // 2), out=y)
torch.cuda.synchronize()
print("ok")
