"""Hand-run launcher for ncu captures of the per-step kernels at BASELINE cfg2 (GPU box).
Usage: python tests/debug_tools/run_rows.py [flags_hex] [reps]   (flags as in td_b200.h, e.g. 0x400 = round-1 kernels)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import bench  # noqa: E402


def main():
    flags = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = bench.Workload(dev, 0, 1, 4, "none")
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        wl.set_stream()
        for i in range(reps):
            wl.scatter(i % 4, flags)
            wl.blend(i % 4, flags)
            wl.blend_mod(i % 4, flags)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
