"""Counts the Blackwell-specific SASS mnemonics per kernel of libtd_b200.so (cuobjdump -sass): the evidence table
profiles/r02_sass_evidence.json (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA
tensor copies, UBLKCP = cp.async.bulk, LDGSTS = cp.async).  Usage: python tests/debug_tools/sass_evidence.py"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
LIB = os.path.join(ROOT, "multidiffusion_upscaler_for_automatic1111_b200", "libtd_b200.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "UTCBAR", "HMMA", "STG.E.ENL2.256", "LDG.E.ENL2.256"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = out.setdefault(name[:160], {})
            continue
        if cur is None:
            continue
        for w in WANT:
            if re.search(r"\b" + re.escape(w), line):
                cur[w] = cur.get(w, 0) + 1
    out = {k: v for k, v in out.items() if v}
    path = os.path.join(ROOT, "profiles", "r02_sass_evidence.json")
    json.dump({"library": "multidiffusion_upscaler_for_automatic1111_b200/libtd_b200.so", "tool": "cuobjdump -sass (CUDA 12.9), counts of instruction lines per kernel",
               "kernels": out}, open(path, "w"), indent=1, sort_keys=True)
    for k in sorted(out):
        if any(x in out[k] for x in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM")):
            print(k[:110], out[k])
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
