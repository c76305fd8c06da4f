#!/usr/bin/env python
"""Per-kernel SASS mnemonic summary of the in-tree library (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA / UTCBAR / LDTM /
STTM), bulk copies (UBLKCP), mbarrier (SYNCS) and packed-fp32 / conversion instructions, and how many UTCHMMA sit under a uniform
predicate (issued by the elect.sync lane of a converged warp, no election loop).   python tools/sass_summary.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nope_nerf_b200", "libnope_nerf_b200.so")
KEEP = re.compile(r"^(UTC|LDTM|STTM|UBLKCP|SYNCS|ELECT|F2FP|FADD2|FMUL2|FFMA2)")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_tc_kernels_r2.txt")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    per = collections.OrderedDict(); uni = collections.Counter(); cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); per.setdefault(cur, collections.Counter()); continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if cur and m:
            op = m.group(2)
            if KEEP.match(op):
                per[cur][op] += 1
                if op.startswith("UTCHMMA") and m.group(1) and m.group(1).lstrip("@!").startswith("UP"):
                    uni[cur] += 1
    with open(out, "w") as f:
        f.write("SASS mnemonics per kernel of nope_nerf_b200/libnope_nerf_b200.so (cuobjdump -sass, sm_100a; tools/sass_summary.py): tcgen05 = UTCHMMA / "
                "UTCBAR / LDTM / STTM, bulk copies = UBLKCP, mbarrier = SYNCS\n\n")
        for k, c in per.items():
            if not any(o.startswith(("UTC", "LDTM", "UBLKCP")) for o in c):
                continue
            f.write(k + "\n   " + ", ".join("%s x%d" % (o, n) for o, n in sorted(c.items())) + "\n")
            n = sum(v for o, v in c.items() if o.startswith("UTCHMMA"))
            f.write("   UTCHMMA under a uniform predicate (elect.sync lane, no election loop): %d of %d\n\n" % (uni[k], n))
    print(out)


if __name__ == "__main__":
    main()
