#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths of libmcarray_b200.so: tcgen05.mma (UTCHMMA), tcgen05.ld (LDTM),
TMA tensor loads (UTMALDG), bulk copies (UBLKCP), warp reductions (REDUX) and packed fp32 (FADD2 / FMUL2 / FFMA2).
usage: python tools/sass_counts.py > profiles/sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mcarray_b200", "libmcarray_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "REDUX", "FADD2", "FMUL2", "FFMA2", "SHFL", "BAR", "SYNCS"]
counts, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        cur["_total"] += 1
        if op in OPS:
            cur[op] += 1
print(f"SASS mnemonic counts of {os.path.relpath(so, ROOT)} (sm_100a), one line per kernel; columns: total instructions, then " + " ".join(OPS))
tot = collections.Counter()
for name, c in counts.items():
    tot.update(c)
    print(f"{name[:110]:110s} {c['_total']:7d} " + " ".join(f"{c[o]:5d}" for o in OPS))
print(f"{'ALL KERNELS':110s} {tot['_total']:7d} " + " ".join(f"{tot[o]:5d}" for o in OPS))
