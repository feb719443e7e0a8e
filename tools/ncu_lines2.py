#!/usr/bin/env python
"""Per source line totals of an ncu --set full --import-source on capture: instructions, stall samples, shared wavefronts.
usage: tools/ncu_lines2.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, agg = "", None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 20 and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
    elif len(r) > 20 and hdr and r[0].strip().isdigit():
        key = (cur_file, int(r[0]), r[1][:90])
        def f(name):
            try:
                return float(r[hdr[name]])
            except (KeyError, ValueError):
                return 0.0
        a = agg.setdefault(key, [0, 0, 0, 0, 0])
        a[0] += f("Instructions Executed"); a[1] += f("# Samples"); a[2] += f("L1 Wavefronts Shared"); a[3] += f("L1 Wavefronts Shared Ideal")
        a[4] += f("stall_mio") + f("stall_short_sb")
tot = [sum(v[i] for v in agg.values()) for i in range(5)]
print(f"total: inst {tot[0]:.3g} samples {tot[1]:.0f} shared wavefronts {tot[2]:.3g} (ideal {tot[3]:.3g})")
for name, idx in (("shared wavefronts", 2), ("samples", 1), ("instructions", 0)):
    print(f"\n-- top lines by {name}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
        print(f"{k[0]}:{k[1]:<5d} inst {v[0]:9.3g} samp {v[1]:6.0f} wf {v[2]:9.3g} ideal {v[3]:9.3g} | {k[2]}")
