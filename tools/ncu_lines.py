#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, out = None, []
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] not in ("Function Name", "Line No") and len(r) > 17 and r[2] == "-" and r[0].isdigit():
        try:
            out.append((int(r[7]), int(r[6]), int(r[17]) if r[17].isdigit() else 0, cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
ti, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"warp instructions {ti}, stall samples {ts}")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0] / ti * 100:5.1f}% inst {o[1] / ts * 100:5.1f}% samples  smem wavefronts {o[2] / 1e6:7.1f} M  {o[3]}:{o[4]}  {o[5]}")
