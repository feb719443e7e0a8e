#!/usr/bin/env python
"""Markdown summary + DRAM traffic of every kernel in .ncu-rep files (ncu --set full captures).
usage: tools/ncu_summary.py out.md traffic.json tag=report.ncu-rep[:workload] ..."""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def rows_of(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    return h, u, rows[2:]


def main():
    out_md, out_json = sys.argv[1], sys.argv[2]
    traffic, md = {}, []
    for spec in sys.argv[3:]:
        tag, rest = spec.split("=", 1)
        rep, _, wl = rest.partition(":")
        h, u, rows = rows_of(rep)
        idx = {n: i for i, n in enumerate(h)}
        md.append(f"\n## {tag} (`{rep.split('/')[-1]}`)\n")
        for r in rows:
            name = r[idx["Kernel Name"]].split("(")[0]
            md.append(f"\n### {name}\n")
            for k in KEYS:
                if k in idx:
                    md.append(f"- `{k}` = {r[idx[k]]} {u[idx[k]]}")
            try:
                rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT[u[idx["dram__bytes_read.sum"]]]
                wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT[u[idx["dram__bytes_write.sum"]]]
                if wl:
                    traffic.setdefault(wl, {})[name] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr,
                                                       "source": rep.split("/")[-1] + " (ncu --set full, one launch)"}
            except (KeyError, ValueError):
                pass
    open(out_md, "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
