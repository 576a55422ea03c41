#!/usr/bin/env python
"""Extracts the judged numbers from ncu reports (gpurun_out/*.ncu-rep) into small CSV/markdown files under profiles/.
Usage: python profiles/summarize.py <tag> <report.ncu-rep> [<report2> ...]   ->  profiles/<tag>_<kernel>.csv"""
import csv
import io
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def opcode_mix(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))[2:]
    tot = sum(int(r[5]) for r in rows) or 1
    by, st = {}, {}
    for r in rows:
        t = r[1].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        by[op] = by.get(op, 0) + int(r[5])
        st[op] = st.get(op, 0) + int(r[2])
    return tot, sorted(by.items(), key=lambda x: -x[1])[:14], st


def main():
    tag = sys.argv[1]
    here = os.path.dirname(os.path.abspath(__file__))
    for rep in sys.argv[2:]:
        hdr, units, vals = raw(rep)
        for v in vals:
            d = dict(zip(hdr, v))
            u = dict(zip(hdr, units))
            kname = d.get("Kernel Name", "kernel").split("(")[0].replace("void ", "").replace("i3d::", "").replace("<", "_").replace(">", "")
            path = os.path.join(here, f"{tag}_{kname}.csv")
            tot, mix, stalls = opcode_mix(rep)
            with open(path, "w") as f:
                f.write(f"# ncu --set full --clock-control none; report {os.path.basename(rep)}; kernel {d.get('Kernel Name')}\n")
                f.write("metric,unit,value\n")
                for k in KEYS:
                    if k in d:
                        f.write(f"{k},{u.get(k, '')},{d[k]}\n")
                f.write(f"# SASS opcode mix (warp instructions executed, total {tot}); stall samples in brackets\n")
                for op, c in mix:
                    f.write(f"opcode_{op},pct,{100.0 * c / tot:.2f} [{stalls.get(op, 0)}]\n")
            print("wrote", path)


if __name__ == "__main__":
    main()
