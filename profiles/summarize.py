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
    rows = [r for r in csv.reader(io.StringIO(out))]
    # a report may hold several kernels: one "Kernel Name" row, one header row ("Address", ...), then the SASS lines
    res, cur = {}, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = res.setdefault(r[1], {"by": {}, "st": {}, "tot": 0})
            continue
        if cur is None or len(r) < 7 or r[0] == "Address":
            continue
        t = r[1].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        cur["by"][op] = cur["by"].get(op, 0) + int(r[5])
        cur["st"][op] = cur["st"].get(op, 0) + int(r[2])
        cur["tot"] += int(r[5])
    return res


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
            mixes = opcode_mix(rep)
            key = next((k for k in mixes if d.get("Kernel Name", "").split("(")[0].replace("void ", "") in k.replace("(int)", "")), None) or next(iter(mixes), None)
            m = mixes.get(key, {"by": {}, "st": {}, "tot": 0})
            tot, mix, stalls = m["tot"] or 1, sorted(m["by"].items(), key=lambda x: -x[1])[:14], m["st"]
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
