#!/usr/bin/env python
"""profiles/traffic.json from the ncu summaries: dram__bytes_read.sum + dram__bytes_write.sum per launch of the two roofline kernels.
   python profiles/make_traffic.py profiles/<tag>_k_eg_apply_0.csv profiles/<tag>_k_eg_rows_0.csv"""
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram(path):
    rd = wr = None
    for line in open(path):
        f = line.strip().split(",")
        if f[0] == "dram__bytes_read.sum":
            rd = float(f[2]) * UNIT[f[1]]
        if f[0] == "dram__bytes_write.sum":
            wr = float(f[2]) * UNIT[f[1]]
    return rd, wr


out = {}
for key, path in (("k_eg_apply", sys.argv[1]), ("k_eg_build", sys.argv[2])):
    rd, wr = dram(path)
    out[key] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                "source": f"profiles/{os.path.basename(path)} (ncu --set full, C3, one launch; dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
