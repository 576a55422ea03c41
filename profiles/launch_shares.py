#!/usr/bin/env python
"""Per-kernel share of the engine's GPU time from an ncu launch list (profiles/capture.sh step 1):
   python profiles/launch_shares.py gpurun_out/<tag>_launches.csv > profiles/<tag>_launch_shares.txt
ncu serialises the launches and runs them cold, so only the SHARES are meaningful (they must agree with the CUDA-event phase times)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = d["Kernel Name"].split("(")[0]
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1000.0 if u in ("ns", "nsecond") else v * 1000.0 if u in ("ms", "msecond") else v
        tot[k] += v
        cnt[k] += 1
s = sum(tot.values()) or 1.0
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {s / 1000.0:.3f} ms summed kernel time (serialised, cold)")
print(f"{'kernel':42s} {'launches':>8s} {'sum us':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in tot.most_common():
    print(f"{k[:42]:42s} {cnt[k]:8d} {v:10.1f} {100 * v / s:6.1f}% {v / cnt[k]:9.1f}")
