"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, mean µs, share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg.setdefault(r[ki].split("(")[0], []).append(v)
tot = sum(sum(v) for v in agg.values())
print("kernel,launches,avg_us,share_pct")
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print(f"{k},{len(v)},{sum(v) / len(v):.1f},{100 * sum(v) / tot:.1f}")
