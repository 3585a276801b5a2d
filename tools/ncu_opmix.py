"""Aggregate an `ncu --page source --csv` dump: executed warp instructions per SASS opcode and the busiest SASS
ranges with their stall samples.  usage: ncu_opmix.py dump.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ithr = hdr.index("Avg. Threads Executed")
ops, smp = Counter(), Counter()
tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= iex:
        continue
    src = r[isrc].strip()
    n = int(r[iex] or 0)
    s = int(r[ismp] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += n
    smp[op] += s
    tot += n
    lines.append((n, s, src, r[ithr]))
print("total warp instructions", tot, " samples", sum(smp.values()))
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {n:12d} {100.0 * n / tot:6.2f}%   samples {smp[op]:7d}")
