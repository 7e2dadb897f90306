"""Summarise an `ncu --page source --csv` dump: executed warp instructions and stall samples per SASS opcode and per
region of the SASS listing (hot blocks).  Usage: python tools/ncu_src.py dump.csv [queries_or_units]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
by_op = collections.Counter()
smp_op = collections.Counter()
tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= iex or not r[iex]:
        continue
    try:
        ex, smp = int(r[iex]), int(r[ismp] or 0)
    except ValueError:
        continue
    src = r[isrc].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    by_op[op] += ex
    smp_op[op] += smp
    tot += ex
    lines.append((r[ia], ex, smp, src))
print(f"total warp instructions {tot}  per unit {tot / units:.1f}")
tsmp = sum(smp_op.values()) or 1
for op, ex in by_op.most_common(25):
    print(f"  {op:12s} {ex:12d}  {ex / units:8.1f}/unit  {100 * ex / tot:5.1f}%   samples {100 * smp_op[op] / tsmp:5.1f}%")
if len(sys.argv) > 3:
    for a, ex, smp, src in lines:
        print(f"{a} {ex:10d} {smp:6d}  {src[:100]}")
