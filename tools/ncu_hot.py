#!/usr/bin/env python
"""Top SASS lines of one kernel from an `ncu --set full --import-source on` report:
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


s, e, src = ci["# Samples"], ci["Instructions Executed"], ci["Source"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_e, tot_s = sum(num(r[e]) for r in data), sum(num(r[s]) for r in data)
print(f"instructions executed {tot_e:.0f}, samples {tot_s:.0f}, SASS lines {len(data)}")
agg = {h: sum(num(r[ci[h]]) for r in data) for h in stalls}
print("stall totals:", ", ".join(f"{h[6:]}={v:.0f}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0))
print(f"{'samples':>8} {'executed':>10}  line  top-stall  SASS")
for r in sorted(data, key=lambda r: -num(r[s]))[:top_n]:
    st = max(stalls, key=lambda h: num(r[ci[h]]))
    print(f"{num(r[s]):8.0f} {num(r[e]):10.0f}  {data.index(r):4d}  {st[6:]:<10} {r[src].strip()[:100]}")
