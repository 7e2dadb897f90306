"""Per-source-line executed warp instructions / stall samples from an ncu report (needs -lineinfo + --import-source).
Usage: python tools/ncu_lines.py report.ncu-rep [units] [launch_index] [min_fraction]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
launch = sys.argv[3] if len(sys.argv) > 3 else "0"
minf = float(sys.argv[4]) if len(sys.argv) > 4 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", launch,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
fname, hdr = None, None
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No" and "Instructions Executed" in r:
        hdr = r
        iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) > iex and r[0]:
        try:
            ex, smp = int(r[iex]), int(r[ismp] or 0)
        except ValueError:
            continue
        a = agg.setdefault((fname, r[0], r[1][:110]), [0, 0])
        a[0] += ex
        a[1] += smp
tot = sum(a[0] for a in agg.values())
tsmp = sum(a[1] for a in agg.values()) or 1
print(f"total {tot}  per unit {tot / units:.1f}")
for (f, l, s), (ex, smp) in agg.items():
    if ex > tot * minf or smp > tsmp * minf * 2:
        print(f"{f[:12]:12s} {l:>4s} {ex / units:9.1f} {100 * smp / tsmp:5.1f}%  {s}")
