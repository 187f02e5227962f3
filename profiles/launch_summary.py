"""Per-kernel totals and shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`).
Usage: python profiles/launch_summary.py gpurun_out/r02_launches.csv > profiles/r02_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
tot = OrderedDict()
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void\s+", "", r[ik])
    name = re.sub(r"^b200::", "", name.split("(")[0])[:44]
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
total = sum(t for _, t in tot.values())
print("%-46s %5s %12s %10s %7s" % ("kernel", "n", "total us", "avg us", "share"))
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-46s %5d %12.1f %10.1f %6.1f%%" % (name, n, t, t / n, 100.0 * t / total))
print("%-46s %5d %12.1f" % ("total", sum(n for n, _ in tot.values()), total))
