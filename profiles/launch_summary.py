"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`).
usage: python profiles/launch_summary.py X.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[d["Metric Unit"]]
    a = agg.setdefault(d["Kernel Name"][:64], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-66s %4d launches %9.3f ms %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
print("total %.3f ms" % tot)
