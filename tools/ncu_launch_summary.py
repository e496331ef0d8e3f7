"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and share per kernel.
Usage: python tools/ncu_launch_summary.py launches.csv"""
import collections
import csv
import sys

tot = collections.Counter()
cnt = collections.Counter()
with open(sys.argv[1]) as f:
    rows = [r for r in csv.reader(f) if len(r) > 10]
hdr = rows[0]
k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[1:]:
    name = r[k].split("(")[0]
    ns = float(r[v].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[u], 1.0)
    tot[name] += ns
    cnt[name] += 1
allns = sum(tot.values())
print("%-62s %9s %12s %7s %10s" % ("kernel", "launches", "total_us", "share", "avg_us"))
for name, ns in tot.most_common(24):
    print("%-62s %9d %12.1f %6.1f%% %10.2f" % (name[:62], cnt[name], ns / 1e3, 100 * ns / allns, ns / 1e3 / cnt[name]))
