"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, launches and total device time of the LAST
call in the file (the first call is warm-up).   python tools/launch_summary.py in.csv [ncalls] > profiles/xxx.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
ncalls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
rows = rows[len(rows) - len(rows) // ncalls:]
tot = sum(v for _, v in rows)
print("# %s: last of %d calls, %d launches, %.1f us of kernel time (ncu: serialised, cold caches -- compare SHARES, not absolutes)" % (path, ncalls, len(rows), tot / 1e3))
agg = collections.OrderedDict()
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
for k, (c, v) in agg.items():
    print("%-78s x%-3d %10.1f us  %5.1f %%" % (k[:78], c, v / 1e3, 100 * v / tot))
