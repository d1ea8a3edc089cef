"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name:  python tools/launch_summary.py file.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(r[ik][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print("%-72s n=%4d %10.1f us %5.1f%%  (%.1f us each)" % (k, c, v / 1e3, 100 * v / tot, v / 1e3 / c))
print("total %.1f us over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
