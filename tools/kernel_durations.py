"""Per-kernel durations from an ncu launch list (--metrics gpu__time_duration.sum --csv), grouped by name and grid:
python tools/kernel_durations.py file.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
d = defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    d[(r[ik][:48], r[ig] if ig is not None else "")].append(v / 1e3)
for k, v in d.items():
    v2 = sorted(v)
    print("%-50s grid %-18s n=%3d median %8.1f us  min %8.1f  max %8.1f" % (k[0], k[1], len(v), v2[len(v2) // 2], v2[0], v2[-1]))
