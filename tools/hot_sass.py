"""Print the hot SASS of one kernel from an ncu report: python tools/hot_sass.py rep.ncu-rep <kernel regex> [launch index] [threshold]"""
import csv
import io
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:150])
hdr = rows[1]
ia, isrc, ie, ist = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Warp Stall Sampling (All Samples)"))
out = []
for r in rows[2:]:
    try:
        out.append((r[ia], r[isrc], int(r[ie]), r[ist]))
    except (ValueError, IndexError):
        pass
tot = sum(o[2] for o in out)
mx = max(o[2] for o in out)
print("total warp-instructions", tot, " max per instruction", mx, " static", len(out))
hot = [o for o in out if o[2] >= thr * mx]
print("hot instructions (>= %.0f%% of max): %d, covering %.1f%% of executed" % (thr * 100, len(hot), 100.0 * sum(o[2] for o in hot) / tot))
for a, s, n, st in hot:
    print("%s %9d %5s  %s" % (a[-5:], n, st, s[:100]))
