"""Group a kernel's SASS by execution count (loop structure) from an ncu report.
python tools/sass_sections.py rep.ncu-rep <kernel regex> [launch skip] [--list MINCOUNT]"""
import csv
import io
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "0"
listing = int(sys.argv[sys.argv.index("--list") + 1]) if "--list" in sys.argv else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:140])
hdr = rows[1]
ia, isrc, ie, ist = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Warp Stall Sampling (All Samples)"))
seen, out = set(), []
for r in rows[2:]:
    try:
        if r[ia] in seen:
            continue
        seen.add(r[ia])
        out.append((r[ia][-4:], r[isrc], int(r[ie]), int(r[ist] or 0)))
    except (ValueError, IndexError):
        pass
tot = sum(o[2] for o in out)
stall = sum(o[3] for o in out)
print("total warp-instructions %d, static %d, stall samples %d" % (tot, len(out), stall))
cur = start = None
cnt = acc = st = 0
for a, s, n, w in out:
    if cur is None or abs(n - cur) > 0.05 * max(cur, 1):
        if cur is not None and acc > 0.005 * tot:
            print("  %s..  x%-9d %4d instr -> %10d (%5.1f%%)  stalls %5.1f%%" % (start, cur, cnt, acc, 100 * acc / tot, 100 * st / max(stall, 1)))
        cur, start, cnt, acc, st = n, a, 0, 0, 0
    cnt += 1; acc += n; st += w
if cur is not None:
    print("  %s..  x%-9d %4d instr -> %10d (%5.1f%%)  stalls %5.1f%%" % (start, cur, cnt, acc, 100 * acc / tot, 100 * st / max(stall, 1)))
if listing is not None:
    for a, s, n, w in out:
        if n >= listing:
            print("%s %9d %5d  %s" % (a, n, w, s[:100]))
