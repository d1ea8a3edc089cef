"""Split one kernel of an ncu report into the stretches between barriers / calls: instructions and stall samples per stretch.
python tools/phase_split.py rep.ncu-rep <kernel regex> [launch-skip]"""
import csv
import io
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:120])
hdr = rows[1]
isrc, ie, ist = (hdr.index(k) for k in ("Source", "Instructions Executed", "Warp Stall Sampling (All Samples)"))
body = rows[2:]
half = len(body) // 2
if half and body[0][isrc] == body[half][isrc]:
    body = body[:half]
tot = sum(int(r[ie]) for r in body)
ts = sum(int(r[ist]) for r in body)
print("static", len(body), "executed", tot, "samples", ts)
seg = acc = sacc = start = 0
for n, r in enumerate(body):
    acc += int(r[ie])
    sacc += int(r[ist])
    if any(k in r[isrc] for k in ("BAR.SYNC", "CALL", "RET")) or n == len(body) - 1:
        print("seg %2d [%5d-%5d] instr %9d (%4.1f%%) samples %5d (%4.1f%%)  %s" % (seg, start, n, acc, 100 * acc / tot, sacc,
                                                                                  100 * sacc / ts, r[isrc].strip()[:50]))
        seg += 1
        acc = sacc = 0
        start = n + 1
