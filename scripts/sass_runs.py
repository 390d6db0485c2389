"""Aggregate an `ncu --page source --csv` dump into runs of SASS instructions with (nearly) the same execution count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Instructions Executed" in r)
data = []
for r in rows[rows.index(hdr) + 1:]:
    if len(r) != len(hdr) or not r[hdr.index("Instructions Executed")].isdigit():
        break   # the next launch of the report starts here
    data.append(r)
ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data); ts = max(1, sum(int(r[smp]) for r in data))
print("total warp instructions", tot, "samples", ts, "SASS lines", len(data))
runs, cur = [], None
for i, r in enumerate(data):
    n = int(r[ie])
    if cur and abs(n - cur[2]) <= 0.03 * max(n, cur[2]):
        cur[1] = i; cur[3] += n; cur[4] += int(r[smp])
    else:
        cur = [i, i, n, n, int(r[smp])]; runs.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
for a, b, n, s, sm in runs:
    if s > thr * tot:
        print(f"{a}-{b} len {b - a + 1} exec {n} share {100 * s / tot:.1f}% samples {100 * sm / ts:.1f}%  {data[a][src].strip()[:70]}")
