"""Split the warp-sampling counts of one kernel (ncu --page source --csv) at its barrier instructions.

usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python scripts/ncu_segments.py src.csv
Each printed line is one barrier-delimited stretch of SASS: its share of all samples, which barrier ends it
and how often that barrier executed.  Used to see which phase of a multi-phase kernel the time goes to.
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr, data = rows[hi], rows[hi + 1:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
val = lambda r: int(r[isamp]) if len(r) > isamp and r[isamp].isdigit() else 0
tot = sum(val(r) for r in data)
print("samples", tot, "instructions", len(data))
start, acc = 0, 0
for i, r in enumerate(data):
    acc += val(r)
    if any(k in r[ia] for k in ("UCGABAR_WAIT", "BAR.SYNC", "WARPSYNC.COLLECTIVE")) or i == len(data) - 1:
        if acc > tot * 0.005:
            print("%5d-%5d %7d (%4.1f%%) ends with %s execs %s" % (start, i, acc, 100.0 * acc / tot, r[ia].strip()[:40], r[iex]))
        start, acc = i + 1, 0
