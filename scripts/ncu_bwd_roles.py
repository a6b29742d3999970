"""Backward-sweep kernel: warp-sampling counts per barrier-delimited phase, split by how many warps execute each instruction
(chain warp: once per step and CTA; all basis warps; one basis warp).  usage: python scripts/ncu_bwd_roles.py src.csv"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr, data = rows[hi], rows[hi + 1:]
ia, isamp, iex, ib = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_barrier")
v = lambda r, i=isamp: int(r[i]) if r[i].isdigit() else 0
tot = sum(v(r) for r in data)
print("samples", tot, "barrier-stall", sum(v(r, ib) for r in data), "instructions", len(data))
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[ia]]
prev = 0
for b in bars + [len(data) - 1]:
    seg = data[prev:b + 1]
    c, n = Counter(), Counter()
    for r in seg:
        c[r[iex]] += v(r); n[r[iex]] += 1
    wait = sum(v(r, ib) for r in seg)
    print("%5d-%5d samples %6d (barrier wait %6d)  by exec count:" % (prev, b, sum(c.values()), wait),
          ", ".join("%s: %d smp / %d ins" % (k, s, n[k]) for k, s in sorted(c.items(), key=lambda x: -x[1])[:6] if s > 20))
    prev = b + 1
