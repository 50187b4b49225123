"""Aggregate the `ncu --page source --csv` dump of one kernel: stall samples by opcode + top lines."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[h]; iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iE = hdr.index("Instructions Executed")
def num(x):
    try: return int(float(x.replace(",", "")))
    except Exception: return 0
data = [(num(r[iS]), num(r[iE]), n, r[iSrc]) for n, r in enumerate(rows[h + 1:]) if len(r) > iS]
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "lines", len(data), "instr executed", sum(d[1] for d in data))
c, ce = Counter(), Counter()
for smp, ex, n, src in data:
    t = src.split()
    if not t: continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    c[op] += smp; ce[op] += ex
for op, v in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    print(f"{op:26s} {100*v/tot:5.1f}%  exec={ce[op]}")
print("top lines")
for smp, ex, n, src in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    print(n, smp, ex, src[:110])
