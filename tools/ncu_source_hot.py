"""Hot instructions and opcode mix of one kernel from `ncu -i rep --page source --csv --print-source sass` output:
python tools/ncu_source_hot.py source.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
S, E, SRC = idx["# Samples"], idx["Instructions Executed"], idx["Source"]
data = [r for r in rows[hi + 1:] if len(r) > E and r[S].isdigit()]
tot = sum(int(r[S]) for r in data)
print(rows[0][1][:100] if len(rows[0]) > 1 else "")
print("total samples", tot, "SASS lines", len(data))
for r in sorted(data, key=lambda r: -int(r[S]))[:top_n]:
    print("%6s %5.1f%%  exec %9s  %s" % (r[S], 100 * int(r[S]) / tot, r[E], r[SRC].strip()[:100]))
c, s = Counter(), Counter()
for r in data:
    parts = r[SRC].strip().split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    c[op] += int(r[E])
    s[op] += int(r[S])
T = sum(c.values())
print("total warp instructions", T)
for op, n in c.most_common(30):
    print("%-10s %11d %5.1f%%  samples %5.1f%%" % (op, n, 100 * n / T, 100 * s[op] / tot))
