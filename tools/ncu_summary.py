"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram bytes]) per kernel name.

usage: python tools/ncu_summary.py launches.csv [skip_launches] > profiles/rNN_ncu_step_launch_summary.txt
`skip_launches`: drop this many leading launches (the warm-up step) so the table covers one steady-state step."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        k = r["ID"]
        d = per.setdefault(k, {"name": r["Kernel Name"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = v * scale
    ids = sorted(per, key=int)[skip:]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for k in ids:
        d = per[k]
        a = agg[d["name"][:66]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.1f ms serialised (compare shares, not absolutes)" % (len(ids), tot))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s n=%4d %9.3f ms %5.1f%%  dram rd %8.2f GB wr %8.2f GB" % (name, a[0], a[1], 100 * a[1] / tot, a[2] / 1e9, a[3] / 1e9))


if __name__ == "__main__":
    main()
