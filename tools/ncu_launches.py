"""Prints the per-launch durations (us) of an `ncu --metrics gpu__time_duration.sum --csv` log; --last N: only the last
N launches; --sum: totals per kernel name."""
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
        rows.append((row["Kernel Name"], v))
if "--last" in sys.argv:
    rows = rows[-int(sys.argv[sys.argv.index("--last") + 1]):]
if "--sum" in sys.argv:
    agg = {}
    for k, v in rows:
        a = agg.setdefault(k.split("(")[0][:60], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:9.1f} us  {100 * v / tot:5.1f} %  x{n:<3d} {k}")
    print(f"{tot:9.1f} us  total")
else:
    for k, v in rows:
        print(f"{v:8.1f} us  {k[:80]}")
