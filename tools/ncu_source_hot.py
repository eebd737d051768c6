"""Stall samples of one ncu report (--import-source on) aggregated per CUDA source line: python tools/ncu_source_hot.py X.ncu-rep [N]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, agg = None, {}
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        cur = row[1].split("/")[-1]
        continue
    if row[0] in ("Function Name", "Line No") or row[0] == "":
        continue
    try:
        ln, samples, inst = int(row[0]), int(row[6]), int(row[7])
    except ValueError:
        continue
    a = agg.setdefault((cur, ln), [0, 0, row[1][:110]])
    a[0] += samples
    a[1] += inst
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{100 * v[0] / max(tot, 1):5.1f}%  inst {v[1]:7d}  {k[0]}:{k[1]:<4d} {v[2]}")
