"""Print selected metrics of an `ncu --page raw --csv` export, one block per launch.
Usage: python scripts/ncu_metrics.py raw.csv substr [substr ...]   (metric-name substrings)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[2:]
for r in data:
    print("==", r[hdr.index("Kernel Name")][:90])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            print(f"   {h:90s} {r[i]:>16s} {units[i]}")
