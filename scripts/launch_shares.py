"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(cold-cache, serialised: compare SHARES, not absolutes).  Usage: launch_shares.py launches.csv[.gz] [out.md]"""
import csv
import gzip
import io
import sys

path = sys.argv[1]
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
lines = [l for l in raw.splitlines() if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
tot, per = 0.0, {}
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("aopt::", "")
    if "native::" in name or "at::" in name or "cub::" in name or "cuda::" in name:
        name = "[torch] " + name[:70]
    d = per.setdefault(name, [0.0, 0])
    d[0] += v
    d[1] += 1
    tot += v
out = [f"launches: {sum(d[1] for d in per.values())}, total {tot/1e3:.3f} ms (serialised, cold cache)", "",
       "| kernel | launches | total us | share |", "|---|---|---|---|"]
for name, (v, c) in sorted(per.items(), key=lambda kv: -kv[1][0])[:40]:
    out.append(f"| {name} | {c} | {v:.1f} | {100*v/tot:.1f}% |")
txt = "\n".join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
print(txt)
