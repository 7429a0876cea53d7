"""Hottest SASS instructions (by warp-stall samples) of an `ncu --page source --csv` export.
Usage: ncu_hot_sass.py src_kernel.csv[.gz] [N]"""
import csv
import gzip
import io
import sys

path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read().splitlines()
hdr_i = [i for i, l in enumerate(raw) if l.startswith('"Address"')][0]
rows = [r for r in csv.DictReader(io.StringIO("\n".join(raw[hdr_i:]))) if (r.get("# Samples") or "").isdigit()]
tot = sum(int(r["# Samples"]) for r in rows)
print(raw[0][:160])
print("total samples", tot, "instructions", len(rows))
stall_cols = [k for k in rows[0] if k.startswith("stall_")]
agg = {k: sum(int(r[k] or 0) for r in rows) for k in stall_cols}
print("stall mix:", ", ".join(f"{k[6:]}={100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:n]:
    st = sorted(((k[6:], int(r[k] or 0)) for k in stall_cols if (r[k] or "0") != "0"), key=lambda kv: -kv[1])[:3]
    print(f"{int(r['# Samples']):6d} {100*int(r['# Samples'])/max(tot,1):5.1f}%  {r['Source'].strip()[:64]:64s} {st}")
