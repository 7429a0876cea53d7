"""Summarise an `ncu --page raw --csv` export: one line per launch with duration, DRAM bytes,
DRAM/L2/L1 throughput %, occupancy, registers.  Usage: python scripts/ncu_summary.py raw.csv [out.md]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name, default=float("nan")):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default


def scaled(r, name):
    v, u = num(r, name), units[col[name]] if name in col else ""
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
            "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return v * mult


lines = ["| # | kernel | grid | regs | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | L2 % | L1 % | L1 hit % | L2 hit % | occ % |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for i, r in enumerate(data):
    t = scaled(r, "gpu__time_duration.sum")
    rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
    name = r[col["Kernel Name"]]
    name = name.split("(")[0].replace("aopt::", "").replace("void ", "")[:60]
    lines.append(f"| {i} | {name} | {num(r,'launch__grid_size'):.0f} | {num(r,'launch__registers_per_thread'):.0f} | {t*1e6:.1f} | "
                 f"{rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/t/1e9:.0f} | {num(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{num(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {num(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{num(r,'l1tex__t_sector_hit_rate.pct'):.1f} | {num(r,'lts__t_sector_hit_rate.pct'):.1f} | "
                 f"{num(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} |")
out = "\n".join(lines)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
print(out)
