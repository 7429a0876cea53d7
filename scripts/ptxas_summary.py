#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs: kernel, registers, spill bytes, smem (ao_b200/lib/*.ptxas.log)."""
import glob, re, subprocess, sys
rows = []
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "ao_b200/lib/*.ptxas.log")):
    name = None
    spill = 0
    for line in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1); spill = 0
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
        m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", line)
        if m and name:
            rows.append((name, int(m.group(1)), spill, m.group(2) or "0"))
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for (n, regs, spill, smem), dn in zip(rows, names):
    dn = re.sub(r"\(.*", "", dn)
    print(f"{dn:60s} regs={regs:3d} stack/spill={spill} smem={smem}")
