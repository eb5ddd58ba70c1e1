#!/usr/bin/env python
"""Stall-reason totals per warp-role region (regions split at USETMAXREG) from `ncu --page source --csv`.
usage: python tools/ncu_stalls.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
isrc = h.index("Source")
reg, acc = 0, {}
for r in rows[hi + 1:]:
    if "USETMAXREG" in r[isrc]:
        reg += 1
    a = acc.setdefault(reg, [0] * len(cols))
    for j, i in enumerate(cols):
        a[j] += int(r[i] or 0)
for reg, a in acc.items():
    tot = sum(a) or 1
    print("region %d: %d samples | " % (reg, tot) + ", ".join("%s %.0f%%" % (h[i][6:], 100.0 * v / tot) for i, v in sorted(zip(cols, a), key=lambda t: -t[1])[:8]))
