#!/usr/bin/env python
"""Stall samples / executed instructions per CUDA source line of one captured kernel
(`ncu --import-source on`, kernels built with -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr = None, None
    items = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_smp, i_ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or r[0] in ("Function Name",) or not r[0].strip().isdigit():
            continue
        try:
            items.append((int(r[i_smp] or 0), int(r[i_ex] or 0), cur_file, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
    tot_s = sum(x[0] for x in items) or 1
    tot_i = sum(x[1] for x in items) or 1
    print("total stall samples %d, warp instructions %d" % (tot_s, tot_i))
    for s, n, f, ln, src in sorted(items, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * n / tot_i, f, ln, src))


if __name__ == "__main__":
    main()
