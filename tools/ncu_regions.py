#!/usr/bin/env python
"""Executed warp instructions and stall samples of one captured kernel, summed over source-line ranges.
usage: python tools/ncu_regions.py report.ncu-rep file:lo-hi[:label] ...   (lines of other files: 'other')"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    regions = []
    for a in sys.argv[2:]:
        parts = a.split(":")
        lo, hi = parts[1].split("-")
        regions.append((parts[0], int(lo), int(hi), parts[2] if len(parts) > 2 else a))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur, hdr = None, None
    agg = {}
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r; i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed"); continue
        if hdr is None or not r[0].strip().isdigit():
            continue
        try:
            s, n, ln = int(r[i_s] or 0), int(r[i_e] or 0), int(r[0])
        except ValueError:
            continue
        key = cur
        for f, lo, hi, lab in regions:
            if cur == f and lo <= ln <= hi:
                key = lab; break
        a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += n
    ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
    print("total stall samples %d, warp instructions %d" % (ts, ti))
    for k, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%5.1f%% inst (%11d) %5.1f%% smp  %s" % (100.0 * n / ti, n, 100.0 * s / ts, k))

main()
