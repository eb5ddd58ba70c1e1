#!/usr/bin/env python
"""Per-region instruction / stall-sample breakdown of one kernel from `ncu --page source --csv`.
usage: python tools/ncu_hot.py report.ncu-rep [marker-regex ...]
Regions are delimited by SASS lines matching the markers (default: USETMAXREG = warp-role boundaries)."""
import csv, io, re, subprocess, sys

def main():
    rep = sys.argv[1]
    markers = sys.argv[2:] or ["USETMAXREG"]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    regions, cur = [], {"name": "prologue", "inst": 0, "samples": 0, "ops": {}, "top": []}
    for r in rows[hdr_i + 1:]:
        if len(r) <= iex:
            continue
        src = r[isrc]
        if any(re.search(m, src) for m in markers):
            regions.append(cur)
            cur = {"name": src.strip()[:50], "inst": 0, "samples": 0, "ops": {}, "top": []}
        n = int(r[iex] or 0); s = int(r[ismp] or 0)
        cur["inst"] += n; cur["samples"] += s
        op = re.sub(r"^@!?U?P\d+\s+", "", src.strip()).split(" ")[0].split(".")[0]
        cur["ops"][op] = cur["ops"].get(op, 0) + n
        cur["top"].append((s, n, r[ia], src.strip()[:90]))
    regions.append(cur)
    tot = sum(x["inst"] for x in regions) or 1
    tots = sum(x["samples"] for x in regions) or 1
    for x in regions:
        print("== region after [%s]: %.1f%% of instructions (%d), %.1f%% of stall samples" % (x["name"], 100.0 * x["inst"] / tot, x["inst"], 100.0 * x["samples"] / tots))
        ops = sorted(x["ops"].items(), key=lambda kv: -kv[1])[:14]
        print("   ops: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, x["inst"])) for k, v in ops))
        for s, n, a, src in sorted(x["top"], reverse=True)[:8]:
            print("   samples %6d  exec %9d  %s  %s" % (s, n, a, src))

if __name__ == "__main__":
    main()
