"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass).  usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tinyda_b200", "libtinyda_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "MUFU", "FFMA2", "HMMA", "SHFL", "USETMAXREG"]
rows, cur, k = [], None, 0
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = dict(name=names[k], n=0, **{c: 0 for c in cols}); k += 1; rows.append(cur); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur["n"] += 1
        op = m.group(1)
        for c in cols:
            if op.startswith(c):
                cur[c] += 1
print("SASS summary of tinyda_b200/libtinyda_b200.so (cuobjdump -sass; round 2, final tree). Architectures in the fatbin: %s" % ", ".join(arch))
print("Mnemonics: UTCHMMA = tcgen05.mma (kind::f16 / tf32), LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (1-D TMA), UTCBAR = tcgen05.commit,")
print("SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, FFMA2 = fma.rn.f32x2, SHFL = warp shuffles (the warp-per-chain kernels).  No HMMA (mma.sync / wmma)")
print("anywhere; no UTMALDG: operands are pre-packed on the host into the canonical K-major layout, so plain 1-D bulk copies replace tensor-map TMA.\n")
print("%-78s %7s " % ("kernel", "instrs") + " ".join("%7s" % c for c in cols))
for r in sorted(rows, key=lambda r: -r["n"]):
    nm = r["name"].replace("void ", "")
    cut = nm.find(">(")
    nm = nm[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", nm)
    nm = nm.replace("(int)", "").replace("(bool)", "")
    nm = re.sub(r"\(anonymous namespace\)::", "", nm)
    print("%-78s %7d " % (nm[:78], r["n"]) + " ".join("%7d" % r[c] for c in cols))
