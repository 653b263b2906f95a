#!/usr/bin/env python
"""tcgen05 / TMA / multimem mnemonic counts per kernel of flex_dm_b200/libflexdm_mfp.so (cuobjdump -sass, sm_100a):
    python tools/sass_mnemonics.py > profiles/<round>_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys

WANT = ("UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "UTMALDG.2D.2CTA", "UTMALDG.3D.2CTA", "UCGABAR_ARV", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA", "UBLKCP", "MULTIMEM", "REDG", "LDGMC", "STG.E.128", "LDG.E.128")
lib = sys.argv[1] if len(sys.argv) > 1 else "flex_dm_b200/libflexdm_mfp.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = demangle(m.group(1))
        name = re.sub(r"\(.*", "", name)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for w in WANT:
        if re.search(r"\b" + re.escape(w) + r"\b", line) or (w in ("MULTIMEM",) and ".MULTIMEM" in line.upper()) or (w == "MULTIMEM" and "multimem" in line.lower()):
            counts[name][w] += 1
print("# SASS mnemonics per kernel of %s (cuobjdump -sass; sm_100a)" % lib)
print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops;")
print("# *.2CTA = the CTA-pair forms (tcgen05.mma.cta_group::2, pair TMA loads completing on the leader's barrier, multicast commit), UCGABAR = cluster barrier;")
print("# HMMA would be the legacy mma.sync path.  The GEMM epilogue stores with STG.E.128 (no UTMASTG); UTMAREDG = split-K reduce-add.\n")
for k, c in counts.items():
    if c:
        print("%-64s %s" % (k[:64], "  ".join("%s=%d" % (w, c[w]) for w in WANT if c[w])))
