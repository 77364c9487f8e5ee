#!/usr/bin/env python
"""Opcode histogram of the Blackwell-specific instructions per kernel of libneurons_mm.so (cuobjdump -sass), the evidence that the
tensor-core kernels are tcgen05 / TMEM / TMA code and not recompiled mma.sync:
    UTCHMMA / UTCQMMA   tcgen05.mma (.2CTA = cta_group::2)        LDTM / STTM   tcgen05.ld / st (TMEM <-> registers)
    UTCBAR              tcgen05.commit -> mbarrier                 UTMALDG / UTMASTG / UTMAPF   TMA tensor load / store / prefetch
    UBLKCP              1-D bulk copy                              SYNCS         mbarrier operations
    HMMA                legacy mma.sync (the 8 x 8 x 40 temporal attention problems: far below tcgen05's 64-row minimum)
    FFMA2 / FMUL2 / FADD2  packed fp32x2 arithmetic (GEGLU / LayerNorm / GroupNorm epilogues)
Usage: python scripts/sass_summary.py [libneurons_mm.so] > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "neurons_b200", "libneurons_mm.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "FMUL2", "FADD2", "LDSM", "LDGSTS"]
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    kernels[cur]["_total"] += 1
    base = op.split(".")[0]
    if base in ("UTCHMMA", "UTCQMMA") and ".2CTA" in op:
        kernels[cur]["UTCHMMA.2CTA"] += 1
    if base in WATCH:
        kernels[cur][base] += 1


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True)
    return out.stdout.splitlines() if out.returncode == 0 else names


names = list(kernels)
pretty = demangle(names)
print(f"# {os.path.basename(lib)}: {len(names)} kernels, sm_100a SASS (cuobjdump -sass); counts are static instructions")
print(f"# columns: total | " + " ".join(WATCH))
tot = collections.Counter()
for n, p in zip(names, pretty):
    c = kernels[n]
    cut = p.rfind(">(")
    short = (p[:cut + 1] if cut >= 0 else p.split("(")[0]).replace("void ", "").replace("nmm::", "").replace("(int)", "").replace("(bool)", "")
    if not any(c[w] for w in WATCH):
        continue
    print(f"{short[:72]:72s} {c['_total']:6d} | " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
    tot.update({w: c[w] for w in WATCH})
print("# library totals: " + " ".join(f"{w}={tot[w]}" for w in WATCH))
