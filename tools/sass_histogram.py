#!/usr/bin/env python3
"""Opcode histogram of the SASS in libflamed_b200.so, per kernel family: the mnemonics that prove the Blackwell-native
paths (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, FFMA2 = packed fp32) next to the
legacy ones (HMMA = mma.sync, LDGSTS = cp.async).  usage: python tools/sass_histogram.py > profiles/<round>/sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flamed_tts_b200", "libflamed_b200.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "FFMA2", "FADD2", "FMUL2",
         "HMMA", "LDSM", "LDGSTS", "MUFU", "FFMA", "REDUX", "CCTL", "ERRBAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fam = collections.defaultdict(collections.Counter)
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"^void ", "", d)
        cur = re.sub(r"\(.*$", "", d)[:70]
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        fam[cur][base] += 1
        if op.startswith("UTCHMMA.2CTA") or ".2CTA" in op:
            fam[cur][base + ".2CTA"] += 1
tot = collections.Counter()
for k in fam:
    tot.update(fam[k])
print("libflamed_b200.so: %d kernels, %d SASS instructions" % (len(fam), sum(tot.values())))
print("whole library:", ", ".join("%s x%d" % (w, tot[w]) for w in WATCH if tot[w]))
print()
print("%-72s %s" % ("kernel", "watched opcodes"))
for k in sorted(fam, key=lambda k: -sum(fam[k].values())):
    c = fam[k]
    if any(c[w] for w in WATCH[:9] + ["HMMA", "FFMA2"]):
        print("%-72s %s" % (k, ", ".join("%s x%d" % (w, c[w]) for w in WATCH if c[w])))
