#!/usr/bin/env python
"""SASS opcode histogram per kernel of libseer_b200.so: the mnemonics that prove which hardware path a kernel uses
(/opt/skills/guides/B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG /
UTMASTG / UTMAPF = TMA load / store / prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync,
MUFU.EX2, FFMA2 / FADD2 = packed fp32x2, LDGSTS = cp.async, STL / LDL = register spills.

    python tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "seervideoldm_b200", "libseer_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.TANH",
        "FFMA2", "FADD2", "FMUL2", "LDGSTS", "STL", "LDL", "ATOMG", "RED"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True, check=True).stdout
regs = {}
fn = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+)", ln)
    if m and fn:
        regs[fn] = (int(m.group(1)), int(m.group(2)))
demangle = lambda names: dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))
hist = collections.OrderedDict()
total = collections.Counter()
cur = None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("MUFU.EX2", "MUFU.TANH") and op.startswith(k)):
                hist[cur][k] += 1
names = demangle(list(hist))
print(f"SASS opcode histogram per kernel, {os.path.basename(lib)} (cuobjdump -sass; sm_100a).  Columns with no hits are omitted per row.")
for fn, h in sorted(hist.items(), key=lambda kv: names[kv[0]]):
    short = re.sub(r"\(.*", "", names[fn]).replace("void seer::", "").replace("seer::", "")
    r = regs.get(fn, ("?", "?"))
    cells = " ".join(f"{k}={h[k]}" for k in KEYS if h[k])
    print(f"{short:62s} instr={total[fn]:6d} regs={r[0]:>3} stack={r[1]:>4}  {cells}")
