#!/usr/bin/env python
"""Per-kernel SASS opcode statistics of libeuler2d_b200.so (static counts; FP64-pipe = DADD/DMUL/DFMA/DSETP/DMNMX).

usage: python tools/sass_stats.py [path/to/lib.so] [kernel-name-substring]
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "euler2d_kokkos_b200/libeuler2d_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
stats = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        full = op + m.group(2)
        stats[cur][op] += 1
        if op in ("MUFU", "IMAD"):
            stats[cur][full] += 1
for fn, c in stats.items():
    if pat not in fn:
        continue
    short = re.sub(r"^_ZN3e2d\d+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+", "", fn)
    fp64 = sum(c[k] for k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
    total = sum(v for k, v in c.items() if "." not in k)
    print(f"{short[:60]:60s} total={total:5d} fp64={fp64:4d} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DSETP {c['DSETP']})"
          f" MUFU.RCP64H={c['MUFU.RCP64H']} MUFU.RSQ64H={c['MUFU.RSQ64H']} CALL={c['CALL']} IMAD.MOV={c['IMAD.MOV.U32']+c['IMAD.MOV']}"
          f" LDG={c['LDG']} STG={c['STG']} LDS={c['LDS']} STS={c['STS']} BAR={c['BAR']}")
