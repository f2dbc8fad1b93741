#!/usr/bin/env python
"""Warp-stall samples of one kernel of an ncu report aggregated per SOURCE line.

ncu's CSV export of the source page carries per-SASS-instruction samples but no line numbers; `nvdisasm -g` of the
same cubin carries line info per SASS offset.  This joins the two (innermost inlined location).

usage: python tools/ncu_hot_lines.py report.ncu-rep lib.so kernel-mangled-name-substring [top N]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, pat = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "euler2d_kokkos_b200", "csrc")

off2line = {}
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    for cubin in glob.glob(os.path.join(d, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
        start = next((i for i, l in enumerate(txt) if l.startswith("//--------------------- .text.") and pat in l), None)
        if start is None:
            continue
        cur = None
        for l in txt[start + 1:]:
            if l.startswith("//--------------------- "):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+\S", l)
            if m and cur:
                off2line[int(m.group(1), 16)] = cur
        break
if not off2line:
    sys.exit("kernel not found in " + lib)

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, data = rows[1], rows[2:]
iA, iN, iE = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][iA], 16)
agg, tot = collections.defaultdict(lambda: [0, 0, 0]), 0
for r in data:
    if len(r) <= iE or not r[iA].startswith("0x"):
        continue
    key = off2line.get(int(r[iA], 16) - base, ("?", 0))
    n, e = int(r[iN] or 0), int(r[iE] or 0)
    agg[key][0] += n
    agg[key][1] += e
    agg[key][2] += 1
    tot += n
src = {}
print(f"# {rows[0][1]}")
print("# warp-stall samples per source line (innermost inlined location); share of all samples, warp instructions executed (1e6), SASS instructions")
for (f, ln), (n, e, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open(os.path.join(src_dir, f)).read().split("\n")
        except OSError:
            src[f] = []
    code = src[f][ln - 1].strip() if 0 < ln <= len(src[f]) else ""
    print(f"{100 * n / max(tot, 1):6.2f}%  inst {e / 1e6:8.1f}  sass {c:3d}  {f}:{ln:<4d} {code[:105]}")
