#!/usr/bin/env python
"""Dynamic opcode mix + stall-sample summary of one kernel from an ncu report's source page.

usage: python tools/ncu_opmix.py report.ncu-rep [kernel-index]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[which]
h = b["hdr"]
iS, iE, iSamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
ops, samples, stalls = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in b["rows"]:
    toks = r[iS].split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    if not toks:
        continue
    op = toks[0].split(".")[0]
    if op == "MUFU":
        op = toks[0]
    n = int(r[iE] or 0)
    ops[op] += n
    samples[op] += int(r[iSamp] or 0)
    tot += n
    for k in stall_cols:
        stalls[k] += int(r[h.index(k)] or 0)
print(b["name"])
print(f"warp instructions executed: {tot}")
fp64 = sum(ops[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
print(f"FP64-pipe: {fp64} ({100.0 * fp64 / tot:.1f} %)")
ts = sum(samples.values())
for op, n in ops.most_common(40):
    print(f"  {op:14s} {n:12d} {100.0 * n / tot:6.2f} %   samples {100.0 * samples[op] / max(ts, 1):6.2f} %")
print("stall samples:")
t = sum(stalls.values())
for k, n in stalls.most_common(12):
    print(f"  {k:28s} {100.0 * n / max(t, 1):6.2f} %")
