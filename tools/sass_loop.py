#!/usr/bin/env python
"""Static opcode count of the HOT PATH through the marching loop of a k_fused_step instantiation.

The kernel's main loop is the backward branch with the largest span that holds a BAR.SYNC.  Inside it, the fast-path
guards skip their slow paths with forward branches; this walker follows the path a healthy run takes: a forward
conditional branch is taken when the region it jumps over contains a CALL (the IEEE slow path of a division) or is a
whole re-computation of a phase (longer than `--cold` instructions), otherwise it falls through.  The count is per
loop trip (= one grid row of one warp); ncu's dynamic count per OUTPUT warp-row is ~6 % higher (halo lanes, re-traced
rows, prologue).  A development aid to judge a change before spending GPU time — ncu stays the evidence.

usage: python tools/sass_loop.py [lib.so] [--kernel SUBSTR] [--cold N] [--dump]
"""
import argparse
import collections
import re
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("lib", nargs="?", default="euler2d_kokkos_b200/libeuler2d_b200.so")
ap.add_argument("--kernel", default="k_fused_stepILi2ELb1ELb0ELi0E")
ap.add_argument("--cold", type=int, default=80)
ap.add_argument("--dump", action="store_true")
args = ap.parse_args()

txt = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))

FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")
for fn, ins in funcs.items():
    if args.kernel not in fn:
        continue
    addr_index = {a: k for k, (a, _) in enumerate(ins)}

    def opcode(text):
        t = re.sub(r"^@!?U?P\d+\s+", "", text)
        return t.split()[0].split(".")[0], t

    # the main loop: backward branch with the largest span that contains a barrier
    best = None
    for k, (a, text) in enumerate(ins):
        op, t = opcode(text)
        if op == "BRA":
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                tgt = int(m.group(1), 16)
                body = ins[addr_index[tgt]:k + 1]
                if any("BAR.SYNC" in b for _, b in body) and (best is None or len(body) > best[2]):
                    best = (addr_index[tgt], k, len(body))
    if best is None:
        print(fn, ": no loop with a barrier found")
        continue
    lo, hi, _ = best
    cnt = collections.Counter()
    k = lo
    path = []
    skipped = 0
    while k <= hi:
        a, text = ins[k]
        op, t = opcode(text)
        path.append((a, text))
        cnt[op] += 1
        if op == "BRA" and k != hi:
            m = re.search(r"0x([0-9a-f]+)", t)
            tgt = int(m.group(1), 16) if m else None
            if tgt is not None and tgt > a and tgt in addr_index and addr_index[tgt] <= hi + 1:
                region = ins[k + 1:addr_index[tgt]]
                cond = text.startswith("@")
                cold = any("CALL" in r for _, r in region) or len(region) > args.cold
                if not cond or cold:
                    skipped += len(region)
                    k = addr_index[tgt]
                    continue
        k += 1
    total = sum(cnt.values())
    fp64 = sum(cnt[o] for o in FP64)
    short = re.sub(r"^_ZN3e2d\d+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+", "", fn)
    print(f"{short}\n  loop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {hi - lo + 1} static instructions, hot path {total} "
          f"(skipped {skipped} cold), FP64-pipe {fp64}, others {total - fp64}")
    print("  " + "  ".join(f"{o} {n}" for o, n in cnt.most_common()))
    if args.dump:
        for a, text in path:
            print(f"    {a:05x}  {text}")
