#!/usr/bin/env python
"""Static view of a warp-specialised kernel's SASS: cut the code at the roles' tick loops (a backward branch that
follows a `BAR.SYNC 0x0`) and print, per loop, the instruction mix.  Usage:
    cuobjdump -sass -fun <mangled kernel> lib.so > k.sass;  python tools/sass_roles.py k.sass
"""
import collections
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2idx = {a: i for i, (a, _) in enumerate(ins)}
bars = [i for i, (_, t) in enumerate(ins) if ("BAR.SYNC" in t and " 0x0" in t) or "SYNCS.ARRIVE" in t]
loops = []
for b in bars:
    for i in range(b + 1, min(b + 16, len(ins))):
        t = ins[i][1]
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < ins[b][0]:
            loops.append((addr2idx[int(m.group(1), 16)], i))
            break
print(f"{len(ins)} instructions, {len(bars)} tick barriers, {len(loops)} tick loops")
for lo, hi in loops:
    ops = collections.Counter()
    for _, t in ins[lo:hi + 1]:
        t = re.sub(r"^@!?U?P\d\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    n = hi - lo + 1
    top = ", ".join(f"{k} {v}" for k, v in ops.most_common(9))
    print(f"[{ins[lo][0]:#06x}..{ins[hi][0]:#06x}] {n:5d} instr  FFMA2 {ops['FFMA2']:4d} ({100.0 * ops['FFMA2'] / n:4.1f}%)  {top}")
