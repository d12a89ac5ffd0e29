"""Pipe-cost model of a warp-specialised kernel from an .ncu-rep (development aid).
Every SASS line's executed count (source page) is weighted by what one such instruction costs a scheduler that is also
streaming FFMA2 -- measured by tools/ubench/ubench_mix.cu (profiles/r02_ubench_mix.txt): FFMA2 2.0 cycles, scalar
FFMA/FMUL/FADD ~1.45, IMAD 2.0, FMNMX/FSEL/FSETP ~1.15, integer ALU ~0.9, loads/stores/MUFU/F2I/branches ~0 (own units;
they still need an issue cycle: counted 0.25).  Segments are cut at the tick barrier, so one segment = one role.
   python tools/sass_cost.py rep.ncu-rep [ticks]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))

def cost(op):
    o = op.split(".")[0]
    if o == "FFMA2": return 2.0
    if o in ("FFMA", "FMUL", "FADD", "FADD2", "FMUL2"): return 1.45 if not o.endswith("2") else 2.0
    if o in ("IMAD", "HFMA2"): return 2.0
    if o in ("FMNMX", "FMNMX3", "FSEL", "FSETP", "FCHK"): return 1.15
    if o in ("IADD3", "LOP3", "SHF", "LEA", "ISETP", "SEL", "MOV", "PRMT", "VIMNMX", "PLOP3", "IABS", "I2FP", "FRND", "VOTE", "R2P", "P2R", "CS2R", "S2R", "IADD", "VIADD", "LOP", "SGXT", "BMSK", "POPC", "FLO"): return 0.9
    return 0.25

ticks = int(sys.argv[2]) if len(sys.argv) > 2 else max(int(r["Instructions Executed"]) for r in rows if "BAR.SYNC" in r["Source"] or "BAR.ARV" in r["Source"])
bars = [i for i, r in enumerate(rows) if "BAR.SYNC" in r["Source"] and " 0x0" in r["Source"]]
print(f"{len(rows)} SASS lines, ticks (barrier executions, per warp) = {ticks}")
prev = 0
tot_all = 0.0
for b in bars + [len(rows) - 1]:
    seg = rows[prev:b + 1]
    ex = sum(int(r["Instructions Executed"]) for r in seg)
    if ex > ticks * 20:
        c = collections.Counter()
        w = collections.Counter()
        for r in seg:
            src = r["Source"].strip()
            toks = src.split()
            op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
            n = int(r["Instructions Executed"])
            c[op.split(".")[0]] += n
            w[op.split(".")[0]] += n * cost(op)
        total = sum(w.values())
        tot_all += total
        top = ", ".join(f"{k} {v / ticks:.0f}x={w[k] / ticks:.0f}" for k, v in sorted(c.items(), key=lambda kv: -w[kv[0]])[:9])
        print(f"[{prev:5d}:{b + 1:5d}] instr/tick {ex / ticks:7.1f}  modelled cycles/tick {total / ticks:7.1f}   {top}")
    prev = b + 1
print(f"all roles: {tot_all / ticks:.0f} modelled cycles per tick per CTA -> {tot_all / ticks / 4:.0f} per scheduler")
