"""Per-region stall breakdown of one kernel in an .ncu-rep (regions = SASS index ranges, e.g. one role each).
   python tools/ncu_regions.py rep.ncu-rep [start:end ...]   (no ranges: split at BAR.SYNC / big gaps, print top lines)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and not c.endswith("(Not Issued)")]
tot = sum(int(r["# Samples"]) for r in rows)
print(f"{len(rows)} SASS instructions, {tot} samples")
ranges = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:] if ":" in a]
if not ranges:
    # buckets of 100 instructions
    ranges = [(i, min(i + 100, len(rows))) for i in range(0, len(rows), 100)]
for (a, b) in ranges:
    sub = rows[a:b]
    n = sum(int(r["# Samples"]) for r in sub)
    ex = sum(int(r["Instructions Executed"]) for r in sub)
    st = {c: sum(int(r[c] or 0) for r in sub) for c in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{a:5d}:{b:5d}] samples {n:8d} ({100.0*n/tot:5.1f}%)  inst_exec {ex:12d}  " + "  ".join(f"{k[6:]}={100.0*v/max(n,1):.0f}%" for k, v in top if v))
if "--top" in sys.argv:
    k = int(sys.argv[sys.argv.index("--top") + 1])
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"]))[:k]
    for i in sorted(idx):
        r = rows[i]
        st = sorted(((c, int(r[c] or 0)) for c in stall_cols), key=lambda kv: -kv[1])[:3]
        print(f"   [{i:5d}] {int(r['# Samples'])*100.0/tot:5.2f}% exec {r['Instructions Executed']:>10s} {r['Source'].strip()[:70]:70s} " + " ".join(f"{c[6:]}={v}" for c, v in st if v))
