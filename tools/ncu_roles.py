"""Per-role busy fractions of a warp-specialised kernel from an .ncu-rep: the SASS is cut at its BAR.SYNC
instructions (one tick loop per role), stall samples are summed per segment.
   python tools/ncu_roles.py rep.ncu-rep n_warps [--dump annotated.txt]"""
import csv, io, subprocess, sys
rep, nwarps = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and not c.endswith("(Not Issued)")]
tot = sum(int(r["# Samples"]) for r in rows)
per_warp = tot / nwarps
ticks = max(int(r["Instructions Executed"]) for r in rows if "BAR.SYNC" in r["Source"])
print(f"{len(rows)} SASS instructions, {tot} samples, {per_warp:.0f} per warp; barrier executions (max) {ticks}")
if "--dump" in sys.argv:
    with open(sys.argv[sys.argv.index("--dump") + 1], "w") as f:
        for i, r in enumerate(rows):
            st = sorted(((c, int(r[c] or 0)) for c in stall_cols), key=lambda kv: -kv[1])[:3]
            f.write(f"[{i:5d}] {int(r['# Samples']):6d} {int(r['Instructions Executed']):10d} {r['Source'].strip()[:80]:80s} " + " ".join(f"{c[6:]}={v}" for c, v in st if v) + "\n")
bars = [i for i, r in enumerate(rows) if "BAR.SYNC" in r["Source"]]
prev = 0
for b in bars + [len(rows) - 1]:
    seg = rows[prev:b + 3]
    n = sum(int(r["# Samples"]) for r in seg)
    if n > per_warp * 0.02:
        st = {c: sum(int(r[c] or 0) for r in seg) for c in stall_cols}
        ex = sum(int(r["Instructions Executed"]) for r in seg)
        nb = st.get("stall_barrier", 0)
        top = sorted(((k, v) for k, v in st.items() if k != "stall_barrier"), key=lambda kv: -kv[1])[:6]
        print(f"[{prev:5d}:{b + 3:5d}] {b + 3 - prev:4d} instr  warps~{n / per_warp:4.2f}  busy {100.0 * (n - nb) / per_warp:5.1f}% of one warp  "
              f"inst/tick {ex / max(ticks, 1):6.0f}  " + " ".join(f"{k[6:]}={100.0 * v / max(n - nb, 1):.0f}%" for k, v in top))
    prev = b + 3
