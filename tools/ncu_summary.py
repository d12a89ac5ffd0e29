"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed):
   python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 25]  -> prints key metrics + hottest SASS instructions"""
import csv, io, subprocess, sys, json

def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_fmaheavy.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALLS = "smsp__average_warps_issue_stalled_"

def source_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))

if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    m = raw_metrics(rep)
    summary = {}
    for k in KEYS:
        if k in m:
            print(f"{k:75s} {m[k][0]:>18s} {m[k][1]}")
            summary[k] = m[k][0]
    print("-- stall reasons (warps per issue-active cycle)")
    for k, (v, u) in sorted(m.items()):
        if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio"):
            try:
                if float(v) > 0.01:
                    print(f"   {k[len(STALLS):-len('_per_issue_active.ratio')]:32s} {float(v):.3f}")
                    summary["stall_" + k[len(STALLS):-len('_per_issue_active.ratio')]] = float(v)
            except ValueError:
                pass
    rows = source_rows(rep)
    tot = sum(int(r["# Samples"]) for r in rows)
    print(f"-- SASS instructions: {len(rows)}, stall samples {tot}")
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"]))[:top]
    for i in sorted(idx):
        r = rows[i]
        print(f"   [{i:5d}] {int(r['# Samples'])*100.0/tot:5.2f}%  exec {r['Instructions Executed']:>10s}  {r['Source'].strip()[:90]}")
    if "--json" in sys.argv:
        json.dump(summary, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
