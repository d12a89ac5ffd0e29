"""Ad-hoc GPU probe (development aid): smoke parity + per-variant kernel timing."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdrpp_tetra_demodulator_b200 as pkg
import __graft_entry__ as g

def timeit(C_, N, variant, reps=3, symbols=False):
    iq, _ = pkg.synth_capture(C_, N)
    torch.cuda.synchronize()
    dm = pkg.Demodulator(C_, N)
    dm.set_kernel_variant(variant)
    dm.use_torch_stream()
    out = None
    best = 1e9
    for r in range(reps):
        out = dm.process(iq, symbols=symbols, dibits=True, out=out)
        torch.cuda.synchronize()
        best = min(best, dm.last_kernel_ms())
    cnt = out.counts.cpu().numpy()
    m = dm.metrics()
    dm.close()
    return best, cnt, m

if __name__ == "__main__":
    t = time.time(); g.smoke(); print("smoke time", time.time() - t)
    print(torch.cuda.get_device_name(0))
    res = []
    for (C_, N) in [(4096, 32768), (512, 65536), (256, 65536), (32768, 8192)]:
        for v in [2, 4, 5]:
            ms, cnt, m = timeit(C_, N, v)
            gs = C_ * N / ms / 1e6
            print(f"C={C_} N={N} variant={v}: {ms:.2f} ms  {gs:.2f} Gsamples/s  {gs*8/6492.4*100:.2f}% HBM  mean syms {cnt.mean():.1f} sync {m['sync'].mean():.2f}")
            res.append(dict(C=C_, N=N, variant=v, ms=ms, gsps=gs))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/probe.json", "w"))
