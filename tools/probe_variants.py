"""Ad-hoc GPU probe (development aid): kernel time per variant at a few shapes.
   python tools/probe_variants.py "5,6,7,8,9" "4096x262144,256x262144"  """
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdrpp_tetra_demodulator_b200 as pkg

def timeit(C_, N, variant, reps=3):
    iq, _ = pkg.synth_capture(C_, N)
    torch.cuda.synchronize()
    dm = pkg.Demodulator(C_, N)
    dm.set_kernel_variant(variant)
    dm.use_torch_stream()
    out, best = None, 1e9
    for r in range(reps):
        out = dm.process(iq, dibits=True, out=out)
        torch.cuda.synchronize()
        best = min(best, dm.last_kernel_ms())
    cnt = out.counts.cpu().numpy()
    m = dm.metrics()
    dm.close()
    return best, cnt, m

if __name__ == "__main__":
    variants = [int(v) for v in sys.argv[1].split(",")]
    shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[2].split(",")]
    res = []
    for (C_, N) in shapes:
        for v in variants:
            ms, cnt, m = timeit(C_, N, v)
            gs = C_ * N / ms / 1e6
            cyc = ms * 1e-3 * 1.965e9 / (N / 8)
            print(f"C={C_} N={N} variant={v}: {ms:.2f} ms  {gs:.2f} Gsamples/s  {cyc:.0f} cycles/tick  mean syms {cnt.mean():.1f} sync {m['sync'].mean():.2f}", flush=True)
            res.append(dict(C=C_, N=N, variant=v, ms=ms, gsps=gs))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/probe_variants.json", "w"))
