import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdrpp_tetra_demodulator_b200 as pkg
from oracle import oracle as O
C_, N = 4, 6000
iq = O.generate(C_, N)
iq[0, :500] = 0.0
iq[1, :700] *= 1e-38
iq[2, 100:300] *= 8.0
iq[3, 1000:1010] = 0.0
for v in (2, 4):
    for n in (64, 400, 800, 1200, 6000):
        ob = O.OracleB(C_)
        cb, sb, db, _ = ob.process(np.ascontiguousarray(iq[:, :n]))
        dm = pkg.Demodulator(C_, n); dm.set_kernel_variant(v)
        r = dm.process(torch.from_numpy(np.ascontiguousarray(iq[:, :n])).cuda(), symbols=True, dibits=True)
        torch.cuda.synchronize()
        st = dm.get_state()
        for c in range(C_):
            bad = [f for f in O.EXACT_STATE_FIELDS if not np.array_equal(st[f][c], ob.states[f][c])]
            if bad:
                print("variant", v, "n", n, "ch", c, "bad", bad[:6], "agc", st["agc_gain"][c], ob.states["agc_gain"][c],
                      "fll", st["fll_phase"][c], ob.states["fll_phase"][c])
        dm.close()
print("done")
n = N
ob = O.OracleB(C_)
cb, sb, db, _ = ob.process(iq)
dm = pkg.Demodulator(C_, n); dm.set_kernel_variant(2)
r = dm.process(torch.from_numpy(iq).cuda(), symbols=True, dibits=True)
torch.cuda.synchronize()
dib = r.dibits.cpu().numpy(); sym = r.symbols.cpu().numpy(); cnt = r.counts.cpu().numpy()
print("counts", cnt, cb)
for c in range(C_):
    m = min(cnt[c], cb[c])
    d = np.flatnonzero(dib[c, :m] != db[c, :m])
    s = np.flatnonzero((sym[c, :m].view(np.uint32) != sb[c, :m].view(np.uint32)).any(axis=1))
    print("ch", c, "dibit diffs", len(d), d[:5], "sym diffs", len(s), s[:5])
    for i in list(d[:3]):
        print("    i", i, "gpu sym", sym[c, i - 1:i + 1].tolist(), sym[c, i-1:i+1].view(np.uint32).tolist(), "ora", sb[c, i - 1:i + 1].tolist(), sb[c,i-1:i+1].view(np.uint32).tolist(), "dib", dib[c, i], db[c, i])
