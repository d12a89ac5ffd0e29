import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdrpp_tetra_demodulator_b200 as pkg
from oracle import oracle as O
C_ = 4
iqfull = O.generate(C_, 4096)
for N in [1, 2, 3, 4, 5, 8, 9, 16, 63, 64, 65, 128, 129, 1000, 4096]:
    iq = np.ascontiguousarray(iqfull[:, :N])
    ob = O.OracleB(C_)
    cb, sb, db, _ = ob.process(iq)
    dm = pkg.Demodulator(C_, N)
    r = dm.process(torch.from_numpy(iq).cuda(), symbols=True, dibits=True)
    torch.cuda.synchronize()
    st = dm.get_state()
    bad = [f for f in O.EXACT_STATE_FIELDS if not np.array_equal(st[f], ob.states[f])]
    cnt = r.counts.cpu().numpy()
    sym = r.symbols.cpu().numpy()
    print("N", N, "counts", cnt, cb, "bad fields", bad)
    if bad:
        for f in bad[:6]:
            a, b = st[f], ob.states[f]
            if a.ndim > 1:
                idx = np.argwhere(a != b)[:3]
                print("   ", f, "first diffs", [(tuple(i), a[tuple(i)], b[tuple(i)]) for i in idx])
            else:
                print("   ", f, a, b)
        for c in range(C_):
            n = min(cnt[c], cb[c])
            d = np.flatnonzero((sym[c, :n].view(np.uint32) != sb[c, :n].view(np.uint32)).any(axis=1))
            print("    ch", c, "first sym diff at", d[:3], "of", n)
        break
    dm.close()
