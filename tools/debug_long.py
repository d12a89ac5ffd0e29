import sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import sdrpp_tetra_demodulator_b200 as pkg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
iq, tx = pkg.synth_capture(1, N, device=0, want_tx=True)
dm = pkg.Demodulator(S, 1024); dm.use_torch_stream()
W = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
d, info = dm.process_long(iq[0], warmup=W)
torch.cuda.synchronize()
print(info)
n = info["n_dibits"]; skip = n // 4; m = n - 64
best = None
for lag in range(12, 26):
    e = (d[lag + skip:lag + m] != tx[0, skip:m])
    c = int(e.sum())
    if best is None or c < best[0]: best = (c, lag, e)
c, lag, e = best
pos = torch.nonzero(e).flatten().cpu().numpy() + skip
L2 = info["segment_samples"] // 2
print("errors", c, "lag", lag)
seg = pos // L2
import collections
cnt = collections.Counter(seg.tolist())
print("segments with errors:", len(cnt), list(cnt.items())[:20])
for s_, k in list(cnt.items())[:6]:
    p = pos[seg == s_]
    print("seg", s_, "n", k, "offsets in segment (symbols):", (p - s_ * L2)[:8], "...", (p - s_ * L2)[-3:])
# sequential check on one bad segment region: run the plain chain over [start-200k, end] and compare
if len(cnt):
    s_ = list(cnt.keys())[0]
    lo = max(0, int(s_ * L2 * 2 - 400_000)); hi = min(N, int((s_ + 1) * L2 * 2 + 100_000))
    with pkg.Demodulator(1, hi - lo) as one:
        r = one.process(iq[:, lo:hi].contiguous(), dibits=True)
        torch.cuda.synchronize()
        k = int(r.counts[0]); dd = r.dibits[0, :k]
        t = tx[0, lo // 2: lo // 2 + k]
        bestc = min((int((dd[l + 150_000:k - 10] != t[150_000:k - 10 - l]).sum()), l) for l in range(12, 26))
        print("sequential chain over the same region: errors after 150k symbols:", bestc)
