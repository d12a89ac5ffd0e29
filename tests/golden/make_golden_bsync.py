"""Writes tests/golden/bsync_c6.npz: bit streams and what the REFERENCE's burst synchroniser
(phy/tetra_burst.c + phy/tetra_burst_sync.c + tetra_tdma.c compiled unmodified -> oracle/_ref/libtetra_bsync_ref.so)
delivered for them.  Run here (needs /root/reference):  python tests/golden/make_golden_bsync.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_bsync as B  # noqa: E402

B.build()
assert B.have_ref()
CALL_BITS, MAX_BURSTS = 432, 40
rows = [B.downlink_stream(100 + c, 30, ber=[0, 0, 2e-3, 0, 1e-2, 0][c], glitch_at=[(), (11,), (), (7, 19), (), ()][c])
        for c in range(5)]
rows.append(np.random.default_rng(7).integers(0, 2, 9000).astype(np.uint8))          # noise only: never locks
n_bits = np.array([len(r) for r in rows], dtype=np.int32)
bits = np.zeros((len(rows), int(n_bits.max())), dtype=np.uint8)
for c, r in enumerate(rows):
    bits[c, :len(r)] = r
R = B.RefBsync(len(rows))
nb, bursts, _ = R.feed(bits, n_bits, CALL_BITS, MAX_BURSTS)
bursts["reserved"] = 0
final = {f: np.array([R.state(c)[f] for c in range(len(rows))]) for f in
         ["state", "bits_in_buf", "bitbuf_start_bitnum", "next_frame_start_bitnum", "tn", "fn", "mn"]}
R.close()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bsync_c6.npz"), bits=np.packbits(bits, axis=1), n_bits=n_bits,
                    call_bits=CALL_BITS, n_bursts=nb, bursts=bursts, **{"final_" + k: v for k, v in final.items()})
print("bursts per channel:", nb, "final states:", final["state"])
