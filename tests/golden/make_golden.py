"""Generate the golden fixtures from the REFERENCE ITSELF (Oracle A = the reference's own
src/dsp/*.cpp built by `make -C oracle ref`).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

The reference ships no tests or vectors, so these fixtures are what pins parity:
inputs come from the deterministic generator (oracle/siggen.c), expected outputs from Oracle A.
Each fixture stores a SHA-256 of the exact input bytes so a drifting libm is caught, not absorbed.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ILL_MARGIN = 0.1


def pack2(d):
    """dibits (values 0..3) -> 4 per byte, first in the top bits."""
    n = (len(d) + 3) // 4 * 4
    p = np.zeros(n, np.uint8)
    p[:len(d)] = d
    p = p.reshape(-1, 4)
    return (p[:, 0] << 6 | p[:, 1] << 4 | p[:, 2] << 2 | p[:, 3]).astype(np.uint8)


def make_case(name, n_channels, n_samples, snr_db, first_channel=0, store_input=False, store_syms=False, re_only=False):
    sp = O.default_sg_params(snr_db=snr_db)
    iq = O.generate(n_channels, n_samples, sp, first_channel=first_channel)
    a = O.OracleA(n_channels, fastamp_re_only=re_only)
    counts, syms, dibits, _ = a.process(iq, want_syms=True)
    out = {
        "fastamp_re_only": int(re_only),
        "n_channels": n_channels, "n_samples": n_samples, "snr_db": snr_db, "first_channel": first_channel,
        "input_sha256": np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8),
        "counts": counts,
        "ref_standarderr": np.array([a.loop_state(c).standarderr for c in range(n_channels)], np.float32),
        "ref_sync": np.array([a.loop_state(c).sync for c in range(n_channels)], np.int32),
    }
    lock = np.zeros(n_channels, np.int64)
    for c in range(n_channels):
        n = int(counts[c])
        out[f"dibits_packed_{c}"] = pack2(dibits[c, :n])
        # Decisions the reference itself takes within ILL_MARGIN of a quadrant boundary: any change in
        # float operation order (a different VOLK kernel, FMA contraction) can flip these, most of them
        # before the loops have locked.  Stored so the parity tests can tell "differs where the reference
        # is ill-conditioned" from "differs".
        mag = np.hypot(syms[c, :n, 0], syms[c, :n, 1]) + 1e-30
        ill = np.minimum(np.abs(syms[c, :n, 0]), np.abs(syms[c, :n, 1])) < ILL_MARGIN * mag
        out[f"ill_packed_{c}"] = np.packbits(ill)
        # first symbol index from which the reference's output equals the transmitted dibits to the end
        tx = O.tx_dibits(first_channel + c, n + 64, seed_data=sp.seed_data)
        best = None
        for lag in range(10, 30):
            e = np.flatnonzero(dibits[c, lag:n] != tx[:n - lag])
            last = int(e[-1]) + lag + 1 if len(e) else lag
            if best is None or last < best:
                best = last
        lock[c] = best
    out["lock_index"] = lock
    out["ill_margin"] = ILL_MARGIN
    if store_input:
        out["iq"] = iq
    if store_syms:
        out["ref_syms"] = np.stack([syms[c, :counts.min()] for c in range(n_channels)])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    a.close()
    print(name, "counts", counts, "sync", out["ref_sync"], "lock", lock)


def make_design():
    a = O.OracleA(1)
    nt, rrc, lbe, hbe, bank, P, T = a.taps()
    np.savez_compressed(os.path.join(HERE, "design_default.npz"), rrc=rrc, lbe=lbe, hbe=hbe, bank=bank,
                        coeffs=a.coeffs())
    a.close()


if __name__ == "__main__":
    assert O.have_ref(), "build Oracle A first: make -C oracle ref"
    make_design()
    # BASELINE.json configs[0]: 1 channel x 1e6 samples, the golden capture
    make_case("cfg1_c1_n1e6_snr30", 1, 1_000_000, 30.0)
    # a batch at the survey's impairment spread, and one at the README's 20 dB lock threshold
    make_case("batch_c8_n60000_snr30", 8, 60_000, 30.0)
    make_case("batch_c4_n60000_snr20", 4, 60_000, 20.0, first_channel=100)
    # a small case that carries its own input and the reference's symbols
    make_case("small_c2_n4096", 2, 4096, 30.0, store_input=True, store_syms=True)
    # the same captures through the reference built with the OTHER reading of complex_t::fastAmplitude()
    # (oracle/_ref/libtetra_ref_reonly.so; tdm_config.flags & TDM_CFG_FASTAMP_RE_ONLY on the product side)
    make_case("cfg1_c1_n1e6_snr30_reonly", 1, 1_000_000, 30.0, re_only=True)
    make_case("batch_c8_n60000_snr30_reonly", 8, 60_000, 30.0, re_only=True)
    make_case("batch_c4_n60000_snr20_reonly", 4, 60_000, 20.0, first_channel=100, re_only=True)
