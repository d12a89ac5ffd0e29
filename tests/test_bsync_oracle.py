"""CPU tests of the burst-sync checkers: the restatement (oracle/oracle_bsync.c) against the reference's own
phy/tetra_burst.c + phy/tetra_burst_sync.c + tetra_tdma.c compiled unmodified (oracle/_ref/libtetra_bsync_ref.so),
against the committed golden fixture, and against the constants the reference tabulates."""
import os

import numpy as np
import pytest

from oracle import oracle_bsync as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "bsync_c6.npz")

needs_ref = pytest.mark.skipif(not B.have_ref(), reason="oracle/_ref/libtetra_bsync_ref.so not built (needs /root/reference)")


@pytest.fixture(scope="module", autouse=True)
def _built():
    B.build()


def _planted_buffer(rng):
    n = int(rng.integers(1, 700))
    buf = rng.integers(0, 2, n).astype(np.uint8)
    for _ in range(int(rng.integers(0, 3))):
        s = B.SEQ[str(rng.choice(list("npqxy")))]
        p = int(rng.integers(0, n))
        m = min(len(s), n - p)
        buf[p:p + m] = s[:m]
    return buf


@needs_ref
def test_find_train_seq_matches_reference():
    """incl. the reference's look-ahead filter quirk: sequences starting in the first 21 positions are normally missed"""
    rng = np.random.default_rng(1)
    P, R = B.PortBsync(1), B.RefBsync(1)
    early = 0
    for _ in range(4000):
        buf = _planted_buffer(rng)
        if rng.random() < 0.3:                               # plant one in the blind zone on purpose
            s = B.SEQ[str(rng.choice(list("npqxy")))]
            p = int(rng.integers(0, 21))
            if p + len(s) <= len(buf):
                buf[p:p + len(s)] = s
                early += 1
        mask = int(rng.integers(1, 32))
        end = int(rng.integers(0, len(buf) + 1))
        a, b = P.find_train_seq(buf, end, mask), R.find_train_seq(buf, end, mask)
        assert a[0] == b[0] and (a[0] < 0 or a[1] == b[1]), (a, b, len(buf), end, mask)
    assert early > 500
    R.close()


@needs_ref
def test_blind_zone_is_real():
    """KAT for the quirk: y at offset 5 of an otherwise zero buffer is not found, at offset 21 it is"""
    R, P = B.RefBsync(1), B.PortBsync(1)
    for off, expect in [(5, -1), (20, -1), (21, B.TRAIN_SYNC), (214, B.TRAIN_SYNC)]:
        buf = np.zeros(600, dtype=np.uint8)
        buf[off:off + 38] = B.SEQ["y"]
        for O_ in (R, P):
            rc, o = O_.find_train_seq(buf, 600, 1 << B.TRAIN_SYNC)
            assert rc == expect and (rc < 0 or o == off), (off, rc, o)
    R.close()


def _compare_bursts(nb, bu, nr, br):
    assert np.array_equal(nb, nr), (nb, nr)
    for c in range(len(nb)):
        for i in range(min(nb[c], bu.shape[1])):
            for f in ["bitnum", "train_seq", "tn", "fn", "mn", "call_index", "bits"]:
                assert np.array_equal(bu[c, i][f], br[c, i][f]), (f, c, i, bu[c, i][f], br[c, i][f])


@needs_ref
@pytest.mark.parametrize("call_bits", [1, 7, 100, 432, 509, 510])
def test_state_machine_matches_reference(call_bits):
    total = 0
    for seed in range(6):
        bits = B.downlink_stream(seed, 40, ber=[0, 0, 1e-3, 5e-3, 0, 2e-2][seed], glitch_at=(13, 27) if seed % 2 else ())
        P, R = B.PortBsync(1), B.RefBsync(1)
        h = len(bits) // 2 + seed                         # two feeds: state carried across tdm_bsync_in calls
        for part in (bits[:h], bits[h:]):
            nb, bu = P.feed(part[None, :], len(part), call_bits, 64)
            nr, br, bl = R.feed(part[None, :], len(part), call_bits, 64)
            _compare_bursts(nb, bu, nr, br)
            for i in range(nb[0]):
                dm = P.demux(bu[0, i])                    # tetra_burst_rx_cb's split vs what tp_sap_udata_ind received
                assert len(dm) == br[0, i]["reserved"][0]
                for j in range(len(dm)):
                    for f in ["type", "blk_num", "n_bits", "bits"]:
                        assert np.array_equal(dm[j][f], bl[0, i, j][f]), (f, j)
            rs = R.state(0)
            for f in rs:
                assert np.array_equal(P.states[0][f], rs[f]), (f, call_bits, seed, P.states[0][f], rs[f])
            total += int(nb[0])
        R.close()
    assert total > 100                                # the streams really lock and deliver


@needs_ref
def test_state_machine_random_feeds_match_reference():
    """randomised: feeds of random length with a random call length each, streams with bit errors, inserted and
    deleted stretches (slips) -- the restatement must track the reference through every re-acquisition"""
    rng = np.random.default_rng(77)
    for trial in range(12):
        bits = B.downlink_stream(500 + trial, 60, ber=float(rng.choice([0, 1e-3, 1e-2])), glitch_at=tuple(rng.integers(5, 55, 3)))
        for _ in range(int(rng.integers(0, 3))):                       # delete a stretch: the slot grid jumps backwards
            p = int(rng.integers(2000, len(bits) - 2000))
            bits = np.concatenate([bits[:p], bits[p + int(rng.integers(1, 400)):]])
        P, R = B.PortBsync(1), B.RefBsync(1)
        pos, total = 0, 0
        while pos < len(bits):
            n = int(rng.integers(1, 4000))
            part = bits[pos:pos + n]
            pos += len(part)
            cb = int(rng.integers(1, 511))
            nb, bu = P.feed(part[None, :], len(part), cb, 16)
            nr, br, _ = R.feed(part[None, :], len(part), cb, 16)
            _compare_bursts(nb, bu, nr, br)
            rs = R.state(0)
            for f in rs:
                assert np.array_equal(P.states[0][f], rs[f]), (f, trial, pos, cb)
            total += int(nb[0])
        R.close()
        assert total > 10


@needs_ref
def test_noise_and_constant_input_match_reference():
    rng = np.random.default_rng(5)
    for kind in range(3):
        bits = [rng.integers(0, 2, 30000), np.zeros(30000), np.ones(30000)][kind].astype(np.uint8)
        P, R = B.PortBsync(1), B.RefBsync(1)
        nb, bu = P.feed(bits[None, :], len(bits), 333, 128)
        nr, br, _ = R.feed(bits[None, :], len(bits), 333, 128)
        _compare_bursts(nb, bu, nr, br)
        rs = R.state(0)
        for f in rs:
            assert np.array_equal(P.states[0][f], rs[f]), (f, kind)
        R.close()


@needs_ref
def test_own_burst_builder_matches_reference_layout():
    """the test-stream generator vs the reference's build_sync_c_d_burst / build_norm_c_d_burst
    (phy/tetra_burst.c:171-269) everywhere except the four phase-adjustment bits (the reference indexes its
    phase2bits table with a possibly negative value there, tetra_burst.c:163)"""
    import ctypes as C
    L = B.ref_lib()
    rng = np.random.default_rng(9)
    mine = B.sync_burst(rng)
    sb, bb, bkn = mine[94:214].copy(), mine[252:282].copy(), mine[282:498].copy()
    buf = np.zeros(512, dtype=np.uint8)
    assert L.rbs_build_sync_burst(buf.ctypes.data_as(C.c_void_p), sb.ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p),
                                  bkn.ctypes.data_as(C.c_void_p)) == 510
    keep = np.ones(510, bool)
    keep[[12, 13, 498, 499]] = False
    assert np.array_equal(buf[:510][keep], mine[keep])
    for two in (0, 1):
        mine = B.norm_burst(rng, bool(two))
        bkn1, bkn2 = mine[14:230].copy(), mine[282:498].copy()
        bb = np.concatenate([mine[230:244], mine[266:282]])
        buf[:] = 0
        assert L.rbs_build_norm_burst(buf.ctypes.data_as(C.c_void_p), bkn1.ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p),
                                      bkn2.ctypes.data_as(C.c_void_p), two) == 510
        assert np.array_equal(buf[:510][keep], mine[keep])


def test_training_sequences_are_the_reference_constants():
    """KAT: the offsets the state machine insists on (tetra_burst_sync.c:123,133) are where the builders put the
    sequences, and the sequences have the lengths the reference declares"""
    rng = np.random.default_rng(2)
    assert np.array_equal(B.sync_burst(rng)[214:252], B.SEQ["y"])
    assert np.array_equal(B.norm_burst(rng, False)[244:266], B.SEQ["n"])
    assert np.array_equal(B.norm_burst(rng, True)[244:266], B.SEQ["p"])
    assert {k: len(v) for k, v in B.SEQ.items()} == {"n": 22, "p": 22, "q": 22, "N": 33, "P": 33, "x": 30, "X": 45, "y": 38}


def test_ts_detector_restatement():
    """src/main.cpp:385-414 against a direct numpy reading of it (this function cannot be compiled from the reference)"""
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, 9000).astype(np.uint8)
    for name, pos in [("n", 100), ("X", 2500), ("y", 2548 + 2047), ("P", 8990)]:
        s = B.SEQ[name]
        m = min(len(s), len(bits) - pos)
        bits[pos:pos + m] = s[:m]
    # direct model
    w = np.zeros(45, np.uint8)
    found, expire, trace = 0, 0, []
    for b in bits:
        w[:-1] = w[1:]
        w[-1] = b
        if any(np.array_equal(w[:len(s)], s) for s in B.SEQ.values()):
            found, expire = 1, 2048
        if expire > 0:
            expire -= 1
            if expire == 0:
                found = 0
        trace.append((found, expire))
    for split in (9000, 1, 144, 4500):
        P = B.PortBsync(1)
        pos = 0
        while pos < len(bits):
            part = bits[pos:pos + split]
            P.feed(part[None, :], len(part), 510, 8, detect_ts=True)
            pos += len(part)
            assert (int(P.states[0]["ts_found"]), int(P.states[0]["ts_expire"])) == trace[pos - 1], (split, pos)


def test_golden_fixture():
    """bursts the REFERENCE delivered for committed bit streams (tests/golden/make_golden_bsync.py)"""
    g = np.load(GOLDEN)
    n_bits = g["n_bits"]
    bits = np.ascontiguousarray(np.unpackbits(g["bits"], axis=1)[:, :int(n_bits.max())])
    P = B.PortBsync(bits.shape[0])
    nb, bu = P.feed(bits, n_bits, int(g["call_bits"]), g["bursts"].shape[1])
    assert np.array_equal(nb, g["n_bursts"])
    for c in range(bits.shape[0]):
        for i in range(nb[c]):
            for f in ["bitnum", "train_seq", "tn", "fn", "mn", "call_index", "bits"]:
                assert np.array_equal(bu[c, i][f], g["bursts"][c, i][f]), (f, c, i)
    for f in ["state", "bits_in_buf", "bitbuf_start_bitnum", "next_frame_start_bitnum", "tn", "fn", "mn"]:
        assert np.array_equal(P.states[f], g["final_" + f]), f
