"""TEST INFRASTRUCTURE -- ctypes loaders for the burst-sync checkers and a protocol-valid bit-stream generator.

Only tests/, __graft_entry__.smoke() and the CPU-baseline legs of the bench scripts may import this module.

  RefBsync   oracle/_ref/libtetra_bsync_ref.so   the reference's own phy/tetra_burst.c, phy/tetra_burst_sync.c,
                                                 tetra_tdma.c compiled unmodified (authority)
  PortBsync  oracle/_build/liboracle_bsync.so    restatement (oracle_bsync.c), travels without the reference
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libtetra_bsync_ref.so")
PORT_SO = os.path.join(HERE, "_build", "liboracle_bsync.so")

BITS_PER_TS = 510
TRAIN_NORM_1, TRAIN_NORM_2, TRAIN_NORM_3, TRAIN_SYNC, TRAIN_EXT = 0, 1, 2, 3, 4
RX_UNLOCKED, RX_KNOW_FSTART, RX_LOCKED = 0, 1, 2

# numpy views of include/tdm_burst_b200.h
BURST_DTYPE = np.dtype([("bitnum", "<u4"), ("train_seq", "<i4"), ("tn", "<u4"), ("fn", "<u4"), ("mn", "<u4"),
                        ("call_index", "<u4"), ("reserved", "<u4", (2,)), ("bits", "u1", (512,))], align=True)
STATE_DTYPE = np.dtype([("state", "<i4"), ("bits_in_buf", "<u4"), ("bitbuf_start_bitnum", "<u4"),
                        ("next_frame_start_bitnum", "<u4"), ("tn", "<u4"), ("fn", "<u4"), ("mn", "<u4"),
                        ("ts_found", "<u4"), ("ts_expire", "<u4"), ("ts_window_lo", "<u4"), ("ts_window_hi", "<u4"),
                        ("searched_upto", "<u4"), ("n_bits", "<u8"), ("n_bursts", "<u8"), ("bitbuf", "<u4", (128,))],
                       align=True)
BLOCK_DTYPE = np.dtype([("type", "<i4"), ("blk_num", "<i4"), ("n_bits", "<i4"), ("bits", "u1", (432,))], align=True)
# what must agree between the CUDA path and the checkers (searched_upto is an internal search cursor)
STATE_COMPARE = ["state", "bits_in_buf", "bitbuf_start_bitnum", "next_frame_start_bitnum", "tn", "fn", "mn", "n_bits",
                 "n_bursts", "bitbuf"]
TS_COMPARE = ["ts_found", "ts_expire", "ts_window_lo", "ts_window_hi"]
assert BURST_DTYPE.itemsize == 544 and STATE_DTYPE.itemsize == 48 + 16 + 512, (BURST_DTYPE.itemsize, STATE_DTYPE.itemsize)

# ETSI EN 300 392-2 9.4.4.3 training sequences (the reference tabulates the same: phy/tetra_burst.c:61-72)
SEQ = {
    "n": [1,1, 0,1, 0,0, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,1, 0,0],
    "p": [0,1, 1,1, 1,0, 1,0, 0,1, 0,0, 0,0, 1,1, 0,1, 1,1, 1,0],
    "q": [1,0, 1,1, 0,1, 1,1, 0,0, 0,0, 0,1, 1,0, 1,0, 1,1, 0,1],
    "N": [1,1,1, 0,0,1, 1,0,1, 1,1,1, 0,0,0, 1,1,1, 1,0,0, 0,1,1, 1,1,0, 0,0,0, 0,0,0],
    "P": [1,0,1, 0,1,1, 1,1,1, 1,0,1, 0,1,0, 1,0,1, 1,1,0, 0,0,1, 1,0,0, 0,1,0, 0,1,0],
    "x": [1,0, 0,1, 1,1, 0,1, 0,0, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,1, 0,0, 0,0, 1,1],
    "X": [0,1,1,1,0,0,1,1,0,1,0,0,0,0,1,0,0,0,1,1,1,0,1,1,0,1,0,1,0,1,1,1,1,1,0,1,0,0,0,0,0,1,1,1,0],
    "y": [1,1, 0,0, 0,0, 0,1, 1,0, 0,1, 1,1, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,0, 0,0, 0,1, 1,0, 0,1, 1,1],
}
SEQ = {k: np.array(v, dtype=np.uint8) for k, v in SEQ.items()}
# 9.4.4.3.1 frequency correction field f1..f80: 8 ones, 64 zeros, 8 ones
F_BITS = np.concatenate([np.ones(8, np.uint8), np.zeros(64, np.uint8), np.ones(8, np.uint8)])


def build():
    subprocess.run(["make", "-C", HERE, "oracle_bsync"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/src/decoder/src/phy"):
        subprocess.run(["make", "-C", HERE, "ref_bsync"], check=True, capture_output=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        L.obs_find_train_seq.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.obs_find_train_seq.restype = C.c_int
        L.obs_in.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        L.obs_in.restype = C.c_int
        L.obs_ts_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.obs_ts_detect.restype = None
        L.obs_burst_demux.argtypes = [C.c_void_p, C.c_void_p]
        L.obs_burst_demux.restype = C.c_int
        _port = L
    return _port


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.rbs_new.restype = C.c_void_p
        L.rbs_free.argtypes = [C.c_void_p]
        L.rbs_in.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.rbs_in.restype = C.c_int
        L.rbs_get_time.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.rbs_set_time.argtypes = [C.c_uint32] * 3
        L.rbs_get_state.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), C.c_void_p]
        L.rbs_find_train_seq.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.rbs_find_train_seq.restype = C.c_int
        L.rbs_build_sync_burst.argtypes = [C.c_void_p] * 4
        L.rbs_build_sync_burst.restype = C.c_int
        L.rbs_build_norm_burst.argtypes = [C.c_void_p] * 4 + [C.c_int]
        L.rbs_build_norm_burst.restype = C.c_int
        _ref = L
    return _ref


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """one bit per byte -> 128 MSB-first uint32 words (the tdm_bsync_state.bitbuf layout)"""
    out = np.zeros(128, dtype=np.uint32)
    b = np.zeros(4096, dtype=np.uint8)
    b[:len(bits)] = bits & 1
    out[:] = np.packbits(b).view(">u4").astype(np.uint32)
    return out


class PortBsync:
    """oracle_bsync.c for C channels; state has the product's tdm_bsync_state layout."""

    def __init__(self, n_channels: int):
        self.C = n_channels
        self.states = np.zeros(n_channels, dtype=STATE_DTYPE)
        self.L = port_lib()

    def feed(self, bits: np.ndarray, n_bits, call_bits: int, max_bursts: int, detect_ts: bool = False):
        """bits [C][stride] one bit per byte; n_bits int or [C].  Returns (n_bursts [C], bursts [C][max_bursts])."""
        n_bits = np.broadcast_to(np.asarray(n_bits, dtype=np.int64), (self.C,))
        bursts = np.zeros((self.C, max_bursts), dtype=BURST_DTYPE)
        nb = np.zeros(self.C, dtype=np.int32)
        for c in range(self.C):
            row = np.ascontiguousarray(bits[c, :n_bits[c]])
            st = self.states[c:c + 1]
            nb[c] = self.L.obs_in(_vp(st), _vp(row), int(n_bits[c]), call_bits, _vp(bursts[c]), max_bursts)
            if detect_ts:
                self.L.obs_ts_detect(_vp(st), _vp(row), int(n_bits[c]))
        return nb, bursts

    def find_train_seq(self, buf: np.ndarray, end: int, mask: int):
        off = C.c_uint32(0)
        buf = np.ascontiguousarray(buf)
        rc = self.L.obs_find_train_seq(_vp(buf), end, mask, C.byref(off))
        return rc, off.value

    def demux(self, burst: np.ndarray):
        blocks = np.zeros(3, dtype=BLOCK_DTYPE)
        b = np.ascontiguousarray(burst.reshape(1))
        n = self.L.obs_burst_demux(_vp(b), _vp(blocks))
        return blocks[:n]


class RefBsync:
    """The reference's tetra_burst_sync_in, one context per channel; the process-global slot counter is swapped in/out."""

    def __init__(self, n_channels: int):
        self.L = ref_lib()
        self.C = n_channels
        self.ctx = [self.L.rbs_new() for _ in range(n_channels)]
        self.time = [(0, 0, 0)] * n_channels

    def close(self):
        for c in self.ctx:
            self.L.rbs_free(c)
        self.ctx = []

    def feed(self, bits: np.ndarray, n_bits, call_bits: int, max_bursts: int):
        n_bits = np.broadcast_to(np.asarray(n_bits, dtype=np.int64), (self.C,))
        bursts = np.zeros((self.C, max_bursts), dtype=BURST_DTYPE)      # rbs_burst has the same layout (n_blocks in reserved[0])
        blocks = np.zeros((self.C, max_bursts, 3), dtype=BLOCK_DTYPE)
        nb = np.zeros(self.C, dtype=np.int32)
        for c in range(self.C):
            row = np.ascontiguousarray(bits[c, :n_bits[c]])
            self.L.rbs_set_time(*self.time[c])
            nb[c] = self.L.rbs_in(self.ctx[c], _vp(row), int(n_bits[c]), call_bits, _vp(bursts[c]), _vp(blocks[c]), max_bursts)
            t = [C.c_uint32(), C.c_uint32(), C.c_uint32()]
            self.L.rbs_get_time(*[C.byref(x) for x in t])
            self.time[c] = tuple(x.value for x in t)
        return nb, bursts, blocks

    def state(self, c: int):
        st, nb, sb, nf = C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        buf = np.zeros(4096 + 64, dtype=np.uint8)
        self.L.rbs_get_state(self.ctx[c], C.byref(st), C.byref(nb), C.byref(sb), C.byref(nf), _vp(buf))
        return {"state": st.value, "bits_in_buf": nb.value, "bitbuf_start_bitnum": sb.value,
                "next_frame_start_bitnum": nf.value, "tn": self.time[c][0], "fn": self.time[c][1], "mn": self.time[c][2],
                "bitbuf": pack_bits(buf[:nb.value])}

    def find_train_seq(self, buf: np.ndarray, end: int, mask: int):
        padded = np.zeros(len(buf) + 32, dtype=np.uint8)        # the reference's look-ahead reads 21 bytes past `end`
        padded[:len(buf)] = buf
        off = C.c_uint32(0)
        rc = self.L.rbs_find_train_seq(_vp(padded), end, mask, C.byref(off))
        return rc, off.value


# ------------------------------------------------------------------------------------------------
# protocol-valid downlink bit streams (EN 300 392-2 9.4.4.2.5/6 continuous downlink bursts; field layout as in
# the reference's build_sync_c_d_burst / build_norm_c_d_burst, phy/tetra_burst.c:171-269)
# ------------------------------------------------------------------------------------------------
def sync_burst(rng) -> np.ndarray:
    b = np.empty(BITS_PER_TS, dtype=np.uint8)
    b[0:12] = SEQ["q"][10:22]
    b[12:14] = rng.integers(0, 2, 2)                 # phase adjustment bits hc (payload for this stage)
    b[14:94] = F_BITS
    b[94:214] = rng.integers(0, 2, 120)              # sb(1..120)
    b[214:252] = SEQ["y"]
    b[252:282] = rng.integers(0, 2, 30)              # bb(1..30)
    b[282:498] = rng.integers(0, 2, 216)             # bkn2
    b[498:500] = rng.integers(0, 2, 2)               # hd
    b[500:510] = SEQ["q"][0:10]
    return b


def norm_burst(rng, two_log_chan: bool) -> np.ndarray:
    b = np.empty(BITS_PER_TS, dtype=np.uint8)
    b[0:12] = SEQ["q"][10:22]
    b[12:14] = rng.integers(0, 2, 2)
    b[14:230] = rng.integers(0, 2, 216)              # bkn1
    b[230:244] = rng.integers(0, 2, 14)              # bb(1..14)
    b[244:266] = SEQ["p"] if two_log_chan else SEQ["n"]
    b[266:282] = rng.integers(0, 2, 16)              # bb(15..30)
    b[282:498] = rng.integers(0, 2, 216)             # bkn2
    b[498:500] = rng.integers(0, 2, 2)
    b[500:510] = SEQ["q"][0:10]
    return b


def downlink_stream(seed: int, n_slots: int, lead_bits: int | None = None, sync_every: int = 4, ber: float = 0.0,
                    glitch_at: tuple = ()) -> np.ndarray:
    """random lead-in, then n_slots continuous downlink bursts (a SYNC burst every sync_every-th slot, NORM_1/NORM_2
    otherwise); `glitch_at` slots get 1..509 extra garbage bits in front (forces re-acquisition); bits flip with
    probability `ber`."""
    rng = np.random.default_rng(seed)
    if lead_bits is None:
        lead_bits = int(rng.integers(0, 1500))
    parts = [rng.integers(0, 2, lead_bits).astype(np.uint8)]
    for s in range(n_slots):
        if s in glitch_at:
            parts.append(rng.integers(0, 2, int(rng.integers(1, BITS_PER_TS))).astype(np.uint8))
        parts.append(sync_burst(rng) if s % sync_every == 0 else norm_burst(rng, bool(rng.integers(0, 2))))
    bits = np.concatenate(parts)
    if ber > 0:
        bits ^= (rng.random(len(bits)) < ber).astype(np.uint8)
    return bits


def bits_to_dibits(bits: np.ndarray) -> np.ndarray:
    """BitUnpacker inverse: dibit = b[2i] << 1 | b[2i+1] (src/dsp/bit_unpacker.cpp:6-7)"""
    n = len(bits) // 2
    return ((bits[0:2 * n:2] << 1) | bits[1:2 * n:2]).astype(np.uint8)
