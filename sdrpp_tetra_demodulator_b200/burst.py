"""Host-side mirror of the burst synchroniser (include/tdm_burst_b200.h) for C channels.

Reference surface (paths relative to the reference tree): `tetra_burst_sync_in(trs, bits, len)`
(src/decoder/src/phy/tetra_burst_sync.c:54), `tetra_find_train_seq(in, end_of_in, mask, &offset)`
(src/decoder/src/phy/tetra_burst.c:271), the bursts `tetra_burst_rx_cb` receives (phy/tetra_burst.c:343), and
the network mode's training-sequence detector (src/main.cpp:385-414).  All work happens in libtdm_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import BURST_DTYPE, BURST_UNPACKED_DTYPE, BSYNC_STATE_DTYPE, TP_SAP_BLOCK_DTYPE, check, lib


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return t.ctypes.data_as(C.c_void_p)


class BurstSync:
    """C independent `struct tetra_rx_state` receivers, stepped in lock step on one GPU.

    feed(bits, ...) = for every channel, `tetra_burst_sync_in(trs, bits + k*call_bits, call_bits)` for k = 0, 1, ...;
    returns the bursts the reference would have passed to tetra_burst_rx_cb."""

    def __init__(self, n_channels: int, max_units: int, device: int = 0):
        self.n_channels, self.max_units, self.device = int(n_channels), int(max_units), int(device)
        h = C.c_void_p()
        check(lib().tdm_bsync_create(self.n_channels, self.max_units, self.device, C.byref(h)), "tdm_bsync_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().tdm_bsync_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def use_torch_stream(self):
        check(lib().tdm_bsync_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "tdm_bsync_set_stream")

    def reset(self):
        check(lib().tdm_bsync_reset(self._h), "tdm_bsync_reset")

    def feed(self, data, n_units=None, *, dibits: bool = False, call_bits: int = 432, max_bursts: int = 0,
             detect_ts: bool = False, out=None):
        """data: [C][stride] uint8 -- a CUDA tensor (asynchronous on the handle's stream) or a numpy array (host path).
        n_units: None (all of every row), an int, or a per-channel int32 array/tensor living where `data` lives.
        Returns (n_bursts [C] int32, bursts [C][max_bursts] as BURST_DTYPE-viewable uint8 rows)."""
        on_dev = isinstance(data, torch.Tensor)
        assert data.ndim == 2 and data.shape[0] == self.n_channels
        stride = data.stride(0) if on_dev else data.strides[0]
        units_all, units_ptr = int(data.shape[1]), None
        if isinstance(n_units, (int, np.integer)):
            units_all = int(n_units)
        elif n_units is not None:
            assert isinstance(n_units, torch.Tensor) == on_dev
            assert n_units.dtype in (torch.int32, np.int32)
            units_ptr, units_all = _ptr(n_units), 0
        if out is not None:
            nb, bursts = out
        elif on_dev:
            nb = torch.empty(self.n_channels, dtype=torch.int32, device=data.device)
            bursts = torch.empty((self.n_channels, max(max_bursts, 1), BURST_DTYPE.itemsize), dtype=torch.uint8, device=data.device)
        else:
            nb = np.zeros(self.n_channels, dtype=np.int32)
            bursts = np.zeros((self.n_channels, max(max_bursts, 1)), dtype=BURST_DTYPE)
        check(lib().tdm_bsync_in(self._h, _ptr(data), stride, units_ptr, units_all,
                                 capi.TDM_BSYNC_IN_DIBITS if dibits else capi.TDM_BSYNC_IN_BITS, call_bits,
                                 _ptr(bursts) if max_bursts > 0 else None, max_bursts, _ptr(nb), int(detect_ts),
                                 capi.TDM_MEM_DEVICE if on_dev else capi.TDM_MEM_HOST), "tdm_bsync_in")
        return nb, bursts

    def get_state(self) -> np.ndarray:
        st = np.zeros(self.n_channels, dtype=BSYNC_STATE_DTYPE)
        check(lib().tdm_bsync_get_state(self._h, st.ctypes.data_as(C.c_void_p), self.n_channels), "tdm_bsync_get_state")
        return st

    def set_state(self, st: np.ndarray):
        st = np.ascontiguousarray(st, dtype=BSYNC_STATE_DTYPE)
        check(lib().tdm_bsync_set_state(self._h, st.ctypes.data_as(C.c_void_p), self.n_channels), "tdm_bsync_set_state")

    def last_kernel_ms(self):
        """(pack, detect, sync) device milliseconds of the most recent feed()"""
        ms = (C.c_float * 3)()
        check(lib().tdm_bsync_last_kernel_ms(self._h, ms), "tdm_bsync_last_kernel_ms")
        return tuple(float(x) for x in ms)

    def launch_count(self) -> int:
        return int(lib().tdm_bsync_launch_count(self._h))


def bursts_raw(bursts) -> np.ndarray:
    """[C][max_bursts] structured view (BURST_DTYPE, bits packed) of what feed() returned"""
    if isinstance(bursts, torch.Tensor):
        bursts = bursts.cpu().numpy()
    if bursts.dtype == BURST_DTYPE:
        return bursts
    return np.ascontiguousarray(bursts).view(BURST_DTYPE).reshape(bursts.shape[0], bursts.shape[1])


def bursts_view(bursts) -> np.ndarray:
    """[C][max_bursts] records with the burst one bit per byte (BURST_UNPACKED_DTYPE): the form
    tetra_burst_rx_cb(burst, 510, type, priv) receives (phy/tetra_burst.c:343)"""
    raw = bursts_raw(bursts)
    out = np.zeros(raw.shape, dtype=BURST_UNPACKED_DTYPE)
    for f in ("bitnum", "train_seq", "tn", "fn", "mn", "call_index"):
        out[f] = raw[f]
    be = raw["bits"].astype(">u4")                                   # MSB-first words -> bytes in stream order
    out["bits"] = np.unpackbits(be.view(np.uint8).reshape(raw.shape + (64,)), axis=-1)
    return out


def find_train_seq(bufs, end_of_in: int, mask: int, device: int = 0):
    """tetra_find_train_seq for every row of `bufs` ([C][>= end_of_in] uint8, numpy or CUDA tensor):
    returns (type [C] int32, -1 = none; offset [C] uint32)."""
    on_dev = isinstance(bufs, torch.Tensor)
    n = bufs.shape[0]
    stride = bufs.stride(0) if on_dev else bufs.strides[0]
    if on_dev:
        typ = torch.empty(n, dtype=torch.int32, device=bufs.device)
        off = torch.empty(n, dtype=torch.int32, device=bufs.device)
        stream = C.c_void_p(torch.cuda.current_stream(bufs.device).cuda_stream)
        device = bufs.device.index
    else:
        typ, off, stream = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.uint32), None
    check(lib().tdm_find_train_seq(device, stream, _ptr(bufs), stride, n, end_of_in, mask, _ptr(typ), _ptr(off),
                                   capi.TDM_MEM_DEVICE if on_dev else capi.TDM_MEM_HOST), "tdm_find_train_seq")
    return typ, off


def burst_demux(burst: np.ndarray) -> np.ndarray:
    """tetra_burst_rx_cb's split of one burst record into the blocks it passes to tp_sap_udata_ind"""
    if burst.dtype != BURST_DTYPE:                                   # unpacked record -> the ABI's packed one
        rec = np.zeros(1, dtype=BURST_DTYPE)
        for f in ("bitnum", "train_seq", "tn", "fn", "mn", "call_index"):
            rec[f] = burst[f]
        rec["bits"][0] = np.packbits(np.asarray(burst["bits"]).reshape(512) & 1).view(">u4").astype(np.uint32)
        burst = rec[0]
    b = np.ascontiguousarray(burst.reshape(1))
    blocks = np.zeros(3, dtype=TP_SAP_BLOCK_DTYPE)
    n = lib().tdm_burst_demux(b.ctypes.data_as(C.c_void_p), blocks.ctypes.data_as(C.c_void_p))
    return blocks[:n]
