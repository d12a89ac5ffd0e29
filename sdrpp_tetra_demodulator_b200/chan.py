"""Host-side mirror of the channeliser C ABI (include/tdm_chan_b200.h): one wideband capture -> M channels at 36 kS/s,
channel-major in HBM, the layout Demodulator.process() takes in place."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class TdmChanConfig(C.Structure):
    _fields_ = [("n_channels", C.c_int32), ("decimation", C.c_int32), ("taps_per_branch", C.c_int32), ("reserved", C.c_int32),
                ("passband", C.c_double), ("stopband", C.c_double), ("stop_atten_db", C.c_double)]


def _lib():
    L = capi.lib()
    if not getattr(L, "_chan_ready", False):
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.tdm_chan_default_config.argtypes = [i32, C.POINTER(TdmChanConfig)]
        L.tdm_chan_design.argtypes = [C.POINTER(TdmChanConfig), vp]
        L.tdm_chan_create.argtypes = [C.POINTER(TdmChanConfig), i32, C.POINTER(vp)]
        L.tdm_chan_destroy.argtypes = [vp]
        L.tdm_chan_reset.argtypes = [vp]
        L.tdm_chan_process.argtypes = [vp, vp, i64, vp, i64, vp]
        L.tdm_chan_process_instant_major.argtypes = [vp, vp, i64, vp, i64, vp]
        L.tdm_chan_process_ex.argtypes = [vp, vp, i32, i64, vp, i64, i32, vp]
        L.tdm_chan_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L._chan_ready = True
    return L


def chan_default_config(g: int) -> TdmChanConfig:
    cfg = TdmChanConfig()
    capi.check(_lib().tdm_chan_default_config(int(g), C.byref(cfg)), "tdm_chan_default_config")
    return cfg


def chan_design(cfg: TdmChanConfig) -> np.ndarray:
    taps = np.zeros(cfg.taps_per_branch * cfg.n_channels, np.float32)
    capi.check(_lib().tdm_chan_design(C.byref(cfg), taps.ctypes.data_as(C.c_void_p)), "tdm_chan_design")
    return taps


class Channelizer:
    def __init__(self, config: TdmChanConfig | None = None, g: int = 4, device: int = 0):
        self._lib = _lib()
        self.config = config if config is not None else chan_default_config(g)
        self.device = int(device)
        h = C.c_void_p()
        capi.check(self._lib.tdm_chan_create(C.byref(self.config), self.device, C.byref(h)), "tdm_chan_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tdm_chan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        capi.check(self._lib.tdm_chan_reset(self._h), "tdm_chan_reset")

    def process(self, wide, out=None, instant_major: bool = False):
        """wide: CUDA float32 (or int16: CS16, value s / 32768) tensor [N][2], N a multiple of the decimation -> [M][N/D][2], or with instant_major=True
        [N/D][M][2] (the DFT's own order: no transposing pass; Demodulator.process(..., instant_major=True) reads it in
        place).  Asynchronous on torch's current stream."""
        import torch
        assert wide.is_cuda and wide.dtype in (torch.float32, torch.int16) and wide.dim() == 2 and wide.shape[1] == 2 and wide.is_contiguous()
        n = int(wide.shape[0])
        n_out = n // self.config.decimation
        M = self.config.n_channels
        if out is None:
            out = torch.empty((n_out, M, 2) if instant_major else (M, n_out, 2), dtype=torch.float32, device=wide.device)
        st = torch.cuda.current_stream(wide.device).cuda_stream
        # int16 input = TDM_CHAN_IN_CS16 (interleaved int16 pairs, value s / 32768)
        capi.check(self._lib.tdm_chan_process_ex(self._h, C.c_void_p(wide.data_ptr()), 1 if wide.dtype == torch.int16 else 0, n,
                                                 C.c_void_p(out.data_ptr()), out.stride(0) // 2, 1 if instant_major else 0, C.c_void_p(st)),
                   "tdm_chan_process_ex")
        return out

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        capi.check(self._lib.tdm_chan_last_kernel_ms(self._h, C.byref(a), C.byref(b)), "tdm_chan_last_kernel_ms")
        return float(a.value), float(b.value)
