"""Host-side mirror of the reference's operator surface, on top of the C ABI.

Two layers:

* ``Demodulator`` -- a batch of C independent channels behind one ``tdm_handle``
  (the analogue of C plugin instances, src/main.cpp:51).  Device tensors go through the
  zero-copy path, numpy arrays through the library's own host<->device staging.

* ``PI4DQPSK`` / ``DQPSKSymbolExtractor`` / ``BitUnpacker`` -- single-channel classes with
  the reference's names, ``init(...)`` argument order, ``process(count, in, out)`` return
  convention, setters and ``reset()`` (src/dsp/pi4dqpsk.h:27-81, src/dsp/dqpsk_sym_extr.h:19-46,
  src/dsp/bit_unpacker.h:16-34), so the parity tests read like tests of the reference.
  The GPU kernel is fused: one launch produces symbols, dibits and bits; the extractor and
  unpacker mirrors hand out the fused results of the demodulator they are attached to.

PyTorch is used only as the owner of device memory and streams.  All arithmetic happens
in libtdm_b200.so; nothing here computes demodulator outputs on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi


def _torch():
    import torch
    return torch


@dataclass
class DemodResult:
    counts: object          # [C] int32 symbols written per channel
    symbols: object | None  # [C][S][2] float32
    dibits: object | None   # [C][S] uint8, values 0..3
    bits: object | None     # [C][2S] uint8, values 0/1
    packed: object | None = None   # [C][S/4] uint8, four dibits per byte, first symbol in bits 7..6 (TDM_OUT_PACKED)


class Demodulator:
    """C channels x (AGC -> FLL -> RRC -> timing -> Costas -> slicer -> diff decoder), state carried across calls."""

    def __init__(self, n_channels: int = 1, max_chunk: int = capi.TDM_STREAM_BUFFER_SIZE, device: int = 0,
                 config: capi.TdmConfig | None = None):
        self._lib = capi.lib()
        self.n_channels = int(n_channels)
        self.max_chunk = int(max_chunk)
        self.device = int(device)
        self.config = config if config is not None else capi.default_config()
        h = C.c_void_p()
        capi.check(self._lib.tdm_create(C.byref(self.config), self.n_channels, self.max_chunk, self.device,
                                        C.byref(h)), "tdm_create")
        self._h = h

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.tdm_destroy(self._h)
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

    # -- configuration / control
    def set_stream(self, cuda_stream_ptr: int | None) -> None:
        """cuda_stream_ptr: a cudaStream_t as an int (0 = CUDA's legacy default stream); None = the handle's
        own private stream (TDM_OWN_STREAM)."""
        ptr = C.c_void_p(-1) if cuda_stream_ptr is None else C.c_void_p(int(cuda_stream_ptr))
        capi.check(self._lib.tdm_set_stream(self._h, ptr), "tdm_set_stream")

    def use_torch_stream(self) -> None:
        torch = _torch()
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def set_config(self, config: capi.TdmConfig) -> None:
        capi.check(self._lib.tdm_set_config(self._h, C.byref(config)), "tdm_set_config")
        self.config = config

    def set_params(self, config: capi.TdmConfig, what: int) -> None:
        """one of PI4DQPSK's setters with the reference's partial effect (tdm_set_params, capi.TDM_SET_*)"""
        capi.check(self._lib.tdm_set_params(self._h, C.byref(config), int(what)), "tdm_set_params")
        self.config = config

    def set_kernel_variant(self, variant: int) -> None:
        capi.check(self._lib.tdm_set_kernel_variant(self._h, int(variant)), "tdm_set_kernel_variant")

    def design(self) -> capi.TdmDesign:
        d = capi.TdmDesign()
        capi.check(self._lib.tdm_get_design(self._h, C.byref(d)), "tdm_get_design")
        return d

    def max_symbols(self, count: int) -> int:
        return int(self._lib.tdm_max_symbols(self._h, int(count)))

    def reset(self) -> None:
        capi.check(self._lib.tdm_reset(self._h), "tdm_reset")

    def reset_all(self) -> None:
        capi.check(self._lib.tdm_reset_all(self._h), "tdm_reset_all")

    def get_state(self) -> np.ndarray:
        st = np.zeros(self.n_channels, dtype=capi.STATE_DTYPE)
        capi.check(self._lib.tdm_get_state(self._h, st.ctypes.data_as(C.c_void_p), self.n_channels), "tdm_get_state")
        return st

    def set_state(self, states: np.ndarray) -> None:
        st = np.ascontiguousarray(states, dtype=capi.STATE_DTYPE)
        assert st.shape == (self.n_channels,)
        capi.check(self._lib.tdm_set_state(self._h, st.ctypes.data_as(C.c_void_p), self.n_channels), "tdm_set_state")

    def metrics(self) -> np.ndarray:
        m = np.zeros(self.n_channels, dtype=capi.METRICS_DTYPE)
        capi.check(self._lib.tdm_get_metrics(self._h, m.ctypes.data_as(C.c_void_p), self.n_channels),
                   "tdm_get_metrics")
        return m

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        capi.check(self._lib.tdm_last_kernel_ms(self._h, C.byref(ms)), "tdm_last_kernel_ms")
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self._lib.tdm_launch_count(self._h))

    # -- the hot call
    def process(self, iq, symbols: bool = False, dibits: bool = True, bits: bool = False, packed: bool = False,
                out: DemodResult | None = None, instant_major: bool = False) -> DemodResult:
        """iq: [C][N][2] float32 -- a CUDA torch tensor (zero-copy, asynchronous on the handle's stream)
        or a numpy array (staged through the library, synchronous).  instant_major=True (CUDA tensors only): iq is
        [N][C][2], the order the channeliser's DFT leaves (tdm_io.sample_stride)."""
        flags = (capi.TDM_OUT_SYMBOLS if symbols else 0) | (capi.TDM_OUT_DIBITS if dibits else 0) | \
                (capi.TDM_OUT_BITS if bits else 0) | (capi.TDM_OUT_PACKED if packed else 0)
        if isinstance(iq, np.ndarray):
            if instant_major:
                raise ValueError("instant-major input needs a CUDA tensor")
            return self._process_host(iq, flags)
        return self._process_device(iq, flags, out, instant_major)

    def _process_host(self, iq: np.ndarray, flags: int) -> DemodResult:
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        if iq.ndim != 3 or iq.shape[0] != self.n_channels or iq.shape[2] != 2:
            raise ValueError(f"iq must be [C={self.n_channels}][N][2] float32, got {iq.shape}")
        n = iq.shape[1]
        s = self.max_symbols(n)
        syms = np.zeros((self.n_channels, s, 2), np.float32) if flags & capi.TDM_OUT_SYMBOLS else None
        dib = np.zeros((self.n_channels, s), np.uint8) if flags & capi.TDM_OUT_DIBITS else None
        bit = np.zeros((self.n_channels, 2 * s), np.uint8) if flags & capi.TDM_OUT_BITS else None
        pk = np.zeros((self.n_channels, s // 4), np.uint8) if flags & capi.TDM_OUT_PACKED else None
        counts = np.zeros(self.n_channels, np.int32)
        p = lambda a: None if a is None else a.ctypes.data
        io = capi.TdmIo(p(iq), n, n, capi.TDM_MEM_HOST, p(syms), p(dib), p(bit), p(pk), s, s // 4, p(counts), flags, 0)
        capi.check(self._lib.tdm_process_io(self._h, C.byref(io)), "tdm_process_io")
        return DemodResult(counts, syms, dib, bit, pk)

    def _process_device(self, iq, flags: int, out: DemodResult | None, instant_major: bool = False) -> DemodResult:
        torch = _torch()
        if instant_major:
            if not (iq.is_cuda and iq.dtype == torch.float32 and iq.dim() == 3 and iq.shape[2] == 2 and
                    iq.shape[1] == self.n_channels and iq.stride(2) == 1 and iq.stride(1) == 2 and iq.stride(0) % 2 == 0):
                raise ValueError("iq must be a CUDA float32 tensor [N][C][2] with contiguous instants")
            n, in_stride, sample_stride = iq.shape[0], 1, iq.stride(0) // 2
        else:
            if not (iq.is_cuda and iq.dtype == torch.float32 and iq.dim() == 3 and iq.shape[2] == 2 and
                    iq.shape[0] == self.n_channels and iq.stride(2) == 1 and iq.stride(1) == 2):
                raise ValueError("iq must be a CUDA float32 tensor [C][N][2] with contiguous rows")
            n, in_stride, sample_stride = iq.shape[1], iq.stride(0) // 2, 0
        s = self.max_symbols(n)
        dev = iq.device
        # device tensors belong to torch's stream-ordered allocator: enqueue on torch's current stream so
        # the launch is ordered after the producers of `iq` and before any consumer of the outputs
        self.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        if out is None:
            out = DemodResult(
                torch.empty(self.n_channels, dtype=torch.int32, device=dev),
                torch.empty((self.n_channels, s, 2), dtype=torch.float32, device=dev) if flags & capi.TDM_OUT_SYMBOLS else None,
                torch.empty((self.n_channels, s), dtype=torch.uint8, device=dev) if flags & capi.TDM_OUT_DIBITS else None,
                torch.empty((self.n_channels, 2 * s), dtype=torch.uint8, device=dev) if flags & capi.TDM_OUT_BITS else None,
                torch.empty((self.n_channels, s // 4), dtype=torch.uint8, device=dev) if flags & capi.TDM_OUT_PACKED else None)
        stride = None
        for t, div in ((out.symbols, 1), (out.dibits, 1), (out.bits, 2)):
            if t is not None:
                st = t.shape[1] // div
                stride = st if stride is None else min(stride, st)
        if stride is None:
            stride = s
        p = lambda t: None if t is None else t.data_ptr()
        io = capi.TdmIo(p(iq), in_stride, n, capi.TDM_MEM_DEVICE, p(out.symbols), p(out.dibits), p(out.bits), p(out.packed),
                        stride, 0 if out.packed is None else out.packed.shape[1], p(out.counts), flags, sample_stride)
        capi.check(self._lib.tdm_process_io(self._h, C.byref(io)), "tdm_process_io")
        return out

    def process_long(self, iq, warmup: int = 65536, out=None):
        """ONE long capture of one channel ([N][2] float32, CUDA tensor or numpy array) demodulated as up to
        n_channels overlapping time segments in parallel (tdm_process_long; decoded dibits only).
        Returns (dibits [n] uint8 -- same kind of array as `iq` --, info dict)."""
        info = capi.TdmLongInfo()
        n = int(iq.shape[0])
        cap = n // 2 + 64
        if isinstance(iq, np.ndarray):
            iq = np.ascontiguousarray(iq, dtype=np.float32)
            dib = np.zeros(cap, np.uint8) if out is None else out
            capi.check(self._lib.tdm_process_long(self._h, iq.ctypes.data_as(C.c_void_p), n, warmup, dib.ctypes.data_as(C.c_void_p),
                                                  len(dib), C.byref(info), capi.TDM_MEM_HOST), "tdm_process_long")
        else:
            torch = _torch()
            if not (iq.is_cuda and iq.dtype == torch.float32 and iq.dim() == 2 and iq.shape[1] == 2 and iq.is_contiguous()):
                raise ValueError("iq must be a contiguous CUDA float32 tensor [N][2]")
            self.set_stream(torch.cuda.current_stream(iq.device).cuda_stream)
            dib = torch.empty(cap, dtype=torch.uint8, device=iq.device) if out is None else out
            capi.check(self._lib.tdm_process_long(self._h, C.c_void_p(iq.data_ptr()), n, warmup, C.c_void_p(dib.data_ptr()),
                                                  dib.numel(), C.byref(info), capi.TDM_MEM_DEVICE), "tdm_process_long")
        d = {f: int(getattr(info, f)) for f, _ in capi.TdmLongInfo._fields_}
        return dib[:d["n_dibits"]], d

    def process_long_batch(self, iq, warmup: int = 65536, out=None):
        """C long captures at once ([C][N][2] float32, CUDA tensor or numpy array), every channel cut into
        n_channels // C overlapping time segments (tdm_process_long_batch; decoded dibits only).
        Returns (dibits [C][N // 2 + 64] uint8, counts [C] int64, info dict) -- same kind of arrays as `iq`."""
        info = capi.TdmLongInfo()
        Cn, n = int(iq.shape[0]), int(iq.shape[1])
        stride = n // 2 + 64
        if isinstance(iq, np.ndarray):
            iq = np.ascontiguousarray(iq, dtype=np.float32)
            dib = np.zeros((Cn, stride), np.uint8) if out is None else out
            cnt = np.zeros(Cn, np.int64)
            capi.check(self._lib.tdm_process_long_batch(self._h, iq.ctypes.data_as(C.c_void_p), n, n, Cn, warmup, dib.ctypes.data_as(C.c_void_p),
                                                        dib.shape[1], cnt.ctypes.data_as(C.c_void_p), C.byref(info), capi.TDM_MEM_HOST),
                       "tdm_process_long_batch")
        else:
            torch = _torch()
            if not (iq.is_cuda and iq.dtype == torch.float32 and iq.dim() == 3 and iq.shape[2] == 2 and iq.stride(2) == 1 and iq.stride(1) == 2):
                raise ValueError("iq must be a CUDA float32 tensor [C][N][2] with contiguous rows")
            self.set_stream(torch.cuda.current_stream(iq.device).cuda_stream)
            dib = torch.empty((Cn, stride), dtype=torch.uint8, device=iq.device) if out is None else out
            cnt = torch.empty(Cn, dtype=torch.int64, device=iq.device)
            capi.check(self._lib.tdm_process_long_batch(self._h, C.c_void_p(iq.data_ptr()), iq.stride(0) // 2, n, Cn, warmup,
                                                        C.c_void_p(dib.data_ptr()), dib.shape[1], C.c_void_p(cnt.data_ptr()), C.byref(info),
                                                        capi.TDM_MEM_DEVICE), "tdm_process_long_batch")
        d = {f: int(getattr(info, f)) for f, _ in capi.TdmLongInfo._fields_}
        return dib, cnt, d

    def pack_dibits(self, dibits, counts):
        """4 dibits per byte (first symbol in bits 7..6): the form shipped over NVLink by the multi-GPU gather."""
        torch = _torch()
        s = dibits.shape[1]
        packed = torch.empty((self.n_channels, (s + 3) // 4), dtype=torch.uint8, device=dibits.device)
        capi.check(self._lib.tdm_pack_dibits(self._h, C.c_void_p(dibits.data_ptr()), s, C.c_void_p(counts.data_ptr()),
                                             C.c_void_p(packed.data_ptr()), packed.shape[1]), "tdm_pack_dibits")
        return packed


    def unpack_dibits(self, packed, counts, max_symbols: int | None = None, dibits: bool = True, bits: bool = False):
        """Packed rows (this handle's, or the rows gathered from every rank: any row count) -> one dibit per byte
        and / or one bit per byte, the stream BitUnpacker / the NETSYMS sink emit (tdm_unpack_dibits)."""
        torch = _torch()
        rows = int(packed.shape[0])
        m = int(packed.shape[1]) * 4 if max_symbols is None else int(max_symbols)
        m16 = (m + 15) // 16 * 16
        self.set_stream(torch.cuda.current_stream(packed.device).cuda_stream)
        d = torch.empty((rows, m16), dtype=torch.uint8, device=packed.device) if dibits else None
        b = torch.empty((rows, 2 * m16), dtype=torch.uint8, device=packed.device) if bits else None
        p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        capi.check(self._lib.tdm_unpack_dibits(self._h, p(packed), packed.stride(0), p(counts), rows, p(d), m16, p(b), 2 * m16, m),
                   "tdm_unpack_dibits")
        return d, b


def synth_capture(n_channels: int, n_samples: int, device: int = 0, snr_db: float = 30.0,
                  max_freq_off_hz: float = 300.0, min_amp: float = 0.05, max_amp: float = 2.0,
                  seed_data: int = 12345, seed_noise: int = 777, first_channel: int = 0, want_tx: bool = False,
                  want_iq: bool = True):
    """Synthetic TETRA-mapped pi/4-DQPSK capture generated in HBM (SURVEY.md 8d recipe).
    Returns (iq [C][N][2] float32 cuda | None, tx_dibits [C][N/2+64] uint8 cuda | None)."""
    torch = _torch()
    dev = torch.device("cuda", device)
    iq = torch.empty((n_channels, n_samples, 2), dtype=torch.float32, device=dev) if want_iq else None
    tx = torch.zeros((n_channels, n_samples // 2 + 64), dtype=torch.uint8, device=dev) if want_tx else None
    sp = capi.TdmSynthParams(snr_db, max_freq_off_hz, min_amp, max_amp, seed_data, seed_noise)
    stream = torch.cuda.current_stream(dev).cuda_stream
    capi.check(capi.lib().tdm_synth_capture(device, C.c_void_p(stream), C.byref(sp), n_channels, n_samples, n_samples,
                                            first_channel, C.c_void_p(iq.data_ptr() if iq is not None else 0),
                                            C.c_void_p(tx.data_ptr() if tx is not None else 0),
                                            tx.shape[1] if tx is not None else 0), "tdm_synth_capture")
    return iq, tx


# ------------------------------------------------------------------------------------------
# Reference-shaped single-channel classes
# ------------------------------------------------------------------------------------------
class PI4DQPSK:
    """dsp::demod::PI4DQPSK (src/dsp/pi4dqpsk.h:27-81): same init() arguments, process() returns the
    number of symbols written to `out`.  Streams/threads (run(), start(), stop()) belong to the C++
    block in host/pi4dqpsk_b200.h; this mirror covers the arithmetic surface."""

    def __init__(self, device: int = 0):
        self._device = device
        self._dm: Demodulator | None = None
        self._cfg: capi.TdmConfig | None = None
        self.last: DemodResult | None = None

    def init(self, in_=None, symbolrate=18000.0, samplerate=36000.0, rrcTapCount=65, rrcBeta=0.35, agcRate=0.02,
             costasBandwidth=0.01, fllBandwidth=0.006, omegaGain=None, muGain=None, omegaRelLimit=0.01):
        d = capi.default_config()
        cfg = capi.TdmConfig(symbolrate, samplerate, int(rrcTapCount), 0, rrcBeta, agcRate, costasBandwidth,
                             fllBandwidth, d.omega_gain if omegaGain is None else omegaGain,
                             d.mu_gain if muGain is None else muGain, omegaRelLimit)
        self._cfg = cfg
        self._dm = Demodulator(1, capi.TDM_STREAM_BUFFER_SIZE, self._device, cfg)

    def init_default(self):
        """init() with exactly what src/main.cpp:84 passes."""
        self._cfg = capi.default_config()
        self._dm = Demodulator(1, capi.TDM_STREAM_BUFFER_SIZE, self._device, self._cfg)

    def _need(self) -> Demodulator:
        assert self._dm is not None, "block not initialised"   # assert(base_type::_block_init)
        return self._dm

    def process(self, count: int, in_: np.ndarray, out: np.ndarray) -> int:
        """in_: [count][2] float32; out: [>=count][2] float32 receives complex symbols.  Returns #symbols."""
        dm = self._need()
        iq = np.ascontiguousarray(in_[:count], dtype=np.float32).reshape(1, count, 2)
        r = dm.process(iq, symbols=True, dibits=True, bits=True)
        n = int(r.counts[0])
        out[:n] = r.symbols[0, :n]
        self.last = r
        return n

    def reset(self):
        self._need().reset()

    # setters, src/dsp/pi4dqpsk.h:52-63 -- each with the reference's own partial effect (src/dsp/pi4dqpsk.cpp:31-118)
    def _set(self, what, **kw):
        dm = self._need()
        for k, v in kw.items():
            setattr(self._cfg, k, v)
        dm.set_params(self._cfg, what)

    def setSymbolrate(self, symbolrate): self._set(capi.TDM_SET_RATES, symbolrate=symbolrate)
    def setSamplerate(self, samplerate): self._set(capi.TDM_SET_RATES, samplerate=samplerate)
    def setRRCParams(self, rrcTapCount, rrcBeta): self._set(capi.TDM_SET_RRC, rrc_tap_count=int(rrcTapCount), rrc_beta=rrcBeta)
    def setRRCTapCount(self, rrcTapCount): self._set(capi.TDM_SET_RRC, rrc_tap_count=int(rrcTapCount))
    def setRRCBeta(self, rrcBeta): self._set(capi.TDM_SET_RRC, rrc_beta=int(rrcBeta))          # int, like the reference ([A.9])
    def setAGCRate(self, agcRate): self._set(capi.TDM_SET_AGC_RATE, agc_rate=agcRate)
    def setCostasBandwidth(self, bandwidth): self._set(capi.TDM_SET_COSTAS_BW, costas_bandwidth=bandwidth)
    def setFllBandwidth(self, fllBandwidth): self._set(capi.TDM_SET_FLL_BW, fll_bandwidth=fllBandwidth)
    def setMMParams(self, omegaGain, muGain, omegaRelLimit=0.01):
        self._set(capi.TDM_SET_TIMING_GAINS, omega_gain=omegaGain, mu_gain=muGain, omega_rel_limit=omegaRelLimit)
    def setOmegaGain(self, omegaGain): self._set(capi.TDM_SET_TIMING_GAINS, omega_gain=omegaGain)
    def setMuGain(self, muGain): self._set(capi.TDM_SET_TIMING_GAINS, mu_gain=muGain)
    def setOmegaRelLimit(self, omegaRelLimit): self._set(capi.TDM_SET_TIMING_GAINS, omega_rel_limit=omegaRelLimit)


class DQPSKSymbolExtractor:
    """dsp::DQPSKSymbolExtractor (src/dsp/dqpsk_sym_extr.h:19-46).  Attached to a PI4DQPSK: the slicer and
    differential decoder run inside the same fused kernel launch, this hands out that launch's dibits."""

    def __init__(self):
        self._src: PI4DQPSK | None = None
        self.sync = False
        self.standarderr = 0.0

    def init(self, demod: PI4DQPSK):
        self._src = demod

    def process(self, count: int, in_, out: np.ndarray) -> int:
        r = self._src.last
        n = int(r.counts[0])
        assert count == n, "extractor must be fed the symbols of the demodulator's last process() call"
        out[:n] = r.dibits[0, :n]
        m = self._src._need().metrics()[0]
        self.sync = bool(m["sync"])
        self.standarderr = float(m["standarderr"])
        return n


class BitUnpacker:
    """dsp::BitUnpacker (src/dsp/bit_unpacker.h:16-34): dibit/byte -> 2 x bit/byte, MSB first."""

    def __init__(self):
        self._src: PI4DQPSK | None = None

    def init(self, extractor: DQPSKSymbolExtractor):
        self._src = extractor._src

    def process(self, count: int, in_, out: np.ndarray) -> int:
        r = self._src.last
        n = int(r.counts[0])
        assert count == n
        out[:2 * n] = r.bits[0, :2 * n]
        return 2 * n
