"""TEST INFRASTRUCTURE -- ctypes loaders for the two CPU checkers and the capture generator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (sdrpp_tetra_demodulator_b200) never does.

  OracleA  oracle/_ref/libtetra_ref.so   the reference's own src/dsp/*.cpp (authority: bits)
  OracleB  oracle/_build/liboracle_b_*.so canonical-order restatement (authority: float state)
  siggen   deterministic synthetic pi/4-DQPSK captures
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libtetra_ref.so")
REF_SO_REONLY = os.path.join(HERE, "_ref", "libtetra_ref_reonly.so")   # fastAmplitude() with both operands from |re|
REF_SO_SIMD = os.path.join(HERE, "_ref", "libtetra_ref_simd.so")       # AVX2+FMA dot products in the VOLK stand-in: timing only
TDM_CFG_FASTAMP_RE_ONLY = 1

TDM_MAX_TAPS = 65
TDM_HIST = 64
TDM_INTERP_PHASES = 128
TDM_INTERP_TAPS = 8
TDM_SYNC_BLOCKS = 16


class TdmConfig(C.Structure):
    _fields_ = [
        ("symbolrate", C.c_double), ("samplerate", C.c_double),
        ("rrc_tap_count", C.c_int32), ("flags", C.c_int32),
        ("rrc_beta", C.c_double), ("agc_rate", C.c_double), ("costas_bandwidth", C.c_double),
        ("fll_bandwidth", C.c_double), ("omega_gain", C.c_double), ("mu_gain", C.c_double),
        ("omega_rel_limit", C.c_double),
    ]


class TdmDesign(C.Structure):
    _fields_ = [
        ("ntaps", C.c_int32), ("fastamp_re_only", C.c_int32),
        ("rrc", C.c_float * TDM_MAX_TAPS), ("be_a", C.c_float * TDM_MAX_TAPS), ("be_b", C.c_float * TDM_MAX_TAPS),
        ("bank", (C.c_float * TDM_INTERP_TAPS) * TDM_INTERP_PHASES),
        ("agc_rate", C.c_float), ("agc_set_point", C.c_float), ("agc_max_gain", C.c_float), ("agc_init_gain", C.c_float),
        ("fll_beta", C.c_float), ("fll_min_freq", C.c_float), ("fll_max_freq", C.c_float), ("fll_init_freq", C.c_float),
        ("tr_alpha", C.c_float), ("tr_beta", C.c_float), ("tr_min_omega", C.c_float), ("tr_max_omega", C.c_float),
        ("tr_init_omega", C.c_float),
        ("costas_alpha", C.c_float), ("costas_beta", C.c_float), ("costas_min_freq", C.c_float),
        ("costas_max_freq", C.c_float),
        ("reserved1", C.c_float * 3),
    ]


# numpy view of tdm_channel_state (include/tdm_b200.h); itemsize must equal sizeof(tdm_channel_state)
STATE_DTYPE = np.dtype([
    ("agc_gain", "<f4"), ("fll_phase", "<f4"), ("fll_freq", "<f4"), ("tr_mu", "<f4"), ("tr_omega", "<f4"),
    ("tr_offset", "<i4"), ("costas_phase", "<f4"), ("costas_freq", "<f4"), ("costas_ph2", "<f4"),
    ("prev_sym", "<u4"), ("err_ptr", "<u4"), ("err_disp", "<u4"), ("err_partial", "<f4"),
    ("standarderr", "<f4"), ("sync", "<u4"), ("fll_quad", "<u4"), ("n_samples", "<u8"), ("n_symbols", "<u8"),
    ("err_blocks", "<f4", (TDM_SYNC_BLOCKS,)), ("x_hist", "<f4", (2 * TDM_HIST,)),
    ("r_hist", "<f4", (2 * (TDM_INTERP_TAPS - 1),)), ("fll_r", "<f4"), ("reserved1", "<f4"),
], align=True)

# fields that must match bit-for-bit between the CUDA path and Oracle B
EXACT_STATE_FIELDS = ["agc_gain", "fll_phase", "fll_freq", "tr_mu", "tr_omega", "tr_offset", "costas_phase",
                      "costas_freq", "costas_ph2", "prev_sym", "err_ptr", "err_disp", "n_samples", "n_symbols",
                      "x_hist", "r_hist", "fll_quad", "fll_r"]
# atan2f-derived GUI metric: libm vs CUDA differ in the last place -> tolerance
METRIC_STATE_FIELDS = ["err_partial", "standarderr", "err_blocks"]


class SgParams(C.Structure):
    _fields_ = [("snr_db", C.c_double), ("max_freq_off_hz", C.c_double), ("min_amp", C.c_double),
                ("max_amp", C.c_double), ("seed_data", C.c_uint64), ("seed_noise", C.c_uint64)]


class TrefParams(C.Structure):
    _fields_ = [("symbolrate", C.c_double), ("samplerate", C.c_double), ("rrc_taps", C.c_int),
                ("rrc_beta", C.c_double), ("agc_rate", C.c_double), ("costas_bw", C.c_double),
                ("fll_bw", C.c_double), ("omega_gain", C.c_double), ("mu_gain", C.c_double),
                ("omega_rel_limit", C.c_double)]


class TrefLoopState(C.Structure):
    _fields_ = [("agc_gain", C.c_float), ("fll_phase", C.c_float), ("fll_freq", C.c_float), ("tr_mu", C.c_float),
                ("tr_omega", C.c_float), ("tr_offset", C.c_int32), ("costas_phase", C.c_float),
                ("costas_freq", C.c_float), ("costas_ph2", C.c_float), ("prev_sym", C.c_uint32),
                ("standarderr", C.c_float), ("sync", C.c_int32)]


def _has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return False


def build(ref: bool = True) -> None:
    """make the checkers (Oracle B always; Oracle A only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "oracle_b"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


_LIB_B = None


def lib_b() -> C.CDLL:
    global _LIB_B
    if _LIB_B is None:
        name = "liboracle_b_fma.so" if _has_fma() else "liboracle_b_generic.so"
        path = os.path.join(HERE, "_build", name)
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.ob_default_config.argtypes = [C.POINTER(TdmConfig)]
        L.ob_design.argtypes = [C.POINTER(TdmConfig), C.POINTER(TdmDesign)]
        L.ob_design.restype = C.c_int
        L.ob_state_init.argtypes = [C.POINTER(TdmDesign), C.c_void_p]
        L.ob_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ob_fll_nco.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ob_fll_fallback_count.restype = C.c_long
        L.ob_process.argtypes = [C.POINTER(TdmDesign), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.ob_process.restype = C.c_int64
        L.ob_process_multi.argtypes = [C.POINTER(TdmDesign), C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.sg_generate.argtypes = [C.POINTER(SgParams), C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.sg_tx_dibits.argtypes = [C.c_uint64, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.sg_hash.argtypes = [C.c_uint64, C.c_uint64]
        L.sg_hash.restype = C.c_uint64
        _LIB_B = L
    return _LIB_B


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- generator
def default_sg_params(snr_db=30.0, max_freq_off_hz=300.0, min_amp=0.05, max_amp=2.0, seed_data=12345,
                      seed_noise=777) -> SgParams:
    """SURVEY.md 8d defaults: SNR 30 dB, df in U(-300,300) Hz, amplitude logU(0.05,2), seeds 12345+c / 777+c."""
    return SgParams(snr_db, max_freq_off_hz, min_amp, max_amp, seed_data, seed_noise)


def generate(n_channels: int, n_samples: int, params: SgParams | None = None, first_channel: int = 0,
             n0: int = 0) -> np.ndarray:
    """[C][N][2] float32 capture, channel-major."""
    p = params or default_sg_params()
    out = np.empty((n_channels, n_samples, 2), dtype=np.float32)
    L = lib_b()
    for c in range(n_channels):
        L.sg_generate(C.byref(p), first_channel + c, n0, n_samples, _ptr(out[c]))
    return out


def tx_dibits(channel: int, n_symbols: int, seed_data: int = 12345, k0: int = 0) -> np.ndarray:
    out = np.empty(n_symbols, dtype=np.uint8)
    lib_b().sg_tx_dibits(seed_data, channel, k0, n_symbols, _ptr(out))
    return out


# --------------------------------------------------------------------------- Oracle B
class OracleB:
    """Canonical-order restatement, C channels with carried state."""

    def __init__(self, n_channels: int = 1, config: TdmConfig | None = None, fastamp_re_only: bool = False):
        self.L = lib_b()
        self.cfg = config or self.default_config()
        if fastamp_re_only:
            self.cfg.flags |= TDM_CFG_FASTAMP_RE_ONLY
        self.design = TdmDesign()
        rc = self.L.ob_design(C.byref(self.cfg), C.byref(self.design))
        if rc != 0:
            raise ValueError(f"ob_design failed: {rc}")
        self.n_channels = n_channels
        self.states = np.zeros(n_channels, dtype=STATE_DTYPE)
        for c in range(n_channels):
            self.L.ob_state_init(C.byref(self.design), self.states[c:c + 1].ctypes.data_as(C.c_void_p))

    def set_params(self, cfg: TdmConfig, what: int):
        """tdm_set_params' contract (include/tdm_b200.h TDM_SET_*), restated: which parts of the design a setter of
        the reference replaces (src/dsp/pi4dqpsk.cpp:31-118), plus COMPLEX_FD::setOmega's restart for the rates."""
        d = TdmDesign()
        rc = self.L.ob_design(C.byref(cfg), C.byref(d))
        if rc != 0:
            raise ValueError(f"ob_design failed: {rc}")
        cur = self.design
        if what & (1 | 2):
            cur.rrc, cur.ntaps = d.rrc, d.ntaps
        if what & 4:
            cur.agc_rate = d.agc_rate
        if what & 8:
            cur.costas_alpha, cur.costas_beta = d.costas_alpha, d.costas_beta
        if what & 16:
            cur.fll_beta = d.fll_beta
        if what & 32:
            cur.tr_alpha, cur.tr_beta = d.tr_alpha, d.tr_beta
            cur.tr_min_omega, cur.tr_max_omega = d.tr_min_omega, d.tr_max_omega
        if what & 1:
            cur.tr_init_omega, cur.tr_min_omega, cur.tr_max_omega = d.tr_init_omega, d.tr_min_omega, d.tr_max_omega
            self.states["tr_offset"] = 0
            self.states["tr_mu"] = 0
            self.states["tr_omega"] = d.tr_init_omega
        self.cfg = cfg

    def reset(self):
        """tdm_reset's contract, restated: PI4DQPSK::reset() (src/dsp/pi4dqpsk.cpp:120-130) = FLL phase/frequency
        (fll.cpp:120-127), the matched filter's delay line, the AGC gain, the Costas loop's phase/frequency and the
        timing loop (complex_fd.cpp:78-87) back to their initial values; ph2, the slicer's memory, the lock metric
        and the interpolator's history stay."""
        st, d = self.states, self.design
        st["fll_phase"] = 0
        st["fll_freq"] = d.fll_init_freq
        st["fll_quad"] = 0
        st["fll_r"] = 0
        st["x_hist"] = 0
        st["agc_gain"] = d.agc_init_gain
        st["costas_phase"] = 0
        st["costas_freq"] = 0
        st["tr_offset"] = 0
        st["tr_mu"] = 0
        st["tr_omega"] = d.tr_init_omega

    @staticmethod
    def default_config() -> TdmConfig:
        cfg = TdmConfig()
        lib_b().ob_default_config(C.byref(cfg))
        return cfg

    def process(self, iq: np.ndarray, want_syms=True, want_bits=False, nthreads: int = 1):
        """iq [C][N][2] float32 -> (counts[C], syms[C][S][2] | None, dibits[C][S], bits[C][2S] | None)."""
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        Cn, N = iq.shape[0], iq.shape[1]
        assert Cn == self.n_channels
        S = int(N / 1.9) + 4
        syms = np.zeros((Cn, S, 2), dtype=np.float32) if want_syms else None
        dibits = np.zeros((Cn, S), dtype=np.uint8)
        bits = np.zeros((Cn, 2 * S), dtype=np.uint8) if want_bits else None
        counts = np.zeros(Cn, dtype=np.int32)
        self.L.ob_process_multi(C.byref(self.design), _ptr(self.states), Cn, _ptr(iq), N, N, _ptr(syms), _ptr(dibits),
                                _ptr(bits), S, _ptr(counts), nthreads)
        return counts, syms, dibits, bits


def count_fll_fallbacks(iq: np.ndarray, config: TdmConfig | None = None, fastamp_re_only: bool = False) -> int:
    """How many samples of `iq` take the classic range reduction in the FLL's NCO (oracle_b.c ob_fll_reduce)."""
    L = lib_b()
    n0 = L.ob_fll_fallback_count()
    OracleB(iq.shape[0], config, fastamp_re_only=fastamp_re_only).process(iq, want_syms=False)
    return int(L.ob_fll_fallback_count() - n0)


# --------------------------------------------------------------------------- Oracle A
def have_ref(fastamp_re_only: bool = False) -> bool:
    return os.path.exists(REF_SO_REONLY if fastamp_re_only else REF_SO)


def have_ref_simd() -> bool:
    """the timing build of the reference (VOLK stand-in with 256-bit FMA kernels) and a host that can run it"""
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((line for line in f if line.startswith("flags")), "")
    except OSError:
        flags = ""
    return os.path.exists(REF_SO_SIMD) and " avx2 " in flags + " " and " fma " in flags + " "


_LIB_A = {}


def lib_a(fastamp_re_only: bool = False, simd: bool = False) -> C.CDLL:
    key = "simd" if simd else bool(fastamp_re_only)
    if key not in _LIB_A:
        path = REF_SO_SIMD if simd else (REF_SO_REONLY if fastamp_re_only else REF_SO)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(path)
        L.tref_default_params.argtypes = [C.POINTER(TrefParams)]
        L.tref_create.argtypes = [C.POINTER(TrefParams)]
        L.tref_create.restype = C.c_void_p
        L.tref_destroy.argtypes = [C.c_void_p]
        L.tref_reset.argtypes = [C.c_void_p]
        L.tref_process.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tref_process.restype = C.c_int64
        L.tref_get_state.argtypes = [C.c_void_p, C.POINTER(TrefLoopState)]
        L.tref_set.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.tref_get_taps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.tref_get_taps.restype = C.c_int
        L.tref_get_coeffs.argtypes = [C.c_void_p, C.c_void_p]
        L.tref_process_multi.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int]
        _LIB_A[key] = L
    return _LIB_A[key]


class OracleA:
    """The reference's own PI4DQPSK -> DQPSKSymbolExtractor -> BitUnpacker chain, one instance per channel."""

    def __init__(self, n_channels: int = 1, params: TrefParams | None = None, fastamp_re_only: bool = False, simd: bool = False):
        self.L = lib_a(fastamp_re_only, simd)
        self.n_channels = n_channels
        self.handles = [self.L.tref_create(C.byref(params) if params is not None else None)
                        for _ in range(n_channels)]

    def close(self):
        for h in self.handles:
            self.L.tref_destroy(h)
        self.handles = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, iq: np.ndarray, want_syms=True, want_bits=False):
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        Cn, N = iq.shape[0], iq.shape[1]
        assert Cn == self.n_channels
        S = int(N / 1.9) + 4
        syms = np.zeros((Cn, S, 2), dtype=np.float32) if want_syms else None
        dibits = np.zeros((Cn, S), dtype=np.uint8)
        bits = np.zeros((Cn, 2 * S), dtype=np.uint8) if want_bits else None
        counts = np.zeros(Cn, dtype=np.int32)
        for c, h in enumerate(self.handles):
            counts[c] = self.L.tref_process(h, N, _ptr(iq[c]), _ptr(syms[c]) if want_syms else None,
                                            _ptr(dibits[c]), _ptr(bits[c]) if want_bits else None)
        return counts, syms, dibits, bits

    def process_multi(self, iq: np.ndarray, nthreads: int):
        """dibits only, channels spread over nthreads std::threads (all-cores CPU baseline)."""
        iq = np.ascontiguousarray(iq, dtype=np.float32)
        Cn, N = iq.shape[0], iq.shape[1]
        S = int(N / 1.9) + 4
        dibits = np.zeros((Cn, S), dtype=np.uint8)
        counts = np.zeros(Cn, dtype=np.int64)
        arr = (C.c_void_p * Cn)(*self.handles)
        self.L.tref_process_multi(arr, Cn, N, _ptr(iq), _ptr(dibits), S, _ptr(counts), nthreads)
        return counts, dibits

    def reset(self):
        """PI4DQPSK::reset() of every chain (src/dsp/pi4dqpsk.cpp:120-130)"""
        for h in self.handles:
            self.L.tref_reset(h)

    def set(self, what: int, a: float, b: float = 0.0, c: float = 0.0):
        """the reference's own setters on every chain (ref_driver.cpp tref_set: 1 setSymbolrate .. 10 setMuGain)"""
        for h in self.handles:
            self.L.tref_set(h, int(what), float(a), float(b), float(c))

    def loop_state(self, c: int = 0) -> TrefLoopState:
        s = TrefLoopState()
        self.L.tref_get_state(self.handles[c], C.byref(s))
        return s

    def taps(self):
        n = TDM_MAX_TAPS
        rrc = np.zeros(n, np.float32)
        lbe = np.zeros((n, 2), np.float32)
        hbe = np.zeros((n, 2), np.float32)
        bank = np.zeros((TDM_INTERP_PHASES, TDM_INTERP_TAPS), np.float32)
        P, T = C.c_int(), C.c_int()
        nt = self.L.tref_get_taps(self.handles[0], _ptr(rrc), _ptr(lbe), _ptr(hbe), _ptr(bank), C.byref(P), C.byref(T))
        return nt, rrc[:nt], lbe[:nt], hbe[:nt], bank, P.value, T.value

    def coeffs(self) -> np.ndarray:
        out = np.zeros(16, np.float32)
        self.L.tref_get_coeffs(self.handles[0], _ptr(out))
        return out
