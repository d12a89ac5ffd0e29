"""ctypes binding of libtdm_b200.so (include/tdm_b200.h).

The shared library is the product; this file only declares its signatures.  Loading
fails loudly if the library has not been built -- there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TDM_LIB_OVERRIDE") or os.path.join(PKG_DIR, "libtdm_b200.so")   # override: development A/B builds only

TDM_MAX_TAPS = 65
TDM_HIST = 64
TDM_INTERP_PHASES = 128
TDM_INTERP_TAPS = 8
TDM_SYNC_BLOCKS = 16
TDM_STREAM_BUFFER_SIZE = 1000000

TDM_OK = 0
TDM_ERR_ARG = -1
TDM_ERR_NO_DEVICE = -2
TDM_ERR_CUDA = -3
TDM_ERR_UNSUPPORTED = -4
TDM_ERR_NOMEM = -5

TDM_MEM_HOST = 0
TDM_MEM_DEVICE = 1

TDM_OUT_SYMBOLS = 1
TDM_OUT_DIBITS = 2
TDM_OUT_BITS = 4
TDM_OUT_PACKED = 8
TDM_CFG_FASTAMP_RE_ONLY = 1
TDM_SET_RATES, TDM_SET_RRC, TDM_SET_AGC_RATE, TDM_SET_COSTAS_BW, TDM_SET_FLL_BW, TDM_SET_TIMING_GAINS = 1, 2, 4, 8, 16, 32

# every symbol include/tdm_b200.h declares (tests check the library exports each one)
EXPORTED_SYMBOLS = [
    "tdm_default_config", "tdm_design_from_config", "tdm_create", "tdm_destroy", "tdm_set_stream",
    "tdm_max_symbols", "tdm_process", "tdm_reset", "tdm_reset_all", "tdm_get_state", "tdm_set_state",
    "tdm_get_metrics", "tdm_set_config", "tdm_get_design", "tdm_set_kernel_variant", "tdm_last_kernel_ms",
    "tdm_launch_count", "tdm_pack_dibits", "tdm_synth_capture", "tdm_last_error", "tdm_abi_version",
    "tdm_process_long", "tdm_process_long_batch", "tdm_process_io", "tdm_unpack_dibits",
    "tdm_set_params", "tdm_comm_unique_id", "tdm_comm_create", "tdm_comm_adopt", "tdm_comm_destroy", "tdm_gather_packed",
]
# ... and include/tdm_burst_b200.h
EXPORTED_BURST_SYMBOLS = [
    "tdm_bsync_create", "tdm_bsync_destroy", "tdm_bsync_set_stream", "tdm_bsync_reset", "tdm_bsync_in",
    "tdm_bsync_get_state", "tdm_bsync_set_state", "tdm_bsync_launch_count", "tdm_bsync_last_kernel_ms", "tdm_find_train_seq", "tdm_burst_demux", "tdm_burst_unpack",
]

TDM_BITS_PER_TS = 510
TDM_BSYNC_MAX_CALL_BITS = 510
TDM_TRAIN_NORM_1, TDM_TRAIN_NORM_2, TDM_TRAIN_NORM_3, TDM_TRAIN_SYNC, TDM_TRAIN_EXT = 0, 1, 2, 3, 4
TDM_RX_S_UNLOCKED, TDM_RX_S_KNOW_FSTART, TDM_RX_S_LOCKED = 0, 1, 2
TDM_BSYNC_IN_BITS, TDM_BSYNC_IN_DIBITS = 0, 1


class TdmConfig(C.Structure):
    """tdm_config: the arguments of dsp::demod::PI4DQPSK::init (src/dsp/pi4dqpsk.h:36)."""
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


class TdmSynthParams(C.Structure):
    _fields_ = [("snr_db", C.c_double), ("max_freq_off_hz", C.c_double), ("min_amp", C.c_double),
                ("max_amp", C.c_double), ("seed_data", C.c_uint64), ("seed_noise", C.c_uint64)]


class TdmLongInfo(C.Structure):
    _fields_ = [("n_dibits", C.c_int64), ("n_segments", C.c_int32), ("n_rerun", C.c_int32),
                ("segment_samples", C.c_int32), ("warmup", C.c_int32), ("n_forced", C.c_int32), ("n_extended", C.c_int32)]


class TdmMetrics(C.Structure):
    _fields_ = [("standarderr", C.c_float), ("sync", C.c_uint32), ("n_samples", C.c_uint64),
                ("n_symbols", C.c_uint64)]


# numpy view of tdm_channel_state; itemsize == sizeof(tdm_channel_state) == 720
STATE_DTYPE = np.dtype([
    ("agc_gain", "<f4"), ("fll_phase", "<f4"), ("fll_freq", "<f4"), ("tr_mu", "<f4"), ("tr_omega", "<f4"),
    ("tr_offset", "<i4"), ("costas_phase", "<f4"), ("costas_freq", "<f4"), ("costas_ph2", "<f4"),
    ("prev_sym", "<u4"), ("err_ptr", "<u4"), ("err_disp", "<u4"), ("err_partial", "<f4"),
    ("standarderr", "<f4"), ("sync", "<u4"), ("fll_quad", "<u4"), ("n_samples", "<u8"), ("n_symbols", "<u8"),
    ("err_blocks", "<f4", (TDM_SYNC_BLOCKS,)), ("x_hist", "<f4", (2 * TDM_HIST,)),
    ("r_hist", "<f4", (2 * (TDM_INTERP_TAPS - 1),)), ("fll_r", "<f4"), ("reserved1", "<f4"),
], align=True)
# numpy views of tdm_burst / tdm_bsync_state / tdm_tp_sap_block (include/tdm_burst_b200.h)
BURST_DTYPE = np.dtype([("bitnum", "<u4"), ("train_seq", "<i4"), ("tn", "<u4"), ("fn", "<u4"), ("mn", "<u4"),
                        ("call_index", "<u4"), ("reserved", "<u4", (2,)), ("bits", "<u4", (16,))], align=True)      # 96 bytes, bits packed
# the same record with the burst one bit per byte (what tetra_burst_rx_cb receives); bursts_view() converts
BURST_UNPACKED_DTYPE = np.dtype([("bitnum", "<u4"), ("train_seq", "<i4"), ("tn", "<u4"), ("fn", "<u4"), ("mn", "<u4"),
                                 ("call_index", "<u4"), ("reserved", "<u4", (2,)), ("bits", "u1", (512,))], align=True)
BSYNC_STATE_DTYPE = np.dtype([("state", "<i4"), ("bits_in_buf", "<u4"), ("bitbuf_start_bitnum", "<u4"),
                              ("next_frame_start_bitnum", "<u4"), ("tn", "<u4"), ("fn", "<u4"), ("mn", "<u4"),
                              ("ts_found", "<u4"), ("ts_expire", "<u4"), ("ts_window_lo", "<u4"), ("ts_window_hi", "<u4"),
                              ("searched_upto", "<u4"), ("n_bits", "<u8"), ("n_bursts", "<u8"), ("bitbuf", "<u4", (128,))],
                             align=True)
TP_SAP_BLOCK_DTYPE = np.dtype([("type", "<i4"), ("blk_num", "<i4"), ("n_bits", "<i4"), ("bits", "u1", (432,))], align=True)
METRICS_DTYPE = np.dtype([("standarderr", "<f4"), ("sync", "<u4"), ("n_samples", "<u8"), ("n_symbols", "<u8")],
                         align=True)


class TdmIo(C.Structure):
    """tdm_io (include/tdm_b200.h)"""
    _fields_ = [("iq", C.c_void_p), ("in_stride", C.c_int64), ("count", C.c_int32), ("mem_kind", C.c_int32),
                ("syms", C.c_void_p), ("dibits", C.c_void_p), ("bits", C.c_void_p), ("packed", C.c_void_p),
                ("out_stride", C.c_int64), ("packed_stride", C.c_int64), ("out_counts", C.c_void_p),
                ("out_flags", C.c_uint32), ("sample_stride", C.c_uint32)]


class TdmError(RuntimeError):
    def __init__(self, code: int, where: str, text: str):
        super().__init__(f"{where} failed with tdm_status {code}: {text}")
        self.code = code


_LIB = None


def lib() -> C.CDLL:
    """Load libtdm_b200.so.  Raises if it is missing: the CUDA library IS the product."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {os.path.join(PKG_DIR, 'csrc')}` (needs nvcc; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    sig = {
        "tdm_default_config": (C.c_int, [C.POINTER(TdmConfig)]),
        "tdm_design_from_config": (C.c_int, [C.POINTER(TdmConfig), C.POINTER(TdmDesign)]),
        "tdm_create": (C.c_int, [C.POINTER(TdmConfig), i32, i32, i32, C.POINTER(vp)]),
        "tdm_destroy": (C.c_int, [vp]),
        "tdm_set_stream": (C.c_int, [vp, vp]),
        "tdm_max_symbols": (i64, [vp, i64]),
        "tdm_process": (C.c_int, [vp, vp, i64, i32, vp, vp, vp, i64, vp, u32, i32]),
        "tdm_process_io": (C.c_int, [vp, C.POINTER(TdmIo)]),
        "tdm_unpack_dibits": (C.c_int, [vp, vp, i64, vp, i32, vp, i64, vp, i64, i64]),
        "tdm_comm_unique_id": (C.c_int, [vp]),
        "tdm_comm_create": (C.c_int, [vp, i32, i32, i32, C.POINTER(vp)]),
        "tdm_comm_adopt": (C.c_int, [vp, i32, i32, i32, C.POINTER(vp)]),
        "tdm_comm_destroy": (C.c_int, [vp]),
        "tdm_gather_packed": (C.c_int, [vp, i32, i32, vp, i64, vp, vp, vp, vp]),
        "tdm_reset": (C.c_int, [vp]),
        "tdm_reset_all": (C.c_int, [vp]),
        "tdm_get_state": (C.c_int, [vp, vp, i32]),
        "tdm_set_state": (C.c_int, [vp, vp, i32]),
        "tdm_get_metrics": (C.c_int, [vp, vp, i32]),
        "tdm_set_config": (C.c_int, [vp, C.POINTER(TdmConfig)]),
        "tdm_set_params": (C.c_int, [vp, C.POINTER(TdmConfig), u32]),
        "tdm_get_design": (C.c_int, [vp, C.POINTER(TdmDesign)]),
        "tdm_set_kernel_variant": (C.c_int, [vp, i32]),
        "tdm_last_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "tdm_launch_count": (i64, [vp]),
        "tdm_pack_dibits": (C.c_int, [vp, vp, i64, vp, vp, i64]),
        "tdm_synth_capture": (C.c_int, [i32, vp, C.POINTER(TdmSynthParams), i32, i64, i64, i32, vp, vp, i64]),
        "tdm_last_error": (C.c_char_p, []),
        "tdm_abi_version": (C.c_int, []),
        "tdm_process_long": (C.c_int, [vp, vp, i64, i32, vp, i64, C.POINTER(TdmLongInfo), i32]),
        "tdm_process_long_batch": (C.c_int, [vp, vp, i64, i64, i32, i32, vp, i64, vp, C.POINTER(TdmLongInfo), i32]),
        "tdm_bsync_create": (C.c_int, [i32, i64, i32, C.POINTER(vp)]),
        "tdm_bsync_destroy": (C.c_int, [vp]),
        "tdm_bsync_set_stream": (C.c_int, [vp, vp]),
        "tdm_bsync_reset": (C.c_int, [vp]),
        "tdm_bsync_in": (C.c_int, [vp, vp, i64, vp, i32, i32, i32, vp, i32, vp, i32, i32]),
        "tdm_bsync_get_state": (C.c_int, [vp, vp, i32]),
        "tdm_bsync_set_state": (C.c_int, [vp, vp, i32]),
        "tdm_bsync_launch_count": (i64, [vp]),
        "tdm_bsync_last_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "tdm_find_train_seq": (C.c_int, [i32, vp, vp, i64, i32, u32, u32, vp, vp, i32]),
        "tdm_burst_demux": (C.c_int, [vp, vp]),
        "tdm_burst_unpack": (C.c_int, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def check(rc: int, where: str) -> None:
    if rc != TDM_OK:
        raise TdmError(rc, where, lib().tdm_last_error().decode(errors="replace"))


def default_config() -> TdmConfig:
    cfg = TdmConfig()
    check(lib().tdm_default_config(C.byref(cfg)), "tdm_default_config")
    return cfg


def design_from_config(cfg: TdmConfig) -> TdmDesign:
    d = TdmDesign()
    check(lib().tdm_design_from_config(C.byref(cfg), C.byref(d)), "tdm_design_from_config")
    return d
