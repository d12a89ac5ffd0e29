"""B200-native TETRA pi/4-DQPSK demodulation chain (AGC -> FLL -> RRC -> ML timing recovery ->
Costas -> slicer -> differential decoder) behind the SDR++ block surface of
cropinghigh/sdrpp-tetra-demodulator's src/dsp.  The product is libtdm_b200.so (hand-written
sm_100a CUDA behind the C ABI in include/tdm_b200.h); this package is its host-side mirror."""
from . import capi
from .capi import TdmConfig, TdmDesign, TdmError, default_config, design_from_config
from .burst import BurstSync, burst_demux, bursts_raw, bursts_view, find_train_seq
from .chan import Channelizer, TdmChanConfig, chan_default_config, chan_design
from .demod import BitUnpacker, Demodulator, DemodResult, DQPSKSymbolExtractor, PI4DQPSK, synth_capture

__all__ = ["capi", "TdmConfig", "TdmDesign", "TdmError", "default_config", "design_from_config", "Demodulator",
           "DemodResult", "PI4DQPSK", "DQPSKSymbolExtractor", "BitUnpacker", "synth_capture",
           "Channelizer", "TdmChanConfig", "chan_default_config", "chan_design", "BurstSync", "burst_demux", "bursts_raw", "bursts_view", "find_train_seq"]
