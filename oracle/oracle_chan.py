"""TEST INFRASTRUCTURE -- float64 restatement of the channeliser's DEFINITION (include/tdm_chan_b200.h):

    y_c[m] = sum_{n < T M} h[n] x[t_m - n] exp(-j 2 pi c (t_m - n) / M),   t_m = (m + 1) D - 1,  x[<0] = 0

PARITY UNPINNED BY THE REFERENCE: the reference does not channelise (it asks SDR++ for one VFO per instance,
/root/reference/src/main.cpp:75, and SDR++ core is not vendored), so there is nothing of the reference's to pin this
against.  The restatement evaluates the defining sum directly (no polyphase split, no FFT), which is what makes it an
independent check of the polyphase/FFT factorisation on the device.  Also here: a wideband test-signal builder that
places narrowband 36 kS/s captures on the 25 kHz raster (rational upsampling + mixing, float64).
Only tests/ and bench.py's checking legs import this."""
from __future__ import annotations

import numpy as np


def channelize_direct(x: np.ndarray, h: np.ndarray, M: int, D: int, channels, instants) -> np.ndarray:
    """x complex128 [N], h float64 [T M] -> y[len(channels)][len(instants)] by the defining sum."""
    L = len(h)
    out = np.zeros((len(channels), len(instants)), np.complex128)
    n = np.arange(L)
    for j, m in enumerate(instants):
        t = (m + 1) * D - 1
        idx = t - n
        ok = idx >= 0
        seg = np.zeros(L, np.complex128)
        seg[ok] = x[idx[ok]]
        for i, c in enumerate(channels):
            out[i, j] = np.sum(h * seg * np.exp(-2j * np.pi * c * (idx % M) / M))
    return out


def place_on_raster(narrow: np.ndarray, channels, M: int, D: int) -> np.ndarray:
    """narrow complex [K][n] at fs_out = fs_wide / D -> wideband complex128 [n D]: each capture interpolated by D
    (windowed-sinc, 8 output-rate periods each side) and mixed to channel c's centre c fs_wide / M."""
    from scipy.signal import resample_poly
    K, n = narrow.shape
    wide = np.zeros(n * D, np.complex128)
    t = np.arange(n * D)
    for k in range(K):
        up = resample_poly(narrow[k].astype(np.complex128), D, 1, window=("kaiser", 9.0))
        wide += up[:n * D] * np.exp(2j * np.pi * channels[k] * (t % M) / M)
    return wide
