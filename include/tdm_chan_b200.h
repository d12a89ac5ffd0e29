/* tdm_chan_b200.h -- C ABI of the front-end channeliser of libtdm_b200.so:
 * ONE wideband complex capture -> M channels at 36 kS/s, channel-major, ready for tdm_process(TDM_MEM_DEVICE).
 *
 * What it stands in for.  The reference plugin does not channelise: it asks SDR++ for one VFO per plugin instance
 * (sigpath::vfoManager.createVFO(... 30 kHz bandwidth, 36 kS/s), /root/reference/src/main.cpp:75) and SDR++ core
 * (not vendored by the reference) runs one frequency translation + decimating FIR per VFO on the CPU.  For thousands
 * of channels of ONE wideband receiver that is one pass over the wideband stream per channel; here it is one pass in
 * total: an oversampled polyphase filterbank (prototype low-pass of T*M taps, M branches, decimation D, M/D = 36/25
 * for TETRA's 25 kHz raster at 36 kS/s) -- a hand-written polyphase kernel + an M-point inverse DFT per output instant
 * (cuFFT, loaded at run time).  PARITY UNPINNED BY THE REFERENCE: there is no reference code for this stage; the
 * checker is a float64 restatement of the defining sum (oracle/oracle_chan.py) with a stated tolerance, plus an
 * end-to-end test (wideband capture -> channeliser -> demodulator -> transmitted dibits).
 *
 * Definition (x = wideband samples, global index n; h = prototype; t_m = (m + 1) D - 1 = newest sample of output m):
 *     y_c[m] = sum_{n < T M} h[n] x[t_m - n] exp(-j 2 pi c (t_m - n) / M),      c = 0 .. M-1
 * i.e. channel c is the band centred at c * fs / M (c > M/2: negative frequencies), mixed to 0 Hz against the capture's
 * sample 0, low-pass filtered, decimated by D.  Calls carry their history: consecutive calls continue one stream.
 */
#ifndef TDM_CHAN_B200_H
#define TDM_CHAN_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tdm_chan_config {
    int32_t n_channels;        /* M: channels = polyphase branches (fs_wide = M * spacing)                  */
    int32_t decimation;        /* D: wideband samples per output sample (fs_out = fs_wide / D), D <= M      */
    int32_t taps_per_branch;   /* T: prototype length is T * M; 4 .. 32                                     */
    int32_t reserved;
    double passband;           /* prototype pass-band edge, as a fraction of the channel spacing (e.g. 0.6)  */
    double stopband;           /* stop-band edge, same unit (e.g. 0.88); Kaiser window from the two          */
    double stop_atten_db;      /* wanted stop-band attenuation (sets the Kaiser beta), e.g. 70               */
} tdm_chan_config;

typedef struct tdm_chan tdm_chan;

/* TETRA at 36 kS/s per channel (src/main.cpp:35-36,75): spacing 25 kHz, M = 36 g, D = 25 g, fs_wide = 0.9 g MHz. */
int tdm_chan_default_config(int32_t g, tdm_chan_config* cfg);
/* Host-side prototype design (no GPU): taps[T * M], unit DC gain. */
int tdm_chan_design(const tdm_chan_config* cfg, float* taps);

int tdm_chan_create(const tdm_chan_config* cfg, int32_t device, tdm_chan** out);
int tdm_chan_destroy(tdm_chan* c);
int tdm_chan_reset(tdm_chan* c);            /* forget the history: the next call starts a new capture at sample 0 */

/* wide: n_wide interleaved float32 pairs in DEVICE memory, n_wide a multiple of D;
 * out : [M][out_stride] float32 pairs in device memory, n_wide / D samples written per channel (out_stride >= that);
 * asynchronous on cuda_stream (a cudaStream_t).  The first T*M - 1 samples before the first call are zeros. */
int tdm_chan_process(tdm_chan* c, const float* wide, int64_t n_wide, float* out, int64_t out_stride, void* cuda_stream);

/* The same samples INSTANT-major: out [n_wide / D][row_pitch] float32 pairs, sample m of channel k at
 * out + 2 * (m * row_pitch + k), row_pitch >= M.  This is the order the batched DFT produces, so the transposing pass of
 * tdm_chan_process (a third of its time) is not needed; tdm_process_io reads it in place with in_stride = 1,
 * sample_stride = row_pitch (tdm_b200.h).  Calls of the two kinds can be mixed on one handle. */
int tdm_chan_process_instant_major(tdm_chan* c, const float* wide, int64_t n_wide, float* out, int64_t row_pitch, void* cuda_stream);

/* Both of the above with the input format as an argument.  TDM_CHAN_IN_CS16: `wide` is n_wide interleaved int16 pairs
 * (re, im) -- what SDR hardware and baseband recordings deliver -- taken as s / 32768 (the scale of the reference's own
 * volk_16i_s32f_convert_32f call, src/dsp/osmotetra_dec.h:219): half the bytes across PCIe and from HBM; results are
 * bit-identical to feeding the converted floats.  Formats can be mixed from call to call (the history is kept as floats).
 * pitch: out_stride (channel-major) or row_pitch (instant-major). */
#define TDM_CHAN_IN_CF32 0
#define TDM_CHAN_IN_CS16 1
#define TDM_CHAN_OUT_CHANNEL_MAJOR 0
#define TDM_CHAN_OUT_INSTANT_MAJOR 1
int tdm_chan_process_ex(tdm_chan* c, const void* wide, int32_t in_format, int64_t n_wide, float* out, int64_t pitch, int32_t out_layout, void* cuda_stream);

/* kernel time of the last call's two stages in ms (synchronises) */
int tdm_chan_last_kernel_ms(tdm_chan* c, float* polyphase_ms, float* dft_ms);

#ifdef __cplusplus
}
#endif
#endif
