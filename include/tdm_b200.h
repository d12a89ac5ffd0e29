/* tdm_b200.h -- C ABI of the B200-native TETRA pi/4-DQPSK demodulator library
 * (libtdm_b200.so, built from sdrpp_tetra_demodulator_b200/csrc by nvcc for sm_100a).
 *
 * The reference plugin has NO foreign-function interface on this path: its seam
 * is a C++ class surface.  Each entry point below names
 * the reference interface it stands in for (paths relative to the reference
 * tree).  The C++ block that keeps SDR++'s dsp::Processor / dsp::stream surface
 * on top of this ABI is sdrpp_tetra_demodulator_b200/host/pi4dqpsk_b200.h.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types;
 *   - every function returns 0 (TDM_OK) or a negative tdm_status;
 *   - one handle = one batch of C independent channels processed in lock step,
 *     the analogue of C plugin instances (src/main.cpp:51 "Max instances -1");
 *   - a handle is single-caller; distinct handles may be used concurrently;
 *   - all device work of a handle is enqueued on ONE CUDA stream (its own, or
 *     the one given to tdm_set_stream);
 *   - there is no CPU fallback: without a usable sm_100 device tdm_create fails.
 */
#ifndef TDM_B200_H
#define TDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDM_ABI_VERSION 2

/* Fixed design sizes of the reference chain (src/main.cpp:35-44,
 * src/dsp/complex_fd.h:30: 128 phases x 8 taps). */
#define TDM_MAX_TAPS 65        /* RRC / band-edge FIR length the kernels are built for   */
#define TDM_HIST (TDM_MAX_TAPS - 1)
#define TDM_INTERP_PHASES 128
#define TDM_INTERP_TAPS 8
#define TDM_SYNC_BUF 4096      /* SYNC_DETECT_BUF,     src/dsp/dqpsk_sym_extr.h:14 */
#define TDM_SYNC_DISPLAY 256   /* SYNC_DETECT_DISPLAY, src/dsp/dqpsk_sym_extr.h:15 */
#define TDM_SYNC_BLOCKS (TDM_SYNC_BUF / TDM_SYNC_DISPLAY)
#define TDM_STREAM_BUFFER_SIZE 1000000 /* SDR++ STREAM_BUFFER_SIZE: largest count per process() call */

typedef enum tdm_status {
    TDM_OK = 0,
    TDM_ERR_ARG = -1,        /* null pointer, negative size, count > max_chunk ...            */
    TDM_ERR_NO_DEVICE = -2,  /* no CUDA device / not sm_100 / driver error at create time     */
    TDM_ERR_CUDA = -3,       /* a CUDA call failed; tdm_last_error() has the text             */
    TDM_ERR_UNSUPPORTED = -4,/* configuration outside what the kernels are built for          */
    TDM_ERR_NOMEM = -5
} tdm_status;

/* Where the caller's buffers live. */
typedef enum tdm_mem_kind {
    TDM_MEM_HOST = 0,   /* pageable or pinned host memory: the library copies in/out itself  */
    TDM_MEM_DEVICE = 1  /* device memory on the handle's device: zero-copy                    */
} tdm_mem_kind;

/* Output selection flags for tdm_process. */
#define TDM_OUT_SYMBOLS 1u  /* complex symbols, PI4DQPSK `out` stream  (src/dsp/pi4dqpsk.h:67)          */
#define TDM_OUT_DIBITS  2u  /* one dibit per byte, DQPSKSymbolExtractor `out` (src/dsp/dqpsk_sym_extr.cpp:34-51) */
#define TDM_OUT_BITS    4u  /* one bit per byte, BitUnpacker `out`     (src/dsp/bit_unpacker.cpp:4-10)   */
#define TDM_OUT_PACKED  8u  /* the same dibits four per byte, first symbol in bits 7..6, last byte zero padded:
                               what travels between GPUs (tdm_gather_packed) -- written by the slicer itself,
                               no second pass over the dibits.  tdm_process_io only.                          */

/* tdm_config.flags.  complex_t::fastAmplitude() (used by the FLL's band-edge error, src/dsp/fll.cpp:143) lives
 * in SDR++ core, which the reference neither vendors nor pins.  Its recalled form is a=|re|, b=|im|,
 * a>b ? a+0.4b : b+0.4a; the other reading of upstream takes BOTH operands from |re| (so it returns
 * 1.4|re|).  The bits that come out differ only while the FLL pulls in, but the float trajectories
 * differ everywhere, so both are built, pinned against the reference compiled the same way
 * (oracle/_ref/libtetra_ref.so / libtetra_ref_reonly.so) and selectable here. */
#define TDM_CFG_FASTAMP_RE_ONLY 1

/* The arguments of dsp::demod::PI4DQPSK::init (src/dsp/pi4dqpsk.h:36), same order
 * and meaning.  tdm_default_config fills in what src/main.cpp:35-44,78-84 passes. */
typedef struct tdm_config {
    double symbolrate;
    double samplerate;
    int32_t rrc_tap_count;
    int32_t flags;           /* TDM_CFG_* bits; 0 = the behaviour SURVEY.md Appendix A records */
    double rrc_beta;
    double agc_rate;
    double costas_bandwidth;
    double fll_bandwidth;
    double omega_gain;
    double mu_gain;
    double omega_rel_limit;
} tdm_config;

/* Everything the kernels read that is derived from tdm_config on the host in
 * double precision (the reference designs its taps on the CPU at init time:
 * src/dsp/pi4dqpsk.cpp:17-22, src/dsp/fll.cpp:61-95, src/dsp/complex_fd.cpp:153-158).
 * Exposed so tests can pin the design against the reference's own tables. */
typedef struct tdm_design {
    int32_t ntaps;                       /* rrc_tap_count, <= TDM_MAX_TAPS                     */
    int32_t fastamp_re_only;             /* tdm_config.flags & TDM_CFG_FASTAMP_RE_ONLY          */
    float rrc[TDM_MAX_TAPS];             /* matched filter taps, oldest-sample-first           */
    float be_a[TDM_MAX_TAPS];            /* band-edge taps: hbe = a + j b, lbe = a - j b        */
    float be_b[TDM_MAX_TAPS];
    float bank[TDM_INTERP_PHASES][TDM_INTERP_TAPS];
    float agc_rate, agc_set_point, agc_max_gain, agc_init_gain;
    float fll_beta, fll_min_freq, fll_max_freq, fll_init_freq;
    float tr_alpha /* mu gain */, tr_beta /* omega gain */, tr_min_omega, tr_max_omega, tr_init_omega;
    float costas_alpha, costas_beta, costas_min_freq, costas_max_freq;
    float reserved1[3];
} tdm_design;

/* Per-channel carried state: what the reference keeps in class members between
 * process() calls (SURVEY.md 8a "per-channel carried state").  This struct IS
 * the checkpoint format of tdm_get_state / tdm_set_state.  All float fields are
 * compared bit-for-bit against the canonical-order oracle in the tests. */
typedef struct tdm_channel_state {
    float agc_gain;                      /* FastAGC::_gain                                      */
    float fll_phase, fll_freq;           /* FLL pcl.phase / pcl.freq   (src/dsp/fll.h:58)        */
    float tr_mu, tr_omega;               /* COMPLEX_FD pcl.phase/freq  (src/dsp/complex_fd.h:56) */
    int32_t tr_offset;                   /* COMPLEX_FD::offset         (src/dsp/complex_fd.h:71) */
    float costas_phase, costas_freq;     /* PLL pcl.phase / pcl.freq                             */
    float costas_ph2;                    /* PI4DQPSK_COSTAS::ph2 (src/dsp/pi4dqpsk_costas.h:32) */
    uint32_t prev_sym;                   /* DQPSKSymbolExtractor::prev                          */
    /* sync metric (src/dsp/dqpsk_sym_extr.cpp:9-31), kept as 16 block sums of 256 */
    uint32_t err_ptr;                    /* errorptr 0..4095                                    */
    uint32_t err_disp;                   /* errordisplayptr 0..255                              */
    float err_partial;                   /* sum of the current (unfinished) block of 256        */
    float standarderr;                   /* DQPSKSymbolExtractor::standarderr                   */
    uint32_t sync;                       /* DQPSKSymbolExtractor::sync                          */
    uint32_t fll_quad;                   /* with fll_r: the FLL phase reduced for the next sample, */
                                         /* fll_phase = fll_quad * pi/2 + fll_r (mod 2 pi); see   */
                                         /* DESIGN.md "canonical order" (the NCO's range reduction */
                                         /* is prepared one sample ahead)                          */
    uint64_t n_samples;                  /* lifetime input samples                              */
    uint64_t n_symbols;                  /* lifetime output symbols                             */
    float err_blocks[TDM_SYNC_BLOCKS];   /* completed block sums, slot = err_ptr / 256          */
    float x_hist[2 * TDM_HIST];          /* last 64 FLL outputs (re,im), oldest first: the ONE   */
                                         /* delay line behind lbe/hbe/RRC FIRs (fll.cpp:141-142, */
                                         /* pi4dqpsk.cpp:135-136)                                */
    float r_hist[2 * (TDM_INTERP_TAPS - 1)]; /* last 7 RRC outputs (complex_fd.cpp:148)          */
    float fll_r;
    float reserved1;
} tdm_channel_state;

/* GUI-facing numbers the plugin reads from the slicer (src/main.cpp:211-217). */
typedef struct tdm_metrics {
    float standarderr;
    uint32_t sync;
    uint64_t n_samples;
    uint64_t n_symbols;
} tdm_metrics;

typedef struct tdm_handle tdm_handle;

/* src/main.cpp:35-44,78-84 -- the plugin's compile-time constants and gain arithmetic. */
int tdm_default_config(tdm_config* cfg);

/* Host-side design (no GPU needed): PI4DQPSK::init's tap/gain derivations. */
int tdm_design_from_config(const tdm_config* cfg, tdm_design* out);

/* PI4DQPSK::init + DQPSKSymbolExtractor::init + BitUnpacker::init for C channels
 * (src/dsp/pi4dqpsk.cpp:11-30, src/main.cpp:84,90-91).  max_chunk = largest
 * `count` a later tdm_process may pass (<= TDM_STREAM_BUFFER_SIZE unless
 * device buffers are used, where it only sizes scratch).  device = CUDA ordinal. */
int tdm_create(const tdm_config* cfg, int32_t n_channels, int32_t max_chunk, int32_t device, tdm_handle** out);

/* ~PI4DQPSK (src/dsp/pi4dqpsk.cpp:5-9). */
int tdm_destroy(tdm_handle* h);

/* Enqueue subsequent work on `cuda_stream` (a cudaStream_t cast to void*).  NULL is
 * the CUDA legacy default stream, exactly as in the runtime API; TDM_OWN_STREAM goes
 * back to the private non-blocking stream every handle is created with. */
#define TDM_OWN_STREAM ((void*)(intptr_t)-1)
int tdm_set_stream(tdm_handle* h, void* cuda_stream);

/* Symbols a call with `count` input samples can emit at most, per channel:
 * the row stride (in elements) callers must allocate for each output. */
int64_t tdm_max_symbols(const tdm_handle* h, int64_t count);

/* PI4DQPSK::process -> DQPSKSymbolExtractor::process -> BitUnpacker::process
 * (src/dsp/pi4dqpsk.cpp:132-140, src/dsp/dqpsk_sym_extr.cpp:4-55,
 * src/dsp/bit_unpacker.cpp:4-10) for all channels, state carried across calls
 * exactly as the reference's members carry it.
 *   iq       : [C][in_stride] interleaved float32 (re,im); channel c starts at
 *              iq + 2*c*in_stride floats; `count` samples are consumed per channel
 *   syms     : [C][out_stride] float32 pairs     (needs TDM_OUT_SYMBOLS) or NULL
 *   dibits   : [C][out_stride] bytes, values 0..3 (needs TDM_OUT_DIBITS) or NULL
 *   bits     : [C][2*out_stride] bytes, values 0/1 (needs TDM_OUT_BITS)  or NULL
 *   out_counts : [C] int32, symbols written per channel this call
 *   out_stride >= tdm_max_symbols(h, count)
 * With TDM_MEM_DEVICE the call is asynchronous on the handle's stream; with
 * TDM_MEM_HOST it returns after the results are in the caller's buffers. */
int tdm_process(tdm_handle* h, const float* iq, int64_t in_stride, int32_t count,
                float* syms, uint8_t* dibits, uint8_t* bits, int64_t out_stride,
                int32_t* out_counts, uint32_t out_flags, int32_t mem_kind);

/* tdm_process with its arguments in a struct, plus the packed output (TDM_OUT_PACKED):
 *   packed : [C][packed_stride] bytes, packed_stride >= tdm_max_symbols(count) / 4 (tdm_max_symbols is a multiple of 16);
 *            byte j of a row holds symbols 4j .. 4j+3 of this call.
 *   sample_stride : 0 or 1: the samples of a channel are contiguous (channel-major rows, as above).  s > 1 (TDM_MEM_DEVICE
 *            only): sample n of channel c is the float pair at iq + 2 * (c * in_stride + n * s) -- with in_stride = 1 and
 *            s = the number of channels this is INSTANT-major input [sample][channel], what tdm_chan_process_instant_major
 *            (tdm_chan_b200.h) leaves: the front-end channeliser then needs no transposing pass, and a warp's 32
 *            channels read one 256-byte row per sample.  (This member was `reserved`, must-be-zero, before: old
 *            callers are unaffected.)
 * Unused members must be zero. */
typedef struct tdm_io {
    const float* iq; int64_t in_stride; int32_t count; int32_t mem_kind;
    float* syms; uint8_t* dibits; uint8_t* bits; uint8_t* packed;
    int64_t out_stride, packed_stride;
    int32_t* out_counts;
    uint32_t out_flags; uint32_t sample_stride;
} tdm_io;
int tdm_process_io(tdm_handle* h, const tdm_io* io);

/* ONE long capture of ONE channel, demodulated as up to n_channels overlapping time segments in parallel (SURVEY.md
 * section 8f rank 4; BASELINE.json configs[1] "1 channel, 1e9 complex samples").  The chain is a recurrence in
 * time and the reference can only walk it sample by sample (PI4DQPSK::process, src/dsp/pi4dqpsk.cpp:132-140);
 * here segment c covers samples [c L, (c+1) L + warmup), the first `warmup` samples of every segment but the
 * first only let its loops converge, and the dibit streams are joined where 128 consecutive dibits agree
 * (a segment that has not converged in time is demodulated again as the sequential continuation of its
 * predecessor).  Segment 0 starts from the state the previous tdm_process_long left, so consecutive calls
 * continue one stream.
 *
 * CONTRACT: decoded dibits only.  They equal the sequential chain's (tdm_process on one channel, hence the
 * reference's) from the point where the sequential chain has locked; before that, and in the float loop states,
 * the two may differ; across a stretch without signal (neither run locks) the streams are joined at the nominal
 * place and may gain or lose a symbol (info->n_forced).  iq: [n_samples] interleaved float pairs; dibits: [dibits_cap] bytes, info->n_dibits
 * written; n_samples / 2 + 64 is always enough.  Uses the handle's per-channel states as scratch: do not mix with
 * tdm_process on the same handle. */
typedef struct tdm_long_info {
    int64_t n_dibits;          /* dibits written                                                  */
    int32_t n_segments;        /* segments actually used (fewer for short captures)              */
    int32_t n_rerun;           /* segments whose join was not found at the first attempt          */
    int32_t segment_samples;   /* L                                                               */
    int32_t warmup;            /* W actually used (0 when a single segment was enough)            */
    int32_t n_forced;          /* segments joined at the nominal place: predecessor not locked there (no signal),
                                  or its continuation contradicted by two independent later runs             */
    int32_t n_extended;        /* of n_rerun, segments that joined after their predecessor ran on a little (no redo) */
} tdm_long_info;
int tdm_process_long(tdm_handle* h, const float* iq, int64_t n_samples, int32_t warmup, uint8_t* dibits, int64_t dibits_cap,
                     tdm_long_info* info, int32_t mem_kind);
/* The same for n_channels long captures at once (BASELINE.json configs[2], "256 channels x 4e6 samples"): every channel
 * is cut into floor(handle rows / n_channels) segments, so a few hundred channels fill the GPU like a few thousand do.
 *   iq         : [n_channels][in_stride] float pairs, n_samples used per channel
 *   dibits     : [n_channels][out_stride] bytes; out_counts [n_channels] int64 dibits written per channel
 *                (host or device pointers according to mem_kind); n_samples / 2 + 64 is always a sufficient out_stride
 *   info       : n_dibits = sum over channels; n_segments = segments per channel
 * Carried state is per channel; calling with a different n_channels than last time starts every channel from reset. */
int tdm_process_long_batch(tdm_handle* h, const float* iq, int64_t in_stride, int64_t n_samples, int32_t n_channels, int32_t warmup,
                           uint8_t* dibits, int64_t out_stride, int64_t* out_counts, tdm_long_info* info, int32_t mem_kind);

/* PI4DQPSK::reset (src/dsp/pi4dqpsk.cpp:120-130): loop scalars back to their
 * initial values, RRC history cleared, FLL/timing histories KEPT (fll.cpp:120-127,
 * complex_fd.cpp:78-87).  tdm_reset_all additionally clears every history and
 * the slicer, i.e. returns the handle to its just-created state. */
int tdm_reset(tdm_handle* h);
int tdm_reset_all(tdm_handle* h);

/* Checkpoint / resume: [C] tdm_channel_state in host memory. */
int tdm_get_state(tdm_handle* h, tdm_channel_state* host_states, int32_t n_channels);
int tdm_set_state(tdm_handle* h, const tdm_channel_state* host_states, int32_t n_channels);

/* DQPSKSymbolExtractor::sync / ::standarderr for every channel (src/dsp/dqpsk_sym_extr.h:35-36). */
int tdm_get_metrics(tdm_handle* h, tdm_metrics* host_metrics, int32_t n_channels);

/* The setters of PI4DQPSK (src/dsp/pi4dqpsk.h:52-63), with the reference's own (partial) effects
 * (src/dsp/pi4dqpsk.cpp:31-118).  `cfg` is the complete new configuration, `what` says which setter ran:
 *   TDM_SET_RATES        setSymbolrate / setSamplerate: new RRC taps AND the timing loop restarted as
 *                        COMPLEX_FD::setOmega does (offset 0, mu 0, omega = samplerate/symbolrate, limits from
 *                        the current omega_rel_limit; complex_fd.cpp:31-42).  The band-edge filters of the FLL
 *                        are NOT redesigned -- the reference never calls fll.setSymbolrate/setSamplerate
 *                        from these setters (pi4dqpsk.cpp:31-55)
 *   TDM_SET_RRC          setRRCParams / setRRCTapCount / setRRCBeta: new RRC taps only (pi4dqpsk.cpp:57-75)
 *   TDM_SET_AGC_RATE     setAGCRate (agc.setRate)
 *   TDM_SET_COSTAS_BW    setCostasBandwidth (PLL::setBandwidth: alpha, beta)
 *   TDM_SET_FLL_BW       setFllBandwidth (fll.cpp setBandwidth: beta; alpha stays 0)
 *   TDM_SET_TIMING_GAINS setMMParams / setOmegaGain / setMuGain / setOmegaRelLimit: loop gains and the omega limits
 *                        omega (1 -+ omega_rel_limit) (complex_fd.cpp:44-61); the loop's state is not touched
 * Loop state other than the timing restart above is left alone.  tdm_set_config = everything at once, as a fresh
 * init() would design it (band-edge filters included), with the timing restart when the rates changed. */
#define TDM_SET_RATES 1u
#define TDM_SET_RRC 2u
#define TDM_SET_AGC_RATE 4u
#define TDM_SET_COSTAS_BW 8u
#define TDM_SET_FLL_BW 16u
#define TDM_SET_TIMING_GAINS 32u
int tdm_set_params(tdm_handle* h, const tdm_config* cfg, uint32_t what);
int tdm_set_config(tdm_handle* h, const tdm_config* cfg);
int tdm_get_design(const tdm_handle* h, tdm_design* out);

/* Kernel-variant control, for benchmarking / profiling only (0 = auto). */
int tdm_set_kernel_variant(tdm_handle* h, int32_t variant);
/* Device time in ms of the most recent demod kernel launch sequence of this
 * handle, measured with CUDA events on the handle's stream (sync'ing). */
int tdm_last_kernel_ms(tdm_handle* h, float* ms);
/* How many kernels this library has launched through this handle so far. */
int64_t tdm_launch_count(const tdm_handle* h);

/* Pack dibits 4-per-byte (first symbol in the two most significant bits) for
 * the multi-GPU gather: in [C][in_stride] bytes, out [C][out_stride] bytes,
 * counts [C].  Device pointers; asynchronous on the handle's stream. */
int tdm_pack_dibits(tdm_handle* h, const uint8_t* dibits, int64_t in_stride, const int32_t* counts,
                    uint8_t* packed, int64_t out_stride);

/* The inverse, on the receiving side of the gather: packed rows -> one dibit per byte (DQPSKSymbolExtractor's
 * stream, dibits may be NULL) and/or one bit per byte, high bit first (BitUnpacker::process,
 * src/dsp/bit_unpacker.cpp:4-10: the stream the plugin's network sink sends, src/main.cpp:385-389; bits may be
 * NULL).  packed [n_rows][in_stride], counts [n_rows] symbols per row, dibits [n_rows][dibit_stride], bits
 * [n_rows][bit_stride], max_symbols = upper bound of counts.  n_rows need not be the handle's channel count
 * (rank 0 unpacks every rank's rows).  Device pointers; asynchronous on the handle's stream. */
int tdm_unpack_dibits(tdm_handle* h, const uint8_t* packed, int64_t in_stride, const int32_t* counts, int32_t n_rows,
                      uint8_t* dibits, int64_t dibit_stride, uint8_t* bits, int64_t bit_stride, int64_t max_symbols);

/* ---- multi-GPU epilogue: gather of the decoded symbol streams over NCCL (BASELINE.json configs[3]) ----
 * Channels are sharded by contiguous, equal-sized channel blocks, one handle per GPU; nothing is exchanged while
 * demodulating.  Each rank's TDM_OUT_PACKED rows + symbol counts are gathered to rank `dst`, rank-major (= channel-
 * major).  NCCL (libnccl.so.2, 2.x) is loaded at run time; without it these four calls return TDM_ERR_UNSUPPORTED
 * and everything else works.  The reference has no counterpart: it runs one plugin instance per channel in one
 * process (src/main.cpp:51) and its only network output is the UDP symbol sink (src/main.cpp:174,385-389), which
 * rank `dst` can feed after tdm_unpack_dibits.
 *   tdm_comm_unique_id : rank 0 makes the 128-byte id (ncclGetUniqueId) and ships it to the other ranks by any means
 *   tdm_comm_create    : ncclCommInitRank on `device` (collective: every rank calls it)
 *   tdm_comm_adopt     : wrap an ncclComm_t the host application already has (not destroyed by tdm_comm_destroy)
 *   tdm_gather_packed  : packed [n_rows][packed_stride] + counts [n_rows] of every rank -> packed_all
 *                        [world * n_rows][packed_stride], counts_all [world * n_rows] on rank dst (NULL elsewhere);
 *                        device pointers, asynchronous on cuda_stream (a cudaStream_t; give a stream other than the
 *                        handle's to overlap the gather of one call with the demodulation of the next). */
#define TDM_COMM_ID_BYTES 128
typedef struct tdm_comm tdm_comm;
int tdm_comm_unique_id(uint8_t id[TDM_COMM_ID_BYTES]);
int tdm_comm_create(const uint8_t id[TDM_COMM_ID_BYTES], int32_t rank, int32_t world, int32_t device, tdm_comm** out);
int tdm_comm_adopt(void* nccl_comm, int32_t rank, int32_t world, int32_t device, tdm_comm** out);
int tdm_comm_destroy(tdm_comm* c);
int tdm_gather_packed(tdm_comm* c, int32_t dst, int32_t n_rows, const uint8_t* packed, int64_t packed_stride, const int32_t* counts,
                      uint8_t* packed_all, int32_t* counts_all, void* cuda_stream);

/* Deterministic synthetic TETRA-mapped pi/4-DQPSK capture, generated on the
 * device: channel c uses data seed seed_data+c and noise seed
 * seed_noise+c.  iq_dev: [C][stride] device floats pairs.  tx_dibits_dev
 * ([C][n/2+64] bytes, may be NULL) receives the transmitted dibits; iq_dev may be
 * NULL when only those are wanted.
 * Test/bench signal source; not part of the reference's surface. */
typedef struct tdm_synth_params {
    double snr_db;          /* Es/N0 in dB                          */
    double max_freq_off_hz; /* per-channel offset drawn from U(-x, x) */
    double min_amp, max_amp;/* amplitude drawn log-uniformly          */
    uint64_t seed_data, seed_noise;
} tdm_synth_params;
int tdm_synth_capture(int32_t device, void* cuda_stream, const tdm_synth_params* p, int32_t n_channels,
                      int64_t n_samples, int64_t stride, int32_t first_channel, float* iq_dev,
                      uint8_t* tx_dibits_dev, int64_t tx_stride);

const char* tdm_last_error(void);
int tdm_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TDM_B200_H */
