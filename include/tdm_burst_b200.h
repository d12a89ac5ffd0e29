/* tdm_burst_b200.h -- C ABI of the burst synchroniser that follows the demodulator
 * ("burst sync / training-sequence search on GPU over
 * the decoded bit streams").  Same library as tdm_b200.h (libtdm_b200.so).
 *
 * What it stands in for, per channel (paths relative to the reference tree):
 *   - tetra_find_train_seq()   src/decoder/src/phy/tetra_burst.c:271-341
 *   - tetra_burst_sync_in()    src/decoder/src/phy/tetra_burst_sync.c:54-155
 *     (the UNLOCKED -> KNOW_FSTART -> LOCKED state machine; the 510-bit bursts it
 *     hands to tetra_burst_rx_cb(), src/decoder/src/phy/tetra_burst.c:343-393)
 *   - tetra_tdma_time_add_tn() src/decoder/src/tetra_tdma.c:70-74 (the slot counter
 *     tetra_burst_sync_in advances once per received slot; the reference keeps it in
 *     the process-global t_phy_state, src/decoder/src/phy/tetra_burst_sync.c:34 --
 *     here every channel has its own)
 *   - the "training sequence seen" detector of the plugin's network mode,
 *     TetraDemodulatorModule::_demodSinkHandler, src/main.cpp:385-414
 *
 * Everything downstream of the burst callback (lower MAC, Viterbi, PDU parsing,
 * codec) stays out of scope; a burst record carries what tetra_burst_rx_cb gets.
 *
 * Integer / byte work throughout: results are bit-exact against the reference's
 * own C files compiled unmodified (oracle/_ref/libtetra_bsync_ref.so).
 */
#ifndef TDM_BURST_B200_H
#define TDM_BURST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDM_BITS_PER_TS 510          /* TETRA_BITS_PER_TS, src/decoder/src/tetra_common.h:238   */
#define TDM_BSYNC_BITBUF 4096        /* sizeof(tetra_rx_state::bitbuf), phy/tetra_burst_sync.h:15 */
/* Largest `len` of one emulated tetra_burst_sync_in call: one slot.  The state machine consumes at most one
 * slot per call (tetra_burst_sync.c:106-150), so longer calls only let the buffer run full; worse, the reference
 * is then undefined: make_bitbuf_space() (tetra_burst_sync.c:38-51) can push the buffer start past
 * next_frame_start_bitnum, and the memmove at :97 is called with a negative offset (observed: abort with
 * len = 2048).  Up to one slot per call this cannot happen except through the look-ahead quirk described at
 * tdm_find_train_seq; for that corner the offset is defined as 0 here. */
#define TDM_BSYNC_MAX_CALL_BITS 510

/* enum tetra_train_seq, src/decoder/src/phy/tetra_burst.h:27-33 */
#define TDM_TRAIN_NORM_1 0
#define TDM_TRAIN_NORM_2 1
#define TDM_TRAIN_NORM_3 2
#define TDM_TRAIN_SYNC   3
#define TDM_TRAIN_EXT    4

/* enum rx_state, src/decoder/src/phy/tetra_burst_sync.h:6-10 */
#define TDM_RX_S_UNLOCKED    0
#define TDM_RX_S_KNOW_FSTART 1
#define TDM_RX_S_LOCKED      2

/* What one input byte holds. */
#define TDM_BSYNC_IN_BITS   0        /* one bit per byte, BitUnpacker `out` (src/dsp/bit_unpacker.cpp:4-10)      */
#define TDM_BSYNC_IN_DIBITS 1        /* one dibit per byte, DQPSKSymbolExtractor `out`; unpacked MSB first on the fly */

/* struct tetra_rx_state without its bit buffer (phy/tetra_burst_sync.h:12-20), the slot
 * counter, and the detector state of src/main.cpp:470-472.  Checkpoint format of
 * tdm_bsync_get_state / tdm_bsync_set_state together with the carried bits. */
typedef struct tdm_bsync_state {
    int32_t  state;                     /* enum rx_state                                          */
    uint32_t bits_in_buf;
    uint32_t bitbuf_start_bitnum;
    uint32_t next_frame_start_bitnum;
    uint32_t tn, fn, mn;                /* t_phy_state.time (struct tetra_tdma_time)              */
    uint32_t ts_found;                  /* tsfound,          src/main.cpp:471                     */
    uint32_t ts_expire;                 /* symsbeforeexpire, src/main.cpp:472                     */
    uint32_t ts_window_lo, ts_window_hi;/* the newest 44 bits seen by the detector (tsfind_buffer[1..44]), newest in bit 0 */
    uint32_t searched_upto;             /* internal: first stream position not yet known to be free of a SYNC match */
    uint64_t n_bits;                    /* lifetime input bits                                    */
    uint64_t n_bursts;                  /* lifetime bursts delivered                              */
    uint32_t bitbuf[TDM_BSYNC_BITBUF / 32]; /* the buffered bits, packed MSB first, bit 0 of the buffer in bit 31 of word 0 */
} tdm_bsync_state;

/* One call of tetra_burst_rx_cb(burst, 510, type, priv) (phy/tetra_burst.c:343). */
typedef struct tdm_burst {
    uint32_t bitnum;                    /* stream position of the burst's first bit (bitbuf_start_bitnum)  */
    int32_t  train_seq;                 /* TDM_TRAIN_SYNC / TDM_TRAIN_NORM_1 / TDM_TRAIN_NORM_2            */
    uint32_t tn, fn, mn;                /* t_phy_state.time when the callback ran                          */
    uint32_t call_index;                /* which emulated tetra_burst_sync_in call of this tdm_bsync_in delivered it */
    uint32_t reserved[2];
    uint32_t bits[16];                  /* the 510 burst bits packed MSB first: bit i = (bits[i/32] >> (31 - i%32)) & 1;
                                           bits 510 and 511 are 0.  96-byte records: a byte per bit would make the
                                           burst records the largest stream of the whole stage (1.07 B written per
                                           input bit against 0.5 B read); tdm_burst_unpack gives the byte form. */
} tdm_burst;

typedef struct tdm_bsync tdm_bsync;

/* n_channels independent tetra_rx_state objects, zero-initialised like the reference's
 * talloc_zero (src/dsp/osmotetra_dec.h).  max_units = most input bytes per channel a
 * later tdm_bsync_in passes.  At most 65535 channels per handle (TDM_ERR_UNSUPPORTED
 * above; same limit per call of tdm_find_train_seq). */
int tdm_bsync_create(int32_t n_channels, int64_t max_units, int32_t device, tdm_bsync** out);
int tdm_bsync_destroy(tdm_bsync* h);
/* cudaStream_t as void*; NULL = legacy default stream, TDM_OWN_STREAM = the handle's own. */
int tdm_bsync_set_stream(tdm_bsync* h, void* cuda_stream);
int tdm_bsync_reset(tdm_bsync* h);

/* Feed every channel's new bits through tetra_burst_sync_in, `call_bits` bits per
 * emulated call (the last call of a channel takes what is left; a channel with no
 * input makes no call).  1 <= call_bits <= TDM_BSYNC_MAX_CALL_BITS.
 *   in        : [C][in_stride] bytes (in_kind says what a byte holds)
 *   n_units   : [C] int32 input bytes per channel, or NULL = `units_all` for every channel
 *               (device pointer with TDM_MEM_DEVICE, host pointer with TDM_MEM_HOST)
 *   bursts    : [C][max_bursts] records; n_bursts [C] = bursts delivered this call (records
 *               beyond max_bursts are counted but not stored)
 *   detect_ts : also run the src/main.cpp:385-414 detector over the new bits
 * With TDM_MEM_DEVICE the call is asynchronous on the handle's stream. */
int tdm_bsync_in(tdm_bsync* h, const uint8_t* in, int64_t in_stride, const int32_t* n_units, int32_t units_all,
                 int32_t in_kind, int32_t call_bits, tdm_burst* bursts, int32_t max_bursts, int32_t* n_bursts,
                 int32_t detect_ts, int32_t mem_kind);

int tdm_bsync_get_state(tdm_bsync* h, tdm_bsync_state* host_states, int32_t n_channels);
int tdm_bsync_set_state(tdm_bsync* h, const tdm_bsync_state* host_states, int32_t n_channels);
int64_t tdm_bsync_launch_count(const tdm_bsync* h);
/* Device time in ms of the three kernels of the most recent tdm_bsync_in (pack, detect, sync), measured with CUDA
 * events on the handle's stream (synchronises).  Benchmarking / profiling only. */
int tdm_bsync_last_kernel_ms(tdm_bsync* h, float* ms3);

/* tetra_find_train_seq for C independent buffers of one bit per byte, including its look-ahead quirk: the
 * 22-bit pre-filter (tetra_burst.c:289-307) is preloaded one bit short, so for the first 21 positions it does
 * not hold in[i .. i+21] and a sequence starting there is normally NOT reported (oracle/oracle_bsync.c spells
 * out what it holds instead).
 * out_type[c] = the enum value found first in in[c][0 .. end_of_in) among the sequences
 * enabled in mask_of_train_seq (bit 1 << TDM_TRAIN_x), or -1; out_offset[c] = its offset. */
int tdm_find_train_seq(int32_t device, void* cuda_stream, const uint8_t* in, int64_t in_stride, int32_t n_channels,
                       uint32_t end_of_in, uint32_t mask_of_train_seq, int32_t* out_type, uint32_t* out_offset,
                       int32_t mem_kind);

/* tetra_burst_rx_cb's split of a burst into the blocks it passes to tp_sap_udata_ind
 * (phy/tetra_burst.c:343-393, offsets :33-49).  Host-side helper, no GPU involved.
 * blocks: up to 3 entries of {tp_sap_data_type, blk_num, n_bits, bits[432]}; returns the count. */
typedef struct tdm_tp_sap_block {
    int32_t type;                       /* enum tp_sap_data_type, phy/tetra_burst.h:9-16 */
    int32_t blk_num;
    int32_t n_bits;
    uint8_t bits[432];
} tdm_tp_sap_block;
int tdm_burst_demux(const tdm_burst* burst, tdm_tp_sap_block* blocks);
/* The `burst` argument of tetra_burst_rx_cb: 510 bytes, one bit each.  Host-side helper. */
int tdm_burst_unpack(const tdm_burst* burst, uint8_t* bits510);

#ifdef __cplusplus
}
#endif
#endif /* TDM_BURST_B200_H */
