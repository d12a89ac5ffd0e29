/* TEST INFRASTRUCTURE -- CPU restatement of the reference's burst synchroniser, for the
 * burst-sync stage of include/tdm_burst_b200.h.  New code, written from the behaviour of
 *   tetra_find_train_seq      /root/reference/src/decoder/src/phy/tetra_burst.c:271-341
 *   make_bitbuf_space         /root/reference/src/decoder/src/phy/tetra_burst_sync.c:38-51
 *   tetra_burst_sync_in       /root/reference/src/decoder/src/phy/tetra_burst_sync.c:54-155
 *   tetra_tdma_time_add_tn    /root/reference/src/decoder/src/tetra_tdma.c:44-74
 *   tetra_burst_rx_cb         /root/reference/src/decoder/src/phy/tetra_burst.c:343-393 (block split)
 *   _demodSinkHandler         /root/reference/src/main.cpp:385-414 (training-sequence detector)
 * Pinned against the reference's own C files compiled unmodified (oracle/_ref/libtetra_bsync_ref.so,
 * tests/test_bsync_oracle.py) for everything but the detector, which lives in the plugin's C++
 * module class and cannot be compiled here: that one function is "parity unpinned" by the
 * reference and checked only against its restatement.
 * Only tests/, smoke() and the CPU-baseline legs of the bench scripts may use this file.
 */
#include <stdint.h>
#include <string.h>

#include "tdm_burst_b200.h"

/* the checkers keep bursts the way the reference passes them around: one bit per byte */
typedef struct obs_burst {
    uint32_t bitnum;
    int32_t train_seq;
    uint32_t tn, fn, mn;
    uint32_t call_index;
    uint32_t reserved[2];
    uint8_t bits[512];
} obs_burst;

/* ETSI EN 300 392-2 9.4.4.3.2-4 training sequences, as the reference tabulates them
 * (phy/tetra_burst.c:61-72, src/main.cpp:457-468). */
static const uint8_t kSeq_n[22] = { 1,1, 0,1, 0,0, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,1, 0,0 };
static const uint8_t kSeq_p[22] = { 0,1, 1,1, 1,0, 1,0, 0,1, 0,0, 0,0, 1,1, 0,1, 1,1, 1,0 };
static const uint8_t kSeq_q[22] = { 1,0, 1,1, 0,1, 1,1, 0,0, 0,0, 0,1, 1,0, 1,0, 1,1, 0,1 };
static const uint8_t kSeq_N[33] = { 1,1,1, 0,0,1, 1,0,1, 1,1,1, 0,0,0, 1,1,1, 1,0,0, 0,1,1, 1,1,0, 0,0,0, 0,0,0 };
static const uint8_t kSeq_P[33] = { 1,0,1, 0,1,1, 1,1,1, 1,0,1, 0,1,0, 1,0,1, 1,1,0, 0,0,1, 1,0,0, 0,1,0, 0,1,0 };
static const uint8_t kSeq_x[30] = { 1,0, 0,1, 1,1, 0,1, 0,0, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,1, 0,0, 0,0, 1,1 };
static const uint8_t kSeq_X[45] = { 0,1,1,1,0,0,1,1,0,1,0,0,0,0,1,0,0,0,1,1,1,0,1,1,0,1,0,1,0,1,1,1,1,1,0,1,0,0,0,0,0,1,1,1,0 };
static const uint8_t kSeq_y[38] = { 1,1, 0,0, 0,0, 0,1, 1,0, 0,1, 1,1, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,0, 0,0, 0,1, 1,0, 0,1, 1,1 };

/* First position i of in[0..end) that passes the reference's 22-bit look-ahead filter AND holds an enabled
 * sequence completely inside the buffer; at one position the order of the tests is SYNC, NORM_1, NORM_2,
 * NORM_3, EXT (tetra_burst.c:309-338).
 *
 * The filter (tetra_burst.c:289-307) is a shift register that should hold in[i .. i+21] and be compared with
 * the first 22 bits of y, n, p, q, x.  It is preloaded with only 20 bits (`i < FILTER_LOOKAHEAD_LEN-2`) and
 * then fed in[i+21], so in[20] never enters it: for i >= 21 it holds in[i .. i+21] as intended (and the test
 * is implied by the full comparison), but for i <= 20 it holds
 *      i == 0 :  0, in[0..19], in[21]
 *      i >= 1 :  in[i-1 .. 19], in[21 .. 21+i]
 * so a sequence that starts within the first 21 positions is normally NOT found.  Kept, bit for bit. */
static uint32_t seq_prefix22(const uint8_t* s)
{
    uint32_t v = 0;
    for (int i = 0; i < 22; ++i) { v = (v << 1) | s[i]; }
    return v;
}

int obs_find_train_seq(const uint8_t* in, uint32_t end, uint32_t mask, uint32_t* offset)
{
    static const struct { int id; const uint8_t* s; uint32_t n; } order[5] = {
        { TDM_TRAIN_SYNC, kSeq_y, 38 }, { TDM_TRAIN_NORM_1, kSeq_n, 22 }, { TDM_TRAIN_NORM_2, kSeq_p, 22 },
        { TDM_TRAIN_NORM_3, kSeq_q, 22 }, { TDM_TRAIN_EXT, kSeq_x, 30 } };
    for (uint32_t i = 0; i < end; ++i) {
        if (end - i < 22) { break; }                    /* no sequence fits any more */
        if (i <= 20) {
            uint32_t f = 0;
            if (i >= 1) { for (uint32_t k = i - 1; k <= 19; ++k) { f = (f << 1) | in[k]; } }
            else { for (uint32_t k = 0; k <= 19; ++k) { f = (f << 1) | in[k]; } }
            for (uint32_t k = 21; k <= 21 + i; ++k) { f = (f << 1) | in[k]; }
            int pass = 0;
            for (int k = 0; k < 5; ++k) { pass |= (f == seq_prefix22(order[k].s)); }
            if (!pass) { continue; }
        }
        for (int k = 0; k < 5; ++k) {
            if ((mask & (1u << order[k].id)) && end - i >= order[k].n && !memcmp(in + i, order[k].s, order[k].n)) {
                *offset = i;
                return order[k].id;
            }
        }
    }
    return -1;
}

static void time_add_tn(tdm_bsync_state* s)      /* tetra_tdma_time_add_tn(&time, 1) with its normalisation chain */
{
    s->tn += 1;
    if (s->tn > 4) { uint32_t d = s->tn / 4; s->tn %= 4; s->fn += d; }
    if (s->fn > 18) { uint32_t d = s->fn / 18; s->fn %= 18; s->mn += d; }
    if (s->mn > 60) { s->mn %= 60; }
}

static void unpack_bitbuf(const tdm_bsync_state* s, uint8_t* buf)
{
    for (uint32_t i = 0; i < s->bits_in_buf; ++i) { buf[i] = (uint8_t)((s->bitbuf[i >> 5] >> (31 - (i & 31))) & 1u); }
}
static void pack_bitbuf(tdm_bsync_state* s, const uint8_t* buf)
{
    memset(s->bitbuf, 0, sizeof(s->bitbuf));
    for (uint32_t i = 0; i < s->bits_in_buf; ++i) { s->bitbuf[i >> 5] |= (uint32_t)(buf[i] & 1u) << (31 - (i & 31)); }
}

/* One channel: n_bits new bits, call_bits per emulated tetra_burst_sync_in call. */
int obs_in(tdm_bsync_state* s, const uint8_t* bits, uint32_t n_bits, uint32_t call_bits, obs_burst* bursts, uint32_t max_bursts)
{
    uint8_t buf[TDM_BSYNC_BITBUF + 64];
    uint32_t nb = 0, call = 0;
    memset(buf, 0, sizeof(buf));
    unpack_bitbuf(s, buf);
    for (uint32_t off = 0; off < n_bits; off += call_bits, ++call) {
        const uint32_t len = n_bits - off < call_bits ? n_bits - off : call_bits;
        /* make_bitbuf_space + append */
        uint32_t space = TDM_BSYNC_BITBUF - s->bits_in_buf;
        if (space < len) {
            const uint32_t delta = len - space;
            memmove(buf, buf + delta, s->bits_in_buf - delta);
            s->bits_in_buf -= delta;
            s->bitbuf_start_bitnum += delta;
        }
        memcpy(buf + s->bits_in_buf, bits + off, len);
        s->bits_in_buf += len;
        s->n_bits += len;

        uint32_t offs = 0;
        int rc;
        if (s->state == TDM_RX_S_UNLOCKED) {
            if (s->bits_in_buf < 2 * TDM_BITS_PER_TS) { continue; }
            rc = obs_find_train_seq(buf, s->bits_in_buf, 1u << TDM_TRAIN_SYNC, &offs);
            if (rc < 0) { continue; }
            s->state = TDM_RX_S_KNOW_FSTART;
            s->next_frame_start_bitnum = s->bitbuf_start_bitnum + offs + 296;
            continue;
        }
        if (s->state == TDM_RX_S_KNOW_FSTART) {
            if (s->bitbuf_start_bitnum + s->bits_in_buf < s->next_frame_start_bitnum) { continue; }
            uint32_t shift = s->next_frame_start_bitnum - s->bitbuf_start_bitnum;
            if ((int32_t)shift < 0) { shift = 0; }      /* undefined in the reference (negative memmove offset); see tdm_burst_b200.h */
            memmove(buf, buf + shift, s->bits_in_buf - shift);
            s->bits_in_buf -= shift;
            s->bitbuf_start_bitnum += shift;
            s->next_frame_start_bitnum += TDM_BITS_PER_TS;
            s->state = TDM_RX_S_LOCKED;                 /* and straight on into the LOCKED case (no break in the reference) */
        }
        if (s->bits_in_buf < TDM_BITS_PER_TS) { continue; }
        time_add_tn(s);
        rc = obs_find_train_seq(buf, s->bits_in_buf,
                                (1u << TDM_TRAIN_NORM_1) | (1u << TDM_TRAIN_NORM_2) | (1u << TDM_TRAIN_SYNC), &offs);
        int deliver = 0;
        if (rc == TDM_TRAIN_SYNC) {
            if (offs == 214) { deliver = 1; } else { s->state = TDM_RX_S_UNLOCKED; }
        } else if (rc == TDM_TRAIN_NORM_1 || rc == TDM_TRAIN_NORM_2) {
            if (offs == 244) { deliver = 1; }
        } else {
            s->state = TDM_RX_S_UNLOCKED;
        }
        if (deliver) {
            if (nb < max_bursts) {
                obs_burst* b = &bursts[nb];
                memset(b, 0, sizeof(*b));
                b->bitnum = s->bitbuf_start_bitnum; b->train_seq = rc;
                b->tn = s->tn; b->fn = s->fn; b->mn = s->mn; b->call_index = call;
                memcpy(b->bits, buf, TDM_BITS_PER_TS);
            }
            ++nb;
            s->n_bursts++;
        }
        s->bits_in_buf -= TDM_BITS_PER_TS;
        memmove(buf, buf + TDM_BITS_PER_TS, s->bits_in_buf);
        s->bitbuf_start_bitnum += TDM_BITS_PER_TS;
        s->next_frame_start_bitnum += TDM_BITS_PER_TS;
    }
    pack_bitbuf(s, buf);
    s->searched_upto = 0;                               /* internal to the CUDA path; not compared */
    return (int)nb;
}

/* src/main.cpp:385-414: a 45-bit shift register; after each new bit the register's OLDEST end is compared
 * with the eight sequences; a hit sets tsfound and re-arms a 2048-bit expiry counter.  (tsfind_buffer is an
 * uninitialised member in the reference: defined as zeros here.) */
void obs_ts_detect(tdm_bsync_state* s, const uint8_t* bits, uint32_t n_bits)
{
    static const struct { const uint8_t* s; uint32_t n; } seqs[8] = {
        { kSeq_n, 22 }, { kSeq_p, 22 }, { kSeq_q, 22 }, { kSeq_N, 33 }, { kSeq_P, 33 }, { kSeq_x, 30 }, { kSeq_X, 45 }, { kSeq_y, 38 } };
    uint8_t w[45];
    const uint64_t hist = ((uint64_t)s->ts_window_hi << 32) | s->ts_window_lo;     /* newest bit in bit 0, 44 bits */
    w[0] = 0;
    for (int i = 0; i < 44; ++i) { w[1 + i] = (uint8_t)((hist >> (43 - i)) & 1u); }
    for (uint32_t j = 0; j < n_bits; ++j) {
        memmove(w, w + 1, 44);
        w[44] = bits[j];
        for (int k = 0; k < 8; ++k) {
            if (!memcmp(w, seqs[k].s, seqs[k].n)) { s->ts_found = 1; s->ts_expire = 2048; break; }
        }
        if (s->ts_expire > 0) {
            s->ts_expire--;
            if (s->ts_expire == 0) { s->ts_found = 0; }
        }
    }
    uint64_t h = 0;
    for (int i = 0; i < 44; ++i) { h = (h << 1) | (w[1 + i] & 1u); }
    s->ts_window_lo = (uint32_t)h; s->ts_window_hi = (uint32_t)(h >> 32);
}

/* tetra_burst_rx_cb's block split (phy/tetra_burst.c:33-49,343-393); DQPSK4_BITS_PER_SYM = 2. */
int obs_burst_demux(const obs_burst* b, tdm_tp_sap_block* out)
{
    enum { SB1 = 0, SB2 = 1, NDB = 2, BBK = 3, SCH_F = 5 };
    const uint8_t* u = b->bits;
    memset(out, 0, 3 * sizeof(*out));
    if (b->train_seq == TDM_TRAIN_SYNC) {
        out[0].type = SB1; out[0].blk_num = 1; out[0].n_bits = 120; memcpy(out[0].bits, u + 94, 120);
        out[1].type = BBK; out[1].blk_num = 0; out[1].n_bits = 30;  memcpy(out[1].bits, u + 252, 30);
        out[2].type = SB2; out[2].blk_num = 2; out[2].n_bits = 216; memcpy(out[2].bits, u + 282, 216);
        return 3;
    }
    if (b->train_seq == TDM_TRAIN_NORM_2) {
        out[0].type = BBK; out[0].blk_num = 0; out[0].n_bits = 30; memcpy(out[0].bits, u + 230, 14); memcpy(out[0].bits + 14, u + 266, 16);
        out[1].type = NDB; out[1].blk_num = 1; out[1].n_bits = 216; memcpy(out[1].bits, u + 14, 216);
        out[2].type = NDB; out[2].blk_num = 2; out[2].n_bits = 216; memcpy(out[2].bits, u + 282, 216);
        return 3;
    }
    if (b->train_seq == TDM_TRAIN_NORM_1) {
        out[0].type = BBK; out[0].blk_num = 0; out[0].n_bits = 30; memcpy(out[0].bits, u + 230, 14); memcpy(out[0].bits + 14, u + 266, 16);
        out[1].type = SCH_F; out[1].blk_num = 0; out[1].n_bits = 432; memcpy(out[1].bits, u + 14, 216); memcpy(out[1].bits + 216, u + 282, 216);
        return 2;
    }
    return 0;
}
