/* TEST INFRASTRUCTURE -- C driver around the reference's OWN burst synchroniser, compiled
 * unmodified from /root/reference/src/decoder/src/{phy/tetra_burst.c, phy/tetra_burst_sync.c,
 * tetra_tdma.c} where they lie (oracle/Makefile target `ref_bsync`; nothing is copied).
 * It is the authority for the burst-sync stage (include/tdm_burst_b200.h).  Only tests/
 * and the CPU-baseline legs of the bench scripts may load oracle/_ref/libtetra_bsync_ref.so.
 *
 * The reference hands finished bursts to the lower MAC through tp_sap_udata_ind()
 * (phy/tetra_burst.c:343-393); the lower MAC is out of scope, so this driver provides that
 * one symbol and records what arrives.  t_phy_state and the rx state are process-global /
 * single-instance in the reference: this driver runs ONE channel at a time.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <tetra_common.h>
#include <phy/tetra_burst.h>
#include <phy/tetra_burst_sync.h>
#include <tetra_tdma.h>

typedef struct rbs_burst {
    uint32_t bitnum;
    int32_t train_seq;          /* inferred from the first TP-SAP block: see rbs_infer_type */
    uint32_t tn, fn, mn;
    uint32_t call_index;
    uint32_t n_blocks;          /* tp_sap_udata_ind calls this burst produced */
    uint32_t reserved;
    uint8_t bits[512];
} rbs_burst;

typedef struct rbs_block {
    int32_t type, blk_num, n_bits;
    uint8_t bits[432];
} rbs_block;

typedef struct rbs_ctx {
    struct tetra_rx_state trs;
    struct tetra_mac_state tms;
    struct tetra_display_state disp;
    rbs_burst* bursts;
    rbs_block* blocks;          /* 3 per burst slot */
    uint32_t max_bursts, n_bursts;
    uint32_t call_index;
    int in_burst;               /* a tp_sap_udata_ind of the current tetra_burst_sync_in call has been seen */
} rbs_ctx;

static rbs_ctx* g_ctx;          /* the reference's callback has no way to reach us but `priv` = tms */

void tp_sap_udata_ind(enum tp_sap_data_type type, int blk_num, const uint8_t* bits, unsigned int len, void* priv)
{
    rbs_ctx* c = g_ctx;
    (void)priv;
    if (!c) { return; }
    if (!c->in_burst) {
        c->in_burst = 1;
        if (c->n_bursts < c->max_bursts) {
            rbs_burst* b = &c->bursts[c->n_bursts];
            memset(b, 0, sizeof(*b));
            b->bitnum = c->trs.bitbuf_start_bitnum;
            b->tn = t_phy_state.time.tn; b->fn = t_phy_state.time.fn; b->mn = t_phy_state.time.mn;
            b->call_index = c->call_index;
            memcpy(b->bits, c->trs.bitbuf, TETRA_BITS_PER_TS);
            /* first block tells the burst type: SYNC starts with SB1; NORM_2 sends BBK then NDB; NORM_1 BBK then SCH_F */
            b->train_seq = (type == TPSAP_T_SB1) ? TETRA_TRAIN_SYNC : -1;
        }
        c->n_bursts++;
    }
    if (c->n_bursts <= c->max_bursts) {
        rbs_burst* b = &c->bursts[c->n_bursts - 1];
        if (b->n_blocks < 3) {
            rbs_block* k = &c->blocks[3 * (c->n_bursts - 1) + b->n_blocks];
            k->type = (int32_t)type; k->blk_num = blk_num; k->n_bits = (int32_t)len;
            memset(k->bits, 0, sizeof(k->bits));
            memcpy(k->bits, bits, len <= sizeof(k->bits) ? len : sizeof(k->bits));
            if (b->train_seq < 0 && b->n_blocks == 1) {
                b->train_seq = (type == TPSAP_T_SCH_F) ? TETRA_TRAIN_NORM_1 : TETRA_TRAIN_NORM_2;
            }
            b->n_blocks++;
        }
    }
}

rbs_ctx* rbs_new(void)
{
    rbs_ctx* c = (rbs_ctx*)calloc(1, sizeof(rbs_ctx));
    c->tms.t_display_st = &c->disp;
    c->trs.burst_cb_priv = &c->tms;
    memset(&t_phy_state, 0, sizeof(t_phy_state));
    return c;
}

void rbs_free(rbs_ctx* c) { if (g_ctx == c) { g_ctx = NULL; } free(c); }

/* Feed n_bits bits through tetra_burst_sync_in, call_bits per call.  Returns bursts delivered. */
int rbs_in(rbs_ctx* c, const uint8_t* bits, uint32_t n_bits, uint32_t call_bits, rbs_burst* bursts, rbs_block* blocks,
           uint32_t max_bursts)
{
    c->bursts = bursts; c->blocks = blocks; c->max_bursts = max_bursts; c->n_bursts = 0; c->call_index = 0;
    g_ctx = c;
    for (uint32_t off = 0; off < n_bits; off += call_bits, c->call_index++) {
        uint32_t len = n_bits - off < call_bits ? n_bits - off : call_bits;
        c->in_burst = 0;
        tetra_burst_sync_in(&c->trs, (uint8_t*)(bits + off), len);
    }
    g_ctx = NULL;
    return (int)c->n_bursts;
}

/* time is process-global in the reference: save / restore it so several contexts can be interleaved */
void rbs_get_time(uint32_t* tn, uint32_t* fn, uint32_t* mn) { *tn = t_phy_state.time.tn; *fn = t_phy_state.time.fn; *mn = t_phy_state.time.mn; }
void rbs_set_time(uint32_t tn, uint32_t fn, uint32_t mn) { t_phy_state.time.tn = tn; t_phy_state.time.fn = fn; t_phy_state.time.mn = mn; }

void rbs_get_state(const rbs_ctx* c, int32_t* state, uint32_t* bits_in_buf, uint32_t* start_bitnum, uint32_t* next_frame, uint8_t* bitbuf)
{
    *state = (int32_t)c->trs.state; *bits_in_buf = c->trs.bits_in_buf; *start_bitnum = c->trs.bitbuf_start_bitnum;
    *next_frame = c->trs.next_frame_start_bitnum;
    if (bitbuf) { memcpy(bitbuf, c->trs.bitbuf, c->trs.bits_in_buf); }
}

/* tetra_find_train_seq on a caller buffer.  The reference's 22-bit look-ahead filter reads up to 21
 * bytes past end_of_in (phy/tetra_burst.c:295,300); the caller must own that much. */
int rbs_find_train_seq(const uint8_t* in, uint32_t end_of_in, uint32_t mask, uint32_t* offset)
{
    unsigned int off = 0;
    int rc = tetra_find_train_seq(in, end_of_in, mask, &off);
    *offset = off;
    return rc;
}

/* The reference's own burst builders (phy/tetra_burst.c:171-269): signal source for protocol-valid test streams. */
int rbs_build_sync_burst(uint8_t* buf, const uint8_t* sb, const uint8_t* bb, const uint8_t* bkn) { return build_sync_c_d_burst(buf, sb, bb, bkn); }
int rbs_build_norm_burst(uint8_t* buf, const uint8_t* bkn1, const uint8_t* bb, const uint8_t* bkn2, int two_log_chan) { return build_norm_c_d_burst(buf, bkn1, bb, bkn2, two_log_chan); }
