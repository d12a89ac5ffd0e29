/* TEST INFRASTRUCTURE -- "Oracle B": canonical-operation-order CPU restatement of
 * the reference demodulation chain.  See oracle_b.c for the contract.  Only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it. */
#ifndef ORACLE_B_H
#define ORACLE_B_H
#include <stdint.h>
#include "tdm_b200.h" /* tdm_config / tdm_design / tdm_channel_state layouts (public ABI structs only) */

#ifdef __cplusplus
extern "C" {
#endif

void ob_default_config(tdm_config* cfg);
int ob_design(const tdm_config* cfg, tdm_design* d);
void ob_state_init(const tdm_design* d, tdm_channel_state* s);
void ob_sincos(float x, float* s, float* c);
void ob_fll_nco(float phi, float f_old, float f_new, float* s, float* c);
long ob_fll_fallback_count(void);

/* One channel.  syms/dibits/bits may be NULL.  Returns symbols emitted. */
int64_t ob_process(const tdm_design* d, tdm_channel_state* s, const float* iq, int64_t count,
                   float* syms, uint8_t* dibits, uint8_t* bits);

/* nch channels laid out [nch][in_stride] / [nch][out_stride], one state each,
 * spread over nthreads pthreads. */
void ob_process_multi(const tdm_design* d, tdm_channel_state* states, int nch, const float* iq,
                      int64_t in_stride, int64_t count, float* syms, uint8_t* dibits, uint8_t* bits,
                      int64_t out_stride, int32_t* out_counts, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
