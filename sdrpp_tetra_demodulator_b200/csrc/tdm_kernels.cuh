// tdm_kernels.cuh -- launch interface between the C-ABI layer (tdm_api.cu) and the
// sm_100a kernels (tdm_kernels.cu, tdm_synth.cu).  Internal; the public surface is
// include/tdm_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tdm_b200.h"

#define TDM_TAP_PAD 88   /* >= 65 + 2*(T-1) for T <= 8, multiple of 4 */

namespace tdm {

// Everything a demod launch needs, passed BY VALUE as a __grid_constant__ kernel
// parameter: it then lives in the constant bank, so the FIR taps are immediate
// constant-bank operands of the FFMAs (no loads, no registers), and two handles with
// different configurations never share mutable __constant__ state.
struct DemodParams {
    // FIR taps, oldest sample first, zero-padded at the old end to TDM_MAX_TAPS.
    float be_a[TDM_MAX_TAPS];   // band-edge pair: hbe taps = a + jb, lbe taps = a - jb
    float be_b[TDM_MAX_TAPS];
    float rrc[TDM_MAX_TAPS];
    // the same three tables with T-1 leading zeros, filled per kernel variant at launch
    float tpad[3][TDM_TAP_PAD];
    float agc_rate, agc_set_point, agc_max_gain;
    float fll_beta, fll_min_freq, fll_max_freq;
    float tr_alpha, tr_beta, tr_min_omega, tr_max_omega;
    float costas_alpha, costas_beta, costas_min_freq, costas_max_freq;
    const float* bank;              // device [128][8] interpolator polyphase bank
    const float2* iq;               // [C][in_stride]
    long long in_stride;
    long long sample_stride;        // samples of a row are this many float2 apart (1: channel-major rows; M with in_stride 1:
                                    // instant-major [sample][channel], what the channeliser leaves without its transposing pass)
    long long channel_stride;       // only with rows_per_channel > 1 (see row_input() in tdm_kernels.cu)
    int rows_per_channel;           // 0 / 1: every row is a channel
    int count;                      // samples per channel this launch
    int n_channels;
    float2* syms;                   // [C][out_stride] or nullptr
    uint8_t* dibits;                // [C][out_stride] or nullptr
    uint8_t* bits;                  // [C][2*out_stride] or nullptr
    uint8_t* packed;                // [C][packed_stride] or nullptr: 4 dibits per byte, first symbol in bits 7..6
    long long out_stride;
    long long packed_stride;
    int fastamp_re_only;            // tdm_design::fastamp_re_only (selects the kernel instantiation)
    int* out_counts;                // [C]
    int accumulate;                 // 0: rows are written from symbol 0 and out_counts is overwritten;
                                    // 1: append after the out_counts[c] symbols already there (time slices of one call)
    tdm_channel_state* states;      // [C]
    int debug_mask;                 // development builds only (-DTDM_ABLATE): bit r set = role r of the ws4 pipeline skips its work
};

// variant: 0 = auto.  Returns the number of kernels launched, or <0 on launch error.
constexpr int kDemodVariants = 11;      // valid explicit variants are 1..kDemodVariants
int launch_demod(const DemodParams& p, int variant, cudaStream_t stream);
const char* demod_variant_name(int variant);

int launch_pack_dibits(const uint8_t* dibits, long long in_stride, const int* counts, uint8_t* packed,
                       long long out_stride, int n_channels, long long max_syms, cudaStream_t stream);

// packed rows (4 dibits per byte) -> one dibit per byte and/or one bit per byte (either may be null)
int launch_unpack_dibits(const uint8_t* packed, long long in_stride, const int* counts, uint8_t* dibits, long long dibit_stride,
                         uint8_t* bits, long long bit_stride, int n_channels, long long max_syms, cudaStream_t stream);

int launch_synth(const tdm_synth_params& sp, int n_channels, long long n_samples, long long stride,
                 int first_channel, float2* iq, uint8_t* tx_dibits, long long tx_stride, cudaStream_t stream);

// tdm_stitch.cu: joining the dibit streams of overlapping time segments (tdm_process_long[_batch]); S = segments per channel
void launch_long_init_states(tdm_channel_state* states, const tdm_channel_state* carried, const tdm_channel_state* fresh, int n, int S, cudaStream_t s);
void launch_long_shift_states(tdm_channel_state* dst, const tdm_channel_state* src, int n, int S, cudaStream_t s);
void launch_long_last_states(tdm_channel_state* packed, tdm_channel_state* rows, int C, int S, int scatter, cudaStream_t s);
void launch_stitch_find(const uint8_t* dib, long long stride, const int* counts, const int* tails, int n_rows, int S, int K, int jlo, int jhi,
                        int* join, int* fixed, int* cut, int late, cudaStream_t s);
void launch_stitch_plan(int* join, int* fixed, const int* counts, int n_rows, int S, int* adopt, int* n_open, int* n_forced, int force_at,
                        int force_all, const tdm_channel_state* final_states, cudaStream_t s);
void launch_stitch_adopt(uint8_t* dib, long long stride, const uint8_t* dib2, long long stride2, int* counts, const int* counts2, int* join,
                         int* fixed, const int* adopt, int* agree, int* mode, int* n_forced, int force_at, int K,
                         tdm_channel_state* final_states, const tdm_channel_state* run_states, int* cut, int n_rows, int S, long long max_len,
                         cudaStream_t s);
void launch_stitch_scan(const int* counts, const int* cut, const int* join, int n_rows, int S, long long* offs, long long* totals, cudaStream_t s);
void launch_stitch_copy(const uint8_t* dib, long long stride, const int* counts, const int* cut, const int* join, const long long* offs, int S,
                        uint8_t* out, long long out_stride, int n_rows, long long max_len, cudaStream_t s);
void launch_stitch_append(const uint8_t* src, long long stride, const int* count, long long* totals, uint8_t* out, long long out_stride, int C,
                          long long max_len, cudaStream_t s);

}  // namespace tdm
