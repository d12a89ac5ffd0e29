// tdm_chain.cuh -- the stages of the demodulation chain as device functions, shared by the kernel mappings
// (tdm_kernels.cu: thread per channel; tdm_ws.cu: warp-specialised role pipeline).  Every function restates a
// piece of the reference (paths relative to the reference tree) in the canonical operation order the CPU
// checker follows (oracle/oracle_b.c); the mappings only decide WHICH warp runs a stage WHEN.
#pragma once
#include "tdm_kernels.cuh"
#include "tdm_math.cuh"

namespace tdm {

constexpr int kHist = TDM_HIST;            // 64
constexpr int kTaps = TDM_MAX_TAPS;        // 65
constexpr int kTapPad = TDM_TAP_PAD;
constexpr int kITaps = TDM_INTERP_TAPS;    // 8
constexpr int kIPhases = TDM_INTERP_PHASES;

// First input sample of row `ch`.  Ordinarily rows are channels, in_stride apart.  For time-segmented captures
// (tdm_process_long_batch) a row is segment (ch % rows_per_channel) of channel (ch / rows_per_channel): channels are
// channel_stride apart, the segments of a channel in_stride apart (they overlap: in_stride < count).
__device__ __forceinline__ const float2* row_input(const DemodParams& p, int ch) {
    if (p.rows_per_channel <= 1) { return p.iq + (long long)ch * p.in_stride; }
    return p.iq + (long long)(ch / p.rows_per_channel) * p.channel_stride + (long long)(ch % p.rows_per_channel) * p.in_stride;
}

// Loop gains/limits of the symbol-rate loops, pinned in registers by the role that runs them.
struct SymConsts {
    float tr_alpha, tr_beta, tr_min, tr_max;
    float c_alpha, c_beta, c_min, c_max;
};
__device__ __forceinline__ SymConsts load_sym_consts(const DemodParams& p) {
    SymConsts k;
    k.tr_alpha = pin(p.tr_alpha); k.tr_beta = pin(p.tr_beta); k.tr_min = pin(p.tr_min_omega); k.tr_max = pin(p.tr_max_omega);
    k.c_alpha = pin(p.costas_alpha); k.c_beta = pin(p.costas_beta); k.c_min = pin(p.costas_min_freq); k.c_max = pin(p.costas_max_freq);
    return k;
}

// Constants of the sample-rate recurrences (AGC, FLL), pinned in registers for the serial loop.
struct LoopConsts {
    float agc_rate, agc_set, agc_max, fll_beta, fll_min, fll_max;
};
__device__ __forceinline__ LoopConsts load_loop_consts(const DemodParams& p) {
    LoopConsts k;
    k.agc_rate = pin(p.agc_rate); k.agc_set = pin(p.agc_set_point); k.agc_max = pin(p.agc_max_gain);
    k.fll_beta = pin(p.fll_beta); k.fll_min = pin(p.fll_min_freq); k.fll_max = pin(p.fll_max_freq);
    return k;
}

// FastAGC [A.3] for one sample: returns the scaled sample, advances the gain.  Branch free.
__device__ __forceinline__ float2 agc_step(const LoopConsts& lc, float2 in, float& g) {
    const float yr = mul_rn(in.x, g), yi = mul_rn(in.y, g);
    const float amp = sqrt_rn_nobranch(fma_rn(yr, yr, mul_rn(yi, yi)));
    g = fma_rn(sub_rn(lc.agc_set, amp), lc.agc_rate, g);
    g = g > lc.agc_max ? lc.agc_max : g;
    return make_float2(yr, yi);
}

// The FLL's carried loop state: phase/frequency of the PhaseControlLoop (fll.h:58) plus the phase reduced for the
// NEXT sample (tdm_math.cuh "the FLL's NCO"): q quadrants + r.
struct FllState {
    float ph, fr, r;
    uint32_t q;
};
__device__ __forceinline__ FllState fll_load(const tdm_channel_state* sp) {
    FllState s;
    s.ph = sp->fll_phase; s.fr = sp->fll_freq; s.r = sp->fll_r; s.q = sp->fll_quad;
    return s;
}
__device__ __forceinline__ void fll_store(tdm_channel_state* sp, const FllState& s) {
    sp->fll_phase = s.ph; sp->fll_freq = s.fr; sp->fll_r = s.r; sp->fll_quad = s.q;
}
// de-rotation of one gain-scaled sample by the loop phase (fll.cpp:137-138)
__device__ __forceinline__ float2 fll_derotate(const FllState& s, float2 y) {
    float sn, cs;
    fll_poly(s.r, sn, cs);
    return fll_rotate(quarter_turns(s.q, y), sn, cs);
}
// band-edge error and loop update from the finished P/Q chains of one output (fll.cpp:143-145), exact per-sample
// semantics: the prepared reduction is replaced by the classic one where it left the polynomials' range.
template <bool RE_ONLY>
__device__ __forceinline__ void fll_update(const LoopConsts& lc, float2 P, float2 Q, FllState& s) {
    uint32_t nq; float r0;
    fll_prepare(s.ph, s.fr, nq, r0);                       // from the state BEFORE the update
    const float hbe = fast_amplitude<RE_ONLY>(sub_rn(P.x, Q.y), add_rn(P.y, Q.x));
    const float lbe = fast_amplitude<RE_ONLY>(add_rn(P.x, Q.y), sub_rn(P.y, Q.x));
    const float ferr = sub_rn(hbe, lbe);
    s.fr = clampf(fma_rn(lc.fll_beta, ferr, s.fr), lc.fll_min, lc.fll_max);
    float r = add_rn(r0, s.fr);
    s.ph = wrap_pi(add_rn(s.ph, s.fr));
    if (!fll_r_ok(r)) { fll_reduce_classic(s.ph, nq, r); }
    s.q = nq; s.r = r;
}

// Timing recovery for one output symbol (complex_fd.cpp:96-143): interpolate the matched-filter output at
// `offset` with polyphase row floor(mu*128), derivative from the neighbouring rows, sign-decision-directed
// error, PI update of (omega, mu), integer advance of `offset`.  Returns the interpolated symbol.
//   rs     : matched-filter ring, [RE][32] float2, linear index offset+k (7 history entries first)
//   bank4  : the 128 x 8 polyphase bank as float4 halves, REP copies interleaved so that entry
//            (phase*2 + half) of copy c sits at bank4[(phase*2 + half) * REP + c].  A 128-bit shared-memory load is
//            served a quarter warp at a time: with REP = 8 and c = lane & 7 the eight lanes of a quarter warp hit
//            eight different 16-byte bank groups whatever their phases are (no conflicts, 32 KB); REP = 1 is the
//            plain table (conflicts when lanes sit on phases that are equal mod 4).
// The clamp is done on the float and the edge cases are folded into one expression on purpose: with an integer
// min/max clamp followed by `if (ph == 0) .. else if (ph == 127) ..`, ptxas 12.9 for sm_100a derived the
// `ph == 127` test from the predicate output of VIMNMX.RELU and took the last-phase branch for ph == 0.
template <int RE, int REP>
__device__ __forceinline__ float2 timing_step(const SymConsts& kc, const float4* __restrict__ bank4,
                                              const float2* rs, int lane, float& mu, float& om, int& offset) {
    const int ph = __float2int_rd(fminf(fmaxf(mul_rn(mu, (float)kIPhases), 0.0f), (float)(kIPhases - 1)));
    const int plo = max(ph - 1, 0);
    const int phi = min(ph + 1, kIPhases - 1);
    const int c = (REP > 1) ? (lane & (REP - 1)) : 0;
    const float4 t0a = bank4[(ph * 2) * REP + c], t0b = bank4[(ph * 2 + 1) * REP + c];
    const float4 t1a = bank4[(phi * 2) * REP + c], t1b = bank4[(phi * 2 + 1) * REP + c];
    const float4 t2a = bank4[(plo * 2) * REP + c], t2b = bank4[(plo * 2 + 1) * REP + c];
    const float t0[8] = { t0a.x, t0a.y, t0a.z, t0a.w, t0b.x, t0b.y, t0b.z, t0b.w };
    const float t1[8] = { t1a.x, t1a.y, t1a.z, t1a.w, t1b.x, t1b.y, t1b.z, t1b.w };
    const float t2[8] = { t2a.x, t2a.y, t2a.z, t2a.w, t2b.x, t2b.y, t2b.z, t2b.w };
    float2 y = make_float2(0.f, 0.f), a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kITaps; ++k) {
        const float2 v = rs[((offset + k) & (RE - 1)) * 32 + lane];
        y = fma2_rn(t0[k], v, y);
        a = fma2_rn(t1[k], v, a);
        b = fma2_rn(t2[k], v, b);
    }
    // derivative (complex_fd.cpp:107-123): first phase fT1 - y, last phase y - fT_1, otherwise (fT1 - fT_1) * 0.5.
    // At the edges the clamped neighbour row IS the centre row, so its dot product equals y bit for bit and all
    // three cases are (a - b) * scale with scale 1 or 0.5 (x1 is exact).
    const float dscale = (phi - plo == 2) ? 0.5f : 1.0f;
    const float dre = mul_rn(sub_rn(a.x, b.x), dscale);
    const float dim = mul_rn(sub_rn(a.y, b.y), dscale);
    float terr = add_rn(y.x > 0.f ? dre : -dre, y.y > 0.f ? dim : -dim);
    terr = clampf(terr, -1.0f, 1.0f);
    om = clampf(fma_rn(kc.tr_beta, terr, om), kc.tr_min, kc.tr_max);
    mu = add_rn(mu, fma_rn(kc.tr_alpha, terr, om));
    float delta = floorf(mu);
    // Non-finite guard (unreachable for finite input: delta is 1..3 then; tdm_design_from_config rejects
    // configurations whose omega could let mu stay below 1).  The reference would spin or hit UB in
    // `offset += delta` on NaN/Inf; a GPU must not: NaN and negative advances become 1, huge ones 2^20.
    delta = (delta >= 0.0f) ? delta : 1.0f;
    delta = fminf(delta, 1048576.0f);
    offset += (int)delta;
    mu = sub_rn(mu, delta);
    return y;
}

// The carrier-recovery step (pi4dqpsk_costas.cpp:5-28), branch free: returns the de-rotated symbol (PI4DQPSK's
// `out`), advances (phase, freq, ph2).  ph2's wrap forms both candidates and selects.
__device__ __forceinline__ float2 costas_loop_step(const SymConsts& kc, float2 y, float& cph, float& cfr, float& ph2) {
    float sn, cs;
    sincos_canon(cph, sn, cs);
    const float zr = fma_rn(y.x, cs, mul_rn(y.y, sn));
    const float zi = fma_rn(y.y, cs, -mul_rn(y.x, sn));
    const float two_pi_c = 2 * TDM_FL_M_PI;
    const float q0 = add_rn(ph2, -(TDM_FL_M_PI / 4.0f));
    const float qd = sub_rn(q0, two_pi_c), qu = add_rn(q0, two_pi_c);
    const float q = (q0 >= two_pi_c) ? qd : ((q0 <= -two_pi_c) ? qu : q0);
    ph2 = q;
    float s2, c2;
    sincos_canon(q, s2, c2);
    const float ur = fma_rn(zr, c2, -mul_rn(zi, s2));
    const float ui = fma_rn(zi, c2, mul_rn(zr, s2));
    float cerr = sub_rn(ur > 0.f ? ui : -ui, ui > 0.f ? ur : -ur);
    cerr = clampf(cerr, -1.0f, 1.0f);
    cfr = clampf(fma_rn(kc.c_beta, cerr, cfr), kc.c_min, kc.c_max);
    cph = wrap_pi(add_rn(cph, fma_rn(kc.c_alpha, cerr, cfr)));
    return make_float2(ur, ui);
}

// Slicer + lock metric + differential decoder + output formats (dqpsk_sym_extr.cpp:4-55, bit_unpacker.cpp:6-7)
// for up to NS symbols in straight-line code, so the independent per-symbol work (lock metric polynomial,
// decisions, address arithmetic, stores) of several symbols overlaps.  The 256-symbol block rotation of the
// lock metric can fire at most once in NS <= 255 symbols: it is captured with selects and carried out once at
// the end (nothing in between reads it).
struct SlicerState {
    uint32_t prev, err_ptr, err_disp;
    float err_partial, standarderr;
    uint32_t sync;
    int nsym;          // symbols emitted by this launch so far
    int nsym0;         // symbols already in the output rows when the launch began (time-sliced calls append)
    int out_room;      // symbols this launch may still write into the channel's rows
    uint32_t pk;       // TDM_OUT_PACKED: the dibits of the byte being filled (first symbol in the top bits)
    // this row's outputs, positioned at symbol 0 of this launch (null = not selected): one add per store instead of
    // 64-bit row arithmetic per symbol
    float2* o_syms;
    uint8_t* o_dibits;
    uchar2* o_bits;
    uint8_t* o_packed;   // positioned at byte 0 of the row
};
__device__ __forceinline__ void slicer_load(const DemodParams& p, const tdm_channel_state* sp, int ch, SlicerState& sl,
                                            float (&err_blocks)[TDM_SYNC_BLOCKS]) {
    sl.prev = sp->prev_sym; sl.err_ptr = sp->err_ptr; sl.err_disp = sp->err_disp;
    sl.err_partial = sp->err_partial; sl.standarderr = sp->standarderr; sl.sync = sp->sync;
    sl.nsym0 = p.accumulate ? p.out_counts[ch] : 0;
    sl.nsym = 0; sl.out_room = (int)p.out_stride - sl.nsym0;
    sl.pk = 0;
    const long long row = (long long)ch * p.out_stride + sl.nsym0;
    sl.o_syms = p.syms ? p.syms + row : nullptr;
    sl.o_dibits = p.dibits ? p.dibits + row : nullptr;
    sl.o_bits = p.bits ? reinterpret_cast<uchar2*>(p.bits) + row : nullptr;
    sl.o_packed = p.packed ? p.packed + (long long)ch * p.packed_stride : nullptr;
    if (p.packed && (sl.nsym0 & 3)) {      // appending in the middle of a byte: take back the dibits already there
        sl.pk = (uint32_t)p.packed[(long long)ch * p.packed_stride + (sl.nsym0 >> 2)] >> (2 * (4 - (sl.nsym0 & 3)));
    }
#pragma unroll
    for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { err_blocks[j] = sp->err_blocks[j]; }
}
__device__ __forceinline__ void slicer_store(const DemodParams& p, tdm_channel_state* sp, int ch, const SlicerState& sl,
                                             const float (&err_blocks)[TDM_SYNC_BLOCKS], bool active) {
    if (!active) { return; }
    sp->prev_sym = sl.prev; sp->err_ptr = sl.err_ptr; sp->err_disp = sl.err_disp;
    sp->err_partial = sl.err_partial; sp->standarderr = sl.standarderr; sp->sync = sl.sync;
    sp->n_symbols += (unsigned long long)sl.nsym;
#pragma unroll
    for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { sp->err_blocks[j] = err_blocks[j]; }
    const int total = sl.nsym0 + sl.nsym;
    if (p.packed && (total & 3) && total <= sl.nsym0 + sl.out_room) {      // last, partly filled byte: zero padded
        p.packed[(long long)ch * p.packed_stride + (total >> 2)] = (uint8_t)(sl.pk << (2 * (4 - (total & 3))));
    }
    p.out_counts[ch] = total;
}
template <int NS, typename GetSym>
__device__ __forceinline__ void slicer_symbols(const DemodParams& p, int ch, int n, SlicerState& sl, float* __restrict__ err_blocks,
                                               bool active, GetSym&& get_symbol) {
    bool crossed = false;
    float saved_partial = 0.f;
    uint32_t saved_ptr = 0;
    const int base = sl.nsym;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const bool v = k < n;
        const float2 u = get_symbol(base + k, v);
        const bool a = u.y < 0.f, b = u.x < 0.f;
        const float dist = quadrant_phase_error(u.x, u.y);     // |ideal.phase() - sym.phase()|, dqpsk_sym_extr.cpp:8-11
        const float ep = add_rn(sl.err_partial, dist);
        sl.err_partial = v ? ep : sl.err_partial;
        sl.err_ptr += v ? 1u : 0u;
        sl.err_disp += v ? 1u : 0u;
        const bool cross = v && sl.err_disp >= TDM_SYNC_DISPLAY;
        saved_partial = cross ? sl.err_partial : saved_partial;
        saved_ptr = cross ? sl.err_ptr : saved_ptr;
        crossed = crossed || cross;
        sl.err_partial = cross ? 0.f : sl.err_partial;
        sl.err_disp = cross ? 0u : sl.err_disp;
        sl.err_ptr = (sl.err_ptr >= TDM_SYNC_BUF) ? 0u : sl.err_ptr;
        const uint32_t sym = ((uint32_t)a << 1) | (uint32_t)(a != b);
        const uint32_t pd = (sym - sl.prev + 4u) & 3u;
        const uint32_t db = pd ^ (pd >> 1);          // 0,1,2,3 -> 0,1,3,2
        sl.prev = v ? sym : sl.prev;
        sl.pk = v ? ((sl.pk << 2) | db) : sl.pk;
        if (v && active && base + k < sl.out_room) {   // rows are sized by tdm_max_symbols(); never write past one
            const int o = base + k;
            if (sl.o_syms) { sl.o_syms[o] = u; }
            if (sl.o_dibits) { sl.o_dibits[o] = (uint8_t)db; }
            if (sl.o_bits) { sl.o_bits[o] = make_uchar2((uint8_t)((db >> 1) & 1u), (uint8_t)(db & 1u)); }
            const int idx = sl.nsym0 + o;
            if (sl.o_packed && (idx & 3) == 3) { sl.o_packed[idx >> 2] = (uint8_t)sl.pk; }
        }
    }
    sl.nsym = base + n;
    if (crossed) {
        err_blocks[(saved_ptr - 1) / TDM_SYNC_DISPLAY] = saved_partial;
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { tot = add_rn(tot, err_blocks[j]); }
        sl.standarderr = __fdiv_rn(tot, (float)TDM_SYNC_BUF);
        sl.sync = sl.standarderr < 0.35f ? 1u : 0u;
    }
}

// launchers of the two mappings (launch_demod in tdm_kernels.cu picks one)
int launch_tpc(const DemodParams& p, cudaStream_t stream, int T, int warps_per_cta);
int launch_ws4(const DemodParams& p, cudaStream_t stream, int placement, int ctas_per_sm);

}  // namespace tdm
