// tdm_kernels.cu -- the fused pi/4-DQPSK demodulation kernel for sm_100a.
//
// One launch runs the WHOLE reference chain for `count` samples of every channel:
//   FastAGC -> band-edge FLL -> RRC matched filter -> ML timing recovery -> pi/4 Costas
//   -> slicer + differential decoder -> (optional) bit unpack
// i.e. dsp::demod::PI4DQPSK::process (src/dsp/pi4dqpsk.cpp:132-140), then
// DQPSKSymbolExtractor::process (src/dsp/dqpsk_sym_extr.cpp:4-55) and
// BitUnpacker::process (src/dsp/bit_unpacker.cpp:4-10), with the per-channel state
// the reference keeps in class members carried in tdm_channel_state.
//
// Mapping (variant "tpc<T>"): one thread per channel, time processed in blocks of T
// samples.  The chain is a strict recurrence at sample rate (AGC gain, FLL phase) and
// at symbol rate (timing, Costas), so time cannot be split; what CAN be hoisted out of
// the recurrence is almost all of the FIR work:
//   * lbe/hbe/RRC all filter the same FLL-output sequence x (fll.cpp:141-142 ->
//     pi4dqpsk.cpp:135-136), so ONE 64-deep delay line (shared memory, [entry][lane] so
//     a warp's accesses are conflict-free) feeds all three;
//   * the band-edge taps are exact conjugates (fll.cpp:89-93): hbe = P + jQ, lbe = P - jQ
//     with P = sum a_k x, Q = sum b_k x  ->  6 real chains per output instead of 10;
//   * for a block of T outputs the contributions of the 64 samples older than the block
//     ("old part") do not depend on the block's own feedback: they are accumulated first,
//     each history sample loaded once from shared memory and used for up to T outputs
//     x 6 chains, taps as constant-bank FFMA operands.  Only the last <= T terms of each
//     chain sit inside the serial loop.
// Every chain still adds its terms in ascending tap order with one fma per term, which
// is the canonical order the CPU checker follows -- the blocking changes WHEN a term is
// added, never the order within a chain.
#include "tdm_kernels.cuh"
#include "tdm_math.cuh"
#include <cstdio>

namespace tdm {

namespace {

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

template <int T>
struct TpcLayout {
    static_assert(kHist % T == 0, "block length must divide the history length");
    static constexpr int kSlots = kHist / T + 1;         // ring of (64/T + 1) blocks of T
    static constexpr int kXEntries = kSlots * T;
    static constexpr int kREntries = (T + kITaps - 1 <= 16) ? 16 : 32;   // RRC-output ring (power of 2)
    static constexpr int kWarpFloat2 = (kXEntries + kREntries) * 32;
    static constexpr size_t kWarpBytes = sizeof(float2) * kWarpFloat2;
};

struct SymbolState {
    float mu, om;
    int offset;
    float cph, cfr, ph2;
    uint32_t prev;
    uint32_t err_ptr, err_disp;
    float err_partial, standarderr;
    uint32_t sync;
    int nsym;
    int out_room;      // symbols this call may still write into the channel's output rows
};

// Loop gains/limits of the symbol-rate loops, pinned in registers by the role that runs do_symbol().
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

// Timing recovery for one output symbol (complex_fd.cpp:96-143): interpolate the matched-filter output at
// `offset` with polyphase row floor(mu*128), derivative from the neighbouring rows, sign-decision-directed
// error, PI update of (omega, mu), integer advance of `offset`.  Returns the interpolated symbol.
template <int RE>
__device__ __forceinline__ float2 timing_step(const SymConsts& kc, const float* __restrict__ bank_s,
                                              const float2* rs, int lane, float& mu, float& om, int& offset) {
    // phase = clamp(floor(mu*128), 0, 127) (complex_fd.cpp:101).  The clamp is done on the float and the
    // edge cases below are folded into one expression on purpose: with an integer min/max clamp followed
    // by `if (ph == 0) .. else if (ph == 127) ..`, ptxas 12.9 for sm_100a derived the `ph == 127` test from
    // the predicate output of VIMNMX.RELU and took the last-phase branch for ph == 0 (seen on hardware).
    const float phf = fminf(fmaxf(floorf(mul_rn(mu, (float)kIPhases)), 0.0f), (float)(kIPhases - 1));
    const int ph = (int)phf;
    const int plo = max(ph - 1, 0);
    const int phi = min(ph + 1, kIPhases - 1);
    const float4* r0 = reinterpret_cast<const float4*>(bank_s + ph * kITaps);
    const float4* r1 = reinterpret_cast<const float4*>(bank_s + phi * kITaps);
    const float4* r2 = reinterpret_cast<const float4*>(bank_s + plo * kITaps);
    const float4 t0a = r0[0], t0b = r0[1], t1a = r1[0], t1b = r1[1], t2a = r2[0], t2b = r2[1];
    const float t0[8] = { t0a.x, t0a.y, t0a.z, t0a.w, t0b.x, t0b.y, t0b.z, t0b.w };
    const float t1[8] = { t1a.x, t1a.y, t1a.z, t1a.w, t1b.x, t1b.y, t1b.z, t1b.w };
    const float t2[8] = { t2a.x, t2a.y, t2a.z, t2a.w, t2b.x, t2b.y, t2b.z, t2b.w };
    float yre = 0.f, yim = 0.f, are = 0.f, aim = 0.f, bre = 0.f, bim = 0.f;
#pragma unroll
    for (int k = 0; k < kITaps; ++k) {
        // RRC outputs offset-7 .. offset live at linear ring index offset+k (7 history entries first)
        const float2 v = rs[((offset + k) & (RE - 1)) * 32 + lane];
        yre = fma_rn(t0[k], v.x, yre); yim = fma_rn(t0[k], v.y, yim);
        are = fma_rn(t1[k], v.x, are); aim = fma_rn(t1[k], v.y, aim);
        bre = fma_rn(t2[k], v.x, bre); bim = fma_rn(t2[k], v.y, bim);
    }
    // derivative (complex_fd.cpp:107-123): first phase fT1 - y, last phase y - fT_1, otherwise
    // (fT1 - fT_1) * 0.5.  At the edges the clamped neighbour row IS the centre row, so its dot product
    // equals y bit for bit and all three cases are (a - b) * scale with scale 1 or 0.5 (x1 is exact).
    const float dscale = (phi - plo == 2) ? 0.5f : 1.0f;
    const float dre = mul_rn(sub_rn(are, bre), dscale);
    const float dim = mul_rn(sub_rn(aim, bim), dscale);
    float terr = add_rn(yre > 0.f ? dre : -dre, yim > 0.f ? dim : -dim);
    terr = clampf(terr, -1.0f, 1.0f);
    om = clampf(fma_rn(kc.tr_beta, terr, om), kc.tr_min, kc.tr_max);
    mu = add_rn(mu, fma_rn(kc.tr_alpha, terr, om));
    float delta = floorf(mu);
    // Non-finite guard (unreachable for finite input: delta is 1..3 then).  The reference would spin or
    // hit UB in `offset += delta` on NaN/Inf; a GPU must not, so the advance is forced into [1, 2^20].
    delta = (delta >= 0.0f) ? delta : 1.0f;
    delta = fminf(delta, 1048576.0f);
    offset += (int)delta;
    mu = sub_rn(mu, delta);
    return make_float2(yre, yim);
}

// pi/4 Costas loop, slicer, lock metric, differential decoder, bit unpack for one symbol
// (pi4dqpsk_costas.cpp:5-28, dqpsk_sym_extr.cpp:4-55, bit_unpacker.cpp:6-7).
__device__ __forceinline__ void costas_step(const DemodParams& p, const SymConsts& kc, float2 y, SymbolState& st,
                                            float* __restrict__ err_blocks, bool active, long long out_base) {
    float sn, cs;
    sincos_canon(st.cph, sn, cs);
    const float zr = fma_rn(y.x, cs, mul_rn(y.y, sn));
    const float zi = fma_rn(y.y, cs, -mul_rn(y.x, sn));
    const float two_pi_c = 2 * TDM_FL_M_PI;
    float ph2 = add_rn(st.ph2, -(TDM_FL_M_PI / 4.0f));
    ph2 = (ph2 >= two_pi_c) ? sub_rn(ph2, two_pi_c) : ((ph2 <= -two_pi_c) ? add_rn(ph2, two_pi_c) : ph2);
    st.ph2 = ph2;
    float s2, c2;
    sincos_canon(ph2, s2, c2);
    const float ur = fma_rn(zr, c2, -mul_rn(zi, s2));
    const float ui = fma_rn(zi, c2, mul_rn(zr, s2));
    float cerr = sub_rn(ur > 0.f ? ui : -ui, ui > 0.f ? ur : -ur);
    cerr = clampf(cerr, -1.0f, 1.0f);
    st.cfr = clampf(fma_rn(kc.c_beta, cerr, st.cfr), kc.c_min, kc.c_max);
    st.cph = wrap_pi(add_rn(st.cph, fma_rn(kc.c_alpha, cerr, st.cfr)));

    // --- slicer, sync metric, differential decode
    const bool a = ui < 0.f, b = ur < 0.f;
    const float dist = quadrant_phase_error(ur, ui);     // |ideal.phase() - sym.phase()|, dqpsk_sym_extr.cpp:8-11
    st.err_partial = add_rn(st.err_partial, dist);
    st.err_ptr++;
    st.err_disp++;
    if (st.err_disp >= TDM_SYNC_DISPLAY) {
        err_blocks[(st.err_ptr - 1) / TDM_SYNC_DISPLAY] = st.err_partial;
        st.err_partial = 0.f;
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { tot = add_rn(tot, err_blocks[j]); }
        st.standarderr = __fdiv_rn(tot, (float)TDM_SYNC_BUF);
        st.sync = st.standarderr < 0.35f ? 1u : 0u;
        st.err_disp = 0;
    }
    if (st.err_ptr >= TDM_SYNC_BUF) { st.err_ptr = 0; }
    const uint32_t sym = ((uint32_t)a << 1) | (uint32_t)(a != b);
    const uint32_t pd = (sym - st.prev + 4u) & 3u;
    const uint32_t db = pd ^ (pd >> 1);          // 0,1,2,3 -> 0,1,3,2
    st.prev = sym;
    if (active && st.nsym < st.out_room) {   // rows are sized by tdm_max_symbols(); never write past one
        const long long o = out_base + st.nsym;
        if (p.syms) { p.syms[o] = make_float2(ur, ui); }
        if (p.dibits) { p.dibits[o] = (uint8_t)db; }
        if (p.bits) { reinterpret_cast<uchar2*>(p.bits)[o] = make_uchar2((uint8_t)((db >> 1) & 1u), (uint8_t)(db & 1u)); }
    }
    st.nsym++;
}

// One output symbol, both halves in the same thread (thread-per-channel variants).
template <int RE>
__device__ __forceinline__ void do_symbol(const DemodParams& p, const SymConsts& kc, const float* __restrict__ bank_s,
                                          const float2* __restrict__ rs, int lane, SymbolState& st,
                                          float* __restrict__ err_blocks, bool active, long long out_base) {
    const float2 y = timing_step<RE>(kc, bank_s, rs, lane, st.mu, st.om, st.offset);
    costas_step(p, kc, y, st, err_blocks, active, out_base);
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

// AGC + de-rotation of one input sample (FastAGC [A.3]; fll.cpp:137-138).  Branch free.
__device__ __forceinline__ float2 agc_derotate(const LoopConsts& lc, float2 in, float& g, float fph) {
    const float yr = mul_rn(in.x, g), yi = mul_rn(in.y, g);
    const float amp = sqrt_rn_nobranch(fma_rn(yr, yr, mul_rn(yi, yi)));
    g = fma_rn(sub_rn(lc.agc_set, amp), lc.agc_rate, g);
    g = g > lc.agc_max ? lc.agc_max : g;
    float sn, cs;
    sincos_canon(fph, sn, cs);
    return make_float2(fma_rn(yr, cs, mul_rn(yi, sn)), fma_rn(yi, cs, -mul_rn(yr, sn)));
}

// band-edge error and FLL loop update from the finished P/Q chains of one output (fll.cpp:143-145)
__device__ __forceinline__ void fll_update(const LoopConsts& lc, float pr, float pi, float qr, float qi, float& fph, float& ffr) {
    const float hbe = fast_amplitude(sub_rn(pr, qi), add_rn(pi, qr));
    const float lbe = fast_amplitude(add_rn(pr, qi), sub_rn(pi, qr));
    const float ferr = sub_rn(hbe, lbe);
    ffr = clampf(fma_rn(lc.fll_beta, ferr, ffr), lc.fll_min, lc.fll_max);
    fph = wrap_pi(add_rn(fph, ffr));
}

// ---------------------------------------------------------------------------------------
// Variant tpc<T>: thread per channel, time in blocks of T samples.
//
// Instruction-cache discipline: the first version of this kernel unrolled everything
// (53 KB of SASS per block iteration) and ncu showed `stall_no_instruction` as the top
// stall at IPC 0.36.  The hot loop is therefore kept ROLLED and small (about 10 KB):
//   * old part: a loop over the 64/T history slots; each iteration loads T history
//     samples and the 2T-1 taps per filter they meet (uniform constant-bank loads from
//     the T-1-zero-padded tap tables, tpad[f][s*T + j-i+T-1]) and does T*T*6 FMAs;
//   * serial part: a loop over the T samples; the T in-flight chains sit in a register
//     shift-register, so position q always meets tap 64-q (an immediate constant-bank
//     operand) and the loop body does not depend on the sample index;
//   * symbol part: do_symbol(), called from a while loop.
// ---------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(128) demod_tpc_kernel(const __grid_constant__ DemodParams p) {
    using L = TpcLayout<T>;
    constexpr int S = L::kSlots;
    constexpr int RE = L::kREntries;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bank_s = reinterpret_cast<float*>(smem_raw);                         // [128][8]
    float2* warp_base = reinterpret_cast<float2*>(smem_raw + sizeof(float) * kIPhases * kITaps) +
                        (size_t)(threadIdx.x >> 5) * L::kWarpFloat2;
    float2* xs = warp_base;                     // [kXEntries][32]
    float2* rs = warp_base + L::kXEntries * 32; // [RE][32]
    const int lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < kIPhases * kITaps; i += blockDim.x) { bank_s[i] = p.bank[i]; }

    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = ch < p.n_channels;
    if (!active) { ch = p.n_channels - 1; }     // compute a duplicate, store nothing
    tdm_channel_state* __restrict__ sp = p.states + ch;

    // ---- load carried state
    float g = sp->agc_gain, fph = sp->fll_phase, ffr = sp->fll_freq;
    SymbolState st;
    st.mu = sp->tr_mu; st.om = sp->tr_omega; st.offset = sp->tr_offset;
    st.cph = sp->costas_phase; st.cfr = sp->costas_freq; st.ph2 = sp->costas_ph2;
    st.prev = sp->prev_sym; st.err_ptr = sp->err_ptr; st.err_disp = sp->err_disp;
    st.err_partial = sp->err_partial; st.standarderr = sp->standarderr; st.sync = sp->sync;
    const int nsym0 = p.accumulate ? p.out_counts[ch] : 0;       // time-sliced calls append to the rows (tdm_api.cu)
    const long long out_base = (long long)ch * p.out_stride + nsym0;
    st.nsym = 0; st.out_room = (int)p.out_stride - nsym0;
    float err_blocks[TDM_SYNC_BLOCKS];
#pragma unroll
    for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { err_blocks[j] = sp->err_blocks[j]; }
    // delay line: linear index q (0..63 = carried history, 64+n = new sample n) lives in ring
    // slot (1 + q/T) mod S, position q%T;  RRC ring: linear q' (0..6 history, 7+n new) at q' & (RE-1)
    {
        const float2* xh = reinterpret_cast<const float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) { xs[((1 + m / T) * T + (m % T)) * 32 + lane] = xh[m]; }
        const float2* rh = reinterpret_cast<const float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { rs[j * 32 + lane] = rh[j]; }
    }
    __syncthreads();   // bank_s visible

    const float2* __restrict__ in = row_input(p, ch);
    const int count = p.count;
    const int nblk = (count + T - 1) / T;

    float2 cur[T], nxt[T];
#pragma unroll
    for (int i = 0; i < T; ++i) { cur[i] = (i < count) ? __ldg(in + i) : make_float2(0.f, 0.f); }
    const SymConsts kc = load_sym_consts(p);
    const LoopConsts lc = load_loop_consts(p);

    int slot = 0;
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        const int n0 = blk * T;
        const int valid = min(T, count - n0);
        // prefetch the next block's input while this one computes
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const int n = n0 + T + i;
            nxt[i] = (n < count) ? __ldg(in + n) : make_float2(0.f, 0.f);
        }

        // ---- old part: history terms of all T outputs (independent of this block's feedback).
        // acc[i][*] is the chain of output i; history sample m = s*T + j meets it at tap m - i, which
        // is entry j - i + T - 1 of the padded table row starting at s*T (negative taps read zeros).
        float acc[T][6];
#pragma unroll
        for (int i = 0; i < T; ++i) {
#pragma unroll
            for (int c = 0; c < 6; ++c) { acc[i][c] = 0.f; }
        }
        {
            int hs = slot + 1;
#pragma unroll 1
            for (int s = 0; s < kHist / T; ++s) {
                if (hs >= S) { hs -= S; }
                float ta[2 * T - 1], tb[2 * T - 1], tr[2 * T - 1];
#pragma unroll
                for (int c = 0; c < 2 * T - 1; ++c) {
                    ta[c] = p.tpad[0][s * T + c];
                    tb[c] = p.tpad[1][s * T + c];
                    tr[c] = p.tpad[2][s * T + c];
                }
#pragma unroll
                for (int j = 0; j < T; ++j) {
                    const float2 h = xs[(hs * T + j) * 32 + lane];
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        const int c = j - i + T - 1;
                        acc[i][0] = fma_rn(ta[c], h.x, acc[i][0]);
                        acc[i][1] = fma_rn(ta[c], h.y, acc[i][1]);
                        acc[i][2] = fma_rn(tb[c], h.x, acc[i][2]);
                        acc[i][3] = fma_rn(tb[c], h.y, acc[i][3]);
                        acc[i][4] = fma_rn(tr[c], h.x, acc[i][4]);
                        acc[i][5] = fma_rn(tr[c], h.y, acc[i][5]);
                    }
                }
                ++hs;
            }
        }

        // ---- serial part: the recurrences, plus the <= T newest terms of each chain.
        // Shift-register form: before step i, acc[q] is the chain of output i+q; the new sample is
        // tap 64-q of that output.  After the step the finished chain acc[0] is consumed and the
        // register file shifts down by one (positions past T-1-i hold don't-care values).
#pragma unroll 1
        for (int i = 0; i < valid; ++i) {
            const float2 xv = agc_derotate(lc, cur[0], g, fph);
            const float xr = xv.x, xi = xv.y;
            xs[(slot * T + i) * 32 + lane] = make_float2(xr, xi);
#pragma unroll
            for (int q = 0; q < T; ++q) {
                acc[q][0] = fma_rn(p.be_a[kHist - q], xr, acc[q][0]);
                acc[q][1] = fma_rn(p.be_a[kHist - q], xi, acc[q][1]);
                acc[q][2] = fma_rn(p.be_b[kHist - q], xr, acc[q][2]);
                acc[q][3] = fma_rn(p.be_b[kHist - q], xi, acc[q][3]);
                acc[q][4] = fma_rn(p.rrc[kHist - q], xr, acc[q][4]);
                acc[q][5] = fma_rn(p.rrc[kHist - q], xi, acc[q][5]);
            }
            fll_update(lc, acc[0][0], acc[0][1], acc[0][2], acc[0][3], fph, ffr);
            // matched-filter output -> interpolator ring
            rs[((kITaps - 1 + n0 + i) & (RE - 1)) * 32 + lane] = make_float2(acc[0][4], acc[0][5]);
#pragma unroll
            for (int q = 0; q < T - 1; ++q) {
#pragma unroll
                for (int c = 0; c < 6; ++c) { acc[q][c] = acc[q + 1][c]; }
                cur[q] = cur[q + 1];
            }
        }

        // ---- symbols that became computable in this block (complex_fd.cpp:96 `while (offset < count)`)
        while (st.offset < n0 + valid) { do_symbol<RE>(p, kc, bank_s, rs, lane, st, err_blocks, active, out_base); }

        slot = (slot + 1 == S) ? 0 : slot + 1;
#pragma unroll
        for (int i = 0; i < T; ++i) { cur[i] = nxt[i]; }
    }

    // ---- carry state out
    if (active) {
        sp->agc_gain = g; sp->fll_phase = fph; sp->fll_freq = ffr;
        sp->tr_mu = st.mu; sp->tr_omega = st.om; sp->tr_offset = st.offset - count;   // complex_fd.cpp:145
        sp->costas_phase = st.cph; sp->costas_freq = st.cfr; sp->costas_ph2 = st.ph2;
        sp->prev_sym = st.prev; sp->err_ptr = st.err_ptr; sp->err_disp = st.err_disp;
        sp->err_partial = st.err_partial; sp->standarderr = st.standarderr; sp->sync = st.sync;
        sp->n_samples += (unsigned long long)count;
        sp->n_symbols += (unsigned long long)st.nsym;
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { sp->err_blocks[j] = err_blocks[j]; }
        float2* xh = reinterpret_cast<float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) {
            const long long q = (long long)count + m;
            xh[m] = xs[(int)(((1 + q / T) % S) * T + (q % T)) * 32 + lane];
        }
        float2* rh = reinterpret_cast<float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { rh[j] = rs[((count + j) & (RE - 1)) * 32 + lane]; }
        p.out_counts[ch] = nsym0 + st.nsym;
    }
}

template <int T>
int launch_tpc(const DemodParams& p_in, cudaStream_t stream, int warps_per_cta) {
    using L = TpcLayout<T>;
    DemodParams p = p_in;
    // tap tables padded with T-1 leading zeros (see the old-part loop)
    const float* src[3] = { p.be_a, p.be_b, p.rrc };
    for (int f = 0; f < 3; ++f) {
        for (int j = 0; j < kTapPad; ++j) {
            const int k = j - (T - 1);
            p.tpad[f][j] = (k >= 0 && k < kTaps) ? src[f][k] : 0.f;
        }
    }
    const int threads = 32 * warps_per_cta;
    const size_t smem = sizeof(float) * kIPhases * kITaps + L::kWarpBytes * warps_per_cta;
    cudaFuncSetAttribute(demod_tpc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int grid = (p.n_channels + threads - 1) / threads;
    demod_tpc_kernel<T><<<grid, threads, smem, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---------------------------------------------------------------------------------------
// Variant ws<8>: warp-specialised software pipeline, 32 channels per CTA, 5 warps.
//
// The thread-per-channel kernel makes ONE warp carry all ~700 instructions per sample of
// its 32 channels, and at 4096 channels only 128 of the GPU's 592 warp schedulers have a
// warp at all.  Here the same arithmetic (same chains, same order) is cut by ROLE, so the
// only warp that sits on the sample-rate recurrence executes ~165 instructions per sample
// and four more schedulers per 32 channels do the rest concurrently:
//
//   warp 0  LOOP   AGC + FLL recurrence; adds the 2T-ish newest terms of the P/Q chains
//   warp 1  P-far  the older 56-i terms of the two P chains (band-edge real-tap sums)
//   warp 2  Q-far  same for the two Q chains
//   warp 3  RRC    the complete matched-filter chains (not on the feedback path at all)
//   warp 4  SYM    timing recovery, Costas, slicer, differential decoder
//
// Time advances in ticks of T = 8 samples with one __syncthreads() per tick; at tick t
//   LOOP works on block t, P/Q-far prepare block t+1 (they need x only up to block t-1),
//   RRC filters block t-1, SYM consumes the matched-filter outputs of block t-2.
// All hand-offs go through shared-memory rings indexed by absolute sample position, so
// the double buffering is implicit.  Chains still add terms in ascending tap order:
// far part (P/Q warp) -> previous block's 8 samples -> own block (LOOP warp).
// ---------------------------------------------------------------------------------------
constexpr int kWsT = 8;
constexpr int kWsXSlots = 16;                        // x ring: 16 blocks of 8 samples
constexpr int kWsXEntries = kWsXSlots * kWsT;        // 128
constexpr int kWsREntries = 32;                      // matched-filter output ring
constexpr int kWsFar = kHist / kWsT - 1;             // 7 blocks of history feed the far part
struct WsSmem {
    float bank[kIPhases * kITaps];
    float2 xs[kWsXEntries][32];
    float2 rs[kWsREntries][32];
    float2 pfar[2][kWsT][32];
    float2 qfar[2][kWsT][32];
};

// two chains (re, im) of one real-tap filter for the T outputs of block b, over x-ring blocks
// [first, first+nblocks) in linear-q block units; table row tp is padded with T-1 leading zeros
template <int NB, int F>
__device__ __forceinline__ void ws_fir_blocks(const DemodParams& p, const float2 (*xs)[32], int lane,
                                              int qblock0, float (&acc)[kWsT][2]) {
    constexpr int T = kWsT;
#pragma unroll 1
    for (int s = 0; s < NB; ++s) {
        float tt[2 * T - 1];
#pragma unroll
        for (int c = 0; c < 2 * T - 1; ++c) { tt[c] = p.tpad[F][s * T + c]; }   // uniform: F is compile time
        const int slot = (qblock0 + s) & (kWsXSlots - 1);
#pragma unroll
        for (int j = 0; j < T; ++j) {
            const float2 h = xs[slot * T + j][lane];
#pragma unroll
            for (int i = 0; i < T; ++i) {
                acc[i][0] = fma_rn(tt[j - i + T - 1], h.x, acc[i][0]);
                acc[i][1] = fma_rn(tt[j - i + T - 1], h.y, acc[i][1]);
            }
        }
    }
}

__global__ void __launch_bounds__(160) demod_ws_kernel(const __grid_constant__ DemodParams p) {
    constexpr int T = kWsT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsSmem& sm = *reinterpret_cast<WsSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int role = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < kIPhases * kITaps; i += blockDim.x) { sm.bank[i] = p.bank[i]; }

    int ch = blockIdx.x * 32 + lane;
    const bool active = ch < p.n_channels;
    if (!active) { ch = p.n_channels - 1; }
    tdm_channel_state* __restrict__ sp = p.states + ch;
    const int count = p.count;
    const int nblk = (count + T - 1) / T;

    // rings: x linear index q (0..63 carried history, 64+n new sample n) -> slot (q/T) & 15, pos q%T;
    //        matched-filter linear index q' (0..6 history, 7+n new) -> q' & 31
    // Ring entries that are read before they are first written (the tail of a partial last block, met
    // only by ZERO taps) must still be finite: 0 * NaN would poison a chain.  Clear everything past the
    // carried history once.
    for (int i = threadIdx.x; i < (kWsXEntries - kHist) * 32; i += blockDim.x) {
        sm.xs[kHist + i / 32][i % 32] = make_float2(0.f, 0.f);
    }
    if (role == 1) {
        const float2* xh = reinterpret_cast<const float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) { sm.xs[m][lane] = xh[m]; }
    }
    if (role == 3) {
        const float2* rh = reinterpret_cast<const float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { sm.rs[j][lane] = rh[j]; }
    }

    // ---- role-private state
    float g = 0.f, fph = 0.f, ffr = 0.f;
    float2 cur[T], nxt[T];
    const float2* __restrict__ in = row_input(p, ch);
    SymbolState st;
    float err_blocks[TDM_SYNC_BLOCKS];
    const int nsym0 = p.accumulate ? p.out_counts[ch] : 0;       // time-sliced calls append to the rows (tdm_api.cu)
    const long long out_base = (long long)ch * p.out_stride + nsym0;
    LoopConsts lc = {};
    SymConsts kc = {};
    float tria[T], trib[T];                 // taps 64-q of the band-edge pair: the newest terms, in registers
#pragma unroll
    for (int q = 0; q < T; ++q) { tria[q] = 0.f; trib[q] = 0.f; }
    if (role == 0) {
        g = sp->agc_gain; fph = sp->fll_phase; ffr = sp->fll_freq;
#pragma unroll
        for (int i = 0; i < T; ++i) { cur[i] = (i < count) ? __ldg(in + i) : make_float2(0.f, 0.f); }
        lc = load_loop_consts(p);
#pragma unroll
        for (int q = 0; q < T; ++q) { tria[q] = pin(p.be_a[kHist - q]); trib[q] = pin(p.be_b[kHist - q]); }
    }
    if (role == 4) {
        st.mu = sp->tr_mu; st.om = sp->tr_omega; st.offset = sp->tr_offset;
        st.cph = sp->costas_phase; st.cfr = sp->costas_freq; st.ph2 = sp->costas_ph2;
        st.prev = sp->prev_sym; st.err_ptr = sp->err_ptr; st.err_disp = sp->err_disp;
        st.err_partial = sp->err_partial; st.standarderr = sp->standarderr; st.sync = sp->sync;
        st.nsym = 0; st.out_room = (int)p.out_stride - nsym0;
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { err_blocks[j] = sp->err_blocks[j]; }
        kc = load_sym_consts(p);
    }
    __syncthreads();

#ifdef TDM_ROLE_TIMING
    long long aux_cycles = 0;
    long long work_cycles = 0;
    const long long t_begin = clock64();
#endif
#pragma unroll 1
    for (int t = -1; t <= nblk + 1; ++t) {
#ifdef TDM_ROLE_TIMING
        const long long c0 = clock64();
#endif
        if (role == 0) {
            // ================= LOOP: block b = t =================
            if (t >= 0 && t < nblk) {
                const int n0 = t * T;
                const int valid = min(T, count - n0);
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const int n = n0 + T + i;
                    nxt[i] = (n < count) ? __ldg(in + n) : make_float2(0.f, 0.f);
                }
                // chains of this block's outputs: far part from the P/Q warps ...
                float acc[T][4];
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const float2 pf = sm.pfar[t & 1][i][lane], qf = sm.qfar[t & 1][i][lane];
                    acc[i][0] = pf.x; acc[i][1] = pf.y; acc[i][2] = qf.x; acc[i][3] = qf.y;
                }
                // ... then the previous block's 8 samples (taps 56+j-i) ...
                {
                    const int slot = (t + 7) & (kWsXSlots - 1);      // q-block of sample block t-1
#pragma unroll 2
                    for (int j = 0; j < T; ++j) {
                        const float2 h = sm.xs[slot * T + j][lane];
                        float ta[T], tb[T];
#pragma unroll
                        for (int i = 0; i < T; ++i) {
                            ta[i] = p.tpad[0][7 * T + j - i + T - 1];
                            tb[i] = p.tpad[1][7 * T + j - i + T - 1];
                        }
#pragma unroll
                        for (int i = 0; i < T; ++i) {
                            acc[i][0] = fma_rn(ta[i], h.x, acc[i][0]);
                            acc[i][1] = fma_rn(ta[i], h.y, acc[i][1]);
                            acc[i][2] = fma_rn(tb[i], h.x, acc[i][2]);
                            acc[i][3] = fma_rn(tb[i], h.y, acc[i][3]);
                        }
                    }
                }
                // ... then the block's own samples inside the recurrence (shift-register form)
                const int xslot = (t + 8) & (kWsXSlots - 1);
#ifdef TDM_ROLE_TIMING
                const long long c1 = clock64();
                aux_cycles += c1 - c0;
#endif
#pragma unroll 1
                for (int i = 0; i < valid; ++i) {
                    const float2 xv = agc_derotate(lc, cur[0], g, fph);
                    const float xr = xv.x, xi = xv.y;
                    sm.xs[xslot * T + i][lane] = make_float2(xr, xi);
#pragma unroll
                    for (int q = 0; q < T; ++q) {
                        acc[q][0] = fma_rn(tria[q], xr, acc[q][0]);
                        acc[q][1] = fma_rn(tria[q], xi, acc[q][1]);
                        acc[q][2] = fma_rn(trib[q], xr, acc[q][2]);
                        acc[q][3] = fma_rn(trib[q], xi, acc[q][3]);
                    }
                    fll_update(lc, acc[0][0], acc[0][1], acc[0][2], acc[0][3], fph, ffr);
#pragma unroll
                    for (int q = 0; q < T - 1; ++q) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) { acc[q][c] = acc[q + 1][c]; }
                        cur[q] = cur[q + 1];
                    }
                }
#pragma unroll
                for (int i = 0; i < T; ++i) { cur[i] = nxt[i]; }
            }
        } else if (role == 1 || role == 2) {
            // ================= P-far / Q-far: block b = t + 1 =================
            const int b = t + 1;
            if (b < nblk) {
                float acc[T][2];
#pragma unroll
                for (int i = 0; i < T; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
                if (role == 1) { ws_fir_blocks<kWsFar, 0>(p, sm.xs, lane, b, acc); }
                else { ws_fir_blocks<kWsFar, 1>(p, sm.xs, lane, b, acc); }
                float2 (*dst)[32] = (role == 1) ? sm.pfar[b & 1] : sm.qfar[b & 1];
#pragma unroll
                for (int i = 0; i < T; ++i) { dst[i][lane] = make_float2(acc[i][0], acc[i][1]); }
            }
        } else if (role == 3) {
            // ================= RRC: block b = t - 1 (all 65 taps; 9 ring blocks) =================
            const int b = t - 1;
            if (b >= 0 && b < nblk) {
                float acc[T][2];
#pragma unroll
                for (int i = 0; i < T; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
                ws_fir_blocks<kWsFar + 2, 2>(p, sm.xs, lane, b, acc);
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    sm.rs[(kITaps - 1 + b * T + i) & (kWsREntries - 1)][lane] = make_float2(acc[i][0], acc[i][1]);
                }
            }
        } else {
            // ================= SYM: symbols whose last input sample lies in block t - 2 =================
            if (t >= 2) {
                const int lim = min(count, (t - 1) * T);
                while (st.offset < lim) {
                    do_symbol<kWsREntries>(p, kc, sm.bank, &sm.rs[0][0], lane, st, err_blocks, active, out_base);
                }
            }
        }
#ifdef TDM_ROLE_TIMING
        work_cycles += clock64() - c0;
#endif
        __syncthreads();
    }
#ifdef TDM_ROLE_TIMING
    if (blockIdx.x == 3 && lane == 0) {
        printf("role %d: work %lld of %lld cycles (%.1f%%), per tick %lld (before sample loop: %lld)\n", role, work_cycles, clock64() - t_begin,
               100.0 * work_cycles / (double)(clock64() - t_begin), work_cycles / (nblk + 3), aux_cycles / (nblk + 3));
    }
#endif

    // ---- carry state out
    if (!active) { return; }
    if (role == 0) {
        sp->agc_gain = g; sp->fll_phase = fph; sp->fll_freq = ffr;
        sp->n_samples += (unsigned long long)count;
    } else if (role == 1) {
        float2* xh = reinterpret_cast<float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) {
            const long long q = (long long)count + m;
            xh[m] = sm.xs[(int)((q / T) & (kWsXSlots - 1)) * T + (int)(q % T)][lane];
        }
    } else if (role == 3) {
        float2* rh = reinterpret_cast<float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { rh[j] = sm.rs[(count + j) & (kWsREntries - 1)][lane]; }
    } else if (role == 4) {
        sp->tr_mu = st.mu; sp->tr_omega = st.om; sp->tr_offset = st.offset - count;
        sp->costas_phase = st.cph; sp->costas_freq = st.cfr; sp->costas_ph2 = st.ph2;
        sp->prev_sym = st.prev; sp->err_ptr = st.err_ptr; sp->err_disp = st.err_disp;
        sp->err_partial = st.err_partial; sp->standarderr = st.standarderr; sp->sync = st.sync;
        sp->n_symbols += (unsigned long long)st.nsym;
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { sp->err_blocks[j] = err_blocks[j]; }
        p.out_counts[ch] = nsym0 + st.nsym;
    }
}

int launch_ws(const DemodParams& p_in, cudaStream_t stream) {
    DemodParams p = p_in;
    const float* src[3] = { p.be_a, p.be_b, p.rrc };
    for (int f = 0; f < 3; ++f) {
        for (int j = 0; j < kTapPad; ++j) {
            const int k = j - (kWsT - 1);
            p.tpad[f][j] = (k >= 0 && k < kTaps) ? src[f][k] : 0.f;
        }
    }
    cudaFuncSetAttribute(demod_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsSmem));
    const int grid = (p.n_channels + 31) / 32;
    demod_ws_kernel<<<grid, 160, sizeof(WsSmem), stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---------------------------------------------------------------------------------------
// Variant ws8b: the ws8 pipeline after profiling it (profiles/r01_*):
//   * SYM was the slowest role (two serial recurrences in one warp): split into TIMING (interpolator +
//     timing loop) and COSTAS (carrier loop + slicer + decoder), one tick apart, linked by a 16-symbol ring;
//   * LOOP's AGC recurrence (through an IEEE sqrt) started late in each iteration, behind the FLL code in
//     program order (the SM issues in order): the gain product of sample i+1 is now formed in iteration
//     i, so both recurrences start at the top of the body; the 15+15 taps that meet the previous
//     block's samples live in registers and that part is straight-line code;
//   * the FIR roles double-buffer their tap/sample loads so no iteration starts by waiting on LDCU/LDS.
// 6 warps: 0 LOOP, 1 P-far, 2 Q-far, 3 RRC, 4 TIMING, 5 COSTAS.
// ---------------------------------------------------------------------------------------
constexpr int kYRing = 16;
struct Ws2Smem {
    float bank[kIPhases * kITaps];
    float2 xs[kWsXEntries][32];
    float2 rs[kWsREntries][32];
    float2 pfar[2][kWsT][32];
    float2 qfar[2][kWsT][32];
    float2 ys[kYRing][32];
    int ycount[2][32];
};

template <int NB, int F>
__device__ __forceinline__ void ws_fir_blocks2(const DemodParams& p, const float2 (*xs)[32], int lane,
                                               int qblock0, float (&acc)[kWsT][2]) {
    constexpr int T = kWsT;
    float ta[2 * T - 1], tb[2 * T - 1];
    float2 ha[T], hb[T];
    auto load = [&](float (&tt)[2 * T - 1], float2 (&h)[T], int s) {
#pragma unroll
        for (int c = 0; c < 2 * T - 1; ++c) { tt[c] = p.tpad[F][s * T + c]; }
        const int slot = (qblock0 + s) & (kWsXSlots - 1);
#pragma unroll
        for (int j = 0; j < T; ++j) { h[j] = xs[slot * T + j][lane]; }
    };
    auto comp = [&](const float (&tt)[2 * T - 1], const float2 (&h)[T]) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
#pragma unroll
            for (int i = 0; i < T; ++i) {
                acc[i][0] = fma_rn(tt[j - i + T - 1], h[j].x, acc[i][0]);
                acc[i][1] = fma_rn(tt[j - i + T - 1], h[j].y, acc[i][1]);
            }
        }
    };
    load(ta, ha, 0);
#pragma unroll 1
    for (int s = 0; s + 1 < NB; s += 2) {
        load(tb, hb, s + 1);
        comp(ta, ha);
        load(ta, ha, s + 2);        // one block past the end on the last trip when NB is even: in-bounds, unused
        comp(tb, hb);
    }
    if (NB & 1) { comp(ta, ha); }
}

// WARPS = 6: roles on consecutive warps.  WARPS = 8: two idle warps, placed so that (with the hardware's
// warp-slot -> scheduler mapping observed on B200, scheduler = (warp+1) & 3) LOOP has a scheduler to itself,
// P-far shares with TIMING, Q-far with COSTAS, RRC is alone.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) demod_ws2_kernel(const __grid_constant__ DemodParams p) {
    constexpr int T = kWsT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ws2Smem& sm = *reinterpret_cast<Ws2Smem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // warp -> role (6 = idle)
    const int role = (WARPS == 6) ? warp : (warp == 0 ? 0 : warp == 1 ? 1 : warp == 2 ? 2 : warp == 3 ? 3 : warp == 5 ? 4 : warp == 6 ? 5 : 6);

    for (int i = threadIdx.x; i < kIPhases * kITaps; i += blockDim.x) { sm.bank[i] = p.bank[i]; }
    for (int i = threadIdx.x; i < (kWsXEntries - kHist) * 32; i += blockDim.x) {
        sm.xs[kHist + i / 32][i % 32] = make_float2(0.f, 0.f);     // see demod_ws_kernel: zero taps must meet finite data
    }
    if (threadIdx.x < 64) { sm.ycount[threadIdx.x >> 5][lane] = 0; }

    int ch = blockIdx.x * 32 + lane;
    const bool active = ch < p.n_channels;
    if (!active) { ch = p.n_channels - 1; }
    tdm_channel_state* __restrict__ sp = p.states + ch;
    const int count = p.count;
    const int nblk = (count + T - 1) / T;
    if (role == 1) {
        const float2* xh = reinterpret_cast<const float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) { sm.xs[m][lane] = xh[m]; }
    }
    if (role == 3) {
        const float2* rh = reinterpret_cast<const float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { sm.rs[j][lane] = rh[j]; }
    }

    // ---- role-private state
    float g = 0.f, fph = 0.f, ffr = 0.f, yr = 0.f, yi = 0.f;
    float2 cur[T], nxt[T];
    const float2* __restrict__ in = row_input(p, ch);
    LoopConsts lc = {};
    float tria[T], trib[T], mida[2 * T - 1], midb[2 * T - 1];
    if (role == 0) {
        g = sp->agc_gain; fph = sp->fll_phase; ffr = sp->fll_freq;
#pragma unroll
        for (int i = 0; i < T; ++i) { cur[i] = (i < count) ? __ldg(in + i) : make_float2(0.f, 0.f); }
        lc = load_loop_consts(p);
#pragma unroll
        for (int q = 0; q < T; ++q) { tria[q] = pin(p.be_a[kHist - q]); trib[q] = pin(p.be_b[kHist - q]); }
#pragma unroll
        for (int c = 0; c < 2 * T - 1; ++c) { mida[c] = pin(p.be_a[kHist - 2 * T + 1 + c]); midb[c] = pin(p.be_b[kHist - 2 * T + 1 + c]); }
        yr = mul_rn(cur[0].x, g); yi = mul_rn(cur[0].y, g);
    }
    SymConsts kc = {};
    float mu = 0.f, om = 0.f;
    int offset = 0, nsym_t = 0;
    if (role == 4) {
        mu = sp->tr_mu; om = sp->tr_omega; offset = sp->tr_offset;
        kc = load_sym_consts(p);
    }
    SymbolState st;
    float err_blocks[TDM_SYNC_BLOCKS];
    const int nsym0 = p.accumulate ? p.out_counts[ch] : 0;       // time-sliced calls append to the rows (tdm_api.cu)
    const long long out_base = (long long)ch * p.out_stride + nsym0;
    if (role == 5) {
        st.mu = 0.f; st.om = 0.f; st.offset = 0;
        st.cph = sp->costas_phase; st.cfr = sp->costas_freq; st.ph2 = sp->costas_ph2;
        st.prev = sp->prev_sym; st.err_ptr = sp->err_ptr; st.err_disp = sp->err_disp;
        st.err_partial = sp->err_partial; st.standarderr = sp->standarderr; st.sync = sp->sync;
        st.nsym = 0; st.out_room = (int)p.out_stride - nsym0;
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { err_blocks[j] = sp->err_blocks[j]; }
        kc = load_sym_consts(p);
    }
    __syncthreads();

#ifdef TDM_ROLE_TIMING
    long long aux_cycles = 0, work_cycles = 0;
    const long long t_begin = clock64();
#endif
#pragma unroll 1
    for (int t = -1; t <= nblk + 2; ++t) {
#ifdef TDM_ROLE_TIMING
        const long long c0 = clock64();
#endif
        if (role == 0) {
            // ================= LOOP: block b = t =================
            if (t >= 0 && t < nblk) {
                const int n0 = t * T;
                const int valid = min(T, count - n0);
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const int n = n0 + T + i;
                    nxt[i] = (n < count) ? __ldg(in + n) : make_float2(0.f, 0.f);
                }
                float acc[T][4];
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const float2 pf = sm.pfar[t & 1][i][lane], qf = sm.qfar[t & 1][i][lane];
                    acc[i][0] = pf.x; acc[i][1] = pf.y; acc[i][2] = qf.x; acc[i][3] = qf.y;
                }
                {   // previous block's 8 samples: taps 56 + j - i = mid[7 + j - i]
                    const int slot = (t + 7) & (kWsXSlots - 1);
                    float2 h[T];
#pragma unroll
                    for (int j = 0; j < T; ++j) { h[j] = sm.xs[slot * T + j][lane]; }
#pragma unroll
                    for (int j = 0; j < T; ++j) {
#pragma unroll
                        for (int i = 0; i < T; ++i) {
                            acc[i][0] = fma_rn(mida[T - 1 + j - i], h[j].x, acc[i][0]);
                            acc[i][1] = fma_rn(mida[T - 1 + j - i], h[j].y, acc[i][1]);
                            acc[i][2] = fma_rn(midb[T - 1 + j - i], h[j].x, acc[i][2]);
                            acc[i][3] = fma_rn(midb[T - 1 + j - i], h[j].y, acc[i][3]);
                        }
                    }
                }
                const int xslot = (t + 8) & (kWsXSlots - 1);
#ifdef TDM_ROLE_TIMING
                aux_cycles += clock64() - c0;
#endif
#pragma unroll 1
                for (int i = 0; i < valid; ++i) {
                    // FLL recurrence on the already-scaled sample y = in * g
                    float sn, cs;
                    sincos_canon(fph, sn, cs);
                    const float xr = fma_rn(yr, cs, mul_rn(yi, sn));
                    const float xi = fma_rn(yi, cs, -mul_rn(yr, sn));
                    sm.xs[xslot * T + i][lane] = make_float2(xr, xi);
#pragma unroll
                    for (int q = 0; q < T; ++q) {
                        acc[q][0] = fma_rn(tria[q], xr, acc[q][0]);
                        acc[q][1] = fma_rn(tria[q], xi, acc[q][1]);
                        acc[q][2] = fma_rn(trib[q], xr, acc[q][2]);
                        acc[q][3] = fma_rn(trib[q], xi, acc[q][3]);
                    }
                    fll_update(lc, acc[0][0], acc[0][1], acc[0][2], acc[0][3], fph, ffr);
                    // AGC recurrence, one sample ahead: gain after this sample, then the next sample's product
                    const float amp = sqrt_rn_nobranch(fma_rn(yr, yr, mul_rn(yi, yi)));
                    g = fma_rn(sub_rn(lc.agc_set, amp), lc.agc_rate, g);
                    g = g > lc.agc_max ? lc.agc_max : g;
#pragma unroll
                    for (int q = 0; q < T - 1; ++q) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) { acc[q][c] = acc[q + 1][c]; }
                        cur[q] = cur[q + 1];
                    }
                    yr = mul_rn(cur[0].x, g); yi = mul_rn(cur[0].y, g);
                }
#pragma unroll
                for (int i = 0; i < T; ++i) { cur[i] = nxt[i]; }
                yr = mul_rn(cur[0].x, g); yi = mul_rn(cur[0].y, g);
            }
        } else if (role == 1 || role == 2) {
            // ================= P-far / Q-far: block b = t + 1 =================
            const int b = t + 1;
            if (b < nblk) {
                float acc[T][2];
#pragma unroll
                for (int i = 0; i < T; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
                if (role == 1) { ws_fir_blocks2<kWsFar, 0>(p, sm.xs, lane, b, acc); }
                else { ws_fir_blocks2<kWsFar, 1>(p, sm.xs, lane, b, acc); }
                float2 (*dst)[32] = (role == 1) ? sm.pfar[b & 1] : sm.qfar[b & 1];
#pragma unroll
                for (int i = 0; i < T; ++i) { dst[i][lane] = make_float2(acc[i][0], acc[i][1]); }
            }
        } else if (role == 3) {
            // ================= RRC: block b = t - 1 =================
            const int b = t - 1;
            if (b >= 0 && b < nblk) {
                float acc[T][2];
#pragma unroll
                for (int i = 0; i < T; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
                ws_fir_blocks2<kWsFar + 2, 2>(p, sm.xs, lane, b, acc);
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    sm.rs[(kITaps - 1 + b * T + i) & (kWsREntries - 1)][lane] = make_float2(acc[i][0], acc[i][1]);
                }
            }
        } else if (role == 6) {
            // idle warp: only keeps the barrier count
        } else if (role == 4) {
            // ================= TIMING: symbols whose newest input sample lies in block t - 2 =================
            if (t >= 2) {
                const int lim = min(count, (t - 1) * T);
                while (offset < lim) {
                    const float2 y = timing_step<kWsREntries>(kc, sm.bank, &sm.rs[0][0], lane, mu, om, offset);
                    sm.ys[nsym_t & (kYRing - 1)][lane] = y;
                    ++nsym_t;
                }
                sm.ycount[t & 1][lane] = nsym_t;
            }
        } else {
            // ================= COSTAS: the symbols TIMING finished during tick t - 1 =================
            if (t >= 3) {
                const int target = sm.ycount[(t - 1) & 1][lane];
                while (st.nsym < target) {
                    const float2 y = sm.ys[st.nsym & (kYRing - 1)][lane];
                    costas_step(p, kc, y, st, err_blocks, active, out_base);
                }
            }
        }
#ifdef TDM_ROLE_TIMING
        work_cycles += clock64() - c0;
#endif
        __syncthreads();
    }
#ifdef TDM_ROLE_TIMING
    if (blockIdx.x == 3 && lane == 0) {
        unsigned wid, smid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("role %d (hw warp slot %u on SM %u): work %lld of %lld cycles (%.1f%%), per tick %lld (before sample loop: %lld)\n", role, wid, smid, work_cycles,
               clock64() - t_begin, 100.0 * work_cycles / (double)(clock64() - t_begin), work_cycles / (nblk + 4), aux_cycles / (nblk + 4));
    }
#endif

    // ---- carry state out
    if (!active) { return; }
    if (role == 0) {
        sp->agc_gain = g; sp->fll_phase = fph; sp->fll_freq = ffr;
        sp->n_samples += (unsigned long long)count;
    } else if (role == 1) {
        float2* xh = reinterpret_cast<float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) {
            const long long q = (long long)count + m;
            xh[m] = sm.xs[(int)((q / T) & (kWsXSlots - 1)) * T + (int)(q % T)][lane];
        }
    } else if (role == 3) {
        float2* rh = reinterpret_cast<float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { rh[j] = sm.rs[(count + j) & (kWsREntries - 1)][lane]; }
    } else if (role == 4) {
        sp->tr_mu = mu; sp->tr_omega = om; sp->tr_offset = offset - count;
    } else if (role == 5) {
        sp->costas_phase = st.cph; sp->costas_freq = st.cfr; sp->costas_ph2 = st.ph2;
        sp->prev_sym = st.prev; sp->err_ptr = st.err_ptr; sp->err_disp = st.err_disp;
        sp->err_partial = st.err_partial; sp->standarderr = st.standarderr; sp->sync = st.sync;
        sp->n_symbols += (unsigned long long)st.nsym;
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { sp->err_blocks[j] = err_blocks[j]; }
        p.out_counts[ch] = nsym0 + st.nsym;
    }
}

int launch_ws2(const DemodParams& p_in, cudaStream_t stream, int warps) {
    DemodParams p = p_in;
    const float* src[3] = { p.be_a, p.be_b, p.rrc };
    for (int f = 0; f < 3; ++f) {
        for (int j = 0; j < kTapPad; ++j) {
            const int k = j - (kWsT - 1);
            p.tpad[f][j] = (k >= 0 && k < kTaps) ? src[f][k] : 0.f;
        }
    }
    const int grid = (p.n_channels + 31) / 32;
    if (warps == 8) {
        cudaFuncSetAttribute(demod_ws2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ws2Smem));
        demod_ws2_kernel<8><<<grid, 256, sizeof(Ws2Smem), stream>>>(p);
    } else {
        cudaFuncSetAttribute(demod_ws2_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ws2Smem));
        demod_ws2_kernel<6><<<grid, 192, sizeof(Ws2Smem), stream>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---------------------------------------------------------------------------------------
// Variant ws3: the ws8b pipeline rebuilt around FFMA2 and a straight-line recurrence.
//
// What profiling ws8b showed (profiles/r01_ws8b_*): 3635 cycles per 8-sample tick against a
// ~1000-cycle dependency floor; the LOOP warp issued ~1000 instructions per tick (64 FFMA per
// sample for the newest terms in a rolled shift-register loop, 35 MOVs per sample to shift it),
// the FIR warps ran at 1.38 cycles per FMA (FFMA with a uniform operand is issue limited), and
// warps that shared a scheduler with LOOP delayed it.  Changes:
//   * every FIR chain advances (re, im) together with FFMA2 (fma2_rn): half the issue slots, and
//     1.15 cycles per FMA measured in isolation (tools/ubench/ubench_ffma2.cu);
//   * LOOP's tick is ONE basic block for a full block of 8 samples: no shift register, the
//     triangular own-block part costs 36 FFMA2 per chain pair instead of 64, and ptxas is free
//     to sink the 128 "previous block" FFMA2 into the latency shadow of the sin/cos -> rotate ->
//     band-edge -> loop-filter chain.  A partial last block takes the old rolled form;
//   * the matched filter is split over two warps (outputs 0..3 / 4..7 of a block);
//   * 8 warps, placed so that LOOP's scheduler carries nothing else (or only the lightest role);
//   * the input is prefetched into L2 four ticks ahead and loaded one tick ahead.
// Chains and their term order are unchanged (far -> previous block -> own block, ascending taps),
// so the results are bit-identical to every other variant and to the canonical-order checker.
// ---------------------------------------------------------------------------------------
enum Ws3Role { kRLoop = 0, kRPfar = 1, kRQfar = 2, kRRrcA = 3, kRRrcB = 4, kRTiming = 5, kRCostas = 6, kRSlicer = 7, kRIdle = 8, kRAgc = 9, kRMid = 10 };

// warp -> role.  Warps w, w+4, w+8 share a scheduler (SMSP = warp slot mod 4 up to a rotation).
template <int PLACEMENT>
struct Ws3Placement {
    static constexpr unsigned long long r0 = kRLoop, rP = kRPfar, rQ = kRQfar, rA = kRRrcA, rB = kRRrcB, rT = kRTiming, rC = kRCostas,
                                        rS = kRSlicer, rI = kRIdle, rG = kRAgc, rM = kRMid;
    // one role per nibble, warp 0 in the lowest (a table indexed by the warp number would live in local memory).
    // 12 warps; columns = schedulers:        SMSP a     SMSP b     SMSP c     SMSP d
    static constexpr unsigned long long tab =
        PLACEMENT == 0 ? (r0 | rP << 4 | rQ << 8 | rA << 12 |  rG << 16 | rT << 20 | rC << 24 | rB << 28 |  rS << 32 | rM << 36 | rI << 40 | rI << 44)
      : PLACEMENT == 1 ? (r0 | rP << 4 | rQ << 8 | rA << 12 |  rG << 16 | rT << 20 | rC << 24 | rB << 28 |  rS << 32 | rI << 36 | rM << 40 | rI << 44)
      : PLACEMENT == 2 ? (r0 | rP << 4 | rQ << 8 | rA << 12 |  rS << 16 | rT << 20 | rC << 24 | rB << 28 |  rM << 32 | rG << 36 | rI << 40 | rI << 44)
      :                  (r0 | rP << 4 | rQ << 8 | rA << 12 |  rG << 16 | rT << 20 | rC << 24 | rB << 28 |  rM << 32 | rS << 36 | rI << 40 | rI << 44);
    static constexpr int warps = 12;
};
template <int PLACEMENT>
__device__ __forceinline__ int ws3_role_of_warp(int warp) {
    return (int)((Ws3Placement<PLACEMENT>::tab >> (4 * warp)) & 0xfull);
}

// (re, im) chains of one real-tap filter for NI consecutive outputs of a block, over NB x-ring blocks
// starting at linear block qblock0.  Row f of tpad (T-1 leading zeros) holds the taps; c0 = T - 1 - (last
// output index) is the lowest table column this output range meets.  f and c0 are RUN-TIME (warp-uniform)
// values on purpose: P-far and Q-far, and the two matched-filter halves, then execute the same instructions,
// which keeps the kernel's hot code inside the instruction cache.
template <int NB, int NI>
__device__ __forceinline__ void ws3_fir_blocks(const DemodParams& p, const float2 (*xs)[32], int lane,
                                               int qblock0, int f, int c0, float2 (&acc)[NI]) {
    constexpr int T = kWsT;
    constexpr int NC = T + NI - 1;                     // columns met: j - i + NI - 1 for j < T, i < NI
    float ta[NC], tb[NC];
    float2 ha[T], hb[T];
    const float* __restrict__ row = p.tpad[0] + f * kTapPad + c0;
    auto load = [&](float (&tt)[NC], float2 (&h)[T], int s) {
#pragma unroll
        for (int c = 0; c < NC; ++c) { tt[c] = row[s * T + c]; }
        const int slot = (qblock0 + s) & (kWsXSlots - 1);
#pragma unroll
        for (int j = 0; j < T; ++j) { h[j] = xs[slot * T + j][lane]; }
    };
    auto comp = [&](const float (&tt)[NC], const float2 (&h)[T]) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
#pragma unroll
            for (int i = 0; i < NI; ++i) { acc[i] = fma2_rn(tt[j - i + NI - 1], h[j], acc[i]); }
        }
    };
    load(ta, ha, 0);
#pragma unroll 1
    for (int s = 0; s + 1 < NB; s += 2) {
        load(tb, hb, s + 1);
        comp(ta, ha);
        load(ta, ha, s + 2);        // one block past the end on the last trip when NB is even: in-bounds, unused
        comp(tb, hb);
    }
    if (NB & 1) { comp(ta, ha); }
}

// timing_step with the three interpolator chains advanced pairwise by FFMA2 (same terms, same order).
// The polyphase bank is read from a per-lane replica, bank4[(phase * 2 + half) * 32 + lane] (float4): lanes
// sit on different phases, and rows of the plain 128 x 8 table collide 8 ways on shared-memory banks, which
// put ~100 cycles of LSU time per symbol on this recurrence; in the replica every quarter-warp access is
// conflict free whatever the phases are.
template <int RE>
__device__ __forceinline__ float2 timing_step2(const SymConsts& kc, const float4* __restrict__ bank4,
                                               const float2* rs, int lane, float& mu, float& om, int& offset) {
    // phase = clamp(floor(mu*128), 0, 127) (complex_fd.cpp:101); clamped as a float first (see timing_step),
    // then ONE conversion: floor commutes with a clamp to integer bounds
    const int ph = __float2int_rd(fminf(fmaxf(mul_rn(mu, (float)kIPhases), 0.0f), (float)(kIPhases - 1)));
    const int plo = max(ph - 1, 0);
    const int phi = min(ph + 1, kIPhases - 1);
    const float4 t0a = bank4[(ph * 2) * 32 + lane], t0b = bank4[(ph * 2 + 1) * 32 + lane];
    const float4 t1a = bank4[(phi * 2) * 32 + lane], t1b = bank4[(phi * 2 + 1) * 32 + lane];
    const float4 t2a = bank4[(plo * 2) * 32 + lane], t2b = bank4[(plo * 2 + 1) * 32 + lane];
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
    const float dscale = (phi - plo == 2) ? 0.5f : 1.0f;
    const float dre = mul_rn(sub_rn(a.x, b.x), dscale);
    const float dim = mul_rn(sub_rn(a.y, b.y), dscale);
    float terr = add_rn(y.x > 0.f ? dre : -dre, y.y > 0.f ? dim : -dim);
    terr = clampf(terr, -1.0f, 1.0f);
    om = clampf(fma_rn(kc.tr_beta, terr, om), kc.tr_min, kc.tr_max);
    mu = add_rn(mu, fma_rn(kc.tr_alpha, terr, om));
    float delta = floorf(mu);
    delta = (delta >= 0.0f) ? delta : 1.0f;            // non-finite guard, see timing_step
    delta = fminf(delta, 1048576.0f);
    offset += (int)delta;
    mu = sub_rn(mu, delta);
    return y;
}

// The carrier-recovery half of costas_step (pi4dqpsk_costas.cpp:5-28), branch free: returns the de-rotated
// symbol (PI4DQPSK's `out`), advances (phase, freq, ph2).  ph2's wrap forms both candidates and selects.
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

// The slicer half of costas_step (dqpsk_sym_extr.cpp:4-55, bit_unpacker.cpp:6-7) for the first n (<= NS)
// symbols of ring `us` starting at symbol index sl.nsym: NS predicated copies in straight-line code, so the
// independent per-symbol work (lock metric polynomial, decisions, address arithmetic, stores) of several
// symbols overlaps.  The 256-symbol block rotation of the lock metric can fire at most once in NS <= 255
// symbols: it is captured with selects and carried out once at the end (nothing in between reads it).
struct SlicerState {
    uint32_t prev, err_ptr, err_disp;
    float err_partial, standarderr;
    uint32_t sync;
    int nsym;
    int out_room;
};
template <int NS, int RING>
__device__ __forceinline__ void slicer_symbols(const DemodParams& p, const float2 (*us)[32], int lane, int n, SlicerState& sl,
                                               float* __restrict__ err_blocks, bool active, long long out_base) {
    bool crossed = false;
    float saved_partial = 0.f;
    uint32_t saved_ptr = 0;
    const int base = sl.nsym;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const bool v = k < n;
        // (predicated: entries past n may be the ones COSTAS is writing during this very tick -- the value would be
        // discarded anyway, but an unguarded read is a shared-memory race as far as compute-sanitizer can tell)
        const float2 u = v ? us[(base + k) & (RING - 1)][lane] : make_float2(0.f, 0.f);
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
        if (v && active && base + k < sl.out_room) {   // rows are sized by tdm_max_symbols(); never write past one
            const long long o = out_base + base + k;
            if (p.syms) { p.syms[o] = u; }
            if (p.dibits) { p.dibits[o] = (uint8_t)db; }
            if (p.bits) { reinterpret_cast<uchar2*>(p.bits)[o] = make_uchar2((uint8_t)((db >> 1) & 1u), (uint8_t)(db & 1u)); }
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

constexpr int kSymRing = 16;      // >= symbols in flight between two symbol-rate roles (<= 5 per tick, two ticks)
struct Ws3Smem {
    float4 bank4[kIPhases * 2][32];   // per-lane replica of the 128 x 8 interpolator bank (see timing_step2)
    float2 xs[kWsXEntries][32];
    float2 rs[kWsREntries][32];
    float2 pfar[2][kWsT][32];
    float2 qfar[2][kWsT][32];
    float2 ysc[2][kWsT][32];       // AGC -> LOOP: gain-scaled input samples of a block
    float2 ys[kSymRing][32];       // TIMING -> COSTAS: interpolated symbols
    float2 us[kSymRing][32];       // COSTAS -> SLICER: carrier-corrected symbols
    int ycount[2][32];
    int ucount[2][32];
};

// All warps of the CTA meet here once per tick.  Spelled as the PTX barrier because every role runs its
// OWN tick loop (no per-tick role dispatch, no registers of other roles live): the warps arrive at barrier 0
// from different program counters, whole warps at a time, the same number of times (ws3_ticks()).
// The thread count is spelled out (every ws3 placement has 12 warps): the barrier is reached from different program
// counters, which is exactly what the counted form is for (and what compute-sanitizer's synccheck accepts).
__device__ __forceinline__ void ws3_tick_barrier() { asm volatile("bar.sync 0, 384;" ::: "memory"); }
__device__ __forceinline__ int ws3_last_tick(int nblk) { return nblk + 3; }   // ticks run t = -1 .. nblk + 3
// Named barrier 1 links MID (arrives, does not wait) and LOOP (waits) once per tick: 64 threads.
__device__ __forceinline__ void ws3_mid_arrive() { asm volatile("bar.arrive 1, 64;" ::: "memory"); }
__device__ __forceinline__ void ws3_mid_wait() { asm volatile("bar.sync 1, 64;" ::: "memory"); }
constexpr int kMidOwn = 2;     // outputs of a block whose previous-block terms LOOP adds itself (needed before MID can deliver)

template <int PLACEMENT>
__global__ void __launch_bounds__(Ws3Placement<PLACEMENT>::warps * 32) demod_ws3_kernel(const __grid_constant__ DemodParams p) {
    constexpr int T = kWsT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ws3Smem& sm = *reinterpret_cast<Ws3Smem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int role = ws3_role_of_warp<PLACEMENT>(threadIdx.x >> 5);

    {
        const float4* __restrict__ b4 = reinterpret_cast<const float4*>(p.bank);
        for (int i = threadIdx.x; i < kIPhases * 2 * 32; i += blockDim.x) { sm.bank4[i / 32][i % 32] = __ldg(b4 + i / 32); }
    }
    for (int i = threadIdx.x; i < (kWsXEntries - kHist) * 32; i += blockDim.x) {
        sm.xs[kHist + i / 32][i % 32] = make_float2(0.f, 0.f);     // see demod_ws_kernel: zero taps must meet finite data
    }
    if (threadIdx.x < 64) { sm.ycount[threadIdx.x >> 5][lane] = 0; sm.ucount[threadIdx.x >> 5][lane] = 0; }
    for (int i = threadIdx.x; i < kSymRing * 32; i += blockDim.x) {
        sm.ys[i / 32][i % 32] = make_float2(0.f, 0.f);
        sm.us[i / 32][i % 32] = make_float2(0.f, 0.f);
    }

    int ch = blockIdx.x * 32 + lane;
    const bool active = ch < p.n_channels;
    if (!active) { ch = p.n_channels - 1; }
    tdm_channel_state* __restrict__ sp = p.states + ch;
    const int count = p.count;
    const int nblk = (count + T - 1) / T;
    const int t_last = ws3_last_tick(nblk);
    if (role == kRPfar) {
        const float2* xh = reinterpret_cast<const float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) { sm.xs[m][lane] = xh[m]; }
    }
    if (role == kRRrcA) {
        const float2* rh = reinterpret_cast<const float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { sm.rs[j][lane] = rh[j]; }
    }
    __syncthreads();

    if (role == kRAgc) {
        // ================= AGC: FastAGC recurrence [A.3], block b = t + 1 (one tick ahead of LOOP) =================
        // The gain loop does not depend on anything downstream, so it runs as its own role: LOOP then carries
        // only the FLL recurrence (no sqrt chain, no global loads, ~25 % fewer instructions per sample).
        float g = sp->agc_gain;
        const float2* __restrict__ in = row_input(p, ch);
        const LoopConsts lc = load_loop_consts(p);
        float2 cur[T], nxt[T];
#pragma unroll
        for (int i = 0; i < T; ++i) { cur[i] = (i < count) ? __ldg(in + i) : make_float2(0.f, 0.f); }
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            const int b = t + 1;
            if (b < nblk) {
                const int n0 = b * T;
                const int valid = min(T, count - n0);
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const int n = n0 + T + i;
                    nxt[i] = (n < count) ? __ldg(in + n) : make_float2(0.f, 0.f);
                }
                if (n0 + 5 * T < count) { asm volatile("prefetch.global.L2 [%0];" :: "l"(in + n0 + 5 * T)); }
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const float yr = mul_rn(cur[i].x, g), yi = mul_rn(cur[i].y, g);
                    sm.ysc[b & 1][i][lane] = make_float2(yr, yi);
                    const float amp = sqrt_rn_nobranch(fma_rn(yr, yr, mul_rn(yi, yi)));
                    float gn = fma_rn(sub_rn(lc.agc_set, amp), lc.agc_rate, g);
                    gn = gn > lc.agc_max ? lc.agc_max : gn;
                    g = (i < valid) ? gn : g;            // samples past the end of the call do not exist
                }
#pragma unroll
                for (int i = 0; i < T; ++i) { cur[i] = nxt[i]; }
            }
            ws3_tick_barrier();
        }
        if (active) {
            sp->agc_gain = g;
            sp->n_samples += (unsigned long long)count;
        }
    } else if (role == kRLoop) {
        // ================= LOOP: FLL recurrence on the gain-scaled samples, block b = t =================
        float fph = sp->fll_phase, ffr = sp->fll_freq;
        const LoopConsts lc = load_loop_consts(p);
        ws3_tick_barrier();                                  // tick t = -1
#pragma unroll 1
        for (int t = 0; t <= t_last; ++t) {
            if (t < nblk) {
                const int valid = min(T, count - t * T);
                float2 ysc[T];
#pragma unroll
                for (int i = 0; i < T; ++i) { ysc[i] = sm.ysc[t & 1][i][lane]; }
                // chains of this block's outputs: far part from the P/Q warps, then the previous block's 8 samples
                // (sample j meets output i at tap 56 + j - i).  LOOP adds those only for the first kMidOwn outputs;
                // the MID warp does the other outputs concurrently and hands them over through pfar/qfar (barrier 1).
                float2 accP[T], accQ[T];
                const int mslot = (t + 7) & (kWsXSlots - 1);     // ring block of sample block t-1
                const int xslot = (t + 8) & (kWsXSlots - 1);
                if (valid == T) {
                    {
                        float2 h[T];
#pragma unroll
                        for (int j = 0; j < T; ++j) { h[j] = sm.xs[mslot * T + j][lane]; }
#pragma unroll
                        for (int i = 0; i < kMidOwn; ++i) { accP[i] = sm.pfar[t & 1][i][lane]; accQ[i] = sm.qfar[t & 1][i][lane]; }
#pragma unroll
                        for (int j = 0; j < T; ++j) {
#pragma unroll
                            for (int i = 0; i < kMidOwn; ++i) {
                                accP[i] = fma2_rn(p.be_a[kHist - T + j - i], h[j], accP[i]);
                                accQ[i] = fma2_rn(p.be_b[kHist - T + j - i], h[j], accQ[i]);
                            }
                        }
                    }
                    float2 xo[kMidOwn];      // the block's first samples: their terms of the later outputs are added after the hand-over
                    // the block's own samples (sample i meets output q >= i at tap 64 + i - q), straight line
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        if (i == kMidOwn) {
                            ws3_mid_wait();
#pragma unroll
                            for (int q = kMidOwn; q < T; ++q) { accP[q] = sm.pfar[t & 1][q][lane]; accQ[q] = sm.qfar[t & 1][q][lane]; }
#pragma unroll
                            for (int j = 0; j < kMidOwn; ++j) {
#pragma unroll
                                for (int q = kMidOwn; q < T; ++q) {
                                    accP[q] = fma2_rn(p.be_a[kHist + j - q], xo[j], accP[q]);
                                    accQ[q] = fma2_rn(p.be_b[kHist + j - q], xo[j], accQ[q]);
                                }
                            }
                        }
                        float sn, cs;
                        sincos_canon(fph, sn, cs);
                        const float yr = ysc[i].x, yi = ysc[i].y;
                        const float2 x = make_float2(fma_rn(yr, cs, mul_rn(yi, sn)), fma_rn(yi, cs, -mul_rn(yr, sn)));
                        sm.xs[xslot * T + i][lane] = x;
                        if (i < kMidOwn) { xo[i] = x; }
#pragma unroll
                        for (int q = i; q < (i < kMidOwn ? kMidOwn : T); ++q) {
                            accP[q] = fma2_rn(p.be_a[kHist + i - q], x, accP[q]);
                            accQ[q] = fma2_rn(p.be_b[kHist + i - q], x, accQ[q]);
                        }
                        fll_update(lc, accP[i].x, accP[i].y, accQ[i].x, accQ[i].y, fph, ffr);
                    }
                } else {
                    // partial last block of a call: compact rolled code.  MID has added the previous block's terms
                    // to outputs kMidOwn.. ; the first outputs get theirs here, one sample per trip; then the
                    // shift-register form in which position q always meets tap 64 - q.
                    ws3_mid_wait();
#pragma unroll
                    for (int i = 0; i < T; ++i) { accP[i] = sm.pfar[t & 1][i][lane]; accQ[i] = sm.qfar[t & 1][i][lane]; }
#pragma unroll 1
                    for (int j = 0; j < T; ++j) {
                        const float2 h = sm.xs[mslot * T + j][lane];
#pragma unroll
                        for (int i = 0; i < kMidOwn; ++i) {
                            accP[i] = fma2_rn(p.be_a[kHist - T + j - i], h, accP[i]);
                            accQ[i] = fma2_rn(p.be_b[kHist - T + j - i], h, accQ[i]);
                        }
                    }
#pragma unroll 1
                    for (int i = 0; i < valid; ++i) {
                        float sn, cs;
                        sincos_canon(fph, sn, cs);
                        const float yr = ysc[0].x, yi = ysc[0].y;
                        const float2 x = make_float2(fma_rn(yr, cs, mul_rn(yi, sn)), fma_rn(yi, cs, -mul_rn(yr, sn)));
                        sm.xs[xslot * T + i][lane] = x;
#pragma unroll
                        for (int q = 0; q < T; ++q) {
                            accP[q] = fma2_rn(p.be_a[kHist - q], x, accP[q]);
                            accQ[q] = fma2_rn(p.be_b[kHist - q], x, accQ[q]);
                        }
                        fll_update(lc, accP[0].x, accP[0].y, accQ[0].x, accQ[0].y, fph, ffr);
#pragma unroll
                        for (int q = 0; q < T - 1; ++q) { accP[q] = accP[q + 1]; accQ[q] = accQ[q + 1]; ysc[q] = ysc[q + 1]; }
                    }
                }
            }
            ws3_tick_barrier();
        }
        if (active) { sp->fll_phase = fph; sp->fll_freq = ffr; }
    } else if (role == kRMid) {
        // ================= MID: previous block's terms of outputs kMidOwn..7 of block b = t, while LOOP runs its first samples =================
        ws3_tick_barrier();                                  // tick t = -1
#pragma unroll 1
        for (int t = 0; t <= t_last; ++t) {
            if (t < nblk) {
                const int mslot = (t + 7) & (kWsXSlots - 1);
                float2 h[T], aP[T - kMidOwn], aQ[T - kMidOwn];
#pragma unroll
                for (int j = 0; j < T; ++j) { h[j] = sm.xs[mslot * T + j][lane]; }
#pragma unroll
                for (int i = 0; i < T - kMidOwn; ++i) { aP[i] = sm.pfar[t & 1][kMidOwn + i][lane]; aQ[i] = sm.qfar[t & 1][kMidOwn + i][lane]; }
#pragma unroll
                for (int j = 0; j < T; ++j) {
#pragma unroll
                    for (int i = 0; i < T - kMidOwn; ++i) {
                        aP[i] = fma2_rn(p.be_a[kHist - T + j - (kMidOwn + i)], h[j], aP[i]);
                        aQ[i] = fma2_rn(p.be_b[kHist - T + j - (kMidOwn + i)], h[j], aQ[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < T - kMidOwn; ++i) { sm.pfar[t & 1][kMidOwn + i][lane] = aP[i]; sm.qfar[t & 1][kMidOwn + i][lane] = aQ[i]; }
                ws3_mid_arrive();
            }
            ws3_tick_barrier();
        }
    } else if (role == kRPfar || role == kRQfar) {
        // ================= P-far / Q-far: the oldest 56 - i terms of block b = t + 1 =================
        const int f = (role == kRPfar) ? 0 : 1;
        float2 (*const dst)[T][32] = (role == kRPfar) ? sm.pfar : sm.qfar;
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            const int b = t + 1;
            if (b < nblk) {
                float2 acc[T];
#pragma unroll
                for (int i = 0; i < T; ++i) { acc[i] = make_float2(0.f, 0.f); }
                ws3_fir_blocks<kWsFar, T>(p, sm.xs, lane, b, f, 0, acc);
#pragma unroll
                for (int i = 0; i < T; ++i) { dst[b & 1][i][lane] = acc[i]; }
            }
            ws3_tick_barrier();
        }
        if (active && role == kRPfar) {
            float2* xh = reinterpret_cast<float2*>(sp->x_hist);
            for (int m = 0; m < kHist; ++m) {
                const long long q = (long long)count + m;
                xh[m] = sm.xs[(int)((q / T) & (kWsXSlots - 1)) * T + (int)(q % T)][lane];
            }
        }
    } else if (role == kRRrcA || role == kRRrcB) {
        // ================= RRC: block b = t - 1, outputs 0..3 (A) or 4..7 (B), all 65 taps =================
        const int i0 = (role == kRRrcA) ? 0 : T / 2;
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            const int b = t - 1;
            if (b >= 0 && b < nblk) {
                float2 acc[T / 2];
#pragma unroll
                for (int i = 0; i < T / 2; ++i) { acc[i] = make_float2(0.f, 0.f); }
                ws3_fir_blocks<kWsFar + 2, T / 2>(p, sm.xs, lane, b, 2, T / 2 - i0, acc);
#pragma unroll
                for (int i = 0; i < T / 2; ++i) {
                    sm.rs[(kITaps - 1 + b * T + i0 + i) & (kWsREntries - 1)][lane] = acc[i];
                }
            }
            ws3_tick_barrier();
        }
        if (active && role == kRRrcA) {
            float2* rh = reinterpret_cast<float2*>(sp->r_hist);
            for (int j = 0; j < kITaps - 1; ++j) { rh[j] = sm.rs[(count + j) & (kWsREntries - 1)][lane]; }
        }
    } else if (role == kRTiming) {
        // ================= TIMING: symbols whose newest input sample lies in block t - 2 =================
        float mu = sp->tr_mu, om = sp->tr_omega;
        int offset = sp->tr_offset, nsym_t = 0;
        const SymConsts kc = load_sym_consts(p);
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            if (t >= 2) {
                const int lim = min(count, (t - 1) * T);
                while (offset < lim) {
                    const float2 y = timing_step2<kWsREntries>(kc, &sm.bank4[0][0], &sm.rs[0][0], lane, mu, om, offset);
                    sm.ys[nsym_t & (kSymRing - 1)][lane] = y;
                    ++nsym_t;
                }
                sm.ycount[t & 1][lane] = nsym_t;
            }
            ws3_tick_barrier();
        }
        if (active) { sp->tr_mu = mu; sp->tr_omega = om; sp->tr_offset = offset - count; }
    } else if (role == kRCostas) {
        // ================= COSTAS: the symbols TIMING finished during tick t - 1 =================
        float cph = sp->costas_phase, cfr = sp->costas_freq, ph2 = sp->costas_ph2;
        int nsym_c = 0;
        const SymConsts kc = load_sym_consts(p);
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            if (t >= 3) {
                const int target = sm.ycount[(t - 1) & 1][lane];
                while (nsym_c < target) {
                    const float2 y = sm.ys[nsym_c & (kSymRing - 1)][lane];
                    sm.us[nsym_c & (kSymRing - 1)][lane] = costas_loop_step(kc, y, cph, cfr, ph2);
                    ++nsym_c;
                }
                sm.ucount[t & 1][lane] = nsym_c;
            }
            ws3_tick_barrier();
        }
        if (active) { sp->costas_phase = cph; sp->costas_freq = cfr; sp->costas_ph2 = ph2; }
    } else if (role == kRIdle) {
        // placeholder warp of the 12-warp placements: keeps a scheduler slot empty, only counts barriers
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) { ws3_tick_barrier(); }
    } else {
        // ================= SLICER: the symbols COSTAS finished during tick t - 1 =================
        SlicerState sl;
        sl.prev = sp->prev_sym; sl.err_ptr = sp->err_ptr; sl.err_disp = sp->err_disp;
        sl.err_partial = sp->err_partial; sl.standarderr = sp->standarderr; sl.sync = sp->sync;
        const int nsym0 = p.accumulate ? p.out_counts[ch] : 0;       // time-sliced calls append to the rows (tdm_api.cu)
        const long long out_base = (long long)ch * p.out_stride + nsym0;
        sl.nsym = 0; sl.out_room = (int)p.out_stride - nsym0;
        float err_blocks[TDM_SYNC_BLOCKS];
#pragma unroll
        for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { err_blocks[j] = sp->err_blocks[j]; }
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            if (t >= 4) {
                const int target = sm.ucount[(t - 1) & 1][lane];
                do {    // one trip for the usual 4 symbols per 8-sample tick; a second when some lane has a fifth
                    slicer_symbols<4, kSymRing>(p, sm.us, lane, min(target - sl.nsym, 4), sl, err_blocks, active, out_base);
                } while (__any_sync(0xffffffffu, sl.nsym < target));
            }
            ws3_tick_barrier();
        }
        if (active) {
            sp->prev_sym = sl.prev; sp->err_ptr = sl.err_ptr; sp->err_disp = sl.err_disp;
            sp->err_partial = sl.err_partial; sp->standarderr = sl.standarderr; sp->sync = sl.sync;
            sp->n_symbols += (unsigned long long)sl.nsym;
#pragma unroll
            for (int j = 0; j < TDM_SYNC_BLOCKS; ++j) { sp->err_blocks[j] = err_blocks[j]; }
            p.out_counts[ch] = nsym0 + sl.nsym;
        }
    }
}

int launch_ws3(const DemodParams& p_in, cudaStream_t stream, int placement) {
    DemodParams p = p_in;
    const float* src[3] = { p.be_a, p.be_b, p.rrc };
    for (int f = 0; f < 3; ++f) {
        for (int j = 0; j < kTapPad; ++j) {
            const int k = j - (kWsT - 1);
            p.tpad[f][j] = (k >= 0 && k < kTaps) ? src[f][k] : 0.f;
        }
    }
    const int grid = (p.n_channels + 31) / 32;
    auto go = [&](auto kern, int warps) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ws3Smem));
        kern<<<grid, warps * 32, sizeof(Ws3Smem), stream>>>(p);
    };
    switch (placement) {
        case 1: go(demod_ws3_kernel<1>, Ws3Placement<1>::warps); break;
        case 2: go(demod_ws3_kernel<2>, Ws3Placement<2>::warps); break;
        case 3: go(demod_ws3_kernel<3>, Ws3Placement<3>::warps); break;
        default: go(demod_ws3_kernel<0>, Ws3Placement<0>::warps); break;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---------------------------------------------------------------------------------------
// dibit packing for the multi-GPU gather: 4 symbols per byte, first symbol in bits 7..6.
// ---------------------------------------------------------------------------------------
__global__ void pack_dibits_kernel(const uint8_t* __restrict__ dibits, long long in_stride,
                                   const int* __restrict__ counts, uint8_t* __restrict__ packed,
                                   long long out_stride, long long max_bytes) {
    const int ch = blockIdx.y;
    const int n = counts[ch];
    const uint8_t* src = dibits + (long long)ch * in_stride;
    uint8_t* dst = packed + (long long)ch * out_stride;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < max_bytes;
         j += (long long)gridDim.x * blockDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long s = 4 * j + k;
            const uint32_t d = (s < n) ? (src[s] & 3u) : 0u;
            v |= d << (6 - 2 * k);
        }
        dst[j] = (uint8_t)v;
    }
}

}  // namespace

const char* demod_variant_name(int variant) {
    switch (variant) {
        case 1: return "tpc4";
        case 2: return "tpc8";
        case 3: return "tpc4x4";
        case 4: return "ws8";
        case 5: return "ws8b";
        case 6: return "ws8b-8w";
        case 7: return "ws3";
        case 8: return "ws3-p1";
        case 9: return "ws3-p2";
        case 10: return "ws3-p3";
        default: return "auto";
    }
}

int launch_demod(const DemodParams& p, int variant, cudaStream_t stream) {
    if (p.n_channels <= 0) { return 0; }
    // auto: the warp-specialised pipeline (ws3: 12 role warps behind every 32 channels, one CTA per SM) wins
    // while channels are scarce; once there are enough channels to fill every scheduler with plain
    // thread-per-channel warps, the monolithic kernel's lower instruction count per sample wins.
    if (variant == 0) { variant = (p.n_channels >= 16384) ? 2 : 8; }
    switch (variant) {
        case 1: return launch_tpc<4>(p, stream, 1);
        case 2: return launch_tpc<8>(p, stream, 1);
        case 3: return launch_tpc<4>(p, stream, 4);
        case 4: return launch_ws(p, stream);
        case 5: return launch_ws2(p, stream, 6);
        case 6: return launch_ws2(p, stream, 8);
        case 7: return launch_ws3(p, stream, 0);
        case 8: return launch_ws3(p, stream, 1);
        case 9: return launch_ws3(p, stream, 2);
        case 10: return launch_ws3(p, stream, 3);
        default: return -1;
    }
}

int launch_pack_dibits(const uint8_t* dibits, long long in_stride, const int* counts, uint8_t* packed,
                       long long out_stride, int n_channels, long long max_syms, cudaStream_t stream) {
    if (n_channels <= 0) { return 0; }
    const long long max_bytes = (max_syms + 3) / 4;
    if (max_bytes <= 0) { return 0; }
    const int threads = 256;
    long long gx = (max_bytes + threads - 1) / threads;
    if (gx > 1024) { gx = 1024; }
    dim3 grid((unsigned)gx, (unsigned)n_channels);
    pack_dibits_kernel<<<grid, threads, 0, stream>>>(dibits, in_stride, counts, packed, out_stride, max_bytes);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace tdm
