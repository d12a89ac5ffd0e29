// tdm_ws.cu -- the demodulation chain as a warp-specialised software pipeline ("ws4"), the mapping `auto` launches.
//
// Why roles.  A channel is a strict recurrence (AGC gain and FLL phase at sample rate, timing and Costas loops at
// symbol rate), so its time axis cannot be split; and there are few channels per SM (4096 channels = 28 per SM).
// Thread-per-channel (tdm_kernels.cu) then leaves one warp to issue all ~700 instructions per sample of its 32
// channels.  Here the same arithmetic -- same chains, same term order, bit-identical results -- is cut by ROLE:
// ten warps stand behind every 32 channels (lane = channel), each running its own loop over ticks of 8 samples,
// all meeting at one CTA-wide barrier per tick, handing data on through shared-memory rings:
//
//   AGC     gain recurrence [A.3], block t+1 (global loads + L2 prefetch; nothing downstream feeds back into it)
//   LOOP    FLL recurrence (fll.cpp:135-149), block t: NCO, de-rotation, the newest terms of the band-edge sums
//           P, Q, band-edge error, loop filter.  The only role ON the sample-rate recurrence; everything else exists
//           to keep its dependent chain short (see "LOOP" below)
//   MID     the previous block's terms of outputs 2..7 of block t, concurrently with LOOP's first two samples
//           (named barrier 1)
//   P-far,  the 56-i oldest terms of the P / Q sums of block t+1 (they only need x up to block t-1): 420 FFMA2 per
//   Q-far   tick each, taps in REGISTERS, fully unrolled, x samples read two at a time (LDS.128) -- one code body,
//           the role's tap set is a register-file content
//   RRC-A,  matched filter (pi4dqpsk.cpp:136), block t-1, outputs 0..3 / 4..7: 260 FFMA2 per tick each, 65 taps in
//   RRC-B   registers, no zero-tap padding (two instantiations)
//   TIMING  interpolator + timing loop (complex_fd.cpp:96-143), symbols whose newest sample lies in block t-2
//   COSTAS  carrier loop (pi4dqpsk_costas.cpp:5-28), one tick behind TIMING
//   SLICER  decisions, lock metric, differential decoder, every output format incl. 4-per-byte packing
//           (dqpsk_sym_extr.cpp:4-55, bit_unpacker.cpp:6-7), one tick behind COSTAS
//
// Every FIR chain adds its terms in ascending tap order from +0 (far -> previous block -> own block), one fma
// per term, (re, im) pairs advanced together by FFMA2 -- the canonical order of oracle/oracle_b.c.
//
// LOOP.  Per sample the recurrence is  err -> freq -> phase -> sin/cos -> rotate -> newest tap -> detector -> err.
// The canonical order prepares the NCO's range reduction one sample ahead (tdm_math.cuh), so the dependent chain
// is: loop filter (fma, clamp) -> r = r0 + freq -> polynomials -> rotate -> one FFMA2 -> detector; quadrant
// selection is applied to the INPUT sample (known in advance) instead of to sin/cos.  The rare case in which the
// prepared reduction is not usable (|r| > 0.8: the frequency jumped, e.g. a burst arriving while the AGC is wide
// open) has exact per-sample semantics too (the classic reduction of the wrapped phase); the straight-line tick
// body speculates that it does not happen, notes any lane where it did, and the tick is then REPLAYED from its
// saved entry state by the rolled, exact per-sample path -- the one that also runs partial last blocks.
//
// Shared memory: 92.5 KB per CTA (the 128 x 8 interpolator bank as an 8-way replica so that the LDS.128 of a
// quarter warp never conflict, 32 KB; x ring 32 KB; the rest hand-over buffers), so two CTAs fit an SM: with more
// than 148 x 32 rows the second CTA fills the first one's stalls.
#include "tdm_chain.cuh"
#include <cstdio>
#include <cstdlib>

namespace tdm {

namespace {

constexpr int kT = 8;                                // samples per tick
constexpr int kXSlots = 16;                          // x ring: 16 blocks of 8 samples
constexpr int kREntries = 32;                        // matched-filter output ring
constexpr int kSymRing = 16;                         // >= symbols in flight between two symbol-rate roles (<= 5 per tick, two ticks)
constexpr int kBankRep = 8;                          // copies of the interpolator bank (see timing_step)
constexpr int kMidOwn = 2;                           // outputs of a block whose previous-block terms LOOP adds itself
constexpr int kTapRow = 96;                          // padded tap row: entry c holds tap c - 7

enum Role { kRLoop = 0, kRMid, kRTiming, kRCostas, kRSlicer, kRAgc, kRPfar, kRQfar, kRRrcA, kRRrcB, kRIdle };

// warp -> role.  Warps w, w+4, w+8 share a scheduler (SMSP = warp slot mod 4, up to a rotation).  Loads below are
// FP32-pipe cycles per tick (an FFMA2 holds the pipe for two): LOOP ~360, MID 192, TIMING ~270, COSTAS ~200,
// SLICER ~80, AGC ~100, P-far 840, Q-far 840, RRC-A 520, RRC-B 520.
template <int PLACEMENT>
struct Placement {
    static constexpr unsigned long long L = kRLoop, M = kRMid, T = kRTiming, C = kRCostas, S = kRSlicer, G = kRAgc, P = kRPfar,
                                        Q = kRQfar, A = kRRrcA, B = kRRrcB;
    // one role per nibble, warp 0 in the lowest; columns = schedulers a b c d
    static constexpr unsigned long long tab =
        //                  a        b         c         d          a         b         c          d          a          b          c          d
        PLACEMENT == 0 ? (L | Q << 4 | P << 8 | A << 12 | M << 16 | G << 20 | C << 24 | B << 28 | T << 32 | S << 36)       // a: L M T | b: Q G S | c: P C | d: A B
      : PLACEMENT == 1 ? (L | Q << 4 | P << 8 | A << 12 | M << 16 | S << 20 | C << 24 | B << 28 | G << 32 | T << 36)       // a: L M G | b: Q S T | c: P C | d: A B
      : PLACEMENT == 2 ? (L | Q << 4 | P << 8 | A << 12 | M << 16 | G << 20 | T << 24 | B << 28 | C << 32 | S << 36)       // a: L M C | b: Q G S | c: P T | d: A B
      :                  (L | Q << 4 | P << 8 | A << 12 | M << 16 | T << 20 | S << 24 | B << 28 | G << 32 | C << 36);      // a: L M G | b: Q T C | c: P S | d: A B
    static constexpr int warps = 10;
};
template <int PLACEMENT>
__device__ __forceinline__ int role_of_warp(int warp) {
    return (int)((Placement<PLACEMENT>::tab >> (4 * warp)) & 0xfull);
}

struct WsSmem {
    float4 bank4[kIPhases * 2 * kBankRep];     // entry (phase * 2 + half) of copy c at [(phase * 2 + half) * 8 + c]
    float4 xs4[kXSlots * 4][32];               // FLL output ring: [slot * 4 + pair][lane] = samples 2 pair, 2 pair + 1 of a block
    float2 rs[kREntries][32];                  // matched-filter output ring
    float2 pfar[2][kT][32];                    // far (+ mid) parts of the P / Q sums of a block, double buffered by tick parity
    float2 qfar[2][kT][32];
    float2 ysc[2][kT][32];                     // AGC -> LOOP: gain-scaled input samples of a block
    float2 ys[kSymRing][32];                   // TIMING -> COSTAS: interpolated symbols
    float2 us[kSymRing][32];                   // COSTAS -> SLICER: carrier-corrected symbols
    int ycount[2][32];
    int ucount[2][32];
    float tpad[3][kTapRow];                    // band-edge a, band-edge b, matched filter: 7 zeros, the 65 taps, zeros
    unsigned long long tickbar[2];             // mbarriers: "tick t complete", by tick parity
};

// sample with linear index q (0..63 = carried history, 64 + n = new sample n) of lane's channel
__device__ __forceinline__ float2& xs_at(WsSmem& sm, long long q, int lane) {
    const int slot = (int)((q >> 3) & (kXSlots - 1)), j = (int)(q & 7);
    return reinterpret_cast<float2*>(&sm.xs4[slot * 4 + (j >> 1)][lane])[j & 1];
}
// the 8 samples of ring block `slot` (4 LDS.128)
__device__ __forceinline__ void xs_load_block(const WsSmem& sm, int slot, int lane, float2 (&h)[kT]) {
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
        const float4 v = sm.xs4[slot * 4 + pr][lane];
        h[2 * pr] = make_float2(v.x, v.y);
        h[2 * pr + 1] = make_float2(v.z, v.w);
    }
}

// The tick hand-over: all warps of the CTA meet once per tick (`bar.sync 0`; every role runs its OWN tick loop, so the
// warps arrive from different program counters, whole warps at a time, the same number of times).
//
// A finer-grained hand-over was built and measured (kept under -DTDM_TICK_EVENTS): "tick t is complete" as an EVENT --
// an mbarrier per tick parity, one release-arrive per warp, and each role waiting for event t-1 only where it first
// touches something another role produced or consumed in tick t-1 (the FIR roles before their LAST block: chains add
// oldest samples first and only the newest block of the window is written during the previous tick; AGC before it
// stores).  The FIR roles then run most of a tick ahead.  Result on B200, 4096 channels: 2300 cycles per tick against
// 2245 for the rendezvous (profiles/README.md): the SM's issue slots and pipes are shared by ten warps that are all
// busy ~90 % of a tick, so letting some run ahead moves no work off the critical resources, and the spin-waits cost
// issue slots of their own.  The rendezvous stays.
struct TickSync {
    uint32_t bar;        // shared-memory address of the two mbarriers (TDM_TICK_EVENTS builds)
    int lane;
    int nthreads;        // threads of the CTA
#ifdef TDM_ABLATE
    mutable long long t_wait = 0, t_mark = 0, t_total = 0;   // development builds: cycles spent waiting for events
#endif
    __device__ __forceinline__ void arrive(int t) const {          // end of tick t (t >= -1)
#ifndef TDM_TICK_EVENTS
        // the non-.aligned form with its thread count: warps arrive from DIFFERENT program counters (one tick loop per
        // role), which `bar.sync` tolerates in hardware but compute-sanitizer's synccheck reads as divergence
        asm volatile("barrier.sync 0, %0;" ::"r"(nthreads) : "memory");
        return;
#endif
        __syncwarp();
        if (lane == 0) {
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar + 8u * ((t + 1) & 1)) : "memory");
        }
    }
    __device__ __forceinline__ void wait(int t) const {            // event t (t >= -1)
#ifndef TDM_TICK_EVENTS
        return;
#endif
        const uint32_t a = bar + 8u * ((t + 1) & 1), parity = ((t + 1) >> 1) & 1;
#ifdef TDM_ABLATE
        const long long c0 = clock64();
#endif
        asm volatile(
            "{\n .reg .pred p;\n"
            "W_%=: mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
            " @p bra D_%=;\n bra W_%=;\n"
            "D_%=:\n}" ::"r"(a), "r"(parity) : "memory");
#ifdef TDM_ABLATE
        t_wait += clock64() - c0;
#endif
    }
    __device__ __forceinline__ void wait_prev(int t) const { if (t > -1) { wait(t - 1); } }
};
__device__ __forceinline__ int last_tick(int nblk) { return nblk + 3; }   // ticks run t = -1 .. nblk + 3
// Named barrier 1 links MID (arrives, does not wait) and LOOP (waits) once per tick: 64 threads.
__device__ __forceinline__ void mid_arrive() { asm volatile("barrier.arrive 1, 64;" ::: "memory"); }
__device__ __forceinline__ void mid_wait() { asm volatile("barrier.sync 1, 64;" ::: "memory"); }

// fll_update without the out-of-range fallback: notes in `bad` that the exact path has to redo the tick.
template <bool RE_ONLY>
__device__ __forceinline__ void fll_update_spec(const LoopConsts& lc, float2 P, float2 Q, FllState& s, bool& bad) {
    uint32_t nq; float r0;
    fll_prepare(s.ph, s.fr, nq, r0);
    const float hbe = fast_amplitude<RE_ONLY>(sub_rn(P.x, Q.y), add_rn(P.y, Q.x));
    const float lbe = fast_amplitude<RE_ONLY>(add_rn(P.x, Q.y), sub_rn(P.y, Q.x));
    const float ferr = sub_rn(hbe, lbe);
    s.fr = clampf(fma_rn(lc.fll_beta, ferr, s.fr), lc.fll_min, lc.fll_max);
    s.r = add_rn(r0, s.fr);
    s.ph = wrap_pi(add_rn(s.ph, s.fr));
    s.q = nq;
    bad = bad || !fll_r_ok(s.r);
}

// ---- the feed-forward FIR roles ------------------------------------------------------------------------------------
// A block of 8 x-ring samples h[0..7] meets NI outputs: acc[il] += tap(8 s + j - i) * h[j], i = first output + il.
// The taps come from a zero-padded copy of the role's table in shared memory (7 zeros, 65 taps, zeros), read four
// at a time as warp-wide broadcasts: tt[c] = tpad[8 s + base + c], so that the tap of (j, il) is tt[j - il + OFF]
// whatever s is -- the block loop can stay ROLLED.  (Fully unrolled bodies with all taps in registers were
// measured first: 46 KB of straight-line code per tick, every warp stalled on instruction fetch, 3x slower.  The
// whole kernel has to live in the SM's instruction cache.)  Loads of block s+1 are issued before the FFMA2 of block
// s (two register sets), so no iteration starts by waiting on shared memory.
template <int NI, int NTT>
struct FirRegs {
    float tt[NTT];
    float2 h[kT];
};
template <int NI, int NTT>
__device__ __forceinline__ void fir_load(const WsSmem& sm, const float* __restrict__ trow, int qb, int s, int lane, FirRegs<NI, NTT>& r) {
    static_assert(NTT % 4 == 0, "taps are read as float4");
    const float4* t4 = reinterpret_cast<const float4*>(trow + 8 * s);
#pragma unroll
    for (int c = 0; c < NTT / 4; ++c) {
        const float4 v = t4[c];
        r.tt[4 * c] = v.x; r.tt[4 * c + 1] = v.y; r.tt[4 * c + 2] = v.z; r.tt[4 * c + 3] = v.w;
    }
    xs_load_block(sm, (qb + s) & (kXSlots - 1), lane, r.h);
}
template <int NI, int NTT, int OFF>
__device__ __forceinline__ void fir_fma(const FirRegs<NI, NTT>& r, float2 (&acc)[NI]) {
#pragma unroll
    for (int j = 0; j < kT; ++j) {
#pragma unroll
        for (int il = 0; il < NI; ++il) { acc[il] = fma2_rn(r.tt[j - il + OFF], r.h[j], acc[il]); }
    }
}
// NB ring blocks qb .. qb+NB-1 into NI chains, oldest block first (= ascending taps).  ONE rolled body for every
// block, every role of its kind: at the edges of a chain (taps below 0 in the first block, above 64 in the matched
// filter's last) the zero padding of the tap row stands in, which costs 28 of 448 (far) / 288 (matched filter half)
// FFMA2 per tick and keeps the code of all four FIR warps under 5 KB.  `newest_ready()` is called before the
// newest block of the window (written during the previous tick) is loaded.
template <int NI, int NTT, int OFF, int NB, typename Sync>
__device__ __forceinline__ void fir_tick(const WsSmem& sm, const float* __restrict__ trow, int qb, int lane, float2 (&acc)[NI], Sync&& newest_ready) {
    static_assert(NB & 1, "odd block counts: the loop handles pairs, the newest block is done after it");
    FirRegs<NI, NTT> ra, rb;
    fir_load(sm, trow, qb, 0, lane, ra);
    // Every load below is unconditional and precedes the FFMA2 of the block before it: ptxas keeps that order (with
    // the loads under `if (s + 1 < NB)` it sank them behind the FFMA2 and the role ran 1.6x slower, measured in
    // tools/ubench/ubench_fir2.cu).
#pragma unroll 1
    for (int s = 0; s + 1 < NB; s += 2) {
        fir_load(sm, trow, qb, s + 1, lane, rb);
        fir_fma<NI, NTT, OFF>(ra, acc);
        if (s + 2 == NB - 1) { newest_ready(); }
        fir_load(sm, trow, qb, s + 2, lane, ra);
        fir_fma<NI, NTT, OFF>(rb, acc);
    }
    fir_fma<NI, NTT, OFF>(ra, acc);
}

#ifdef TDM_ABLATE
#define TDM_ROLE_ON(r) (!(p.debug_mask & (1 << (r))))
#else
#define TDM_ROLE_ON(r) true
#endif

template <int PLACEMENT, int CTAS, bool RE_ONLY, bool STRIDED = false>
__global__ void __launch_bounds__(Placement<PLACEMENT>::warps * 32, CTAS) demod_ws4_kernel(const __grid_constant__ DemodParams p) {
    constexpr int T = kT;
    constexpr int NT = Placement<PLACEMENT>::warps * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsSmem& sm = *reinterpret_cast<WsSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int role = role_of_warp<PLACEMENT>(threadIdx.x >> 5);

    {
        const float4* __restrict__ b4 = reinterpret_cast<const float4*>(p.bank);
        for (int i = threadIdx.x; i < kIPhases * 2 * kBankRep; i += NT) { sm.bank4[i] = __ldg(b4 + i / kBankRep); }
    }
    // everything past the carried history starts as zeros: nothing reads it before it is written except the chains of
    // outputs past the end of a partial last block, whose results are discarded -- they should still not be NaN soup
    for (int i = threadIdx.x; i < (kXSlots / 2) * 4 * 32; i += NT) { sm.xs4[(kXSlots / 2) * 4 + i / 32][i % 32] = make_float4(0.f, 0.f, 0.f, 0.f); }
    if (threadIdx.x < 64) { sm.ycount[threadIdx.x >> 5][lane] = 0; sm.ucount[threadIdx.x >> 5][lane] = 0; }
    for (int i = threadIdx.x; i < 3 * kTapRow; i += NT) {
        const int f = i / kTapRow, k = i % kTapRow - 7;
        const float* __restrict__ src = (f == 0) ? p.be_a : (f == 1) ? p.be_b : p.rrc;
        sm.tpad[f][i % kTapRow] = (k >= 0 && k < kTaps) ? src[k] : 0.f;
    }
    for (int i = threadIdx.x; i < kSymRing * 32; i += NT) {
        sm.ys[i / 32][i % 32] = make_float2(0.f, 0.f);
        sm.us[i / 32][i % 32] = make_float2(0.f, 0.f);
    }

    int ch = blockIdx.x * 32 + lane;
    const bool active = ch < p.n_channels;
    if (!active) { ch = p.n_channels - 1; }     // compute a duplicate, store nothing
    tdm_channel_state* __restrict__ sp = p.states + ch;
    const int count = p.count;
    const int nblk = (count + T - 1) / T;
    const int t_last = last_tick(nblk);
    if (role == kRPfar) {
        const float2* xh = reinterpret_cast<const float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) { xs_at(sm, m, lane) = xh[m]; }
    }
    if (role == kRRrcA) {
        const float2* rh = reinterpret_cast<const float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { sm.rs[j][lane] = rh[j]; }
    }
    TickSync ts = { (uint32_t)__cvta_generic_to_shared(&sm.tickbar[0]), lane, NT };
#ifdef TDM_ABLATE
    ts.t_mark = clock64();
#endif
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ts.bar), "r"(NT / 32));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ts.bar + 8u), "r"(NT / 32));
    }
    __syncthreads();

    if (role == kRAgc) {
        // ================= AGC: FastAGC recurrence [A.3], block b = t + 1 (one tick ahead of LOOP) =================
        float g = sp->agc_gain;
        const float2* __restrict__ in = row_input(p, ch);
        const LoopConsts lc = load_loop_consts(p);
        // STRIDED (a kernel instantiation of its own: the ordinary one keeps its code byte for byte -- a generic body cost
        // it 11 %): the samples of a row are sample_stride apart (instant-major input: the 32 lanes of a sample sit side
        // by side, one 256-byte row per load)
        const long long ss = STRIDED ? p.sample_stride : 1;
        const unsigned sso = (unsigned)ss;                    // sample_stride comes from a 32-bit field
        // instant-major input and a full warp of channels: the warp's 32 samples of an instant are one 256-byte row
        const bool warp_rows_contiguous = STRIDED && p.in_stride == 1 && p.rows_per_channel <= 1 && blockIdx.x * 32 + 32 <= p.n_channels;
        const float2* __restrict__ warp_base = p.iq + (long long)blockIdx.x * 32;
        float2 cur[T], nxt[T];
#pragma unroll
        for (int i = 0; i < T; ++i) { cur[i] = (i < count) ? __ldg(in + i * ss) : make_float2(0.f, 0.f); }
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            const int b = t + 1;
            if (b < nblk && TDM_ROLE_ON(kRAgc)) {
                const int n0 = b * T;
                const int valid = min(T, count - n0);
                if (!STRIDED) {
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        const int n = n0 + T + i;
                        nxt[i] = (n < count) ? __ldg(in + n) : make_float2(0.f, 0.f);
                    }
                } else {
                    // one 64-bit pointer per tick, the eight rows at uniform 32-bit offsets from it
                    const float2* __restrict__ nx = in + (long long)(n0 + T) * ss;
#pragma unroll
                    for (int i = 0; i < T; ++i) { nxt[i] = (n0 + T + i < count) ? __ldg(nx + (unsigned)i * sso) : make_float2(0.f, 0.f); }
                }
                if (!STRIDED) {
                    if (n0 + 5 * T < count) { asm volatile("prefetch.global.L2 [%0];" :: "l"(in + n0 + 5 * T)); }
                } else if (warp_rows_contiguous && n0 + 6 * T <= count) {
                    // one instruction per tick: lane l asks for the 128-byte line (l >> 3) & 1 of row 5 T + (l & 7) of its warp's channels
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(warp_base + 16 * ((lane >> 3) & 1) + (long long)(n0 + 5 * T + (lane & 7)) * ss));
                }
                float2 y[T];
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    float gn = g;
                    y[i] = agc_step(lc, cur[i], gn);
                    g = (i < valid) ? gn : g;            // samples past the end of the call do not exist
                }
                ts.wait_prev(t);                         // LOOP has read the block that lived in this buffer (tick t-1)
#pragma unroll
                for (int i = 0; i < T; ++i) { sm.ysc[b & 1][i][lane] = y[i]; }
#pragma unroll
                for (int i = 0; i < T; ++i) { cur[i] = nxt[i]; }
            } else {
                ts.wait_prev(t);
            }
            ts.arrive(t);
        }
        if (active) {
            sp->agc_gain = g;
            sp->n_samples += (unsigned long long)count;
        }
    } else if (role == kRLoop) {
        // ================= LOOP: FLL recurrence on the gain-scaled samples, block b = t =================
        FllState fs = fll_load(sp);
        const LoopConsts lc = load_loop_consts(p);
        // taps 55..64 of the band-edge pair: what the previous block's and the block's own samples meet here
        float ta[10], tb[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) { ta[k] = pin(p.be_a[55 + k]); tb[k] = pin(p.be_b[55 + k]); }
        ts.arrive(-1);
#pragma unroll 1
        for (int t = 0; t <= t_last; ++t) {
            ts.wait(t - 1);                                  // AGC's samples, the far parts and x of block t-1 are in place
            if (t < nblk && !TDM_ROLE_ON(kRLoop)) { mid_wait(); }
            if (t < nblk && TDM_ROLE_ON(kRLoop)) {
                const int valid = min(T, count - t * T);
                const int mslot = (t + 7) & (kXSlots - 1);     // ring block of sample block t-1
                const int xslot = (t + 8) & (kXSlots - 1);
                bool exact = valid != T;                       // partial last block: rolled path only
                bool waited = false;
                if (!exact) {
                    const FllState entry = fs;
                    float2 ysc[T], accP[T], accQ[T];
#pragma unroll
                    for (int i = 0; i < T; ++i) { ysc[i] = sm.ysc[t & 1][i][lane]; }
                    {   // previous block's samples (sample j meets output i at tap 56 + j - i) for the first kMidOwn outputs;
                        // MID does the others concurrently and hands them over through pfar/qfar (barrier 1)
                        float2 h[T];
                        xs_load_block(sm, mslot, lane, h);
#pragma unroll
                        for (int i = 0; i < kMidOwn; ++i) { accP[i] = sm.pfar[t & 1][i][lane]; accQ[i] = sm.qfar[t & 1][i][lane]; }
#pragma unroll
                        for (int j = 0; j < T; ++j) {
#pragma unroll
                            for (int i = 0; i < kMidOwn; ++i) {
                                accP[i] = fma2_rn(ta[1 + j - i], h[j], accP[i]);
                                accQ[i] = fma2_rn(tb[1 + j - i], h[j], accQ[i]);
                            }
                        }
                    }
                    bool bad = false;
                    float2 xo[T];
                    // the block's own samples (sample i meets output q >= i at tap 64 + i - q), straight line
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        if (i == kMidOwn) {
                            mid_wait();
#pragma unroll
                            for (int q = kMidOwn; q < T; ++q) { accP[q] = sm.pfar[t & 1][q][lane]; accQ[q] = sm.qfar[t & 1][q][lane]; }
#pragma unroll
                            for (int j = 0; j < kMidOwn; ++j) {
#pragma unroll
                                for (int q = kMidOwn; q < T; ++q) {
                                    accP[q] = fma2_rn(ta[9 + j - q], xo[j], accP[q]);
                                    accQ[q] = fma2_rn(tb[9 + j - q], xo[j], accQ[q]);
                                }
                            }
                        }
                        const float2 x = fll_derotate(fs, ysc[i]);
                        xo[i] = x;
                        if (i & 1) { sm.xs4[xslot * 4 + (i >> 1)][lane] = make_float4(xo[i - 1].x, xo[i - 1].y, x.x, x.y); }
#pragma unroll
                        for (int q = i; q < (i < kMidOwn ? kMidOwn : T); ++q) {
                            accP[q] = fma2_rn(ta[9 + i - q], x, accP[q]);
                            accQ[q] = fma2_rn(tb[9 + i - q], x, accQ[q]);
                        }
                        fll_update_spec<RE_ONLY>(lc, accP[i], accQ[i], fs, bad);
                    }
                    waited = true;
                    if (__any_sync(0xffffffffu, bad)) { fs = entry; exact = true; }
                }
                if (exact) {
                    // exact per-sample path, compact rolled code: partial last block of a call, or the replay of a tick in
                    // which some lane's prepared reduction was out of range.  MID has added the previous block's terms to
                    // outputs kMidOwn.. ; the first outputs get theirs here, one sample per trip; then the shift-register
                    // form in which position q always meets tap 64 - q.
                    if (!waited) { mid_wait(); }
                    float2 ysc[T], accP[T], accQ[T];
#pragma unroll
                    for (int i = 0; i < T; ++i) { ysc[i] = sm.ysc[t & 1][i][lane]; accP[i] = sm.pfar[t & 1][i][lane]; accQ[i] = sm.qfar[t & 1][i][lane]; }
#pragma unroll 1
                    for (int j = 0; j < T; ++j) {
                        const float2 h = xs_at(sm, (long long)mslot * T + j, lane);
#pragma unroll
                        for (int i = 0; i < kMidOwn; ++i) {
                            accP[i] = fma2_rn(p.be_a[kHist - T + j - i], h, accP[i]);
                            accQ[i] = fma2_rn(p.be_b[kHist - T + j - i], h, accQ[i]);
                        }
                    }
#pragma unroll 1
                    for (int i = 0; i < valid; ++i) {
                        const float2 x = fll_derotate(fs, ysc[0]);
                        xs_at(sm, (long long)xslot * T + i, lane) = x;
#pragma unroll
                        for (int q = 0; q < T; ++q) {
                            accP[q] = fma2_rn(ta[9 - q], x, accP[q]);
                            accQ[q] = fma2_rn(tb[9 - q], x, accQ[q]);
                        }
                        fll_update<RE_ONLY>(lc, accP[0], accQ[0], fs);
#pragma unroll
                        for (int q = 0; q < T - 1; ++q) { accP[q] = accP[q + 1]; accQ[q] = accQ[q + 1]; ysc[q] = ysc[q + 1]; }
                    }
                }
            }
            ts.arrive(t);
        }
        if (active) { fll_store(sp, fs); }
    } else if (role == kRMid) {
        // ================= MID: previous block's terms of outputs kMidOwn..7 of block b = t, while LOOP runs its first samples =================
        // sample j meets output kMidOwn + i at tap 56 + j - kMidOwn - i = 49 + (5 + j - i)
        float ta[13], tb[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) { ta[k] = pin(p.be_a[49 + k]); tb[k] = pin(p.be_b[49 + k]); }
        ts.arrive(-1);
#pragma unroll 1
        for (int t = 0; t <= t_last; ++t) {
            ts.wait(t - 1);
            if (t < nblk && !TDM_ROLE_ON(kRMid)) { mid_arrive(); }
            if (t < nblk && TDM_ROLE_ON(kRMid)) {
                const int mslot = (t + 7) & (kXSlots - 1);
                float2 h[T], aP[T - kMidOwn], aQ[T - kMidOwn];
                xs_load_block(sm, mslot, lane, h);
#pragma unroll
                for (int i = 0; i < T - kMidOwn; ++i) { aP[i] = sm.pfar[t & 1][kMidOwn + i][lane]; aQ[i] = sm.qfar[t & 1][kMidOwn + i][lane]; }
#pragma unroll
                for (int j = 0; j < T; ++j) {
#pragma unroll
                    for (int i = 0; i < T - kMidOwn; ++i) {
                        aP[i] = fma2_rn(ta[5 + j - i], h[j], aP[i]);
                        aQ[i] = fma2_rn(tb[5 + j - i], h[j], aQ[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < T - kMidOwn; ++i) { sm.pfar[t & 1][kMidOwn + i][lane] = aP[i]; sm.qfar[t & 1][kMidOwn + i][lane] = aQ[i]; }
                mid_arrive();
            }
            ts.arrive(t);
        }
    } else if (role == kRPfar || role == kRQfar) {
        // ================= P-far / Q-far: the oldest 56 - i terms of block b = t + 1 =================
        const float* __restrict__ trow = sm.tpad[(role == kRPfar) ? 0 : 1];
        float2 (*const dst)[T][32] = (role == kRPfar) ? sm.pfar : sm.qfar;
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            const int b = t + 1;
            if (b < nblk && TDM_ROLE_ON(kRPfar)) {
                float2 acc[T];
#pragma unroll
                for (int i = 0; i < T; ++i) { acc[i] = make_float2(0.f, 0.f); }
                // x of block t-1 is the newest the window holds; by then LOOP and MID are also done with the buffer written below
                // sample (s, j) meets output i at tap 8 s + j - i = entry 8 s + (j - i + 7) of the padded row
                fir_tick<T, 16, 7, 7>(sm, trow, b, lane, acc, [&] { ts.wait_prev(t); });
#pragma unroll
                for (int i = 0; i < T; ++i) { dst[b & 1][i][lane] = acc[i]; }
            } else {
                ts.wait_prev(t);
            }
            ts.arrive(t);
        }
        if (active && role == kRPfar) {
            float2* xh = reinterpret_cast<float2*>(sp->x_hist);
            for (int m = 0; m < kHist; ++m) { xh[m] = xs_at(sm, (long long)count + m, lane); }
        }
    } else if (role == kRRrcA || role == kRRrcB) {
        // ================= RRC: block b = t - 1, outputs 0..3 (A) or 4..7 (B), all 65 taps =================
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            const int b = t - 1;
            if (b >= 0 && b < nblk && TDM_ROLE_ON(kRRrcA)) {
                float2 acc[T / 2];
#pragma unroll
                for (int i = 0; i < T / 2; ++i) { acc[i] = make_float2(0.f, 0.f); }
                const int i0 = (role == kRRrcA) ? 0 : T / 2;
                // sample (s, j) meets output i0 + il at tap 8 s + j - i0 - il = entry 8 s + (4 - i0) + (j - il + 3) of the padded row
                fir_tick<T / 2, 12, 3, 9>(sm, &sm.tpad[2][4 - i0], b, lane, acc, [&] { ts.wait_prev(t); });
#pragma unroll
                for (int i = 0; i < T / 2; ++i) { sm.rs[(kITaps - 1 + b * T + i0 + i) & (kREntries - 1)][lane] = acc[i]; }
            } else {
                ts.wait_prev(t);
            }
            ts.arrive(t);
        }
        if (active && role == kRRrcA) {
            float2* rh = reinterpret_cast<float2*>(sp->r_hist);
            for (int j = 0; j < kITaps - 1; ++j) { rh[j] = sm.rs[(count + j) & (kREntries - 1)][lane]; }
        }
    } else if (role == kRTiming) {
        // ================= TIMING: symbols whose newest input sample lies in block t - 2 =================
        float mu = sp->tr_mu, om = sp->tr_omega;
        int offset = sp->tr_offset, nsym_t = 0;
        const SymConsts kc = load_sym_consts(p);
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            ts.wait_prev(t);
            if (t >= 2 && TDM_ROLE_ON(kRTiming)) {
                const int lim = min(count, (t - 1) * T);
                while (offset < lim) {
                    const float2 y = timing_step<kREntries, kBankRep>(kc, sm.bank4, &sm.rs[0][0], lane, mu, om, offset);
                    sm.ys[nsym_t & (kSymRing - 1)][lane] = y;
                    ++nsym_t;
                }
                sm.ycount[t & 1][lane] = nsym_t;
            }
            ts.arrive(t);
        }
        if (active) { sp->tr_mu = mu; sp->tr_omega = om; sp->tr_offset = offset - count; }
    } else if (role == kRCostas) {
        // ================= COSTAS: the symbols TIMING finished during tick t - 1 =================
        float cph = sp->costas_phase, cfr = sp->costas_freq, ph2 = sp->costas_ph2;
        int nsym_c = 0;
        const SymConsts kc = load_sym_consts(p);
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            ts.wait_prev(t);
            if (t >= 3 && TDM_ROLE_ON(kRCostas)) {
                const int target = sm.ycount[(t - 1) & 1][lane];
                while (nsym_c < target) {
                    const float2 y = sm.ys[nsym_c & (kSymRing - 1)][lane];
                    sm.us[nsym_c & (kSymRing - 1)][lane] = costas_loop_step(kc, y, cph, cfr, ph2);
                    ++nsym_c;
                }
                sm.ucount[t & 1][lane] = nsym_c;
            }
            ts.arrive(t);
        }
        if (active) { sp->costas_phase = cph; sp->costas_freq = cfr; sp->costas_ph2 = ph2; }
    } else if (role == kRIdle) {
        // placeholder warp (diagnostic placements): only takes part in the tick hand-over
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) { ts.wait_prev(t); ts.arrive(t); }
    } else {
        // ================= SLICER: the symbols COSTAS finished during tick t - 1 =================
        SlicerState sl;
        float err_blocks[TDM_SYNC_BLOCKS];
        slicer_load(p, sp, ch, sl, err_blocks);
#pragma unroll 1
        for (int t = -1; t <= t_last; ++t) {
            ts.wait_prev(t);
            if (t >= 4 && TDM_ROLE_ON(kRSlicer)) {
                const int target = sm.ucount[(t - 1) & 1][lane];
                do {    // two symbols per trip (rolled: the kernel has to fit the instruction cache); lanes without one idle
                    // (reads are predicated: entries past the target may be the ones COSTAS is writing during this very tick)
                    slicer_symbols<2>(p, ch, min(target - sl.nsym, 2), sl, err_blocks, active, [&](int idx, bool v) {
                        return v ? sm.us[idx & (kSymRing - 1)][lane] : make_float2(0.f, 0.f);
                    });
                } while (__any_sync(0xffffffffu, sl.nsym < target));
            }
            ts.arrive(t);
        }
        slicer_store(p, sp, ch, sl, err_blocks, active);
    }
#ifdef TDM_ABLATE
    if ((p.debug_mask & 0x10000) && blockIdx.x == 1 && lane == 0) {
        const long long total = clock64() - ts.t_mark;
        printf("role %2d warp %2d: waited %6.1f%% of %lld cycles (%lld per tick; busy %lld per tick)\n", role, (int)(threadIdx.x >> 5),
               100.0 * ts.t_wait / total, total, total / (t_last + 2), (total - ts.t_wait) / (t_last + 2));
    }
#endif
}

}  // namespace

int launch_ws4(const DemodParams& p_in, cudaStream_t stream, int placement, int ctas_per_sm) {
    const int grid = (p_in.n_channels + 31) / 32;
#ifdef TDM_ABLATE
    DemodParams pd = p_in;
    if (const char* e = getenv("TDM_DEBUG_MASK")) { pd.debug_mask = atoi(e); }
    const DemodParams& p = pd;
#else
    const DemodParams& p = p_in;
#endif
    auto go = [&](auto kern, int warps) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsSmem));
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // two CTAs need 2 x 92.5 KB
        kern<<<grid, warps * 32, sizeof(WsSmem), stream>>>(p);
    };
#define TDM_WS4_CASE(PL, CT)                                                                                       \
    if (placement == PL && ctas_per_sm == CT) {                                                                      \
        if (p.fastamp_re_only) { go(demod_ws4_kernel<PL, CT, true>, Placement<PL>::warps); }                        \
        else { go(demod_ws4_kernel<PL, CT, false>, Placement<PL>::warps); }                                         \
        return cudaGetLastError() == cudaSuccess ? 1 : -1;                                                           \
    }
    if (p.sample_stride != 1) {       // strided input: placement 0 only
        if (ctas_per_sm == 1) {
            if (p.fastamp_re_only) { go(demod_ws4_kernel<0, 1, true, true>, Placement<0>::warps); } else { go(demod_ws4_kernel<0, 1, false, true>, Placement<0>::warps); }
        } else {
            if (p.fastamp_re_only) { go(demod_ws4_kernel<0, 2, true, true>, Placement<0>::warps); } else { go(demod_ws4_kernel<0, 2, false, true>, Placement<0>::warps); }
        }
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    TDM_WS4_CASE(0, 1) TDM_WS4_CASE(1, 1) TDM_WS4_CASE(2, 1) TDM_WS4_CASE(3, 1)
    TDM_WS4_CASE(0, 2) TDM_WS4_CASE(1, 2) TDM_WS4_CASE(2, 2) TDM_WS4_CASE(3, 2)
#undef TDM_WS4_CASE
    return -1;
}

}  // namespace tdm
