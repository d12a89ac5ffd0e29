// TEST INFRASTRUCTURE -- "Oracle A": a C driver around the reference's OWN
// demodulator classes, compiled unmodified from /root/reference/src/dsp/*.cpp
// (see oracle/Makefile; nothing from the reference is copied into this repo).
//
// It is the authority for DECODED BITS.  Only tests/, __graft_entry__.smoke()
// and bench.py's CPU-baseline / --impl reference legs may load the library this
// builds (oracle/_ref/libtetra_ref.so).  The product never links or calls it.
//
// What is wired here mirrors the reference plugin's own wiring:
//   * gains and init() arguments  -> /root/reference/src/main.cpp:35-44,78-84
//   * chain of blocks             -> /root/reference/src/main.cpp:84-91
//       mainDemodulator (dsp::demod::PI4DQPSK, src/dsp/pi4dqpsk.cpp:132-140)
//       -> symbolExtractor (dsp::DQPSKSymbolExtractor, src/dsp/dqpsk_sym_extr.cpp:4-55)
//       -> bitsUnpacker    (dsp::BitUnpacker, src/dsp/bit_unpacker.cpp:4-10)
// The blocks are driven through their public process() methods on the caller's
// thread (the same methods their run() loops call, src/dsp/pi4dqpsk.h:38-50).
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <fstream>
#include <iomanip>
#include <mutex>
#include <sstream>
#include <thread>
#include <type_traits>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
// IEEE gradual underflow regardless of what -ffast-math libraries the host process loaded (see oracle_b.c).
struct FpEnvGuard {
    unsigned saved;
    FpEnvGuard() : saved(_mm_getcsr()) { _mm_setcsr(saved & ~0x8040u); }
    ~FpEnvGuard() { _mm_setcsr(saved); }
};
#else
struct FpEnvGuard {};
#endif

// Let the driver read protected loop state / tap tables for diagnostics and
// for pinning Oracle B's tap design.  Standard headers are included above so
// the macro only touches the reference's (and the stand-in's) class bodies.
#define protected public
#define private public
#include "bit_unpacker.h"
#include "dqpsk_sym_extr.h"
#include "pi4dqpsk.h"
#undef private
#undef protected

extern "C" {

typedef struct tref_params {
    double symbolrate, samplerate;
    int rrc_taps;
    double rrc_beta, agc_rate, costas_bw, fll_bw, omega_gain, mu_gain, omega_rel_limit;
} tref_params;

typedef struct tref_loop_state {
    float agc_gain;
    float fll_phase, fll_freq;
    float tr_mu, tr_omega;
    int32_t tr_offset;
    float costas_phase, costas_freq, costas_ph2;
    uint32_t prev_sym;
    float standarderr;
    int32_t sync;
} tref_loop_state;

struct tref_chain {
    dsp::demod::PI4DQPSK demod;
    dsp::DQPSKSymbolExtractor extractor;
    dsp::BitUnpacker unpacker;
    std::vector<dsp::complex_t> symtmp;
    std::vector<uint8_t> dibittmp;
    tref_params p;
};

// The plugin's compile-time constants and its mixed float/double gain
// arithmetic, /root/reference/src/main.cpp:35-44 and :78-84.
void tref_default_params(tref_params* p) {
    float recov_bandwidth = 0.00628f;       // CLOCK_RECOVERY_BW
    float recov_dampningFactor = 0.707f;    // CLOCK_RECOVERY_DAMPN_F
    float recov_denominator = (1.0f + 2.0 * recov_dampningFactor * recov_bandwidth + recov_bandwidth * recov_bandwidth);
    float recov_mu = (4.0f * recov_dampningFactor * recov_bandwidth) / recov_denominator;
    float recov_omega = (4.0f * recov_bandwidth * recov_bandwidth) / recov_denominator;
    p->symbolrate = 18000;
    p->samplerate = 36000;          // VFO_SAMPLERATE
    p->rrc_taps = 65;               // RRC_TAP_COUNT
    p->rrc_beta = 0.35f;            // RRC_ALPHA (float literal widened, as in the call)
    p->agc_rate = 0.02f;            // AGC_RATE
    p->costas_bw = 0.01f;           // COSTAS_LOOP_BANDWIDTH
    p->fll_bw = 0.006f;             // FLL_LOOP_BANDWIDTH
    p->omega_gain = recov_omega;
    p->mu_gain = recov_mu;
    p->omega_rel_limit = 0.02f;     // CLOCK_RECOVERY_REL_LIM
}

void* tref_create(const tref_params* params) {
    tref_chain* c = new tref_chain();
    if (params) { c->p = *params; } else { tref_default_params(&c->p); }
    // NULL input streams: the blocks are never start()ed, only process()ed.
    c->demod.init(NULL, c->p.symbolrate, c->p.samplerate, c->p.rrc_taps, c->p.rrc_beta, c->p.agc_rate,
                  c->p.costas_bw, c->p.fll_bw, c->p.omega_gain, c->p.mu_gain, c->p.omega_rel_limit);
    c->extractor.init(NULL);
    c->unpacker.init(NULL);
    // errorbuf is uninitialised in the reference (src/dsp/dqpsk_sym_extr.h:43);
    // both oracles define it as zeros.
    memset(c->extractor.errorbuf, 0, sizeof(c->extractor.errorbuf));
    c->symtmp.resize(STREAM_BUFFER_SIZE);
    c->dibittmp.resize(STREAM_BUFFER_SIZE);
    return c;
}

void tref_destroy(void* h) { delete (tref_chain*)h; }

void tref_reset(void* h) {
    tref_chain* c = (tref_chain*)h;
    c->demod.reset();
}

// iq: count interleaved float32 pairs.  syms (2 floats each), dibits (1/byte),
// bits (1/byte, 2 per dibit) may each be NULL.  Returns the symbol count.
// Feeds the chain in calls of at most STREAM_BUFFER_SIZE samples, like SDR++.
int64_t tref_process(void* h, int64_t count, const float* iq, float* syms, uint8_t* dibits, uint8_t* bits) {
    FpEnvGuard fpenv;
    tref_chain* c = (tref_chain*)h;
    int64_t done = 0, nsym = 0;
    while (done < count) {
        int n = (int)std::min<int64_t>(count - done, STREAM_BUFFER_SIZE);
        int ns = c->demod.process(n, (const dsp::complex_t*)(iq + 2 * done), c->symtmp.data());
        int nd = c->extractor.process(ns, c->symtmp.data(), c->dibittmp.data());
        if (syms) { memcpy(syms + 2 * nsym, c->symtmp.data(), sizeof(dsp::complex_t) * (size_t)ns); }
        if (dibits) { memcpy(dibits + nsym, c->dibittmp.data(), (size_t)nd); }
        if (bits) { c->unpacker.process(nd, c->dibittmp.data(), bits + 2 * nsym); }
        nsym += ns;
        done += n;
    }
    return nsym;
}

// The reference's own setters (src/dsp/pi4dqpsk.h:52-63), for pinning tdm_set_params:
//   1 setSymbolrate(a)  2 setSamplerate(a)  3 setRRCParams(int(a), b)  4 setAGCRate(a)  5 setCostasBandwidth(a)
//   6 setFllBandwidth(a)  7 setMMParams(a, b, c)  8 setOmegaRelLimit(a)  9 setOmegaGain(a)  10 setMuGain(a)
void tref_set(void* h, int what, double a, double b, double c3) {
    tref_chain* c = (tref_chain*)h;
    switch (what) {
        case 1: c->demod.setSymbolrate(a); break;
        case 2: c->demod.setSamplerate(a); break;
        case 3: c->demod.setRRCParams((int)a, b); break;
        case 4: c->demod.setAGCRate(a); break;
        case 5: c->demod.setCostasBandwidth(a); break;
        case 6: c->demod.setFllBandwidth(a); break;
        case 7: c->demod.setMMParams(a, b, c3); break;
        case 8: c->demod.setOmegaRelLimit(a); break;
        case 9: c->demod.setOmegaGain(a); break;
        case 10: c->demod.setMuGain(a); break;
        default: break;
    }
}

void tref_get_state(void* h, tref_loop_state* s) {
    tref_chain* c = (tref_chain*)h;
    s->agc_gain = c->demod.agc._gain;
    s->fll_phase = c->demod.fll.pcl.phase;
    s->fll_freq = c->demod.fll.pcl.freq;
    s->tr_mu = c->demod.recov.pcl.phase;
    s->tr_omega = c->demod.recov.pcl.freq;
    s->tr_offset = c->demod.recov.offset;
    s->costas_phase = c->demod.costas.pcl.phase;
    s->costas_freq = c->demod.costas.pcl.freq;
    s->costas_ph2 = c->demod.costas.ph2;
    s->prev_sym = c->extractor.prev;
    s->standarderr = c->extractor.standarderr;
    s->sync = c->extractor.sync ? 1 : 0;
}

// Tap tables as the reference designed them (for pinning Oracle B / the
// product's host-side design code).  Sizes: rrc[n], lbe[2n], hbe[2n],
// bank[phases*tapsPerPhase].  Returns n (RRC tap count).
int tref_get_taps(void* h, float* rrc, float* lbe, float* hbe, float* bank, int* phases, int* taps_per_phase) {
    tref_chain* c = (tref_chain*)h;
    int n = c->demod.rrcTaps.size;
    if (rrc) { memcpy(rrc, c->demod.rrcTaps.taps, sizeof(float) * n); }
    if (lbe) { memcpy(lbe, c->demod.fll.lbandedgerrcTaps.taps, sizeof(float) * 2 * n); }
    if (hbe) { memcpy(hbe, c->demod.fll.hbandedgerrcTaps.taps, sizeof(float) * 2 * n); }
    int P = c->demod.recov.interpBank.phaseCount, T = c->demod.recov.interpBank.tapsPerPhase;
    if (bank) {
        for (int p = 0; p < P; p++) { memcpy(bank + p * T, c->demod.recov.interpBank.phases[p], sizeof(float) * T); }
    }
    if (phases) { *phases = P; }
    if (taps_per_phase) { *taps_per_phase = T; }
    return n;
}

// Loop coefficients actually in effect (alpha/beta/limits of the three loops).
void tref_get_coeffs(void* h, float* out16) {
    tref_chain* c = (tref_chain*)h;
    auto& f = c->demod.fll.pcl;
    auto& t = c->demod.recov.pcl;
    auto& k = c->demod.costas.pcl;
    float v[16] = { f._alpha, f._beta, f._minFreq, f._maxFreq, t._alpha, t._beta, t._minFreq, t._maxFreq,
                    k._alpha, k._beta, k._minFreq, k._maxFreq, c->demod.agc._rate, c->demod.agc._setPoint,
                    c->demod.agc._maxGain, c->demod.agc._initGain };
    memcpy(out16, v, sizeof(v));
}

// Multi-channel convenience used for the all-cores CPU baseline: channel c of
// `nch` lives at iq + 2*c*count; one independent chain per channel, channels
// dealt round-robin to `nthreads` std::threads (the reference's own model is
// one plugin instance = one thread per channel).
// dibits: [nch][stride] or NULL; nsyms: [nch].
void tref_process_multi(void** chains, int nch, int64_t count, const float* iq, uint8_t* dibits, int64_t stride,
                        int64_t* nsyms, int nthreads) {
    nthreads = std::max(1, std::min(nthreads, nch));
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([=]() {
            for (int ch = t; ch < nch; ch += nthreads) {
                nsyms[ch] = tref_process(chains[ch], count, iq + 2 * (size_t)ch * (size_t)count,
                                         NULL, dibits ? dibits + (size_t)ch * (size_t)stride : NULL, NULL);
            }
        });
    }
    for (auto& x : th) { x.join(); }
}

}  // extern "C"
