// tdm_design.cpp -- host-side (double precision) filter and loop-gain design.
//
// The reference designs everything on the CPU when PI4DQPSK::init runs
// (src/dsp/pi4dqpsk.cpp:11-30) and never touches it again on the hot path, so this
// stays on the host too: the result (tdm_design) is uploaded once per handle and is
// what the kernels read.  No CUDA in this file.
//
// Pinned against the reference: tests/test_oracles.py::test_design_matches_reference
// compares every table produced here with the tables inside the reference's own
// objects (read out of oracle/_ref), element for element, bit for bit.
//
// [A.n] = SURVEY.md Appendix A item n: semantics of the SDR++-core helpers the
// reference calls (they are not vendored in the reference tree).
#include <cmath>
#include <cstring>
#include "tdm_b200.h"

namespace {

constexpr float kFlPi = 3.1415926535f;               // FL_M_PI [A.1]
constexpr double kDbPi = 3.14159265358979323846;     // DB_M_PI
constexpr double kDbSqrt2 = 1.41421356237309504880;

struct LoopGains { float alpha, beta; };

// PhaseControlLoop<float>::criticallyDamped [A.2]; T = float, so the products that
// involve a double literal are evaluated in double and rounded once.
LoopGains critically_damped(float bw) {
    const float damping = std::sqrt(2.0) / 2.0;
    const float denom = (1.0 + 2.0 * damping * bw + bw * bw);
    return { (4 * damping * bw) / denom, (4 * bw * bw) / denom };
}

double sinc_unnormalised(double x) { return x == 0.0 ? 1.0 : std::sin(x) / x; }   // math::sinc [A.1]

double nuttall_window(double n, double N) {                                       // window::nuttall [A.6]
    static const double coef[4] = { 0.355768, 0.487396, 0.144232, 0.012604 };
    double acc = 0.0, sgn = 1.0;
    for (int i = 0; i < 4; ++i, sgn = -sgn) { acc += sgn * coef[i] * std::cos((double)i * 2.0 * kDbPi * n / N); }
    return acc;
}

// taps::rootRaisedCosine<float>(count, beta, symbolrate, samplerate) [A.6], called at pi4dqpsk.cpp:18
void design_rrc(int count, double beta, double Ts, float* out) {
    const double half = count / 2.0, edge = Ts / (4.0 * beta);
    for (int i = 0; i < count; ++i) {
        const double t = (double)i - half + 0.5;
        double v;
        if (t == 0.0) {
            v = (1.0 + beta * (4.0 / kDbPi - 1.0)) / Ts;
        } else if (t == edge || t == -edge) {
            v = ((1.0 + 2.0 / kDbPi) * std::sin(kDbPi / (4.0 * beta)) + (1.0 - 2.0 / kDbPi) * std::cos(kDbPi / (4.0 * beta))) *
                beta / (Ts * kDbSqrt2);
        } else {
            const double u = 4.0 * beta * t / Ts;
            v = ((std::sin((1.0 - beta) * kDbPi * t / Ts) + std::cos((1.0 + beta) * kDbPi * t / Ts) * 4.0 * beta * t / Ts) /
                 ((1.0 - u * u) * kDbPi * t / Ts)) / Ts;
        }
        out[i] = (float)v;
    }
}

// FLL::createBandedgeFilters (src/dsp/fll.cpp:61-95).  Float arithmetic where the
// reference uses float.  The two band-edge tap sets are exact complex conjugates
// (t1 = phasor(-th)*tap, t2 = phasor(+th)*tap), so only a = Re, b = Im of the
// upper one are kept: hbe = a + jb, lbe = a - jb.  Stored reversed like the reference.
void design_bandedge(int count, int symrate, int samprate, float rolloff, float* a, float* b) {
    const double sym = symrate, samp = samprate;      // FLL::init takes ints, stores doubles (fll.h:33, fll.cpp:12-13)
    const float sps = samp / sym;
    const int M = (count / sps);
    float bb[TDM_MAX_TAPS];
    float power = 0;
    for (int i = 0; i < count; ++i) {
        const float k = -M + i * 2.0f / sps;
        const float tap = sinc_unnormalised(rolloff * k - 0.5f) + sinc_unnormalised(rolloff * k + 0.5f);
        power += tap;
        bb[i] = tap;
    }
    const int N = (count - 1.0f) / 2.0f;
    for (int i = 0; i < count; ++i) {
        const float tap = bb[i] / power;
        const float k = (-N + (int)i) / (2.0f * sps);
        const float theta = 2.0f * kFlPi * (1.0f + rolloff) * k;
        a[count - i - 1] = cosf(theta) * tap;
        b[count - i - 1] = sinf(theta) * tap;
    }
}

// COMPLEX_FD::generateInterpTaps (src/dsp/complex_fd.cpp:153-158):
// windowedSinc(P*T, hzToRads(0.5/P, 1), nuttall, P) then buildPolyphaseBank(P, .) [A.6]
void design_interp_bank(float bank[TDM_INTERP_PHASES][TDM_INTERP_TAPS]) {
    constexpr int P = TDM_INTERP_PHASES, T = TDM_INTERP_TAPS, n = P * T;
    const double omega = 2.0 * kDbPi * ((0.5 / (double)P) / 1.0);
    const double half = n / 2.0, corr = (double)P * omega / kDbPi;
    for (int i = 0; i < n; ++i) {
        const double t = (double)i - half + 0.5;
        bank[(P - 1) - (i % P)][i / P] = (float)(sinc_unnormalised(t * omega) * nuttall_window(t - half, n) * corr);
    }
}

}  // namespace

extern "C" int tdm_default_config(tdm_config* cfg) {
    if (!cfg) { return TDM_ERR_ARG; }
    // src/main.cpp:35-44 (#defines) and :78-82 (clock-recovery gains, float/double mix as written there)
    const float bw = 0.00628f, damping = 0.707f;
    const float denom = (1.0f + 2.0 * damping * bw + bw * bw);
    const float mu_gain = (4.0f * damping * bw) / denom;
    const float omega_gain = (4.0f * bw * bw) / denom;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->symbolrate = 18000;
    cfg->samplerate = 36000;
    cfg->rrc_tap_count = 65;
    cfg->rrc_beta = 0.35f;
    cfg->agc_rate = 0.02f;
    cfg->costas_bandwidth = 0.01f;
    cfg->fll_bandwidth = 0.006f;
    cfg->omega_gain = omega_gain;
    cfg->mu_gain = mu_gain;
    cfg->omega_rel_limit = 0.02f;
    return TDM_OK;
}

extern "C" int tdm_design_from_config(const tdm_config* cfg, tdm_design* d) {
    if (!cfg || !d) { return TDM_ERR_ARG; }
    const int nt = cfg->rrc_tap_count;
    if (nt < 1 || nt > TDM_MAX_TAPS) { return TDM_ERR_UNSUPPORTED; }
    if (!(cfg->samplerate > 0) || !(cfg->symbolrate > 0) || !(cfg->rrc_beta > 0)) { return TDM_ERR_ARG; }
    if (!std::isfinite(cfg->samplerate) || !std::isfinite(cfg->symbolrate) || !std::isfinite(cfg->rrc_beta) || !std::isfinite(cfg->agc_rate) ||
        !std::isfinite(cfg->costas_bandwidth) || !std::isfinite(cfg->fll_bandwidth) || !std::isfinite(cfg->omega_gain) ||
        !std::isfinite(cfg->mu_gain) || !std::isfinite(cfg->omega_rel_limit)) { return TDM_ERR_ARG; }
    if (!(cfg->omega_rel_limit >= 0.0) || !(cfg->omega_rel_limit < 1.0)) { return TDM_ERR_ARG; }
    if (cfg->flags & ~TDM_CFG_FASTAMP_RE_ONLY) { return TDM_ERR_ARG; }
    // The kernels emit at most 8 symbols per 8-sample tick into 16-entry hand-over rings and size the output rows from
    // the smallest possible advance per symbol, omega_min - |mu_gain| (complex_fd.cpp:140-143: mu += omega + muGain*e,
    // |e| <= 1, advance = floor(mu)): below ~1.25 samples per symbol both would have to change.  The reference itself
    // runs at 2 samples per symbol (src/main.cpp:35,84).
    {
        const double omega_min = cfg->samplerate / cfg->symbolrate * (1.0 - cfg->omega_rel_limit);
        if (!(omega_min - std::fabs(cfg->mu_gain) >= 1.25)) { return TDM_ERR_UNSUPPORTED; }
    }
    std::memset(d, 0, sizeof(*d));
    d->ntaps = nt;
    d->fastamp_re_only = (cfg->flags & TDM_CFG_FASTAMP_RE_ONLY) ? 1 : 0;
    // A filter shorter than the kernels' 65 taps is zero-padded at the OLD end: fma(0, x, acc)
    // leaves acc untouched, so the padded filter is bit-identical to the short one.
    const int pad = TDM_MAX_TAPS - nt;
    design_rrc(nt, cfg->rrc_beta, cfg->samplerate / cfg->symbolrate, d->rrc + pad);
    design_bandedge(nt, (int)cfg->symbolrate, (int)cfg->samplerate, (float)cfg->rrc_beta, d->be_a + pad, d->be_b + pad);
    design_interp_bank(d->bank);

    // agc.init(NULL, 1.0, 10e6, agcRate)  pi4dqpsk.cpp:20, [A.3]
    d->agc_set_point = 1.0;
    d->agc_max_gain = 10e6;
    d->agc_rate = cfg->agc_rate;
    d->agc_init_gain = 1.0;
    // fll.init(..., 0, -FL_M_PI/2, FL_M_PI/2)  pi4dqpsk.cpp:17; alpha forced to 0 at fll.cpp:25
    d->fll_beta = critically_damped((float)cfg->fll_bandwidth).beta;
    d->fll_init_freq = 0;
    d->fll_min_freq = (double)(-kFlPi / 2.0f);
    d->fll_max_freq = (double)(kFlPi / 2.0f);
    // costas.init(NULL, bw, 0, 0, -FL_M_PI/10, FL_M_PI/10)  pi4dqpsk.cpp:21, [A.5]
    const LoopGains cg = critically_damped((float)cfg->costas_bandwidth);
    d->costas_alpha = cg.alpha;
    d->costas_beta = cg.beta;
    d->costas_min_freq = (double)(-kFlPi / 10.0f);
    d->costas_max_freq = (double)(kFlPi / 10.0f);
    // recov.init(NULL, samplerate/symbolrate, omegaGain, muGain, relLimit)  pi4dqpsk.cpp:22
    //   -> pcl.init(muGain, omegaGain, 0, 0, 1, omega, omega(1-l), omega(1+l))  complex_fd.cpp:22
    const double omega = cfg->samplerate / cfg->symbolrate;
    d->tr_alpha = cfg->mu_gain;
    d->tr_beta = cfg->omega_gain;
    d->tr_init_omega = omega;
    d->tr_min_omega = omega * (1.0 - cfg->omega_rel_limit);
    d->tr_max_omega = omega * (1.0 + cfg->omega_rel_limit);
    return TDM_OK;
}
