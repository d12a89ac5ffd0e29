/* TEST INFRASTRUCTURE -- "Oracle B": canonical-operation-order CPU restatement of
 * the reference's src/dsp demodulation chain.  Checker only: the product path
 * (libtdm_b200.so) never links, loads or calls anything in oracle/.
 *
 * WHY A SECOND ORACLE.  Oracle A (oracle/ref_driver.cpp + the reference's own
 * sources) is the authority for DECODED BITS.  Its float trajectories, however,
 * cannot be reproduced by ANY other build: the loop states are sensitive to
 * 1-ulp changes (VOLK's summation order is machine dependent, libm's sinf/cosf
 * differ from CUDA's in the last place).  Oracle B therefore
 * FIXES an operation order that both a C compiler and a CUDA kernel can execute
 * bit-identically, and is the authority for FLOAT STATE parity (expected exact;
 * 1e-4 relative is the contractual bound).  tests/test_oracles.py shows that on
 * every golden capture Oracle B's decoded bits equal Oracle A's.
 *
 * PARITY PINNING.  The reference has no tests or golden vectors;
 * this restatement is pinned against outputs of the reference itself built here
 * (oracle/_ref, `make ref`) and against fixtures that build generated
 * (tests/golden/, tests/golden/make_golden.py).
 *
 * THE CANONICAL ORDER (everything is IEEE-754 binary32, round-to-nearest-even,
 * no contraction other than the fmaf() calls written out below):
 *   - every dot product is ONE chain per real component, taps in ascending
 *     index (oldest sample first), acc = fmaf(tap, sample, acc) from acc = +0;
 *   - the two band-edge FIRs share their products: with hbe taps a+jb and
 *     lbe taps a-jb (exact conjugates, fll.cpp:89-93)  P = sum a x, Q = sum b x,
 *     hbe = (P.re-Q.im, P.im+Q.re), lbe = (P.re+Q.im, P.im-Q.re);
 *   - sin/cos come from ob_sincos() below (Cody-Waite + degree-7/8 polynomials),
 *     never from libm; the FLL's NCO (one evaluation per input sample, on the
 *     sample-rate recurrence) uses the same polynomials on a reduction that is
 *     PREPARED one sample ahead (ob_fll_prepare / ob_fll_reduce below) so that a
 *     GPU needs one addition, not a range reduction, between the loop filter and
 *     the polynomial;
 *   - sqrtf is IEEE-correct on both sides; floorf, comparisons, min/max exact.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).  [A.n] = SURVEY.md Appendix A item n (SDR++ core semantics,
 * which are not vendored in the reference tree).
 */
#include "oracle_b.h"

#include <math.h>
#include <pthread.h>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
/* The canonical order is IEEE-754 with gradual underflow.  Any shared object built with -ffast-math that the
 * host process happens to load (crtfastmath.o) switches the thread to flush-to-zero / denormals-are-zero, and
 * new threads inherit that: force it off for the duration of a call, restore afterwards. */
#define OB_FP_ENV_ENTER() unsigned ob_saved_csr = _mm_getcsr(); _mm_setcsr(ob_saved_csr & ~0x8040u)
#define OB_FP_ENV_LEAVE() _mm_setcsr(ob_saved_csr)
#else
#define OB_FP_ENV_ENTER() (void)0
#define OB_FP_ENV_LEAVE() (void)0
#endif
#include <stdlib.h>
#include <string.h>

#define OB_FL_M_PI 3.1415926535f /* [A.1] FL_M_PI */
#define OB_DB_M_PI 3.14159265358979323846
#define OB_DB_M_SQRT2 1.41421356237309504880

/* ------------------------------------------------------------------------- */
/* Host-side design, double precision, mirrors the reference's init path.     */
/* ------------------------------------------------------------------------- */

/* src/main.cpp:35-44 constants, :78-82 gain arithmetic (mixed float/double). */
void ob_default_config(tdm_config* cfg) {
    float bw = 0.00628f, damp = 0.707f;
    float den = (1.0f + 2.0 * damp * bw + bw * bw);
    float mu = (4.0f * damp * bw) / den;
    float omega = (4.0f * bw * bw) / den;
    memset(cfg, 0, sizeof(*cfg));
    cfg->symbolrate = 18000;
    cfg->samplerate = 36000;
    cfg->rrc_tap_count = 65;
    cfg->rrc_beta = 0.35f;
    cfg->agc_rate = 0.02f;
    cfg->costas_bandwidth = 0.01f;
    cfg->fll_bandwidth = 0.006f;
    cfg->omega_gain = omega;
    cfg->mu_gain = mu;
    cfg->omega_rel_limit = 0.02f;
}

/* [A.2] PhaseControlLoop<float>::criticallyDamped */
static void ob_critically_damped(float bandwidth, float* alpha, float* beta) {
    float damp = sqrt(2.0) / 2.0;
    float den = (1.0 + 2.0 * damp * bandwidth + bandwidth * bandwidth);
    *alpha = (4 * damp * bandwidth) / den;
    *beta = (4 * bandwidth * bandwidth) / den;
}

static double ob_sinc(double x) { return (x == 0.0) ? 1.0 : (sin(x) / x); } /* [A.1] */

/* [A.6] window::nuttall */
static double ob_nuttall(double n, double N) {
    const double c[4] = { 0.355768, 0.487396, 0.144232, 0.012604 };
    double win = 0.0, sign = 1.0;
    for (int i = 0; i < 4; i++) {
        win += sign * c[i] * cos((double)i * 2.0 * OB_DB_M_PI * n / N);
        sign = -sign;
    }
    return win;
}

int ob_design(const tdm_config* cfg, tdm_design* d) {
    memset(d, 0, sizeof(*d));
    int nt = cfg->rrc_tap_count;
    if (nt < 1 || nt > TDM_MAX_TAPS) { return -4; }
    d->ntaps = nt;
    d->fastamp_re_only = (cfg->flags & TDM_CFG_FASTAMP_RE_ONLY) ? 1 : 0;
    int pad = TDM_MAX_TAPS - nt; /* shorter filters are zero-padded at the OLD end */

    /* --- RRC: taps::rootRaisedCosine<float>(n, beta, symrate, samprate)  [A.6], pi4dqpsk.cpp:18 */
    {
        double beta = cfg->rrc_beta, Ts = cfg->samplerate / cfg->symbolrate;
        double half = (double)nt / 2.0, limit = Ts / (4.0 * beta);
        for (int i = 0; i < nt; i++) {
            double t = (double)i - half + 0.5, v;
            if (t == 0.0) { v = (1.0 + beta * (4.0 / OB_DB_M_PI - 1.0)) / Ts; }
            else if (t == limit || t == -limit) {
                v = ((1.0 + 2.0 / OB_DB_M_PI) * sin(OB_DB_M_PI / (4.0 * beta)) +
                     (1.0 - 2.0 / OB_DB_M_PI) * cos(OB_DB_M_PI / (4.0 * beta))) * beta / (Ts * OB_DB_M_SQRT2);
            }
            else {
                v = ((sin((1.0 - beta) * OB_DB_M_PI * t / Ts) + cos((1.0 + beta) * OB_DB_M_PI * t / Ts) * 4.0 * beta * t / Ts) /
                     ((1.0 - (4.0 * beta * t / Ts) * (4.0 * beta * t / Ts)) * OB_DB_M_PI * t / Ts)) / Ts;
            }
            d->rrc[pad + i] = (float)v;
        }
    }

    /* --- band-edge filters: FLL::createBandedgeFilters, fll.cpp:61-95.
     * FLL::init takes the rates as int (fll.h:33, [A.9]); filt_a is float. */
    {
        double symrate = (int)cfg->symbolrate, samprate = (int)cfg->samplerate;
        float filt_a = (float)cfg->rrc_beta;
        float sps = samprate / symrate;                 /* fll.cpp:62 */
        const int M = (nt / sps);                       /* fll.cpp:64 */
        float power = 0;
        float bb[TDM_MAX_TAPS];
        for (int i = 0; i < nt; i++) {                  /* fll.cpp:69-75 */
            float k = -M + i * 2.0f / sps;
            float tap = ob_sinc(filt_a * k - 0.5f) + ob_sinc(filt_a * k + 0.5f);
            power += tap;
            bb[i] = tap;
        }
        int N = (nt - 1.0f) / 2.0f;                     /* fll.cpp:83 */
        for (int i = 0; i < nt; i++) {                  /* fll.cpp:84-94 */
            float tap = bb[i] / power;
            float k = (-N + (int)i) / (2.0f * sps);
            float th = 2.0f * OB_FL_M_PI * (1.0f + filt_a) * k;
            /* hbe tap t2 = phasor(+th)*tap; lbe tap t1 = phasor(-th)*tap = conj(t2); stored reversed */
            d->be_a[pad + nt - i - 1] = cosf(th) * tap;
            d->be_b[pad + nt - i - 1] = sinf(th) * tap;
        }
    }

    /* --- AGC: agc.init(NULL, 1.0, 10e6, agcRate), pi4dqpsk.cpp:20, [A.3] */
    d->agc_set_point = 1.0;
    d->agc_max_gain = 10e6;
    d->agc_rate = cfg->agc_rate;
    d->agc_init_gain = 1.0;

    /* --- FLL loop: fll.cpp:22-26 with pi4dqpsk.cpp:17 limits; alpha forced to 0 */
    {
        float a, b;
        ob_critically_damped(cfg->fll_bandwidth, &a, &b);
        d->fll_beta = b;
        d->fll_init_freq = 0;
        d->fll_min_freq = (double)(-OB_FL_M_PI / 2.0f);
        d->fll_max_freq = (double)(OB_FL_M_PI / 2.0f);
    }

    /* --- Costas: costas.init(NULL, bw, 0, 0, -pi/10, pi/10), pi4dqpsk.cpp:21, [A.5] */
    {
        float a, b;
        ob_critically_damped(cfg->costas_bandwidth, &a, &b);
        d->costas_alpha = a;
        d->costas_beta = b;
        d->costas_min_freq = (double)(-OB_FL_M_PI / 10.0f);
        d->costas_max_freq = (double)(OB_FL_M_PI / 10.0f);
    }

    /* --- timing: recov.init(NULL, sps, omegaGain, muGain, relLimit), pi4dqpsk.cpp:22,
     * pcl.init(muGain, omegaGain, 0, 0, 1, omega, omega(1-l), omega(1+l)), complex_fd.cpp:22 */
    {
        double omega = cfg->samplerate / cfg->symbolrate;
        d->tr_alpha = cfg->mu_gain;
        d->tr_beta = cfg->omega_gain;
        d->tr_init_omega = omega;
        d->tr_min_omega = omega * (1.0 - cfg->omega_rel_limit);
        d->tr_max_omega = omega * (1.0 + cfg->omega_rel_limit);
    }

    /* --- interpolator bank: complex_fd.cpp:153-158, windowedSinc + nuttall + buildPolyphaseBank [A.6] */
    {
        const int P = TDM_INTERP_PHASES, T = TDM_INTERP_TAPS, count = P * T;
        double bw = 0.5 / (double)P;
        double omega = 2.0 * OB_DB_M_PI * (bw / 1.0);      /* hzToRads(bw, 1.0) */
        double half = (double)count / 2.0, corr = (double)P * omega / OB_DB_M_PI;
        for (int i = 0; i < count; i++) {
            double t = (double)i - half + 0.5;
            float v = (float)(ob_sinc(t * omega) * ob_nuttall(t - half, count) * corr);
            d->bank[(P - 1) - (i % P)][i / P] = v;
        }
    }
    return 0;
}

void ob_state_init(const tdm_design* d, tdm_channel_state* s) {
    memset(s, 0, sizeof(*s));
    s->agc_gain = d->agc_init_gain;
    s->fll_freq = d->fll_init_freq;
    s->tr_omega = d->tr_init_omega;
}

/* ------------------------------------------------------------------------- */
/* Canonical arithmetic                                                        */
/* ------------------------------------------------------------------------- */

static inline uint32_t ob_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* sin/cos for |x| <= ~2 pi (all callers keep their phases wrapped).
 *   q = nearest integer to x*2/pi  (magic-number rounding, exact in IEEE)
 *   r = x - q*pi/2                 (two fused steps, pi/2 split hi+lo)
 *   sin r = r + r^3 (S1 + r^2 (S2 + r^2 S3)),  cos r = 1 + r^2(-1/2 + r^2 (C1 + r^2 (C2 + r^2 C3)))
 * Replaces math::phasor's cosf/sinf ([A.1]; fll.cpp:137, pi4dqpsk_costas.cpp:7,16). */
void ob_sincos(float x, float* s, float* c) {
    const float two_over_pi = 0.636619747f;
    const float magic = 12582912.0f;          /* 1.5 * 2^23 */
    const float pio2_hi = 1.57079637f;        /* float(pi/2) */
    const float pio2_lo = -4.37113883e-8f;    /* pi/2 - pio2_hi */
    const float S1 = -1.6666654611e-1f, S2 = 8.3321608736e-3f, S3 = -1.9515295891e-4f;
    const float C1 = 4.166664568298827e-2f, C2 = -1.388731625493765e-3f, C3 = 2.443315711809948e-5f;
    float t = fmaf(x, two_over_pi, magic);
    uint32_t n = ob_bits(t) & 3u;
    float q = t - magic;
    float r = fmaf(q, -pio2_hi, x);
    r = fmaf(q, -pio2_lo, r);
    float r2 = r * r;
    float sp = fmaf(r2, S3, S2);
    sp = fmaf(sp, r2, S1);
    float sn = fmaf(sp, r2 * r, r);
    float cp = fmaf(r2, C3, C2);
    cp = fmaf(cp, r2, C1);
    cp = fmaf(cp, r2, -0.5f);
    float cs = fmaf(cp, r2, 1.0f);
    float ss = (n & 1u) ? cs : sn;
    float cc = (n & 1u) ? sn : cs;
    if (n & 2u) { ss = -ss; }
    if ((n + 1u) & 2u) { cc = -cc; }
    *s = ss;
    *c = cc;
}

/* [A.1] complex_t::fastAmplitude: a=|re|, b=|im|; a>b ? a+0.4b : b+0.4a.
 * re_only selects the OTHER reading of upstream SDR++ (both operands from |re|, i.e. b = |re| too:
 * TDM_CFG_FASTAMP_RE_ONLY, include/tdm_b200.h); SDR++ core is not vendored by the reference, so both
 * readings are built and pinned (oracle/_ref/libtetra_ref_reonly.so is the reference compiled with the
 * stand-in's -DSDRPP_STANDIN_FASTAMP_RE_ONLY). */
static inline float ob_fastamp(float re, float im, int re_only) {
    float a = fabsf(re), b = re_only ? a : fabsf(im);
    float hi = a > b ? a : b, lo = a > b ? b : a;
    return fmaf(0.4f, lo, hi);
}

/* ---- the FLL's NCO, math::phasor(-pcl.phase) at fll.cpp:137 -------------------------------------------
 * The reference evaluates cosf/sinf of the loop phase once per input sample, and the phase of sample n+1
 * depends on the error of sample n: range reduction + polynomial sit on the sample-rate recurrence.  The
 * canonical order cuts that dependency short without changing what is computed:
 *   ob_fll_prepare (phi_n, f_n) -> quadrant q and r0 = phi_n - q pi/2, with q the nearest quadrant of the
 *       PREDICTED next phase phi_n + f_n (what the next phase would be if the frequency did not move);
 *   after the loop filter:  r = r0 + f_{n+1}   (= the next phase, unwrapped, minus q pi/2);
 *   ob_fll_reduce: if |r| <= OB_FLL_RMAX the pair (q, r) is used as it is (the frequency moves by
 *       beta*err ~ 1e-4 per sample, so r overshoots pi/4 by about that much at most: the polynomials below
 *       are as accurate there); otherwise -- the frequency jumped, e.g. a burst after silence with the AGC
 *       gain wide open -- (q, r) come from the classic reduction of the wrapped phase.
 * Deviation from sin/cos of the wrapped phase: the polynomial error (< 3e-7 absolute for |r| <= 0.8) plus,
 * in the one sample where the phase wraps, 2*pi_f - 2*pi = 1.7e-7 rad (the wrap subtracts the FLOAT 2 pi like
 * the reference, the quadrant arithmetic uses pi/2 to 2^-48).  (q, r) travel in tdm_channel_state
 * (fll_quad, fll_r), so chunking cannot change a bit. */
#define OB_FLL_RMAX 0.8f
static const float ob_two_over_pi = 0.636619747f, ob_magic = 12582912.0f /* 1.5 * 2^23 */;
static const float ob_pio2_hi = 1.57079637f, ob_pio2_lo = -4.37113883e-8f;

static inline void ob_fll_prepare(float phi, float f, uint32_t* q, float* r0) {
    float t = fmaf(phi + f, ob_two_over_pi, ob_magic);
    uint32_t u; memcpy(&u, &t, 4);
    float qf = t - ob_magic;
    float r = fmaf(qf, -ob_pio2_hi, phi);
    *r0 = fmaf(qf, -ob_pio2_lo, r);
    *q = u & 3u;
}

static long ob_fallbacks;   /* how often the classic reduction was taken (tests want to know they exercised it) */
long ob_fll_fallback_count(void) { return __atomic_load_n(&ob_fallbacks, __ATOMIC_RELAXED); }

static inline void ob_fll_reduce(float phi_wrapped, uint32_t* q, float* r) {
    if (!(fabsf(*r) <= OB_FLL_RMAX)) {
        __atomic_fetch_add(&ob_fallbacks, 1, __ATOMIC_RELAXED);
        float t = fmaf(phi_wrapped, ob_two_over_pi, ob_magic);
        uint32_t u; memcpy(&u, &t, 4);
        float qf = t - ob_magic;
        float rr = fmaf(qf, -ob_pio2_hi, phi_wrapped);
        *r = fmaf(qf, -ob_pio2_lo, rr);
        *q = u & 3u;
    }
}

/* sin r, cos r for the reduced argument: sin as in ob_sincos; cos with the same coefficients in Estrin
 * form (one dependent level less): 1 + r2 ((-1/2 + C1 r2) + r4 (C2 + C3 r2)). */
static inline void ob_fll_poly(float r, float* sn, float* cs) {
    const float S1 = -1.6666654611e-1f, S2 = 8.3321608736e-3f, S3 = -1.9515295891e-4f;
    const float C1 = 4.166664568298827e-2f, C2 = -1.388731625493765e-3f, C3 = 2.443315711809948e-5f;
    float r2 = r * r;
    float sp = fmaf(r2, S3, S2);
    sp = fmaf(sp, r2, S1);
    *sn = fmaf(sp, r2 * r, r);
    float r4 = r2 * r2;
    float cl = fmaf(C1, r2, -0.5f), ch = fmaf(C3, r2, C2);
    float cq = fmaf(ch, r4, cl);
    *cs = fmaf(cq, r2, 1.0f);
}

/* The NCO exactly as one step of the chain evaluates it, for the accuracy test: state (phi, f_old) before the
 * loop update, f_new after it; returns sin/cos of the new phase wrap(phi + f_new) as the next sample will use
 * them, with the quadrant folded back in. */
void ob_fll_nco(float phi, float f_old, float f_new, float* s, float* c) {
    const float pi = OB_FL_M_PI, two_pi = pi - (-pi);
    uint32_t q; float r0;
    ob_fll_prepare(phi, f_old, &q, &r0);
    float r = r0 + f_new;
    float ph = phi + f_new;
    while (ph > pi) { ph -= two_pi; }
    while (ph < -pi) { ph += two_pi; }
    ob_fll_reduce(ph, &q, &r);
    float sn, cs;
    ob_fll_poly(r, &sn, &cs);
    float ss = (q & 1u) ? cs : sn, cc = (q & 1u) ? sn : cs;
    if (q & 2u) { ss = -ss; }
    if ((q + 1u) & 2u) { cc = -cc; }
    *s = ss; *c = cc;
}

/* (yr, yi) * (-j)^q: the quadrant part of exp(-j phi) applied to the sample (exact) */
static inline void ob_quarter_turns(uint32_t q, float* yr, float* yi) {
    float a = *yr, b = *yi;
    switch (q & 3u) {
        case 1: *yr = b; *yi = -a; break;
        case 2: *yr = -a; *yi = -b; break;
        case 3: *yr = -b; *yi = a; break;
        default: break;
    }
}

static inline float ob_clampf(float v, float lo, float hi) { return v > hi ? hi : (v < lo ? lo : v); }

/* ------------------------------------------------------------------------- */
/* The chain, fused per sample.  The reference runs the stages block by block   */
/* over each buffer (pi4dqpsk.cpp:132-140); every stage is causal and nothing   */
/* feeds back across stages, so running them sample by sample is the same       */
/* computation, and chunking cannot change the result.                          */
/* ------------------------------------------------------------------------- */

#define OB_CHUNK 16384

typedef struct ob_work {
    float x[2 * (TDM_HIST + OB_CHUNK)];                 /* FLL-output delay line [A.4]      */
    float r[2 * (TDM_INTERP_TAPS - 1 + OB_CHUNK)];      /* RRC-output buffer, complex_fd.cpp:91 */
} ob_work;

static int64_t ob_process_chunk(const tdm_design* d, tdm_channel_state* s, ob_work* w, const float* iq, int count,
                                float* syms, uint8_t* dibits, uint8_t* bits) {
    const float pi = OB_FL_M_PI, two_pi = pi - (-pi);   /* phaseDelta = maxPhase - minPhase [A.2] */
    const float costas_two_pi = 2 * OB_FL_M_PI;         /* pi4dqpsk_costas.cpp:11-15 */
    const float quarter_pi = OB_FL_M_PI / 4.0f;         /* pi4dqpsk_costas.cpp:10 */
    float g = s->agc_gain;
    float fph = s->fll_phase, ffr = s->fll_freq, fr = s->fll_r;
    uint32_t fq = s->fll_quad;
    const int re_only = d->fastamp_re_only != 0;
    float mu = s->tr_mu, om = s->tr_omega;
    int offset = s->tr_offset;
    float cph = s->costas_phase, cfr = s->costas_freq, ph2 = s->costas_ph2;
    uint32_t prev = s->prev_sym;
    int64_t nsym = 0;

    memcpy(w->x, s->x_hist, sizeof(s->x_hist));
    memcpy(w->r, s->r_hist, sizeof(s->r_hist));

    for (int n = 0; n < count; n++) {
        /* ---- FastAGC [A.3]; pi4dqpsk.cpp:134 */
        float yr = iq[2 * n] * g, yi = iq[2 * n + 1] * g;
        float amp = sqrtf(fmaf(yr, yr, yi * yi));
        g = fmaf(d->agc_set_point - amp, d->agc_rate, g);
        if (g > d->agc_max_gain) { g = d->agc_max_gain; }

        /* ---- FLL: fll.cpp:135-149.  shift = phasor(-phase) = (cos, -sin) = (-j)^q (cos r, -sin r) */
        float sn, cs;
        ob_fll_poly(fr, &sn, &cs);
        float tr = yr, ti = yi;
        ob_quarter_turns(fq, &tr, &ti);
        float xr = fmaf(tr, cs, ti * sn);
        float xi = fmaf(ti, cs, -(tr * sn));
        /* next sample's reduction, from the state BEFORE the loop update */
        uint32_t nq; float r0;
        ob_fll_prepare(fph, ffr, &nq, &r0);
        float* win = &w->x[2 * n];          /* win[0..63] history, win[64] = x */
        win[2 * TDM_HIST] = xr;
        win[2 * TDM_HIST + 1] = xi;
        float pr = 0.0f, pim = 0.0f, qr = 0.0f, qi = 0.0f, rr = 0.0f, ri = 0.0f;
        for (int k = 0; k < TDM_MAX_TAPS; k++) {
            float vr = win[2 * k], vi = win[2 * k + 1];
            pr = fmaf(d->be_a[k], vr, pr);
            pim = fmaf(d->be_a[k], vi, pim);
            qr = fmaf(d->be_b[k], vr, qr);
            qi = fmaf(d->be_b[k], vi, qi);
            rr = fmaf(d->rrc[k], vr, rr);   /* RRC matched filter, pi4dqpsk.cpp:136, [A.4] */
            ri = fmaf(d->rrc[k], vi, ri);
        }
        float hbe = ob_fastamp(pr - qi, pim + qr, re_only);
        float lbe = ob_fastamp(pr + qi, pim - qr, re_only);
        float ferr = hbe - lbe;             /* fll.cpp:143 */
        /* pcl.advance, alpha = 0 (fll.cpp:25,145) [A.2] */
        ffr = ob_clampf(fmaf(d->fll_beta, ferr, ffr), d->fll_min_freq, d->fll_max_freq);
        fr = r0 + ffr;
        fph = fph + ffr;
        while (fph > pi) { fph -= two_pi; }
        while (fph < -pi) { fph += two_pi; }
        fq = nq;
        ob_fll_reduce(fph, &fq, &fr);

        float* rb = &w->r[2 * n];           /* rb[0..6] previous RRC outputs, rb[7] = this one */
        rb[2 * (TDM_INTERP_TAPS - 1)] = rr;
        rb[2 * (TDM_INTERP_TAPS - 1) + 1] = ri;

        /* ---- timing recovery: complex_fd.cpp:89-151.  The reference's loop
         * `while (offset < count)` reads buffer[offset .. offset+7], i.e. RRC
         * outputs offset-7 .. offset: the symbol is computable exactly when
         * sample n == offset has been filtered. */
        while (n == offset) {
        int ph = (int)floorf(mu * (float)TDM_INTERP_PHASES);
        ph = ph < 0 ? 0 : (ph > TDM_INTERP_PHASES - 1 ? TDM_INTERP_PHASES - 1 : ph);
        int plo = ph == 0 ? 0 : ph - 1, phi = ph == TDM_INTERP_PHASES - 1 ? ph : ph + 1;
        float yre = 0, yim = 0, are = 0, aim = 0, bre = 0, bim = 0; /* y, f(T+1), f(T-1) */
        for (int k = 0; k < TDM_INTERP_TAPS; k++) {
            float vr = rb[2 * k], vi = rb[2 * k + 1];
            yre = fmaf(d->bank[ph][k], vr, yre);
            yim = fmaf(d->bank[ph][k], vi, yim);
            are = fmaf(d->bank[phi][k], vr, are);
            aim = fmaf(d->bank[phi][k], vi, aim);
            bre = fmaf(d->bank[plo][k], vr, bre);
            bim = fmaf(d->bank[plo][k], vi, bim);
        }
        float dre, dim;
        if (ph == 0) { dre = are - yre; dim = aim - yim; }                                   /* :107-111 */
        else if (ph == TDM_INTERP_PHASES - 1) { dre = yre - bre; dim = yim - bim; }           /* :112-116 */
        else { dre = (are - bre) * 0.5f; dim = (aim - bim) * 0.5f; }                          /* :117-123 */
        float terr = (yre > 0 ? dre : -dre) + (yim > 0 ? dim : -dim);                         /* :126 */
        terr = ob_clampf(terr, -1.0f, 1.0f);                                                   /* :136-137 */
        /* pcl.advance (PhaseControlLoop<float,false>) :140, then :141-143 */
        om = ob_clampf(fmaf(d->tr_beta, terr, om), d->tr_min_omega, d->tr_max_omega);
        mu = mu + fmaf(d->tr_alpha, terr, om);
        float delta = floorf(mu);
        /* non-finite guard shared with the kernel (unreachable for finite input) */
        if (!(delta >= 0.0f)) { delta = 1.0f; }
        if (delta > 1048576.0f) { delta = 1048576.0f; }
        offset += (int)delta;
        mu -= delta;

        /* ---- pi/4 Costas: pi4dqpsk_costas.cpp:5-28 */
        ob_sincos(cph, &sn, &cs);
        float zr = fmaf(yre, cs, yim * sn);
        float zi = fmaf(yim, cs, -(yre * sn));
        ph2 += -quarter_pi;
        if (ph2 >= costas_two_pi) { ph2 -= costas_two_pi; }
        else if (ph2 <= -costas_two_pi) { ph2 += costas_two_pi; }
        float s2, c2;
        ob_sincos(ph2, &s2, &c2);
        float ur = fmaf(zr, c2, -(zi * s2));
        float ui = fmaf(zi, c2, zr * s2);
        float cerr = (ur > 0 ? ui : -ui) - (ui > 0 ? ur : -ur);                               /* :26 */
        cerr = ob_clampf(cerr, -1.0f, 1.0f);
        cfr = ob_clampf(fmaf(d->costas_beta, cerr, cfr), d->costas_min_freq, d->costas_max_freq);
        cph = cph + fmaf(d->costas_alpha, cerr, cfr);
        while (cph > pi) { cph -= two_pi; }
        while (cph < -pi) { cph += two_pi; }
        if (syms) { syms[2 * nsym] = ur; syms[2 * nsym + 1] = ui; }

        /* ---- slicer + differential decoder: dqpsk_sym_extr.cpp:4-55 */
        int a = ui < 0, b = ur < 0;
        float ideal = a ? (b ? -2.35619449f : -0.785398185f) : (b ? 2.35619449f : 0.785398185f);
        float dist = fabsf(ideal - atan2f(ui, ur));                                            /* :11 */
        s->err_partial += dist;
        s->err_ptr++;
        s->err_disp++;
        if (s->err_disp >= TDM_SYNC_DISPLAY) {                                                 /* :17-30 */
            s->err_blocks[(s->err_ptr - 1) / TDM_SYNC_DISPLAY] = s->err_partial;
            s->err_partial = 0;
            float tot = 0;
            for (int j = 0; j < TDM_SYNC_BLOCKS; j++) { tot += s->err_blocks[j]; }
            s->standarderr = tot / (float)TDM_SYNC_BUF;
            s->sync = s->standarderr < 0.35f;
            s->err_disp = 0;
        }
        if (s->err_ptr >= TDM_SYNC_BUF) { s->err_ptr = 0; }
        uint32_t sym = ((uint32_t)a << 1) | (uint32_t)(a != b);                                /* :32 */
        uint32_t pd = (sym - prev + 4) % 4;                                                    /* :33 */
        static const uint8_t remap[4] = { 0, 1, 3, 2 };                                        /* :34-51 */
        uint8_t db = remap[pd];
        prev = sym;
        if (dibits) { dibits[nsym] = db; }
        if (bits) { bits[2 * nsym] = (db >> 1) & 1; bits[2 * nsym + 1] = db & 1; }             /* bit_unpacker.cpp:6-7 */
        nsym++;
        } /* while (n == offset) */
    }

    offset -= count;                                                                           /* complex_fd.cpp:145 */
    memcpy(s->x_hist, &w->x[2 * count], sizeof(s->x_hist));                                    /* [A.4] memmove */
    memcpy(s->r_hist, &w->r[2 * count], sizeof(s->r_hist));                                    /* complex_fd.cpp:148 */
    s->agc_gain = g;
    s->fll_phase = fph; s->fll_freq = ffr; s->fll_quad = fq; s->fll_r = fr;
    s->tr_mu = mu; s->tr_omega = om; s->tr_offset = offset;
    s->costas_phase = cph; s->costas_freq = cfr; s->costas_ph2 = ph2;
    s->prev_sym = prev;
    s->n_samples += (uint64_t)count;
    s->n_symbols += (uint64_t)nsym;
    return nsym;
}

int64_t ob_process(const tdm_design* d, tdm_channel_state* s, const float* iq, int64_t count,
                   float* syms, uint8_t* dibits, uint8_t* bits) {
    OB_FP_ENV_ENTER();
    ob_work* w = (ob_work*)malloc(sizeof(ob_work));
    int64_t done = 0, nsym = 0;
    while (done < count) {
        int n = (int)((count - done) < OB_CHUNK ? (count - done) : OB_CHUNK);
        int64_t k = ob_process_chunk(d, s, w, iq + 2 * done, n, syms ? syms + 2 * nsym : NULL,
                                     dibits ? dibits + nsym : NULL, bits ? bits + 2 * nsym : NULL);
        nsym += k;
        done += n;
    }
    free(w);
    OB_FP_ENV_LEAVE();
    return nsym;
}

typedef struct ob_job {
    const tdm_design* d; tdm_channel_state* states; int nch; const float* iq; int64_t in_stride, count;
    float* syms; uint8_t* dibits; uint8_t* bits; int64_t out_stride; int32_t* out_counts; int t, nthreads;
} ob_job;

static void* ob_worker(void* arg) {
    ob_job* j = (ob_job*)arg;
    for (int c = j->t; c < j->nch; c += j->nthreads) {
        int64_t k = ob_process(j->d, &j->states[c], j->iq + 2 * (size_t)c * (size_t)j->in_stride, j->count,
                               j->syms ? j->syms + 2 * (size_t)c * (size_t)j->out_stride : NULL,
                               j->dibits ? j->dibits + (size_t)c * (size_t)j->out_stride : NULL,
                               j->bits ? j->bits + 2 * (size_t)c * (size_t)j->out_stride : NULL);
        if (j->out_counts) { j->out_counts[c] = (int32_t)k; }
    }
    return NULL;
}

void ob_process_multi(const tdm_design* d, tdm_channel_state* states, int nch, const float* iq, int64_t in_stride,
                      int64_t count, float* syms, uint8_t* dibits, uint8_t* bits, int64_t out_stride,
                      int32_t* out_counts, int nthreads) {
    if (nthreads < 1) { nthreads = 1; }
    if (nthreads > nch) { nthreads = nch; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    ob_job* jobs = (ob_job*)malloc(sizeof(ob_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        ob_job j = { d, states, nch, iq, in_stride, count, syms, dibits, bits, out_stride, out_counts, t, nthreads };
        jobs[t] = j;
        pthread_create(&th[t], NULL, ob_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); }
    free(jobs);
    free(th);
}
