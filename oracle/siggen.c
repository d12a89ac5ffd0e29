/* TEST INFRASTRUCTURE -- deterministic synthetic TETRA-mapped pi/4-DQPSK capture
 * generator, CPU side.  The reference ships no IQ captures and no
 * tests, so golden inputs have to be manufactured; only the symbol
 * mapping comes from the reference: bits2phase,
 * /root/reference/src/decoder/src/phy/tetra_burst.c:99-104
 *   (b1,b2) = 00 -> +pi/4, 01 -> +3pi/4, 11 -> -3pi/4, 10 -> -pi/4.
 *
 * The integer part of the recipe (who transmits which dibit, per-channel
 * impairment draws) is shared with the device generator in
 * sdrpp_tetra_demodulator_b200/csrc/tdm_synth.cu so TX dibits can be recomputed
 * anywhere; the floating-point waveform is NOT expected to match bit-for-bit
 * between the two (different libm), and no test relies on that.
 *
 * Recipe for channel c, sample n (fs = 36 kS/s, 2 samples/symbol):
 *   a_k     = hash(seed_data+c, 16+k) & 3          absolute QPSK index of symbol k
 *   idx_k   = 2 a_k + (k & 1)                      phase of symbol k in units of pi/4
 *   dibit_k = map[(idx_k - idx_{k-1}) & 7]         1->00, 3->01, 5->11, 7->10   (idx_{-1} = 7)
 *   s[n]    = A e^{j(2 pi df n / fs + phi0)} sum_k e^{j pi idx_k / 4} h(n - 2k - 2 tau)  +  w[n]
 *   h       = root raised cosine, beta 0.35, Ts = 2 samples, support |t| <= 33
 *   df, tau, A, phi0 = draws 0..3 of hash(seed_data+c, .)
 *   w[n]    = complex AWGN, Box-Muller on hash(seed_noise+c, n), Es/N0 = snr_db
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SG_PI 3.14159265358979323846
#define SG_SPAN 33 /* pulse support in samples on each side */

static inline uint64_t sg_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* stateless: value #idx of stream `seed` */
uint64_t sg_hash(uint64_t seed, uint64_t idx) {
    return sg_mix(sg_mix(seed * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL) ^ (idx * 0x9E3779B97F4A7C15ULL));
}
static inline double sg_u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

typedef struct sg_params {
    double snr_db, max_freq_off_hz, min_amp, max_amp;
    uint64_t seed_data, seed_noise;
} sg_params;

typedef struct sg_channel {
    double freq_off_hz, tau, amp, phi0;
} sg_channel;

void sg_channel_draw(const sg_params* p, int c, sg_channel* out) {
    uint64_t s = p->seed_data + (uint64_t)c;
    out->freq_off_hz = (2.0 * sg_u01(sg_hash(s, 0)) - 1.0) * p->max_freq_off_hz;
    out->tau = sg_u01(sg_hash(s, 1));
    out->amp = p->min_amp * pow(p->max_amp / p->min_amp, sg_u01(sg_hash(s, 2)));
    out->phi0 = 2.0 * SG_PI * sg_u01(sg_hash(s, 3));
}

static inline int sg_abs_index(uint64_t seed, int64_t k) {
    if (k < 0) { return 7; } /* virtual symbol -1: keeps the first increment odd */
    return (int)(2 * (sg_hash(seed, 16 + (uint64_t)k) & 3) + (uint64_t)(k & 1));
}

/* transmitted dibit of symbol k (values 0..3, b1 in bit 1) */
int sg_tx_dibit(uint64_t seed_data, int c, int64_t k) {
    static const int map[8] = { -1, 0, -1, 1, -1, 3, -1, 2 };
    uint64_t s = seed_data + (uint64_t)c;
    int d = (sg_abs_index(s, k) - sg_abs_index(s, k - 1)) & 7;
    return map[d];
}

void sg_tx_dibits(uint64_t seed_data, int c, int64_t k0, int64_t n, uint8_t* out) {
    for (int64_t i = 0; i < n; i++) { out[i] = (uint8_t)sg_tx_dibit(seed_data, c, k0 + i); }
}

/* RRC pulse, continuous time, Ts = 2 samples, peak ~ 1.1 (A.6 formula times Ts) */
static double sg_rrc(double t, double beta) {
    const double Ts = 2.0;
    double x = t / Ts;
    if (fabs(t) < 1e-12) { return 1.0 + beta * (4.0 / SG_PI - 1.0); }
    if (fabs(fabs(4.0 * beta * x) - 1.0) < 1e-9) {
        return (beta / sqrt(2.0)) * ((1.0 + 2.0 / SG_PI) * sin(SG_PI / (4.0 * beta)) + (1.0 - 2.0 / SG_PI) * cos(SG_PI / (4.0 * beta)));
    }
    double num = sin(SG_PI * x * (1.0 - beta)) + 4.0 * beta * x * cos(SG_PI * x * (1.0 + beta));
    double den = SG_PI * x * (1.0 - (4.0 * beta * x) * (4.0 * beta * x));
    return num / den;
}

/* Generate samples [n0, n0+n) of channel c into iq (interleaved float32). */
void sg_generate(const sg_params* p, int c, int64_t n0, int64_t n, float* iq) {
    sg_channel ch;
    sg_channel_draw(p, c, &ch);
    const double beta = 0.35;
    uint64_t sd = p->seed_data + (uint64_t)c, sn = p->seed_noise + (uint64_t)c;

    /* pulse table: h(j - 2 tau), j = -SPAN .. SPAN+1 */
    double tab[2 * SG_SPAN + 2];
    double e = 0.0;
    for (int j = -SG_SPAN; j <= SG_SPAN + 1; j++) {
        double t = (double)j - 2.0 * ch.tau;
        double v = (fabs(t) <= (double)SG_SPAN) ? sg_rrc(t, beta) : 0.0;
        tab[j + SG_SPAN] = v;
        e += v * v;
    }
    double ps = ch.amp * ch.amp * e / 2.0;                      /* mean |s|^2 per sample        */
    double sigma2 = ps * 2.0 / pow(10.0, p->snr_db / 10.0);     /* Es/N0 with Es = ps*fs/fsym   */
    double sigma = sqrt(sigma2 / 2.0);                          /* per real component           */
    /* carrier: 64-bit fixed-point turns per sample, wraps exactly */
    double turns = ch.freq_off_hz / 36000.0;
    int64_t inc = (int64_t)llround(turns * 18446744073709551616.0);
    uint64_t ph0 = (uint64_t)llround(ch.phi0 / (2.0 * SG_PI) * 4294967296.0) << 32;

    static const double cs8[8][2] = { { 1, 0 }, { 0.70710678118654752440, 0.70710678118654752440 }, { 0, 1 },
                                      { -0.70710678118654752440, 0.70710678118654752440 }, { -1, 0 },
                                      { -0.70710678118654752440, -0.70710678118654752440 }, { 0, -1 },
                                      { 0.70710678118654752440, -0.70710678118654752440 } };
    for (int64_t i = 0; i < n; i++) {
        int64_t nn = n0 + i;
        /* symbols k with |nn - 2k - 2 tau| <= SPAN  ->  j = nn - 2k in [-SPAN, SPAN+1] */
        int64_t kmin = (nn - (SG_SPAN + 1) + 1) / 2;
        if (nn - (SG_SPAN + 1) < 0) { kmin = 0; }
        int64_t kmax = (nn + SG_SPAN) / 2;
        double re = 0.0, im = 0.0;
        for (int64_t k = kmin; k <= kmax; k++) {
            int64_t j = nn - 2 * k;
            if (j < -SG_SPAN || j > SG_SPAN + 1) { continue; }
            int idx = sg_abs_index(sd, k);
            double h = tab[j + SG_SPAN];
            re += cs8[idx][0] * h;
            im += cs8[idx][1] * h;
        }
        uint64_t ph = ph0 + (uint64_t)inc * (uint64_t)nn;
        double ang = (double)(uint32_t)(ph >> 32) * (2.0 * SG_PI / 4294967296.0);
        double cr = cos(ang) * ch.amp, ci = sin(ang) * ch.amp;
        double sre = re * cr - im * ci, sim = re * ci + im * cr;
        uint64_t hn = sg_hash(sn, (uint64_t)nn);
        double u1 = ((double)(uint32_t)(hn >> 32) + 0.5) * (1.0 / 4294967296.0);
        double u2 = ((double)(uint32_t)hn + 0.5) * (1.0 / 4294967296.0);
        double r = sigma * sqrt(-2.0 * log(u1));
        iq[2 * i] = (float)(sre + r * cos(2.0 * SG_PI * u2));
        iq[2 * i + 1] = (float)(sim + r * sin(2.0 * SG_PI * u2));
    }
}
