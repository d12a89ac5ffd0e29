// TEST INFRASTRUCTURE (oracle side) -- stand-in for VOLK's dot products
// (SURVEY.md Appendix A.8).  Real VOLK picks a SIMD kernel at run time, so the
// reference's summation order is machine dependent; Oracle A pins it to the
// plain ascending-index order of VOLK's *_generic kernels, compiled with
// -ffp-contract=off (one rounding per multiply and per add).
//
// -DSDRPP_STANDIN_VOLK_SIMD (the TIMING build, _ref/libtetra_ref_simd.so, never the parity one): 256-bit
// FMA kernels with four partial sums per lane, the shape of what VOLK dispatches to on an AVX2 host -- so that the
// CPU baseline bench.py reports is not handicapped by scalar dot products.  Its floats differ from the generic
// order in the last places (as real VOLK's do from machine to machine); its decoded dibits are checked against the
// generic build's from the lock point on (tests/test_oracles.py).
#pragma once
#include <stdint.h>

typedef struct { float re, im; } lv_32fc_t;

#ifdef SDRPP_STANDIN_VOLK_SIMD
#include <immintrin.h>

static inline float standin_hsum256(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_add_ps(lo, _mm_movehl_ps(lo, lo));
    lo = _mm_add_ss(lo, _mm_shuffle_ps(lo, lo, 1));
    return _mm_cvtss_f32(lo);
}

// complex x real: lanes hold (re0, im0, re1, im1, ...); taps are spread to (t0, t0, t1, t1, ...)
static inline void volk_32fc_32f_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* input, const float* taps,
                                               unsigned int num_points) {
    const float* in = (const float*)input;
    __m256 acc0 = _mm256_setzero_ps(), acc1 = _mm256_setzero_ps();
    const __m256i spread_lo = _mm256_setr_epi32(0, 0, 1, 1, 2, 2, 3, 3), spread_hi = _mm256_setr_epi32(4, 4, 5, 5, 6, 6, 7, 7);
    unsigned int i = 0;
    for (; i + 8 <= num_points; i += 8) {
        const __m256 t = _mm256_loadu_ps(taps + i);
        acc0 = _mm256_fmadd_ps(_mm256_loadu_ps(in + 2 * i), _mm256_permutevar8x32_ps(t, spread_lo), acc0);
        acc1 = _mm256_fmadd_ps(_mm256_loadu_ps(in + 2 * i + 8), _mm256_permutevar8x32_ps(t, spread_hi), acc1);
    }
    const __m256 acc = _mm256_add_ps(acc0, acc1);
    // lanes alternate re, im
    const __m128 s = _mm_add_ps(_mm256_castps256_ps128(acc), _mm256_extractf128_ps(acc, 1));     // (re, im, re, im)
    float re = _mm_cvtss_f32(s) + _mm_cvtss_f32(_mm_shuffle_ps(s, s, 2));
    float im = _mm_cvtss_f32(_mm_shuffle_ps(s, s, 1)) + _mm_cvtss_f32(_mm_shuffle_ps(s, s, 3));
    for (; i < num_points; i++) {
        re += input[i].re * taps[i];
        im += input[i].im * taps[i];
    }
    result->re = re;
    result->im = im;
}

// complex x complex, no conjugate: re += a.re b.re - a.im b.im, im += a.re b.im + a.im b.re
static inline void volk_32fc_x2_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* input, const lv_32fc_t* taps,
                                              unsigned int num_points) {
    const float* a = (const float*)input;
    const float* b = (const float*)taps;
    __m256 acc_rr = _mm256_setzero_ps(), acc_ri = _mm256_setzero_ps();     // a * (b.re spread), a_swapped * (b.im spread)
    unsigned int i = 0;
    for (; i + 4 <= num_points; i += 4) {
        const __m256 av = _mm256_loadu_ps(a + 2 * i), bv = _mm256_loadu_ps(b + 2 * i);
        const __m256 bre = _mm256_moveldup_ps(bv), bim = _mm256_movehdup_ps(bv);
        const __m256 asw = _mm256_permute_ps(av, 0xB1);                     // (im, re, ...)
        acc_rr = _mm256_fmadd_ps(av, bre, acc_rr);                          // (a.re b.re, a.im b.re)
        acc_ri = _mm256_fmadd_ps(asw, bim, acc_ri);                         // (a.im b.im, a.re b.im)
    }
    const __m256 sum = _mm256_addsub_ps(acc_rr, acc_ri);                    // (rr - ii, ir + ri) per pair
    const __m128 s = _mm_add_ps(_mm256_castps256_ps128(sum), _mm256_extractf128_ps(sum, 1));
    float re = _mm_cvtss_f32(s) + _mm_cvtss_f32(_mm_shuffle_ps(s, s, 2));
    float im = _mm_cvtss_f32(_mm_shuffle_ps(s, s, 1)) + _mm_cvtss_f32(_mm_shuffle_ps(s, s, 3));
    for (; i < num_points; i++) {
        re += (input[i].re * taps[i].re) - (input[i].im * taps[i].im);
        im += (input[i].re * taps[i].im) + (input[i].im * taps[i].re);
    }
    result->re = re;
    result->im = im;
}

static inline void volk_32f_x2_dot_prod_32f(float* result, const float* input, const float* taps, unsigned int num_points) {
    __m256 acc = _mm256_setzero_ps();
    unsigned int i = 0;
    for (; i + 8 <= num_points; i += 8) { acc = _mm256_fmadd_ps(_mm256_loadu_ps(input + i), _mm256_loadu_ps(taps + i), acc); }
    float r = standin_hsum256(acc);
    for (; i < num_points; i++) { r += input[i] * taps[i]; }
    *result = r;
}

#else

// result = sum_i input[i] * taps[i]      complex x real
static inline void volk_32fc_32f_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* input, const float* taps,
                                               unsigned int num_points) {
    float re = 0.0f, im = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) {
        re += input[i].re * taps[i];
        im += input[i].im * taps[i];
    }
    result->re = re;
    result->im = im;
}

// result = sum_i input[i] * taps[i]      complex x complex, no conjugate
static inline void volk_32fc_x2_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* input, const lv_32fc_t* taps,
                                              unsigned int num_points) {
    float re = 0.0f, im = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) {
        re += (input[i].re * taps[i].re) - (input[i].im * taps[i].im);
        im += (input[i].re * taps[i].im) + (input[i].im * taps[i].re);
    }
    result->re = re;
    result->im = im;
}

static inline void volk_32f_x2_dot_prod_32f(float* result, const float* input, const float* taps, unsigned int num_points) {
    float acc = 0.0f;
    for (unsigned int i = 0; i < num_points; i++) { acc += input[i] * taps[i]; }
    *result = acc;
}

#endif  // SDRPP_STANDIN_VOLK_SIMD
