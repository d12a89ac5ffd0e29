// TEST INFRASTRUCTURE (oracle side) -- stand-in for VOLK's dot products
// (SURVEY.md Appendix A.8).  Real VOLK picks a SIMD kernel at run time, so the
// reference's summation order is machine dependent; Oracle A pins it to the
// plain ascending-index order of VOLK's *_generic kernels, compiled with
// -ffp-contract=off (one rounding per multiply and per add).
#pragma once
#include <stdint.h>

typedef struct { float re, im; } lv_32fc_t;

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
