// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core <dsp/types.h>.
//
// The reference plugin (/root/reference/src/dsp/*.h) includes SDR++ core headers
// that are NOT vendored in the reference tree.  This file
// is a from-scratch restatement of the semantics listed in SURVEY.md Appendix
// A.1, written so that the reference's own src/dsp/*.cpp compile unmodified.
// It is only ever used to build oracle/_ref (the "Oracle A" checker) and the
// host-wrapper test; the product never includes it.
#pragma once
#include <math.h>
#include <stdint.h>

#define FL_M_PI 3.1415926535f
#define DB_M_PI 3.14159265358979323846
#define DB_M_SQRT2 1.41421356237309504880

namespace dsp {
    // interleaved float32 IQ, 8 bytes.
    struct complex_t {
        float re;
        float im;

        complex_t operator*(const float b) const { return complex_t{ re * b, im * b }; }
        complex_t operator/(const float b) const { return complex_t{ re / b, im / b }; }
        // (re*b.re - im*b.im, im*b.re + re*b.im) -- A.1
        complex_t operator*(const complex_t& b) const {
            return complex_t{ (re * b.re) - (im * b.im), (im * b.re) + (re * b.im) };
        }
        complex_t operator+(const complex_t& b) const { return complex_t{ re + b.re, im + b.im }; }
        complex_t operator-(const complex_t& b) const { return complex_t{ re - b.re, im - b.im }; }
        complex_t& operator+=(const complex_t& b) { re += b.re; im += b.im; return *this; }
        complex_t& operator-=(const complex_t& b) { re -= b.re; im -= b.im; return *this; }
        complex_t conj() const { return complex_t{ re, -im }; }

        float phase() const { return atan2f(im, re); }
        float amplitude() const { return sqrtf((re * re) + (im * im)); }
        // A.1: a=|re|, b=|im|; a>b ? a+0.4b : b+0.4a
        float fastAmplitude() const {
            float re_abs = fabsf(re);
#ifdef SDRPP_STANDIN_FASTAMP_RE_ONLY
            // Variant kept for sensitivity studies only (see DESIGN.md "open
            // questions about upstream"): both operands taken from |re|.
            float im_abs = fabsf(re);
#else
            float im_abs = fabsf(im);
#endif
            if (re_abs > im_abs) { return re_abs + 0.4f * im_abs; }
            return im_abs + 0.4f * re_abs;
        }
    };

    struct stereo_t {
        float l;
        float r;
    };
}
