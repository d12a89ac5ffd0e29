// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core <dsp/math/*.h>
// (step, phasor, sinc, hzToRads); restates SURVEY.md Appendix A.1.
#pragma once
#include <math.h>
#include "../types.h"

namespace dsp::math {
    template <class T>
    inline T step(T x) { return (x > 0.0) ? 1.0 : -1.0; }

    inline complex_t phasor(float x) {
        complex_t c = { cosf(x), sinf(x) };
        return c;
    }

    // unnormalised sinc, double (A.1)
    inline double sinc(double x) { return (x == 0.0) ? 1.0 : (sin(x) / x); }

    inline double hzToRads(double freq, double samplerate) { return 2.0 * DB_M_PI * (freq / samplerate); }

    template <class T>
    inline T normalizePhase(T diff) {
        if (diff > FL_M_PI) { diff -= 2.0f * FL_M_PI; }
        else if (diff <= -FL_M_PI) { diff += 2.0f * FL_M_PI; }
        return diff;
    }
}
