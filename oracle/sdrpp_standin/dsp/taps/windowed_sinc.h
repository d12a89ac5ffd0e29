// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/taps/windowed_sinc.h>; restates SURVEY.md Appendix A.6.
#pragma once
#include <type_traits>
#include "../processor.h"
#include "../math/step.h"
#include "../window/nuttall.h"

namespace dsp::taps {
    template <class T, typename Func>
    inline tap<T> windowedSinc(int count, double omega, Func window, double norm = 1.0) {
        tap<T> taps = taps::alloc<T>(count);
        double half = (double)count / 2.0;
        double corr = norm * omega / DB_M_PI;
        for (int i = 0; i < count; i++) {
            double t = (double)i - half + 0.5;
            if constexpr (std::is_same_v<T, float>) {
                taps.taps[i] = (float)(math::sinc(t * omega) * window(t - half, count) * corr);
            }
            else {
                complex_t c = { (float)(math::sinc(t * omega) * window(t - half, count) * corr), 0.0f };
                taps.taps[i] = c;
            }
        }
        return taps;
    }
}
